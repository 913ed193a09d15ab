import sys, time
sys.path.insert(0, ".")
from fredholm_b200 import Renderer, scenes
s = scenes.standard_surface_scene()
r = Renderer(0); r.set_scene(s)
for i in range(4):
    t=time.time(); r.build_accel(); print("build", i, r.accel_info()["build_ms"], "wall", (time.time()-t)*1e3, file=sys.stderr)
