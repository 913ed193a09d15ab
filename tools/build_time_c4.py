import sys, time
sys.path.insert(0, ".")
from fredholm_b200 import Renderer, scenes
s = scenes.instanced_scene()
r = Renderer(0); r.set_scene(s)
for i in range(7):
    r.build_accel(); print("build", i, r.accel_info()["build_ms"], file=sys.stderr)
