"""Short render of the benchmark scene for ncu captures (not a benchmark: numbers printed
under a profiler are never bench values)."""
import argparse, sys, time
sys.path.insert(0, ".")
from fredholm_b200 import Renderer, Camera, DeviceLayers, scenes, api

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=4)
ap.add_argument("--depth", type=int, default=10)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--spw", type=int, default=0)
ap.add_argument("--stats-out", default="")
a = ap.parse_args()
s = scenes.standard_surface_scene()
L = scenes.STANDARD_LIGHTING; C = scenes.STANDARD_CAMERA
cam = Camera(api.camera_walk(C["origin"], 0.0, 150.0, 0, 0.0), C["fov"], C["F"], C["focus"])
r = Renderer(0); r.set_scene(s); r.build_accel()
r.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"]); r.load_arhosek_sky(L["turbidity"], L["albedo"])
r.set_resolution(a.width, a.height)
if a.spw: r.set_samples_per_warp(a.spw)
lay = DeviceLayers(a.width, a.height, names=("beauty",))
for f in range(a.frames):
    r.reset_statistics(); t0 = time.time(); r.render(cam, (0, 0, 0), lay, a.spp, a.depth); r.wait()
    st = r.statistics()
    print("frame %d: %.3f s" % (f, time.time() - t0), st)
if a.stats_out:
    import json
    st["spp"] = a.spp; st["depth"] = a.depth
    json.dump(st, open(a.stats_out, "w"))
