"""C4 frame time, repeated (flat tree): is the 16-spp depth-16 frame stable from run to run?"""
import sys, time
sys.path.insert(0, ".")
from fredholm_b200 import Camera, DeviceLayers, Renderer, api, scenes
s = scenes.instanced_scene()
c = scenes.INSTANCED_CAMERA
cam = Camera(api.camera_walk(c["origin"], 0.0, 100.0, 0, 0.0), c["fov"], c["F"], c["focus"])
W, H = 1920, 1080
r = Renderer(0); r.set_scene(s)
for b in range(3):
    r.build_accel(); print("build", b, r.accel_info()["build_ms"], r.accel_info()["n_nodes"], flush=True)
r.set_resolution(W, H)
lay = DeviceLayers(W, H, names=("beauty", "depth"))
r.render(cam, (1, 1, 1), lay, 4, 16); r.wait()
for rep in range(6):
    r.reset_statistics()
    e0 = r.record_event(); lay.clear(); r.init_render_states(); r.render(cam, (1, 1, 1), lay, 16, 16); e1 = r.record_event(); r.wait()
    ms = api.event_elapsed_ms(e0, e1); st = r.statistics()
    print("frame %d: %.1f ms, %.0f Mpaths/s, %.0f Mrays/s" % (rep, ms, st["paths"] / ms / 1e3, st["rays"] / ms / 1e3), flush=True)
r.set_stage_timing(True); r.stage_times()
lay.clear(); r.init_render_states(); r.render(cam, (1, 1, 1), lay, 16, 16); r.wait()
print({k: round(v[0], 1) for k, v in r.stage_times().items()})
