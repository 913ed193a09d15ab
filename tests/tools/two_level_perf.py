"""Config 4 through both acceleration-structure layouts: frame time, rays, traversal work (GPU box tool)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from fredholm_b200 import Camera, DeviceLayers, Renderer, api, scenes
s = scenes.instanced_scene()
c = scenes.INSTANCED_CAMERA
cam = Camera(api.camera_walk(c["origin"], 0.0, 100.0, 0, 0.0), c["fov"], c["F"], c["focus"])
W, H = 1920, 1080
for mode in ("two_level", "flat"):
    r = Renderer(0); r.set_accel_mode(mode); r.set_scene(s); r.build_accel(); r.build_accel()
    info = r.accel_info()
    r.set_resolution(W, H)
    lay = DeviceLayers(W, H, names=("beauty",))
    r.render(cam, (1, 1, 1), lay, 16, 16); r.wait()
    r.set_stage_timing(True); r.stage_times(); r.reset_statistics()
    e0 = r.record_event(); lay.clear(); r.init_render_states(); r.render(cam, (1, 1, 1), lay, 16, 16); e1 = r.record_event(); r.wait()
    ms = api.event_elapsed_ms(e0, e1); st = r.statistics(); stages = r.stage_times()
    rays = r.primary_rays(cam, 0).reshape(-1, 6)[::5]
    _, _, cnt = r.trace_closest(rays, counters=True)
    print(mode, "build %.1f ms, %.0f MB, depth %d | frame %.1f ms, %.0f Mpaths/s, %.2f Grays/s | camera rays: %.1f nodes, %.1f tris per ray | %s"
          % (info["build_ms"], info["bytes"] / 1e6, info["depth"], ms, st["paths"] / ms / 1e3, st["rays"] / ms / 1e6,
             cnt[0] / len(rays), cnt[1] / len(rays), {k: round(v[0], 1) for k, v in stages.items()}), flush=True)
    lay.free(); r.close()
