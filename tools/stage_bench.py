"""Per-stage device times of the benchmark scene (development tool; bench.py is the
maintained measurement).  python tools/stage_bench.py [--spp 16] [--reps 3] [--wave N]"""
import argparse, sys, time
sys.path.insert(0, ".")
from fredholm_b200 import Renderer, Camera, DeviceLayers, scenes, api

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--depth", type=int, default=10)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--wave", type=int, default=0)
ap.add_argument("--count", action="store_true", help="also one frame with the traversal counters")
a = ap.parse_args()
s = scenes.standard_surface_scene()
L = scenes.STANDARD_LIGHTING; C = scenes.STANDARD_CAMERA
cam = Camera(api.camera_walk(C["origin"], 0.0, 150.0, 0, 0.0), C["fov"], C["F"], C["focus"])
r = Renderer(0); r.set_scene(s); r.build_accel(); r.build_accel()
print("accel", r.accel_info())
r.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"]); r.load_arhosek_sky(L["turbidity"], L["albedo"])
W, H = 1920, 1080
r.set_resolution(W, H)
if a.wave: r.set_max_wave_paths(a.wave)
lay = DeviceLayers(W, H, names=("beauty",))
r.render(cam, (0, 0, 0), lay, a.spp, a.depth); r.wait()
r.set_stage_timing(True); r.stage_times(); r.reset_statistics()
e0 = r.record_event()
for _ in range(a.reps):
    lay.clear(); r.init_render_states()
    r.render(cam, (0, 0, 0), lay, a.spp, a.depth)
e1 = r.record_event(); r.wait()
ms = api.event_elapsed_ms(e0, e1) / a.reps
st = r.statistics(); stages = r.stage_times()
print("frame %.2f ms (with event overhead)  %.1f Mpaths/s  %.1f Mrays/s  mean %.5f" % (
    ms, st["paths"] / a.reps / ms / 1e3, st["rays"] / a.reps / ms / 1e3, lay.download("beauty")[..., :3].mean()))
print("  ".join("%s %.2f" % (k, v[0] / a.reps) for k, v in stages.items()))
r.set_stage_timing(False)
e0 = r.record_event()
for _ in range(a.reps):
    r.init_render_states()
    r.render(cam, (0, 0, 0), lay, a.spp, a.depth)
e1 = r.record_event(); r.wait()
ms = api.event_elapsed_ms(e0, e1) / a.reps
print("untimed-stages frame %.2f ms  %.1f Mpaths/s  %.1f Mrays/s" % (ms, st["paths"] / a.reps / ms / 1e3, st["rays"] / a.reps / ms / 1e3))
if a.count:
    r.set_traversal_counting(True); r.reset_statistics(); r.init_render_states()
    r.render(cam, (0, 0, 0), lay, a.spp, a.depth); r.wait()
    st = r.statistics(); tc = r.traversal_counters(); r.set_traversal_counting(False)
    for k in ("radiance", "shadow", "light"):
        nr = max(1, st["rays_" + k])
        if k == "shadow" and tc["light"][0] == 0: nr += st["rays_light"]  # no emitters: MIS rays are visibility rays
        print("  %-8s %.2f nodes/ray  %.2f tris/ray" % (k, tc[k][0] / nr, tc[k][1] / nr))
