"""Round-2 experiment: samples-per-warp sweep (FilmGeom slot mapping) + traversal counters on the
benchmark scene.  python tools/exp_spw.py [--spp 64] [--spw 1,2,4,8,16,32]"""
import argparse, json, sys
import numpy as np
sys.path.insert(0, ".")
from fredholm_b200 import Renderer, Camera, DeviceLayers, scenes, api

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--depth", type=int, default=10)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--spw", default="1,2,4,8,16,32")
ap.add_argument("--count", action="store_true")
a = ap.parse_args()
s = scenes.standard_surface_scene()
L = scenes.STANDARD_LIGHTING; C = scenes.STANDARD_CAMERA
cam = Camera(api.camera_walk(C["origin"], 0.0, 150.0, 0, 0.0), C["fov"], C["F"], C["focus"])
r = Renderer(0); r.set_scene(s); r.build_accel()
r.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"]); r.load_arhosek_sky(L["turbidity"], L["albedo"])
W, H = 1920, 1080
r.set_resolution(W, H)
r.set_max_wave_paths(1 << 28)
lay = DeviceLayers(W, H, names=("beauty",))
ref = None
for spw in [int(x) for x in a.spw.split(",")]:
    r.set_samples_per_warp(spw)
    lay.clear(); r.init_render_states()
    r.render(cam, (0, 0, 0), lay, a.spp, a.depth); r.wait()
    img = lay.download("beauty")[..., :3].copy()
    if ref is None:
        ref = img
    same = bool(np.array_equal(img, ref))
    r.set_stage_timing(True); r.stage_times(); r.reset_statistics()
    for _ in range(a.reps):
        lay.clear(); r.init_render_states()
        r.render(cam, (0, 0, 0), lay, a.spp, a.depth)
    r.wait()
    stages = r.stage_times(); st = r.statistics()
    r.set_stage_timing(False)
    e0 = r.record_event()
    for _ in range(a.reps):
        lay.clear(); r.init_render_states()
        r.render(cam, (0, 0, 0), lay, a.spp, a.depth)
    e1 = r.record_event(); r.wait()
    ms = api.event_elapsed_ms(e0, e1) / a.reps
    out = dict(spw=spw, frame_ms=round(ms, 2), mpaths=round(st["paths"] / a.reps / ms / 1e3, 1),
               identical_to_first=same, maxdiff=float(np.abs(img - ref).max()),
               stages={k: round(v[0] / a.reps, 2) for k, v in stages.items()})
    print(json.dumps(out), flush=True)
if a.count:
    for spw in (1,):
        r.set_samples_per_warp(spw)
        r.set_traversal_counting(True); r.reset_statistics()
        lay.clear(); r.init_render_states()
        r.render(cam, (0, 0, 0), lay, a.spp, a.depth); r.wait()
        st = r.statistics(); c = r.traversal_counters()
        r.set_traversal_counting(False)
        rays = dict(radiance=st["rays_radiance"], shadow=st["rays_shadow"], light=st["rays_light"])
        if c["light"][0] == 0:  # no emitters: the MIS rays were traced as visibility rays
            rays["shadow"] += rays["light"]
        print(json.dumps(dict(spw=spw, rays=rays, per_ray={k: (round(c[k][0] / max(rays[k], 1), 2), round(c[k][1] / max(rays[k], 1), 2)) for k in c})), flush=True)
