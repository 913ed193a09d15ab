"""Throughput against the number of paths one wave keeps in flight (bench scene, 64 spp): waves of 64, 32, 16, 8, 4
samples, without and with wave compaction (stragglers of up to eight waves finish together), and with two waves in
flight on two streams.  python tools/wave_sweep.py [compaction depth] > profiles/..."""
import json, sys
sys.path.insert(0, ".")
from fredholm_b200 import Renderer, Camera, DeviceLayers, scenes, api
s = scenes.standard_surface_scene()
L = scenes.STANDARD_LIGHTING; C = scenes.STANDARD_CAMERA
cam = Camera(api.camera_walk(C["origin"], 0.0, 150.0, 0, 0.0), C["fov"], C["F"], C["focus"])
W, H, SPP, DEPTH = 1920, 1080, 64, 10
slots = ((W + 7) // 8) * ((H + 3) // 4) * 32
import numpy as np
ref_img = None
DEPTH_MOVE = int(sys.argv[1]) if len(sys.argv) > 1 else 0
CASES = [(w, 0, 0) for w in (64, 32, 16, 8, 4)] + [(w, 0, 1) for w in (32, 16, 8, 4, 2)] + [(w, 1, 0) for w in (64, 32, 16, 8)]
if len(sys.argv) > 2: CASES = [c for c in CASES if c[2] == 1]
for wave_spp, overlap, compaction in CASES:
    # with overlap the paths in flight are split over two waves on two streams: "in flight" = wave_spp samples
    r = Renderer(0); r.set_scene(s); r.build_accel(); r.set_wave_overlap(bool(overlap))
    r.set_wave_compaction(bool(compaction), DEPTH_MOVE)
    r.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"]); r.load_arhosek_sky(L["turbidity"], L["albedo"])
    r.set_resolution(W, H); r.set_max_wave_paths(slots * wave_spp)
    lay = DeviceLayers(W, H, names=("beauty",))
    for _ in range(2):
        lay.clear(); r.init_render_states(); r.render(cam, (0, 0, 0), lay, SPP, DEPTH)
    r.wait(); r.reset_statistics()
    e0 = r.record_event()
    for _ in range(3):
        lay.clear(); r.init_render_states(); r.render(cam, (0, 0, 0), lay, SPP, DEPTH)
    e1 = r.record_event(); r.wait()
    ms = api.event_elapsed_ms(e0, e1) / 3
    st = r.statistics()
    img = lay.download("beauty")
    if ref_img is None:
        ref_img = img
    print(json.dumps(dict(samples_in_flight=wave_spp, overlap=overlap, compaction=compaction, identical_image=bool(np.array_equal(img, ref_img)), wave_state_gb=round(r.wave_state_bytes() / 1e9, 2),
                          frame_ms=round(ms, 2), mpaths_per_s=round(st["paths"] / 3 / ms / 1e3, 1), launches_per_frame=st["kernel_launches"] // 3)), flush=True)
    lay.free(); r.close()
