import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from fredholm_b200 import Camera, DeviceLayers, Renderer, api, scenes
from oracle import binding as ob
W = H = 64
s = scenes.standard_surface_scene(48, 24, sphere_res=(12, 6))
c = scenes.STANDARD_CAMERA; L = scenes.STANDARD_LIGHTING
cam = Camera(api.camera_walk(c["origin"], 0.0, 30.0, 0, 0.0), c["fov"], c["F"], c["focus"])
r = Renderer(0); o = ob.Oracle()
for x in (r, o):
    x.set_scene(s); x.build_accel(); x.set_resolution(W, H)
    x.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"]); x.load_arhosek_sky(L["turbidity"], L["albedo"])
def ours(q):
    r.set_single_launch(q); lay = DeviceLayers(W, H); r.init_render_states(); r.render(cam, (0, 0, 0), lay, 8, 5); r.wait()
    return lay.download("beauty")[..., :3].copy(), lay.download("depth").copy()
a, da = ours(False); b, db = ours(True)
o.init_render_states(); ref = o.new_layers(); o.render(cam, (0, 0, 0), ref, 8, 5, n_threads=16)
o.init_render_states(); refc, _ = o.render_canonical(cam, (0, 0, 0), 8, 5, n_threads=16)
q = ref["beauty"][..., :3]; qc = refc["beauty"][..., :3]
print("ours canon vs quirk: differing pixels", (np.abs(a - b).max(-1) > 1e-4).sum())
print("oracle canon vs single: differing pixels", (np.abs(q - qc).max(-1) > 1e-4).sum())
print("ours quirk vs oracle single: differing", (np.abs(b - q).max(-1) > 1e-3).sum(), " ours canon vs oracle canon:", (np.abs(a - qc).max(-1) > 1e-3).sum())
d = np.argwhere(np.abs(q - qc).max(-1) > 1e-4)
for y, x in d[:5]:
    print(y, x, "oracle single", q[y, x], "oracle canon", qc[y, x], "ours quirk", b[y, x], "ours canon", a[y, x])
print("depth layers differ ours:", (np.abs(da - db) > 1e-4).sum(), "oracle:", (np.abs(ref["depth"] - refc["depth"]) > 1e-4).sum())
