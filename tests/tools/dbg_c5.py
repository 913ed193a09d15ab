"""Where does the C5 (textured frame) relMSE against the oracle come from?  Per primary-hit material and per
depth; GPU box tool (python tests/tools/dbg_c5.py)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fredholm_b200 import Camera, DeviceLayers, Renderer, api, scenes
from oracle import binding as ob

NT = os.cpu_count() or 1
W, H = 960, 540
s = scenes.textured_scene(tex_res=1024)
c = scenes.STANDARD_CAMERA
cam = Camera(api.camera_walk(c["origin"], 0.0, 150.0, 0, 0.0), c["fov"], c["F"], c["focus"])
L = scenes.STANDARD_LIGHTING
r = Renderer(0); o = ob.Oracle()
for x in (r, o):
    x.set_scene(s); x.build_accel()
    x.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"]); x.load_arhosek_sky(L["turbidity"], L["albedo"])
    x.set_resolution(W, H)
rays = o.primary_rays(cam, 0).reshape(-1, 6)
ids, _ = o.trace_closest(rays)
face = np.where(ids[:, 0] != 0xffffffff, s.submesh_offsets[np.minimum(ids[:, 0], len(s.submesh_offsets) - 1)] + ids[:, 1], 0)
mat = np.where(ids[:, 0] != 0xffffffff, s.material_ids[face], -1).reshape(H, W)
def compare(tag, spp, depth, bg=(0, 0, 0)):
    lay = DeviceLayers(W, H, names=("beauty",))
    r.init_render_states(); r.render(cam, bg, lay, spp, depth); r.wait()
    got = lay.download("beauty")[..., :3].astype(np.float64)
    o.init_render_states()
    ref, _ = o.render_canonical(cam, bg, spp, depth, n_threads=NT)
    ref = ref["beauty"][..., :3].astype(np.float64)
    e = ((got - ref) ** 2 / (ref ** 2 + 1e-2)).mean(axis=2)
    print("%s spp %d depth %d: relMSE %.3e  max pixel %.3e  pixels > 1e-2: %d" % (tag, spp, depth, e.mean(), e.max(), (e > 1e-2).sum()))
    for m in np.unique(mat):
        sel = mat == m
        print("   material %2d: %7d px  relMSE %.3e  outliers %4d  mean got %.4f ref %.4f" % (m, sel.sum(), e[sel].mean(), (e[sel] > 1e-2).sum(), got[sel].mean(), ref[sel].mean()))
    ys, xs = np.unravel_index(np.argsort(e.ravel())[-6:], e.shape)
    for y, x in zip(ys, xs):
        print("   worst (%d,%d) mat %d err %.3e got %s ref %s" % (x, y, mat[y, x], e[y, x], got[y, x], ref[y, x]))
    lay.free()
    return e

e = compare("sun+sky", 4, 1)
# per-sample values at the worst pixels
ys, xs = np.unravel_index(np.argsort(e.ravel())[-8:], e.shape)
r.set_film_mode("sum")
ours = []
for k in range(4):
    lay = DeviceLayers(W, H, names=("beauty",)); r.set_sample_offset(k); r.render(cam, (0, 0, 0), lay, 1, 1); r.wait()
    ours.append(lay.download("beauty")[..., :3].astype(np.float64)); lay.free()
r.set_film_mode("mean")
o.init_render_states(); layers = o.new_layers(); prev = np.zeros((H, W, 3)); theirs = []
for k in range(4):
    o.render(cam, (0, 0, 0), layers, 1, 1, None, NT)
    cur = layers["beauty"][..., :3].astype(np.float64) * (k + 1)
    theirs.append(cur - prev); prev = cur
for y, x in zip(ys, xs):
    print("pixel (%d,%d) material %d" % (x, y, mat[y, x]))
    for k in range(4):
        print("    sample %d ours %s ref %s" % (k, np.round(ours[k][y, x], 4), np.round(theirs[k][y, x], 4)))
for x in (r, o):
    x.clear_arhosek_sky()
compare("sun only (constant black background)", 4, 1)
for x in (r, o):
    x.clear_directional_light(); x.load_arhosek_sky(L["turbidity"], L["albedo"])
compare("sky only", 4, 1)
