import sys, time
import numpy as np
sys.path.insert(0, ".")
from fredholm_b200 import Renderer, Camera, DeviceLayers, scenes, api
from oracle import binding as ob
s = scenes.standard_surface_scene(128, 64)
L = scenes.STANDARD_LIGHTING; C = scenes.STANDARD_CAMERA
cam = Camera(ob.camera_walk(C["origin"], 0.0, 150.0, 0, 0.0), C["fov"], C["F"], C["focus"])
r = Renderer(0); r.set_scene(s); r.build_accel()
o = ob.Oracle(); o.set_scene(s); o.build_accel()
for x in (r, o):
    x.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"]); x.load_arhosek_sky(L["turbidity"], L["albedo"])
W, H = 480, 270
o.set_resolution(W, H); r.set_resolution(W, H)
rays = o.primary_rays(cam, 0).reshape(-1, 6)
ids, _ = o.trace_closest(rays)
face = np.where(ids[:, 0] != 0xffffffff, s.submesh_offsets[np.minimum(ids[:, 0], len(s.submesh_offsets) - 1)] + ids[:, 1], 0)
mat = np.where(ids[:, 0] != 0xffffffff, s.material_ids[face], 99).reshape(H, W)
for depth in (1, 2, 3, 10):
    o.init_render_states(); r.init_render_states(); o.reset_ray_counts(); r.reset_statistics()
    ref = o.new_layers(); o.render(cam, (0, 0, 0), ref, 1, depth, n_threads=8)
    lay = DeviceLayers(W, H); r.render(cam, (0, 0, 0), lay, 1, depth); r.wait()
    g = lay.download("beauty")[..., :3].astype(np.float64); b = ref["beauty"][..., :3].astype(np.float64)
    rel = np.abs(g - b).max(-1) / (np.abs(b).max(-1) + 1e-3)
    print("depth", depth, "pixels rel>1e-3: %.4f  rel>1e-1: %.4f" % ((rel > 1e-3).mean(), (rel > 1e-1).mean()),
          "rays gpu", r.statistics()["rays_radiance"], "ref", o.ray_counts()["rays_radiance"], "mean", g.mean(), b.mean())
    for m in sorted(set(mat.reshape(-1))):
        sel = mat == m
        print("    mat %2d: n=%6d  frac rel>1e-3 %.4f  mean gpu %.4f ref %.4f" % (m, sel.sum(), (rel[sel] > 1e-3).mean(), g[sel].mean(), b[sel].mean()))
    if depth == 2:
        ys, xs = np.nonzero(rel > 1e-2)
        for y, x in list(zip(ys, xs))[:6]:
            print("      px", x, y, "mat", mat[y, x], "gpu", g[y, x], "ref", b[y, x])
