import sys, numpy as np
sys.path.insert(0, ".")
from fredholm_b200 import Camera, Renderer, api, scenes
from oracle import binding as ob
s = scenes.instanced_scene(n_instances=300, mesh_res=(24, 12), terrain_res=64)
c = scenes.INSTANCED_CAMERA
cam = Camera(api.camera_walk(c["origin"], 0.0, 100.0, 0, 0.0), c["fov"], c["F"], c["focus"])
two = Renderer(0); two.set_scene(s); two.build_accel(); two.set_resolution(480, 270)
o = ob.Oracle(); o.set_scene(s); o.build_accel(); o.set_resolution(480, 270)
rays = o.primary_rays(cam, 0).reshape(-1, 6)
ids_t, tuv_t = two.trace_closest(rays); ids_o, tuv_o = o.trace_closest(rays)
same = (ids_t == ids_o).all(axis=1); hit = same & (ids_o[:, 0] != 0xffffffff)
rel = np.abs(tuv_t[hit, 0] - tuv_o[hit, 0]) / tuv_o[hit, 0]
print("ids same", same.mean(), "hits", hit.sum())
print("t rel err: median %.2e p99 %.2e p9999 %.2e max %.2e" % (np.median(rel), np.quantile(rel, .99), np.quantile(rel, .9999), rel.max()))
w = np.argsort(rel)[-5:]
idx = np.where(hit)[0][w]
for i in idx:
    print("inst", ids_o[i], "t two", tuv_t[i], "t ref", tuv_o[i], "rel", abs(tuv_t[i,0]-tuv_o[i,0])/tuv_o[i,0])
uv = np.abs(tuv_t[hit, 1:] - tuv_o[hit, 1:]); print("uv abs err max", uv.max(), "p9999", np.quantile(uv, .9999))
inst = ids_o[hit, 0]
for k in (0, 1):
    sel = (inst == 0) if k == 0 else (inst != 0)
    print("terrain" if k == 0 else "instances", "rel max %.2e" % rel[sel].max(), "n", sel.sum())
