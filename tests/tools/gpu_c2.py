"""Ad-hoc GPU check of the 1M-triangle Standard-Surface scene: build, traversal parity,
image parity on a crop, throughput.  (Development tool; the tests and bench.py are the
maintained versions of these checks.)"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from fredholm_b200 import Renderer, Camera, DeviceLayers, scenes, api
from oracle import binding as ob

small = "--small" in sys.argv
t0 = time.time()
s = scenes.standard_surface_scene(128, 64) if small else scenes.standard_surface_scene()
print("scene: %d faces, %d verts, gen %.1fs" % (s.n_faces, len(s.vertices), time.time() - t0))
L = scenes.STANDARD_LIGHTING
C = scenes.STANDARD_CAMERA
cam_t = api.camera_walk(C["origin"], 0.0, 150.0, 0, 0.0)
cam_o = ob.camera_walk(C["origin"], 0.0, 150.0, 0, 0.0)
print("camera mirror max abs diff", np.abs(cam_t - cam_o).max())
cam = Camera(cam_o, C["fov"], C["F"], C["focus"])

r = Renderer(0)
t0 = time.time(); r.set_scene(s); print("upload %.2fs" % (time.time() - t0))
r.build_accel(); print("accel", r.accel_info())
r.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
r.load_arhosek_sky(L["turbidity"], L["albedo"])

o = ob.Oracle()
t0 = time.time(); o.set_scene(s); o.build_accel(); print("oracle bvh %.1fs" % (time.time() - t0))
o.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
o.load_arhosek_sky(L["turbidity"], L["albedo"])

# ---- traversal parity on primary rays ----
W, H = 480, 270
o.set_resolution(W, H); r.set_resolution(W, H)
rays = o.primary_rays(cam, 0).reshape(-1, 6)
rays_g = r.primary_rays(cam, 0).reshape(-1, 6)
print("primary ray max abs diff (gpu vs oracle)", np.abs(rays - rays_g).max())
t0 = time.time(); ids_o, tuv_o = o.trace_closest(rays); to = time.time() - t0
ids_g, tuv_g, cnt = r.trace_closest(rays, counters=True)
same = (ids_g == ids_o).all(axis=1)
print("primary hits: oracle %.2fs; identical ids %.6f; hit frac %.3f; tuv bit-equal where same: %s; nodes/ray %.1f tris/ray %.1f"
      % (to, same.mean(), (ids_o[:, 0] != 0xffffffff).mean(),
         np.array_equal(tuv_g[same].view(np.uint32), tuv_o[same].view(np.uint32)), cnt[0] / len(rays), cnt[1] / len(rays)))
if not same.all():
    bad = np.nonzero(~same)[0][:5]
    for b in bad: print("  mismatch", b, ids_g[b], ids_o[b], tuv_g[b], tuv_o[b])
# incoherent rays
rng = np.random.default_rng(3)
oo = rng.uniform(-15, 15, (100000, 3)).astype(np.float32); oo[:, 1] = rng.uniform(0.5, 6, 100000)
dd = rng.normal(size=(100000, 3)).astype(np.float32); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
rr = np.concatenate([oo, dd], 1)
ids_o2, tuv_o2 = o.trace_closest(rr)
ids_g2, tuv_g2, cnt2 = r.trace_closest(rr, counters=True)
same2 = (ids_g2 == ids_o2).all(axis=1)
print("random rays: identical ids %.6f; tuv equal %s; nodes/ray %.1f tris/ray %.1f" % (
    same2.mean(), np.array_equal(tuv_g2[same2].view(np.uint32), tuv_o2[same2].view(np.uint32)),
    cnt2[0] / len(rr), cnt2[1] / len(rr)))

# ---- image parity on the low-res frame ----
spp, depth = 16, 10
t0 = time.time(); ref, secs = o.render_canonical(cam, (0, 0, 0), spp, depth, n_threads=8); print("oracle render %.1fs (%.1f cpu-s)" % (time.time() - t0, secs), o.ray_counts())
layers = DeviceLayers(W, H)
r.reset_statistics()
t0 = time.time(); r.render(cam, (0, 0, 0), layers, spp, depth); r.wait(); tg = time.time() - t0
st = r.statistics(); print("gpu render %.3fs" % tg, st)
got = layers.download("beauty")
ref_b = ref["beauty"]
def rel_mse(x, ref): return float(np.mean((x - ref) ** 2 / (ref ** 2 + 1e-2)))
print("relMSE beauty", rel_mse(got[..., :3].astype(np.float64), ref_b[..., :3].astype(np.float64)),
      "means", got[..., :3].mean(), ref_b[..., :3].mean(), "nan", np.isnan(got).sum())
d_g, d_o = layers.download("depth"), ref["depth"]
print("depth agree frac", np.isclose(d_g, d_o, rtol=1e-5).mean())

# ---- throughput at 1080p ----
if not small:
    W, H = 1920, 1080
    r.set_resolution(W, H)
    big = DeviceLayers(W, H, names=("beauty",))
    for spp in (4, 16):
        r.reset_statistics(); r.init_render_states(); big.clear()
        api.lib().fr_device_synchronize()
        t0 = time.time(); r.render(cam, (0, 0, 0), big, spp, 10); r.wait(); t = time.time() - t0
        st = r.statistics()
        print("1080p %d spp: %.3fs  %.1f Mpaths/s  %.1f Mrays/s  launches %d" % (spp, t, st["paths"] / t / 1e6, st["rays"] / t / 1e6, st["kernel_launches"]))
