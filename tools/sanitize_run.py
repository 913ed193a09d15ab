"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): PLOC build, all wavefront stages with
textures and emitters, post-process.  python tools/sanitize_run.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
from fredholm_b200 import Camera, DeviceLayers, Renderer, api, scenes

r = Renderer(0)
for name, s, cam_def, lights in (("cornell", scenes.cornell_box(), scenes.CORNELL_CAMERA, False),
                                 ("standard", scenes.standard_surface_scene(48, 24, sphere_res=(12, 6)), scenes.STANDARD_CAMERA, True)):
    r.set_scene(s)
    r.build_accel()
    if lights:
        L = scenes.STANDARD_LIGHTING
        r.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
        r.load_arhosek_sky(L["turbidity"], L["albedo"])
    W, H = 96, 64
    r.set_resolution(W, H)
    c = cam_def
    cam = Camera(api.camera_walk(c["origin"], 0.0, 150.0 if lights else 0.0, 0, 0.0), c["fov"], c["F"], c["focus"])
    lay = DeviceLayers(W, H)
    r.render(cam, (0, 0, 0), lay, 4, 6)
    r.wait()
    img = lay.download("beauty")
    print(name, r.accel_info()["n_nodes"], float(img[..., :3].mean()), r.statistics()["rays"])
    rays = np.random.default_rng(1).normal(size=(5000, 6)).astype(np.float32)
    ids, _ = r.trace_closest(rays)
    print("  batch hits", int((ids[:, 0] != 0xffffffff).sum()))
r.close()

# ---- round 2: two-level structure (instances, refit + rebuild), wave overlap, sample groups, GPU mesh preparation,
#      sharded render with a one-rank communicator ----
import os, tempfile
s = scenes.instanced_scene(n_instances=40, mesh_res=(12, 6), terrain_res=16)
r = Renderer(0)
r.set_accel_mode("two_level")
r.set_scene(s)
r.build_accel()
c = scenes.INSTANCED_CAMERA
cam = Camera(api.camera_walk(c["origin"], 0.0, 100.0, 0, 0.0), c["fov"], c["F"], c["focus"])
W, H = 96, 64
r.set_resolution(W, H)
lay = DeviceLayers(W, H)
r.set_samples_per_warp(4)
r.set_wave_overlap(True)
r.set_max_wave_paths(2 * ((W + 3) // 4) * ((H + 1) // 2) * 32)      # two waves of one 4-sample group in flight
r.render(cam, (1, 1, 1), lay, 8, 6)
r.wait()
tr = s.transforms.copy().reshape(-1, 4, 4)
tr[3, 3, 1] += 4.0
r.set_transforms(tr.reshape(-1, 16))                                  # refit
tr[1:, 3, 0] *= 9.0
r.set_transforms(tr.reshape(-1, 16))                                  # scattered: rebuild
info = r.accel_info()
lay.clear(); r.init_render_states()
r.render(cam, (1, 1, 1), lay, 4, 6)
r.wait()
rays = np.random.default_rng(2).normal(size=(5000, 6)).astype(np.float32) * np.array([60, 10, 60, 1, 1, 1], np.float32)
ids, _ = r.trace_closest(rays)
print("two-level", info["two_level"], info["tlas_refitted"], float(lay.download("beauty")[..., :3].mean()), int((ids[:, 0] != 0xffffffff).sum()))
r.comm_init(api.comm_unique_id(), 0, 1)
lay.clear(); r.init_render_states()
r.render_sharded(cam, (1, 1, 1), lay, 16, 4)
r.wait()
print("sharded", float(lay.download("beauty")[..., :3].mean()))
r.comm_destroy()
r.close()
with tempfile.TemporaryDirectory() as tmp:
    os.environ["FRD_GPU_MESH_PREP"] = "1"
    for attrs in (True, False):
        p = scenes.write_obj(scenes.standard_surface_scene(24, 12, sphere_res=(8, 4)), tmp, "m%d" % attrs, with_attributes=attrs)
        sc = api.Scene(); sc.load_model(p); a = sc.arrays()
        print("mesh prep", attrs, a.n_faces, len(a.vertices))

# ---- wave compaction: beauty-only render in waves of one sample, every compaction depth (depth 1 leaves more paths
#      alive than the straggler set holds -> waves finish in place; depth 2 fills it in mid-pass), ten waves = two passes
s = scenes.standard_surface_scene(48, 24, sphere_res=(12, 6))
r = Renderer(0)
r.set_scene(s)
r.build_accel()
L = scenes.STANDARD_LIGHTING
r.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
r.load_arhosek_sky(L["turbidity"], L["albedo"])
W, H = 96, 64
r.set_resolution(W, H)
c = scenes.STANDARD_CAMERA
cam = Camera(api.camera_walk(c["origin"], 0.0, 150.0, 0, 0.0), c["fov"], c["F"], c["focus"])
lay = DeviceLayers(W, H, names=("beauty",))
imgs = []
for on, depth in ((False, 0), (True, 1), (True, 2), (True, 3)):
    r.set_wave_compaction(on, depth)
    r.set_max_wave_paths(((W + 7) // 8) * ((H + 3) // 4) * 32)
    lay.clear(); r.init_render_states()
    r.render(cam, (0, 0, 0), lay, 10, 8)
    r.wait()
    imgs.append(lay.download("beauty"))
print("wave compaction identical:", all(np.array_equal(imgs[0], im) for im in imgs[1:]), float(imgs[0][..., :3].mean()))
r.close()
