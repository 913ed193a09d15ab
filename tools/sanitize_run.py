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
