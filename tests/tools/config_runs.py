"""Measured runs of BASELINE.json's five configurations on one B200.
Prints one JSON object per configuration; `python tests/tools/config_runs.py [c1 c2 c3 c4 c5] > profiles/...`.

  C1  Cornell box 256x256, 16 spp, depth 8: throughput + relMSE vs the reference integrator (oracle, whole image)
      + the oracle's single-thread rate (the reported CPU baseline of 8(d)).
  C2  1920x1080, 1 Mi triangles, 64 spp, depth 10: bench.py is the maintained measurement; here only the
      relMSE of a 64-spp window against the oracle.
  C3  the same frame at 4096 spp on ONE GPU (the N>1 runs are bench.py under torchrun): throughput and the
      north-star image gate -- relMSE <= 1e-3 against the oracle at 4096 spp on a window of the frame (the
      sampler's 32-bit index wrap at n_spp >= 2071, pt.cu:383, is inside this run).
  C4  52 428 800-triangle instanced scene, diffuse, depth 16, white background: BVH build time + traversal.
  C5  3840x2160 textured scene, 48-frame camera dolly, 16 spp, depth 5, denoise + bloom / CA / tone map through
      the batch driver (fr_batch_run), frames kept in host memory.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fredholm_b200 import Camera, DeviceLayers, Renderer, api, scenes  # noqa: E402
from oracle import binding as ob  # noqa: E402  (checker only)

NT = os.cpu_count() or 1


def rel_mse(x, ref):
    x, ref = x.astype(np.float64), ref.astype(np.float64)
    return float(np.mean((x - ref) ** 2 / (ref ** 2 + 1e-2)))


def timed_render(r, cam, bg, layers, spp, depth, reps=1):
    r.reset_statistics()
    e0 = r.record_event()
    for _ in range(reps):
        layers.clear()
        r.init_render_states()
        r.render(cam, bg, layers, spp, depth)
    e1 = r.record_event()
    r.wait()
    ms = api.event_elapsed_ms(e0, e1) / reps
    st = r.statistics()
    return ms, dict(ms_per_frame=ms, mpaths_per_s=st["paths"] / reps / ms / 1e3, mrays_per_s=st["rays"] / reps / ms / 1e3,
                    rays_per_path=st["rays"] / max(st["paths"], 1))


def standard(r, o=None, scene=None):
    s = scene if scene is not None else scenes.standard_surface_scene()
    L = scenes.STANDARD_LIGHTING
    c = scenes.STANDARD_CAMERA
    cam = Camera(api.camera_walk(c["origin"], 0.0, 150.0, 0, 0.0), c["fov"], c["F"], c["focus"])
    r.set_scene(s)
    r.build_accel()
    r.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
    r.load_arhosek_sky(L["turbidity"], L["albedo"])
    if o is not None:
        o.set_scene(s)
        o.build_accel()
        o.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
        o.load_arhosek_sky(L["turbidity"], L["albedo"])
    return s, cam


def c1():
    W = H = 256
    spp, depth = 16, 8
    s = scenes.cornell_box()
    c = scenes.CORNELL_CAMERA
    cam = Camera(api.camera_walk(c["origin"], 0.0, 0.0, 0, 0.0), c["fov"], c["F"], c["focus"])
    r = Renderer(0)
    r.set_scene(s)
    r.build_accel()
    r.set_resolution(W, H)
    lay = DeviceLayers(W, H)
    timed_render(r, cam, (0, 0, 0), lay, spp, depth)
    _, res = timed_render(r, cam, (0, 0, 0), lay, spp, depth, reps=20)
    got = lay.download("beauty")
    o = ob.Oracle()
    o.set_scene(s)
    o.set_resolution(W, H)
    t0 = time.time()
    ref, cpu_s = o.render_canonical(cam, (0, 0, 0), spp, depth, n_threads=1)
    wall = time.time() - t0
    rc = o.ray_counts()
    res.update(config="C1 Cornell 256x256, %d tris, 16 spp, depth 8" % s.n_faces, relmse_vs_oracle=rel_mse(got[..., :3], ref["beauty"][..., :3]),
               oracle_single_thread=dict(seconds=wall, mpaths_per_s=W * H * spp / wall / 1e6, mrays_per_s=rc["rays"] / wall / 1e6))
    r.close()
    return res


def c2():
    W, H = 1920, 1080
    r = Renderer(0)
    o = ob.Oracle()
    s, cam = standard(r, o)
    r.set_resolution(W, H)
    o.set_resolution(W, H)
    lay = DeviceLayers(W, H, names=("beauty",))
    ms, res = timed_render(r, cam, (0, 0, 0), lay, 64, 10)
    got = lay.download("beauty")
    win = (840, 472, 1080, 608)
    t0 = time.time()
    ref, _ = o.render_canonical(cam, (0, 0, 0), 64, 10, window=win, n_threads=NT)
    x0, y0, x1, y1 = win
    res.update(config="C2 1920x1080, %d tris, 64 spp, depth 10 (single cold frame; bench.py is the measurement)" % s.n_faces,
               window=win, relmse_vs_oracle_window=rel_mse(got[y0:y1, x0:x1, :3], ref["beauty"][y0:y1, x0:x1, :3]),
               oracle_window_seconds=time.time() - t0, oracle_threads=NT)
    r.close()
    return res


def c3():
    W, H = 1920, 1080
    spp = 4096
    r = Renderer(0)
    o = ob.Oracle()
    s, cam = standard(r, o)
    r.set_resolution(W, H)
    o.set_resolution(W, H)
    lay = DeviceLayers(W, H, names=("beauty",))
    ms, res = timed_render(r, cam, (0, 0, 0), lay, spp, 10)
    got = lay.download("beauty")
    # a window that straddles sky, spheres and terrain; 72x40 px x 4096 spp = 11.8 M paths on the host
    win = (924, 520, 996, 560)
    t0 = time.time()
    ref, _ = o.render_canonical(cam, (0, 0, 0), spp, 10, window=win, n_threads=NT)
    x0, y0, x1, y1 = win
    # the same window at 64 spp against the 4096-spp oracle: how much of the tolerance is Monte-Carlo noise
    lay.clear()
    r.init_render_states()
    r.render(cam, (0, 0, 0), lay, 64, 10)
    r.wait()
    low = lay.download("beauty")
    res.update(config="C3 (1 GPU) 1920x1080, %d tris, 4096 spp, depth 10" % s.n_faces, seconds=ms / 1e3, window=win,
               relmse_4096_vs_oracle_4096_window=rel_mse(got[y0:y1, x0:x1, :3], ref["beauty"][y0:y1, x0:x1, :3]),
               relmse_64_vs_oracle_4096_window=rel_mse(low[y0:y1, x0:x1, :3], ref["beauty"][y0:y1, x0:x1, :3]),
               oracle_window_seconds=time.time() - t0, oracle_threads=NT, finite=bool(np.isfinite(got).all()))
    r.close()
    return res


def c4():
    W, H = 1920, 1080
    t0 = time.time()
    s = scenes.instanced_scene()
    gen_s = time.time() - t0
    c = scenes.INSTANCED_CAMERA
    cam = Camera(api.camera_walk(c["origin"], 0.0, 100.0, 0, 0.0), c["fov"], c["F"], c["focus"])
    r = Renderer(0)
    t0 = time.time()
    r.set_scene(s)
    upload_s = time.time() - t0
    t0 = time.time()
    r.build_accel()
    build_wall = time.time() - t0
    info = r.accel_info()
    r.build_accel()
    info2 = r.accel_info()
    r.set_resolution(W, H)
    lay = DeviceLayers(W, H, names=("beauty", "depth"))
    timed_render(r, cam, (1, 1, 1), lay, 16, 16)       # warm-up at the measured wave size (allocations)
    ms, res = timed_render(r, cam, (1, 1, 1), lay, 16, 16, reps=2)
    img = lay.download("beauty")
    dep = lay.download("depth")
    # traversal statistics on primary + incoherent rays
    rays = r.primary_rays(cam, 0).reshape(-1, 6)[::7]
    ids, tuv, cnt = r.trace_closest(rays, counters=True)
    hit = ids[:, 0] != 0xffffffff
    rng = np.random.default_rng(5)
    n = 200000
    oo = rng.uniform(-150, 150, (n, 3)).astype(np.float32)
    oo[:, 1] = rng.uniform(12, 30, n)
    dd = rng.normal(size=(n, 3)).astype(np.float32)
    dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    _, _, cnt2 = r.trace_closest(np.concatenate([oo, dd], 1), counters=True)
    res.update(config="C4 1920x1080, %d tris in %d sub-meshes (instanced), diffuse, depth 16, white background, 16 spp"
               % (s.n_faces, len(s.submesh_offsets)), scene_generation_s=gen_s, upload_s=upload_s,
               bvh=dict(nodes=info["n_nodes"], depth=info["depth"], bytes=info["bytes"], build_ms_first=info["build_ms"],
                        build_ms_second=info2["build_ms"], build_wall_s_first=build_wall,
                        mtris_per_s=s.n_faces / info2["build_ms"] / 1e3),
               primary_hit_fraction=float(hit.mean()), primary_nodes_per_ray=cnt[0] / len(rays), primary_tris_per_ray=cnt[1] / len(rays),
               random_nodes_per_ray=cnt2[0] / n, random_tris_per_ray=cnt2[1] / n,
               image_mean=float(img[..., :3].mean()), finite=bool(np.isfinite(img).all()), depth_hit_fraction=float((dep > 0).mean()))
    r.close()
    return res


def c5(n_frames=48):
    W, H = 3840, 2160
    t0 = time.time()
    s = scenes.textured_scene(tex_res=4096)
    gen_s = time.time() - t0
    r = Renderer(0)
    _, cam = standard(r, scene=s)
    c = scenes.STANDARD_CAMERA
    o0 = np.array(c["origin"], np.float32)
    path = np.stack([api.camera_walk(o0 + np.array([0.05 * i, 0.0, -0.08 * i], np.float32), 0.1 * i, 150.0, 0, 0.0)
                     for i in range(n_frames)])
    r.set_resolution(W, H)
    kw = dict(camera_path=path, keep_frames=True, denoise=True, use_bloom=True, bloom_threshold=2.0, bloom_sigma=5.0,
              ISO=80.0, chromatic_aberration=1.0, fps=24.0, max_time=1e9, animate=False)
    r.batch_run(cam, W, H, 16, 5, 2, **kw)                      # warm-up (allocations)
    r.reset_statistics()
    t0 = time.time()
    recs, frames, info = r.batch_run(cam, W, H, 16, 5, n_frames, **kw)
    wall = time.time() - t0
    st = r.statistics()
    mean = lambda k: float(np.mean([x[k] for x in recs]))
    res = dict(config="C5 3840x2160 textured (3 x 4096^2 RGBA8), %d tris, %d frames, 16 spp, depth 5, denoise + bloom/CA/tone map, "
               "batch driver, frames to host RGBA8" % (s.n_faces, info["n_frames"]), scene_generation_s=gen_s,
               frames=info["n_frames"], wall_s=info["wall_s"], python_wall_s=wall, frames_per_s=info["n_frames"] / info["wall_s"],
               mpaths_per_s_e2e=st["paths"] / info["wall_s"] / 1e6, mrays_per_s_e2e=st["rays"] / info["wall_s"] / 1e6,
               render_ms=mean("render_ms"), denoise_ms=mean("denoise_ms"), post_ms=mean("post_ms"), transfer_ms=mean("transfer_ms"),
               frame_mean_rgb=float(frames[..., :3].mean()), distinct_frames=len({f[::16, ::16].tobytes() for f in frames}))
    r.close()
    return res


if __name__ == "__main__":
    todo = [a.lower() for a in sys.argv[1:]] or ["c1", "c2", "c3", "c4", "c5"]
    for name in todo:
        t0 = time.time()
        out = globals()[name]()
        out["tool_wall_s"] = time.time() - t0
        print(json.dumps({name: out}), flush=True)
