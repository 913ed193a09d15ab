#!/usr/bin/env python3
"""Generates the committed golden vectors under tests/golden/ from the HOST ORACLE, i.e.
from the reference's own integrator sources compiled for the CPU (oracle/Makefile).

The reference ships no tests or known-answer vectors, so these
fixtures are "outputs of the reference itself run here": they pin both the oracle
(tests/test_oracle_golden.py, CPU) and the CUDA path (tests/test_gpu_golden.py, GPU box,
where /root/reference does not exist).

    python tests/tools/gen_golden.py          # needs /root/reference (builds oracle/_ref)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from bsdf_cases import make_cases, make_cases_below_horizon  # noqa: E402
from fredholm_b200 import Camera, scenes  # noqa: E402
from oracle import binding as ob  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

SAMPLER_KINDS = "221" + "2" + "12" * 3 + "1" + "2212121" * 6
SAMPLER_POINTS = [(256, 256, 0, 0), (256, 256, 12345, 7), (1920, 1080, 2073599, 63),
                  (1920, 1080, 1000000, 4095), (64, 64, 77, 16), (3840, 2160, 8294399, 2071)]
HOSEK_POINTS = [(3.0, 0.3, 1.2), (2.0, 0.1, 0.2), (6.5, 0.8, 0.7), (10.0, 0.0, 1.5)]


def cornell_camera():
    c = scenes.CORNELL_CAMERA
    return Camera(ob.camera_transform(c["origin"]), c["fov"], c["F"], c["focus"])


def standard_camera():
    c = scenes.STANDARD_CAMERA
    return Camera(ob.camera_walk(c["origin"], 0.0, 150.0, 0, 0.0), c["fov"], c["F"], c["focus"])


def small_standard_oracle(o):
    s = scenes.standard_surface_scene(64, 32, sphere_res=(16, 8))
    L = scenes.STANDARD_LIGHTING
    o.set_scene(s)
    o.build_accel()
    o.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
    o.load_arhosek_sky(L["turbidity"], L["albedo"])
    return s


def bsdf_below_horizon():
    """wo below the shading horizon (normal maps): eval f / pdf of the reference BSDF (round 2)."""
    cases, labels = make_cases_below_horizon(32)
    np.savez_compressed(os.path.join(OUT, "bsdf_below_horizon.npz"), cases=cases, labels=np.array(labels),
                        out=ob.bsdf_eval_sample(cases)[:, :4])


def main():
    if not ob.available():
        ob.build()
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["bsdf_below_horizon"]:   # add this fixture without touching the others
        bsdf_below_horizon()
        return
    bsdf_below_horizon()

    # ---- integer sampler: CMJ + Owen-Sobol + xxhash (bit exact) ----
    seqs = np.stack([ob.sampler_sequence(w, h, 1, idx, spp, SAMPLER_KINDS) for (w, h, idx, spp) in SAMPLER_POINTS])
    L = ob.lib()
    import ctypes as C
    for fn in ("orc_xxhash32_1", "orc_xxhash32_4", "orc_cmj_permute", "orc_sobol", "orc_owen"):
        getattr(L, fn).restype = C.c_uint32
    xs = np.arange(0, 64, dtype=np.uint32) * np.uint32(2654435761)
    xx1 = np.array([L.orc_xxhash32_1(C.c_uint32(int(x))) for x in xs], np.uint32)
    xx4 = np.array([L.orc_xxhash32_4(C.c_uint32(int(x)), C.c_uint32(i), C.c_uint32(7 * i), C.c_uint32(int(x) ^ 0xdeadbeef))
                    for i, x in enumerate(xs)], np.uint32)
    perm = np.array([[L.orc_cmj_permute(C.c_uint32(i), C.c_uint32(16), C.c_uint32(int(p))) for i in range(16)]
                     for p in xs[:8]], np.uint32)
    sob = np.array([[L.orc_sobol(C.c_ulonglong(i), C.c_uint32(d), C.c_uint32(0)) for d in (0, 1, 5, 100, 1023)]
                    for i in (0, 1, 2, 3, 1000, 123456789, 0xffffffff)], np.uint32)
    owen = np.array([L.orc_owen(C.c_uint32(int(x)), C.c_uint32(0x1234567 + i)) for i, x in enumerate(xs)], np.uint32)
    np.savez_compressed(os.path.join(OUT, "sampler.npz"), kinds=np.array(SAMPLER_KINDS), points=np.array(SAMPLER_POINTS),
                        sequences=seqs, xs=xs, xxhash32_1=xx1, xxhash32_4=xx4, cmj_permute=perm, sobol=sob, owen=owen)

    # ---- BSDF eval / sample per material class ----
    cases, labels = make_cases(48)
    np.savez_compressed(os.path.join(OUT, "bsdf.npz"), cases=cases, labels=np.array(labels),
                        out=ob.bsdf_eval_sample(cases))

    # ---- Hosek sky: cooked coefficients + radiance ----
    cooks = np.stack([ob.arhosek_cook(*p) for p in HOSEK_POINTS])
    o = ob.Oracle()
    o.set_scene(scenes.cornell_box())
    Ls = scenes.STANDARD_LIGHTING
    o.set_directional_light(Ls["sun_le"], Ls["sun_dir"], Ls["sun_angle"])
    o.load_arhosek_sky(Ls["turbidity"], Ls["albedo"])
    rng = np.random.default_rng(11)
    d = rng.normal(size=(256, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    np.savez_compressed(os.path.join(OUT, "sky.npz"), points=np.array(HOSEK_POINTS, np.float32), cooked=cooks, dirs=d,
                        radiance=o.sky_radiance(d))

    # ---- Cornell box (BASELINE config 1 at reduced size): traversal + all six layers ----
    o = ob.Oracle()
    s = scenes.cornell_box()
    o.set_scene(s)
    cam = cornell_camera()
    o.set_resolution(64, 64)
    rays = o.primary_rays(cam, 0).reshape(-1, 6)
    rng = np.random.default_rng(1)
    ro = rng.uniform(-0.9, 0.9, (4096, 3)).astype(np.float32) + np.float32([0, 1, 0])
    rd = rng.normal(size=(4096, 3)).astype(np.float32)
    rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    rays = np.concatenate([rays, np.concatenate([ro, rd], 1)]).astype(np.float32)
    ids, tuv = o.trace_closest(rays)
    np.savez_compressed(os.path.join(OUT, "cornell_trace.npz"), rays=rays, ids=ids, tuv=tuv)
    ref, _ = o.render_canonical(cam, (0, 0, 0), 8, 6, n_threads=os.cpu_count() or 1)
    np.savez_compressed(os.path.join(OUT, "cornell_smoke.npz"), spp=8, max_depth=6,
                        **{k: v.astype(np.float32) for k, v in ref.items()})
    o.set_resolution(32, 32)
    o.init_render_states()
    ref, _ = o.render_canonical(cam, (0, 0, 0), 16, 8, n_threads=os.cpu_count() or 1)
    np.savez_compressed(os.path.join(OUT, "cornell_32.npz"), spp=16, max_depth=8,
                        **{k: v.astype(np.float32) for k, v in ref.items()})

    # ---- Standard-Surface scene (BASELINE config 2 at reduced size): ids, t, image ----
    o = ob.Oracle()
    small_standard_oracle(o)
    cam = standard_camera()
    W, H = 96, 54
    o.set_resolution(W, H)
    rays = o.primary_rays(cam, 0).reshape(-1, 6)
    ids, tuv = o.trace_closest(rays)
    ref, _ = o.render_canonical(cam, (0, 0, 0), 16, 10, n_threads=os.cpu_count() or 1)
    np.savez_compressed(os.path.join(OUT, "standard_small.npz"), width=W, height=H, spp=16, max_depth=10, rays=rays,
                        ids=ids, tuv=tuv, **{k: v.astype(np.float32) for k, v in ref.items()})
    for f in sorted(os.listdir(OUT)):
        print("%-24s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
