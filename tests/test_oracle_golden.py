"""CPU tests of the ORACLE (test infrastructure): the reference's own integrator sources
compiled for the host (oracle/Makefile -> oracle/_ref/libfredholm_oracle.so) must reproduce
the committed golden vectors (tests/tools/gen_golden.py), and the parts the oracle has to restate
itself -- traversal -- are pinned against a brute-force loop over all triangles.

The reference has no tests or known-answer vectors of its own."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden, rel_mse
from fredholm_b200 import Camera, scenes

sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
import gen_golden as gg  # noqa: E402


def test_struct_sizes(oracle_mod):
    """Layout contract with the reference: the C ABI passes
    fredholm::Material records of 180 bytes."""
    L = oracle_mod.lib()
    for fn, size in (("orc_sizeof_material", 180), ("orc_sizeof_shading_params", 120), ("orc_sizeof_launch_params", 288)):
        getattr(L, fn).restype = C.c_uint32
        assert getattr(L, fn)() == size
    from fredholm_b200.types import MATERIAL_DTYPE
    assert MATERIAL_DTYPE.itemsize == 180


def test_sampler_golden(oracle_mod):
    g = golden("sampler.npz")
    kinds = str(g["kinds"])
    for (w, h, idx, spp), want in zip(g["points"], g["sequences"]):
        got = oracle_mod.sampler_sequence(int(w), int(h), 1, int(idx), int(spp), kinds)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    L = oracle_mod.lib()
    for fn in ("orc_xxhash32_1", "orc_xxhash32_4", "orc_cmj_permute", "orc_sobol", "orc_owen"):
        getattr(L, fn).restype = C.c_uint32
    xs = g["xs"]
    assert [L.orc_xxhash32_1(C.c_uint32(int(x))) for x in xs] == g["xxhash32_1"].tolist()
    assert [L.orc_owen(C.c_uint32(int(x)), C.c_uint32(0x1234567 + i)) for i, x in enumerate(xs)] == g["owen"].tolist()


def test_sampler_properties(oracle_mod):
    """Domain properties of the reference sampler (cmj.cu:12-80): cmj_permute is a permutation;
    the 16 samples of one CMJ pattern are stratified 4x4 and 16x1 / 1x16 (multi-jittered)."""
    L = oracle_mod.lib()
    L.orc_cmj_permute.restype = C.c_uint32
    for p in (1, 77, 0xabcdef01):
        for n in (4, 16, 7):
            assert sorted(L.orc_cmj_permute(C.c_uint32(i), C.c_uint32(n), C.c_uint32(p)) for i in range(n)) == list(range(n))
    pts = np.array([oracle_mod.sampler_sequence(64, 64, 1, 123, s, "2") for s in range(16)])
    assert pts.min() >= 0.0 and pts.max() < 1.0
    cells = set((int(x * 4), int(y * 4)) for x, y in pts)
    assert len(cells) == 16
    assert len(set(int(x * 16) for x, _ in pts)) == 16 and len(set(int(y * 16) for _, y in pts)) == 16
    # 1-D draws: Owen-scrambled Sobol, values in [0, 1]
    u = np.array([oracle_mod.sampler_sequence(64, 64, 1, i, 3, "1111") for i in range(256)])
    assert u.min() >= 0.0 and u.max() <= 1.0 and abs(u.mean() - 0.5) < 0.05


def test_bsdf_golden(oracle_mod):
    g = golden("bsdf.npz")
    got = oracle_mod.bsdf_eval_sample(g["cases"])
    assert np.array_equal(np.isnan(got), np.isnan(g["out"]))
    assert np.allclose(got, g["out"], rtol=1e-6, atol=1e-7, equal_nan=True)


def test_bsdf_properties(oracle_mod):
    """eval is reciprocal-free but must be non-negative with a non-negative pdf; a sampled
    direction evaluates to a positive pdf for reflective classes."""
    g = golden("bsdf.npz")
    out = g["out"]
    ok = ~np.isnan(out).any(axis=1)
    assert (out[ok, :4] >= 0).all()
    assert (out[ok, 10] >= 0).all()
    lam = g["labels"] == "lambert"
    c = g["cases"][lam]
    f = out[lam][c[:, 35] > 0, :3]
    nz = f.sum(axis=1) > 0
    # Lambert: f = rho / pi on the upper hemisphere (bxdf.cu:119-149)
    assert nz.mean() > 0.7
    assert np.allclose(f[nz], np.float32([0.7, 0.6, 0.5]) / np.pi, rtol=2e-3)


def test_sky_golden(oracle_mod):
    g = golden("sky.npz")
    for p, want in zip(g["points"], g["cooked"]):
        assert np.allclose(oracle_mod.arhosek_cook(*[float(v) for v in p]), want, rtol=1e-6)
    o = oracle_mod.Oracle()
    o.set_scene(scenes.cornell_box())
    L = scenes.STANDARD_LIGHTING
    o.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
    o.load_arhosek_sky(L["turbidity"], L["albedo"])
    got = o.sky_radiance(g["dirs"])
    assert np.array_equal(np.isnan(got), np.isnan(g["radiance"]))
    assert np.allclose(got, g["radiance"], rtol=1e-5, equal_nan=True)
    # below the horizon the reference's closed form is NaN (sqrt of a negative cosine, quirk a9)
    assert np.isnan(got[g["dirs"][:, 1] < -1e-3]).all()


def test_traversal_matches_bruteforce(oracle_mod):
    """The oracle's BVH traversal (its own restatement of optixTrace) returns exactly what a
    loop over every triangle returns: same (instance, primitive), same t, u, v bits."""
    o = oracle_mod.Oracle()
    gg.small_standard_oracle(o)
    rng = np.random.default_rng(5)
    n = 3000
    org = rng.uniform(-12, 12, (n, 3)).astype(np.float32)
    org[:, 1] = rng.uniform(0.3, 6.0, n)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([org, d], 1)
    ids, tuv = o.trace_closest(rays)
    ids_b, tuv_b = o.trace_closest(rays, bruteforce=True)
    assert np.array_equal(ids, ids_b)
    assert np.array_equal(tuv.view(np.uint32), tuv_b.view(np.uint32))
    assert 0.2 < (ids[:, 0] != 0xffffffff).mean() < 1.0


def test_cornell_trace_golden(oracle_mod):
    g = golden("cornell_trace.npz")
    o = oracle_mod.Oracle()
    o.set_scene(scenes.cornell_box())
    ids, tuv = o.trace_closest(g["rays"])
    assert np.array_equal(ids, g["ids"])
    assert np.array_equal(tuv.view(np.uint32), g["tuv"].view(np.uint32))


@pytest.mark.parametrize("name,res", [("cornell_32.npz", 32), ("cornell_smoke.npz", 64)])
def test_cornell_image_golden(oracle_mod, name, res):
    g = golden(name)
    o = oracle_mod.Oracle()
    o.set_scene(scenes.cornell_box())
    o.set_resolution(res, res)
    ref, _ = o.render_canonical(gg.cornell_camera(), (0, 0, 0), int(g["spp"]), int(g["max_depth"]), n_threads=4)
    for k in ("beauty", "position", "depth", "normal", "texcoord", "albedo"):
        assert np.allclose(ref[k], g[k], rtol=1e-5, atol=1e-6), k
    assert ref["beauty"][..., :3].mean() > 0.05


def test_standard_small_golden(oracle_mod):
    g = golden("standard_small.npz")
    o = oracle_mod.Oracle()
    gg.small_standard_oracle(o)
    W, H = int(g["width"]), int(g["height"])
    o.set_resolution(W, H)
    cam = gg.standard_camera()
    rays = o.primary_rays(cam, 0).reshape(-1, 6)
    assert np.allclose(rays, g["rays"], rtol=0, atol=1e-6)
    ids, tuv = o.trace_closest(g["rays"])
    assert np.array_equal(ids, g["ids"])
    assert np.array_equal(tuv.view(np.uint32), g["tuv"].view(np.uint32))
    ref, _ = o.render_canonical(cam, (0, 0, 0), int(g["spp"]), int(g["max_depth"]), n_threads=4)
    assert rel_mse(ref["beauty"][..., :3], g["beauty"][..., :3]) < 1e-8


def test_sample_slices_compose(oracle_mod):
    """The multi-GPU decomposition is exact at the reference level: rendering samples
    [0,8) and [8,16) separately (sample_count preset to the slice start)
    and averaging equals the 16-sample render up to fp32 rounding of the running mean."""
    o = oracle_mod.Oracle()
    o.set_scene(scenes.cornell_box())
    o.set_resolution(24, 24)
    cam = gg.cornell_camera()
    full, _ = o.render_canonical(cam, (0, 0, 0), 16, 6, n_threads=4)
    parts = []
    for first in (0, 8):
        o.set_sample_count(first)
        layers = o.new_layers()
        # running mean starting at n = first with zeroed layers: recover the slice sum
        for s in range(8):
            o.render(cam, (0, 0, 0), layers, 1, 6, n_threads=4)
        parts.append(layers["beauty"].astype(np.float64) * (first + 8))
    # slice 0 holds mean of samples 0..7; slice 1's running mean started from zeroed layers
    # at n=8, i.e. it holds (sum of samples 8..15) / 16
    mean0 = parts[0] / 8.0
    sum1 = parts[1]
    combined = (mean0 * 8.0 + sum1) / 16.0
    assert np.allclose(combined[..., :3], full["beauty"][..., :3], rtol=1e-4, atol=1e-5)


def test_bsdf_golden_below_the_shading_horizon(oracle_mod):
    """wo.y < 0 (reachable through normal / bump maps): the committed eval table is what the reference BSDF gives."""
    g = golden("bsdf_below_horizon.npz")
    got = oracle_mod.bsdf_eval_sample(g["cases"])[:, :4]
    assert np.array_equal(np.isnan(got), np.isnan(g["out"]))
    assert np.allclose(got, g["out"], rtol=1e-6, atol=0, equal_nan=True)
    assert (g["cases"][:, 31] < 0).all() and np.isfinite(g["out"]).mean() > 0.9
