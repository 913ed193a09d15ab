"""Acceleration-structure edge cases on the GPU: both builders, tiny and
degenerate inputs, exact t ties, instance transforms and alpha cut-outs -- (instance, primitive) and the
bits of t/u/v against the host oracle's traversal, which itself is pinned on a brute-force loop
(tests/test_oracle_golden.py::test_traversal_matches_bruteforce)."""
import os

import numpy as np
import pytest

from fredholm_b200 import Camera, DeviceLayers, api, scenes
from fredholm_b200.scenes import _assemble, _quad
from fredholm_b200.types import make_material

pytestmark = pytest.mark.gpu

MISS = 0xffffffff


def random_rays(n, seed, lo=-3.0, hi=3.0):
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], 1).astype(np.float32)


def rays_at(targets, seed):
    """Rays from random origins through given points (so that small scenes are actually hit)."""
    rng = np.random.default_rng(seed)
    t = np.asarray(targets, np.float32)
    o = t + rng.normal(size=t.shape).astype(np.float32) * 2.0
    d = t - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], 1).astype(np.float32)


def both(renderer, oracle, scene, rays):
    renderer.set_scene(scene)
    renderer.build_accel()
    oracle.set_scene(scene)
    oracle.build_accel()
    ids_g, tuv_g = renderer.trace_closest(rays)
    ids_o, tuv_o = oracle.trace_closest(rays)
    return ids_g, tuv_g, ids_o, tuv_o


def assert_identical(ids_g, tuv_g, ids_o, tuv_o, min_hits=1):
    assert np.array_equal(ids_g, ids_o)
    hit = ids_o[:, 0] != MISS
    assert hit.sum() >= min_hits
    assert np.array_equal(tuv_g[hit].view(np.uint32), tuv_o[hit].view(np.uint32))


def soup(n, seed, size=0.4):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-2, 2, (n, 1, 3))
    p = (c + rng.normal(size=(n, 3, 3)) * size).astype(np.float32)
    mat = make_material(base_color=(0.7, 0.7, 0.7))
    return _assemble([[(tri, 0) for tri in p]], [mat]), p


@pytest.mark.parametrize("builder", ["lbvh", "ploc"])
def test_builders_agree_with_oracle(renderer, oracle, builder, monkeypatch):
    monkeypatch.setenv("FRD_BVH_BUILDER", builder)
    s = scenes.standard_surface_scene(48, 24, sphere_res=(12, 6))
    rays = random_rays(60000, 11, -12, 12)
    rays[:, 1] = np.abs(rays[:, 1]) * 0.4 + 0.2
    ids_g, tuv_g, ids_o, tuv_o = both(renderer, oracle, s, rays)
    assert_identical(ids_g, tuv_g, ids_o, tuv_o, min_hits=10000)
    info = renderer.accel_info()
    assert info["n_faces"] == s.n_faces and info["n_nodes"] >= 1


@pytest.mark.parametrize("radius", ["1", "3", "32"])
def test_ploc_radius_extremes(renderer, oracle, radius, monkeypatch):
    monkeypatch.setenv("FRD_BVH_BUILDER", "ploc")
    monkeypatch.setenv("FRD_PLOC_RADIUS", radius)
    s, p = soup(3000, 5)
    rays = np.concatenate([rays_at(p.mean(axis=1), 6), random_rays(20000, 7)])
    assert_identical(*both(renderer, oracle, s, rays), min_hits=3000)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 9, 33])
def test_tiny_scenes(renderer, oracle, n):
    s, p = soup(n, 100 + n, size=0.8)
    rays = np.concatenate([np.repeat(rays_at(p.mean(axis=1), n), 40, axis=0) + 0, random_rays(4000, n)])
    assert_identical(*both(renderer, oracle, s, rays), min_hits=n)


def test_empty_scene_is_rejected_like_the_reference(renderer):
    """Renderer::load_scene throws std::runtime_error("invalid scene") when Scene::is_valid() fails
    (renderer.h:354-359): no faces is not a scene.  The renderer stays usable afterwards."""
    mat = make_material(base_color=(0.5, 0.5, 0.5))
    s = _assemble([[]], [mat])
    assert s.n_faces == 0
    with pytest.raises(Exception, match="invalid scene"):
        renderer.set_scene(s)
    one, p = soup(1, 3)
    renderer.set_scene(one)
    renderer.build_accel()
    ids, _ = renderer.trace_closest(rays_at(p.mean(axis=1), 1))
    assert (ids[:, 0] == 0).all()


def test_exact_ties_take_the_lowest_face(renderer, oracle):
    """Coincident triangles give exactly equal t: the lower global face index wins on both sides
    (the tie rule DESIGN.md states; OptiX leaves it unspecified)."""
    base, p = soup(200, 21)
    tris = [(tri, 0) for tri in p]
    dup = [tris[i] for i in (5, 5, 17, 60, 60, 60, 199)]
    mat = make_material(base_color=(0.7, 0.7, 0.7))
    s = _assemble([tris + dup, dup], [mat])        # duplicates inside the sub-mesh and in a second one
    rays = np.concatenate([np.repeat(rays_at(p.mean(axis=1), 3), 8, axis=0), random_rays(20000, 4)])
    ids_g, tuv_g, ids_o, tuv_o = both(renderer, oracle, s, rays)
    assert_identical(ids_g, tuv_g, ids_o, tuv_o, min_hits=500)
    hit = ids_o[:, 0] != MISS
    assert (ids_g[hit, 0] == 0).all() and (ids_g[hit, 1] < 200).all()   # never a duplicate


def test_degenerate_triangles_are_never_hit(renderer, oracle):
    s0, p = soup(300, 31)
    q = p.copy()
    q[::3, 2] = q[::3, 1]                      # zero area: two equal vertices
    q[1::7, 1] = (q[1::7, 0] + q[1::7, 2]) / 2  # zero area: collinear
    mat = make_material(base_color=(0.7, 0.7, 0.7))
    s = _assemble([[(tri, 0) for tri in q]], [mat])
    rays = np.concatenate([rays_at(p.mean(axis=1), 8), random_rays(20000, 9)])
    ids_g, tuv_g, ids_o, tuv_o = both(renderer, oracle, s, rays)
    assert_identical(ids_g, tuv_g, ids_o, tuv_o, min_hits=100)
    hit = ids_g[:, 0] != MISS
    dead = np.zeros(300, bool)
    dead[::3] = True
    assert not dead[ids_g[hit, 1]].any()


def test_instance_transforms(renderer, oracle):
    """Sub-mesh i is instance i with transform i (renderer.h:509-527): rotation, non-uniform scale and
    translation; ids are (instance, primitive-in-sub-mesh)."""
    mat = make_material(base_color=(0.7, 0.7, 0.7))
    _, p = soup(150, 41, size=0.3)
    shapes = [[(tri, 0) for tri in p[:50]], [(tri, 0) for tri in p[50:100]], [(tri, 0) for tri in p[100:]]]
    s = _assemble(shapes, [mat])
    def trs(angle, scale, t):
        c, sn = np.cos(angle), np.sin(angle)
        R = np.array([[c, 0, sn], [0, 1, 0], [-sn, 0, c]], np.float32) @ np.diag(scale).astype(np.float32)
        M = np.eye(4, dtype=np.float32)
        M[:3, :3] = R
        M[:3, 3] = t
        return M.T.reshape(16)              # column-major mat4
    s.transforms = np.stack([trs(0.0, (1, 1, 1), (0, 0, 0)), trs(0.7, (1.5, 0.5, 2.0), (3, -1, 0.5)),
                             trs(-2.1, (0.3, 2.0, 1.0), (-4, 2, 1))]).astype(np.float32)
    s.instance_ids = np.repeat(np.arange(3, dtype=np.uint32), 50)
    rays = random_rays(80000, 42, -7, 7)
    ids_g, tuv_g, ids_o, tuv_o = both(renderer, oracle, s, rays)
    assert_identical(ids_g, tuv_g, ids_o, tuv_o, min_hits=300)
    hit = ids_g[:, 0] != MISS
    assert set(np.unique(ids_g[hit, 0])) == {0, 1, 2} and ids_g[hit, 1].max() < 50


def test_alpha_cutout_matches_oracle(renderer, oracle):
    """Any-hit programs (pt.cu:545-678): texels with alpha < 0.5 in the base-colour map are holes for
    radiance, shadow and light rays alike -- first-hit depth and the image against the reference.  The
    open wall is also hit on its BACK face by bounce rays: every lobe weight is 0 there, the lobe
    distribution is 0/0 and the NEE / MIS weights are NaN, which the reference's device build saturates
    to 0 (oracle/Makefile, pt.cu:375)."""
    tex = np.zeros((16, 16, 4), np.uint8)
    tex[..., :3] = 200
    yy, xx = np.mgrid[0:16, 0:16]
    tex[..., 3] = np.where((xx // 4 + yy // 4) % 2 == 0, 255, 0)      # checkerboard of holes
    front = make_material(base_color=(0.8, 0.8, 0.8), base_color_texture_id=0)
    back = make_material(base_color=(0.2, 0.6, 0.9))
    wall = [(t, 0) for t in _quad((-1, 0, 0), (1, 0, 0), (1, 2, 0), (-1, 2, 0))]
    behind = [(t, 1) for t in _quad((-2, -1, -1), (2, -1, -1), (2, 3, -1), (-2, 3, -1))]
    s = _assemble([wall, behind], [front, back])
    s.textures = [(tex, True)]
    W, H = 64, 64
    c = dict(scenes.CORNELL_CAMERA)
    cam = Camera(api.camera_walk(c["origin"], 0.0, 0.0, 0, 0.0), c["fov"], c["F"], c["focus"])
    renderer.set_scene(s)
    renderer.build_accel()
    renderer.set_resolution(W, H)
    layers = DeviceLayers(W, H)
    renderer.render(cam, (1, 1, 1), layers, 4, 4)
    renderer.wait()
    oracle.set_scene(s)
    oracle.build_accel()
    oracle.set_resolution(W, H)
    ref, _ = oracle.render_canonical(cam, (1, 1, 1), 4, 4, n_threads=os.cpu_count() or 1)
    d_g, d_o = layers.download("depth").reshape(H, W), ref["depth"].reshape(H, W)
    assert np.isclose(d_g, d_o, rtol=1e-5, atol=1e-6).mean() >= 0.999
    assert d_g.max() > d_g[d_g > 0].min() * 1.15           # both the wall and the quad behind it are seen
    from conftest import rel_mse
    assert rel_mse(layers.download("beauty")[..., :3], ref["beauty"][..., :3]) < 1e-3


def test_back_face_of_an_opaque_surface(renderer, oracle):
    """A single quad seen from behind: is_entering = false switches every reflection lobe off
    (bsdf.cu:56-63), the path continues with NaN weights that saturate to 0 -- black against the
    background, and the same ray counts as the reference."""
    mat = make_material(base_color=(0.8, 0.8, 0.8))
    back = _quad((-1, 0, 0), (-1, 2, 0), (1, 2, 0), (1, 0, 0))     # wound so that the normal points away
    s = _assemble([[(t, 0) for t in back]], [mat])
    W = H = 48
    c = scenes.CORNELL_CAMERA
    cam = Camera(api.camera_walk(c["origin"], 0.0, 0.0, 0, 0.0), c["fov"], c["F"], c["focus"])
    renderer.set_scene(s)
    renderer.build_accel()
    renderer.set_resolution(W, H)
    renderer.reset_statistics()
    layers = DeviceLayers(W, H)
    renderer.render(cam, (1, 1, 1), layers, 2, 3)
    renderer.wait()
    oracle.set_scene(s)
    oracle.build_accel()
    oracle.set_resolution(W, H)
    oracle.reset_ray_counts()
    ref, _ = oracle.render_canonical(cam, (1, 1, 1), 2, 3, n_threads=os.cpu_count() or 1)
    got = layers.download("beauty")[..., :3]
    assert np.allclose(got, ref["beauty"][..., :3], atol=1e-5)
    hit = layers.download("depth").reshape(H, W) > 0
    # pixels whose two samples both hit are black, silhouette pixels are half background
    assert hit.any() and (got[hit] <= 0.5 + 1e-6).all() and (got[hit] == 0.0).mean() > 0.8
    assert np.allclose(got[~hit], 1.0)
    st, rc = renderer.statistics(), oracle.ray_counts()
    assert st["rays_radiance"] == rc["rays_radiance"]


@pytest.mark.parametrize("F,focus", [(100.0, 10000.0), (1.4, 3.0), (0.8, 1.2)])
def test_thin_lens_rays_match_reference(renderer, oracle, F, focus):
    """sample_ray_thinlens_camera (camera.cu:24-53) with a wide-open lens: lens radius 2f/F, concentric
    disk sample, focus plane -- origins and directions of every pixel's first sample, and the image
    with depth of field."""
    s = scenes.cornell_box()
    c = scenes.CORNELL_CAMERA
    cam = Camera(api.camera_walk(c["origin"], 40.0, 30.0, 0, 0.0), c["fov"], F, focus)
    W, H = 64, 48
    for x in (renderer, oracle):
        x.set_scene(s)
        x.build_accel()
        x.set_resolution(W, H)
    for n_spp in (0, 5, 21):
        a, b = renderer.primary_rays(cam, n_spp), oracle.primary_rays(cam, n_spp)
        assert np.allclose(a, b, rtol=1e-5, atol=1e-5)
    if F < 100.0:
        o = renderer.primary_rays(cam, 0).reshape(-1, 6)[:, :3]
        assert np.ptp(o, axis=0).max() > 1e-2          # the lens really has an aperture
    layers = DeviceLayers(W, H)
    renderer.render(cam, (0, 0, 0), layers, 8, 5)
    renderer.wait()
    ref, _ = oracle.render_canonical(cam, (0, 0, 0), 8, 5, n_threads=os.cpu_count() or 1)
    from conftest import rel_mse
    assert rel_mse(layers.download("beauty")[..., :3], ref["beauty"][..., :3]) < 1e-3


def test_emitter_seen_from_behind(renderer, oracle):
    """has_emission / __closesthit__light (pt.cu:125-139, 952-998): an emissive quad facing away from
    the camera above a floor -- what the back of an emitter shows and what it gives to next-event
    estimation and MIS rays, against the reference."""
    lamp = make_material(base_color=(0.5, 0.5, 0.5), specular_color=(0, 0, 0), emission=1.0, emission_color=(9, 8, 7))
    floor = make_material(base_color=(0.7, 0.7, 0.7))
    panel = _quad((-0.6, 0.4, 0.0), (-0.6, 1.6, 0.0), (0.6, 1.6, 0.0), (0.6, 0.4, 0.0))     # normal -z: away
    ground = _quad((-2, 0, 2), (2, 0, 2), (2, 0, -2), (-2, 0, -2))
    s = _assemble([[(t, 0) for t in panel], [(t, 1) for t in ground]], [lamp, floor])
    c = scenes.CORNELL_CAMERA
    cam = Camera(api.camera_walk(c["origin"], 0.0, 60.0, 0, 0.0), c["fov"], c["F"], c["focus"])
    W, H = 64, 64
    for x in (renderer, oracle):
        x.set_scene(s)
        x.build_accel()
        x.set_resolution(W, H)
    layers = DeviceLayers(W, H)
    renderer.render(cam, (0.05, 0.05, 0.05), layers, 16, 4)
    renderer.wait()
    ref, _ = oracle.render_canonical(cam, (0.05, 0.05, 0.05), 16, 4, n_threads=os.cpu_count() or 1)
    got, want = layers.download("beauty")[..., :3], ref["beauty"][..., :3]
    from conftest import rel_mse
    assert np.isfinite(got).all() and rel_mse(got, want) < 1e-3
    assert abs(got.mean() - want.mean()) < 0.01 * want.mean()


def test_degenerate_chain_falls_back_to_the_radix_tree(renderer, oracle, monkeypatch):
    """Triangles whose spacing grows geometrically: every cluster's nearest neighbour is the one to its left,
    so agglomerative clustering merges one pair per round and yields a chain far deeper than the traversal
    stack; the builder must notice and rebuild with the (depth-bounded) radix tree, and the hits must still
    be the reference's."""
    monkeypatch.setenv("FRD_BVH_BUILDER", "ploc")
    k = np.arange(420, dtype=np.float64)
    x = 1.2 ** k
    sz = 0.04 * x
    tris = np.stack([np.stack([x - sz, 0 * x, -sz], -1), np.stack([x + sz, 0 * x, -sz], -1), np.stack([x, 0 * x, sz], -1)], 1)
    mat = make_material(base_color=(0.7, 0.7, 0.7))
    s = _assemble([[(t.astype(np.float32), 0) for t in tris]], [mat])
    o = np.stack([x, 3.0 * sz, 0 * x], -1)
    d = np.tile(np.array([0.0, -1.0, 0.0]), (len(x), 1))
    rays = np.concatenate([o, d], 1).astype(np.float32)
    rays = np.concatenate([rays, random_rays(5000, 5, -50, 50)])
    ids_g, tuv_g, ids_o, tuv_o = both(renderer, oracle, s, rays)
    assert_identical(ids_g, tuv_g, ids_o, tuv_o, min_hits=100)   # the far triangles are beyond tmax = 1e9
    assert renderer.accel_info()["depth"] <= 46


def test_fallback_path_rebuilds_with_the_radix_tree(renderer, oracle, monkeypatch):
    """The too-deep fallback itself (forced through its test hook): same tree as FRD_BVH_BUILDER=lbvh."""
    s = scenes.standard_surface_scene(32, 16, sphere_res=(12, 6))
    rays = random_rays(20000, 3, -10, 10)
    monkeypatch.setenv("FRD_BVH_BUILDER", "lbvh")
    renderer.set_scene(s)
    renderer.build_accel()
    want_nodes = renderer.accel_info()["n_nodes"]
    monkeypatch.setenv("FRD_BVH_BUILDER", "ploc")
    renderer.build_accel()
    assert renderer.accel_info()["n_nodes"] != want_nodes
    monkeypatch.setenv("FRD_PLOC_FORCE_FALLBACK", "1")
    ids_g, tuv_g, ids_o, tuv_o = both(renderer, oracle, s, rays)
    assert renderer.accel_info()["n_nodes"] == want_nodes
    assert_identical(ids_g, tuv_g, ids_o, tuv_o, min_hits=1000)
