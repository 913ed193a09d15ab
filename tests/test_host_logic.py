"""Host-side logic of the boundary that needs no GPU: the fredholm::Camera mirror
(camera.h:51-135), the Hosek-sky coefficient cook (arhosek.h:145-323), the film <-> path-slot
mapping of the wavefront and the multi-GPU sample slicing."""
import numpy as np
import pytest

from fredholm_b200 import api, scenes
from fredholm_b200 import parallel


@pytest.mark.parametrize("origin", [(0, 1, 5), (0.0, 6.0, 22.0), (-3.5, 0.25, 1e3)])
def test_camera_constructor(oracle_mod, origin):
    ours = api.Camera.from_origin(origin).transform
    # analytic inverse vs glm::inverse(glm::lookAt(...)): equal up to fp32 rounding
    assert np.allclose(ours, oracle_mod.camera_transform(origin), rtol=0, atol=2e-6)
    # looks down -z: third column of the camera-to-world block is +z (camera forward = -z)
    m = ours.reshape(3, 4)
    assert np.allclose(m[:, 3], origin)


def test_camera_walk(oracle_mod):
    rng = np.random.default_rng(3)
    for _ in range(40):
        origin = rng.uniform(-10, 10, 3)
        d_phi, d_theta = rng.uniform(-200, 200, 2)
        movement = int(rng.integers(0, 7))
        dt = float(rng.uniform(0, 2))
        a = api.camera_walk(origin, d_phi, d_theta, movement, dt)
        b = oracle_mod.camera_walk(origin, d_phi, d_theta, movement, dt)
        assert np.allclose(a, b, rtol=1e-5, atol=1e-5), (origin, d_phi, d_theta, movement, dt)


def test_hosek_cook(oracle_mod):
    for t, a, e in [(3.0, 0.3, 1.2), (2.0, 0.1, 0.2), (6.5, 0.8, 0.7), (10.0, 0.0, 1.5), (1.0, 1.0, 0.01)]:
        assert np.allclose(api.arhosek_cook(t, a, e), oracle_mod.arhosek_cook(t, a, e), rtol=2e-6), (t, a, e)


def test_sample_slices():
    """disjoint contiguous slices that cover [0, spp); whole CMJ patterns
    (multiples of 16 samples) stay on one rank whenever spp allows it."""
    for spp in (16, 64, 100, 4096, 7):
        for world in (1, 2, 3, 4, 8):
            sl = [parallel.sample_slice(spp, r, world) for r in range(world)]
            assert sl[0][0] == 0 and sl[-1][0] + sl[-1][1] == spp
            for (a, n), (b, _) in zip(sl, sl[1:]):
                assert a + n == b
            if spp % (16 * world) == 0:
                assert all(a % 16 == 0 and n == spp // world for a, n in sl)
            assert all(n >= 0 for _, n in sl)
    assert parallel.frames_for_rank(48, 3, 8) == list(range(3, 48, 8))


def test_procedural_scenes_are_deterministic():
    a = scenes.standard_surface_scene(16, 8, sphere_res=(8, 4))
    b = scenes.standard_surface_scene(16, 8, sphere_res=(8, 4))
    assert np.array_equal(a.vertices, b.vertices) and np.array_equal(a.indices, b.indices)
    assert a.materials.tobytes() == b.materials.tobytes()
    c = scenes.cornell_box()
    assert c.n_faces == 32
    # the benchmark scene has exactly 2^20 triangles (BASELINE.json config 2)
    assert 2 * 512 * 512 + 512 * (2 * 32 * 16) == 1048576


def test_core_sample_slice_matches_the_python_rule():
    """fr_sample_slice (C++ core, multi_gpu.cpp) and parallel.sample_slice (gloo tests) are the same partition:
    contiguous, disjoint, covering, whole 16-sample CMJ patterns."""
    from fredholm_b200 import api, parallel
    for total in (0, 1, 15, 16, 17, 64, 100, 512, 4096, 4100):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for rank in range(world):
                first, cnt = api.sample_slice(total, rank, world)
                assert (first, cnt) == parallel.sample_slice(total, rank, world)
                assert first == covered
                if rank < world - 1 or total % 16 == 0:
                    assert cnt % 16 == 0 or first + cnt == total
                covered += cnt
            assert covered == total
    import pytest
    with pytest.raises(api.FredholmError):
        api.sample_slice(64, 2, 2)
