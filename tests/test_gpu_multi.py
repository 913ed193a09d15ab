"""Multi-GPU path of the C++ core (include/fredholm/multi_gpu.h): sample slices rendered as
sums, ONE ncclReduce of the accumulation buffers inside the library, division by the sample count on the
root.  On a one-GPU box the world has one rank (the NCCL calls still run); with >= 2 devices the
single-process MultiGpuRenderer is compared with the single-GPU render of the same samples."""
import numpy as np
import pytest

from fredholm_b200 import Camera, DeviceLayers, Renderer, api, parallel, scenes
from conftest import rel_mse

pytestmark = pytest.mark.gpu

W, H, SPP, DEPTH = 192, 128, 32, 6


def _setup(r):
    s = scenes.standard_surface_scene(32, 16, sphere_res=(12, 6))
    L = scenes.STANDARD_LIGHTING
    r.set_scene(s)
    r.build_accel()
    r.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
    r.load_arhosek_sky(L["turbidity"], L["albedo"])
    r.set_resolution(W, H)


def _camera():
    c = scenes.STANDARD_CAMERA
    return Camera(api.camera_walk(c["origin"], 0.0, 150.0, 0, 0.0), c["fov"], c["F"], c["focus"])


@pytest.fixture(scope="module")
def single():
    r = Renderer(0)
    _setup(r)
    lay = DeviceLayers(W, H, names=("beauty", "albedo", "depth"))
    r.render(_camera(), (0, 0, 0), lay, SPP, DEPTH)
    r.wait()
    out = {k: lay.download(k) for k in ("beauty", "albedo", "depth")}
    r.close()
    return out


def test_one_rank_world_reduces_in_place(single):
    """fr_comm_init / fr_render_sharded with world = 1: the slice is the whole frame, ncclReduce runs in place
    on the renderer's stream, the sums become means -- the result is the plain render."""
    r = Renderer(0)
    _setup(r)
    r.comm_init(api.comm_unique_id(), 0, 1)
    lay = DeviceLayers(W, H, names=("beauty", "albedo", "depth"))
    lay.clear()
    r.render_sharded(_camera(), (0, 0, 0), lay, SPP, DEPTH, root=0)
    r.wait()
    for k in ("beauty", "albedo", "depth"):
        got = lay.download(k)
        ref = single[k]
        a, b = (got[..., :3], ref[..., :3]) if got.ndim == 3 else (got, ref)
        assert rel_mse(a, b) < 1e-9, k
    assert lay.download("beauty")[..., 3].min() == 1.0 == lay.download("beauty")[..., 3].max()
    assert r.sample_count() == SPP
    r.comm_destroy()
    r.close()


def test_render_sharded_needs_a_communicator():
    r = Renderer(0)
    _setup(r)
    lay = DeviceLayers(W, H, names=("beauty",))
    with pytest.raises(api.FredholmError, match="fr_comm_init"):
        r.render_sharded(_camera(), (0, 0, 0), lay, SPP, DEPTH)
    r.close()


def test_multi_renderer_matches_single_gpu(single):
    """fredholm::MultiGpuRenderer over every device of the box (one on the CI box): scene replicated, sample
    slices per device, one ncclReduce onto the first device."""
    m = api.MultiRenderer()
    n = len(m)
    assert n == api.lib().fr_device_count()
    m.for_each(_setup)
    api.set_device(0)     # the result layers live on the first device (every renderer call switches the thread's device)
    lay = DeviceLayers(W, H, names=("beauty", "albedo", "depth"))
    lay.clear()
    m.render(_camera(), (0, 0, 0), lay, SPP, DEPTH)
    m.wait()
    # the slices are the core's own partition of the samples
    covered = []
    for rank in range(n):
        first, cnt = api.sample_slice(SPP, rank, n)
        covered += list(range(first, first + cnt))
        assert (first, cnt) == parallel.sample_slice(SPP, rank, n)
    assert covered == list(range(SPP))
    for k in ("beauty", "albedo", "depth"):
        got, ref = lay.download(k), single[k]
        a, b = (got[..., :3], ref[..., :3]) if got.ndim == 3 else (got, ref)
        assert rel_mse(a, b) < 1e-9, k
    m.close()
