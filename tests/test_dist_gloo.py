"""N > 1 path on CPU: world_size-2 gloo run of the sample-sharded render (fredholm_b200.parallel).
The per-rank "renderer" here is the host oracle (the checker) because there is no GPU in this
container; what is under test is the decomposition: slices, SUM accumulators, one reduce, the
division -- and that the result equals the single-process render of all samples."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

W = H = 20
SPP, DEPTH = 32, 5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slice_sums(o, cam, first, n):
    """Sums of samples [first, first+n) from the reference's streaming mean: start the
    running mean at n_spp = first with zeroed layers; after n launches the layer holds
    (sum of the slice) / (first + n)."""
    import torch
    o.set_sample_count(first)
    layers = o.new_layers()
    for _ in range(n):
        o.render(cam, (0, 0, 0), layers, 1, DEPTH, n_threads=2)
    return torch.from_numpy(layers["beauty"].astype(np.float64) * float(first + n))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import gen_golden as gg
    from fredholm_b200 import parallel, scenes
    from oracle import binding as ob
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = ob.Oracle()
        o.set_scene(scenes.cornell_box())
        o.set_resolution(W, H)
        cam = gg.cornell_camera()
        img = parallel.render_sharded(lambda first, n: _slice_sums(o, cam, first, n), dist, SPP)
        if rank == 0:
            np.save(out_path, img.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sample_sharding(oracle_mod, tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
    import gen_golden as gg
    from fredholm_b200 import scenes
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)

    o = oracle_mod.Oracle()
    o.set_scene(scenes.cornell_box())
    o.set_resolution(W, H)
    ref, _ = o.render_canonical(gg.cornell_camera(), (0, 0, 0), SPP, DEPTH, n_threads=2)
    assert got.shape == ref["beauty"].shape
    assert np.allclose(got[..., :3], ref["beauty"][..., :3], rtol=1e-4, atol=1e-5)
    assert got[..., :3].mean() > 0.05
