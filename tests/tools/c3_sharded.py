"""BASELINE config 3 across N GPUs (run under torchrun): the 1080p scene at 4096 spp, sample-sharded with
parallel.sample_slice, SUM accumulators, ONE NCCL reduce of the beauty sums to rank 0, then the north-star
image gate on rank 0 -- relMSE <= 1e-3 against the reference integrator (oracle) at 4096 spp on a window of the
frame.  Prints one JSON object on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/tools/c3_sharded.py
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fredholm_b200 import Camera, Renderer, api, parallel, scenes  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, SPP, DEPTH = 1920, 1080, int(os.environ.get("C3_SPP", 4096)), 10
s = scenes.standard_surface_scene()
L, C = scenes.STANDARD_LIGHTING, scenes.STANDARD_CAMERA
cam = Camera(api.camera_walk(C["origin"], 0.0, 150.0, 0, 0.0), C["fov"], C["F"], C["focus"])
r = Renderer(local)
r.set_scene(s)
r.build_accel()
r.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
r.load_arhosek_sky(L["turbidity"], L["albedo"])
r.set_resolution(W, H)
# the data path is the C++ core's: slice render + ONE ncclReduce + scale (fr_render_sharded)
if world > 1:
    parallel.init_core_communicator(r, dist)
else:
    r.comm_init(api.comm_unique_id(), 0, 1)
from fredholm_b200 import DeviceLayers  # noqa: E402
lay = DeviceLayers(W, H, names=("beauty",))
# warm-up at the REAL wave size (the default 64 Mi-path wave = 32 samples of this frame), so that no allocation
# lands in the timed region (round 1 warmed up with 16 samples and timed a 24 GB cudaMalloc)
lay.clear()
r.init_render_states()
r.render_sharded(cam, (0, 0, 0), lay, min(SPP, 64 * world), DEPTH, root=0)
r.wait()
lay.clear()
r.init_render_states()
r.reset_statistics()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
r.render_sharded(cam, (0, 0, 0), lay, SPP, DEPTH, root=0)
r.wait()
first, n = api.sample_slice(SPP, rank, world)
secs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
rays = torch.tensor([float(r.statistics()["rays"])], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(secs, op=dist.ReduceOp.MAX)
    dist.all_reduce(rays, op=dist.ReduceOp.SUM)
if rank == 0:
    got = lay.download("beauty")
    out = {"config": "C3 1920x1080, %d tris, %d spp, depth %d, sample-sharded x%d, one NCCL reduce" % (s.n_faces, SPP, DEPTH, world),
           "n_gpus": world, "samples_per_gpu": n, "seconds": float(secs.item()),
           "mpaths_per_s": W * H * SPP / float(secs.item()) / 1e6, "mrays_per_s": float(rays.item()) / float(secs.item()) / 1e6,
           "finite": bool(np.isfinite(got).all())}
    from oracle import binding as ob  # checker only
    if ob.available():
        o = ob.Oracle()
        o.set_scene(s)
        o.build_accel()
        o.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
        o.load_arhosek_sky(L["turbidity"], L["albedo"])
        o.set_resolution(W, H)
        win = (924, 520, 996, 560)
        ref, _ = o.render_canonical(cam, (0, 0, 0), SPP, DEPTH, window=win, n_threads=os.cpu_count() or 1)
        x0, y0, x1, y1 = win
        a, b = got[y0:y1, x0:x1, :3].astype(np.float64), ref["beauty"][y0:y1, x0:x1, :3].astype(np.float64)
        out["window"] = win
        out["relmse_vs_oracle_window"] = float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))
    print(json.dumps({"c3_sharded": out}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
