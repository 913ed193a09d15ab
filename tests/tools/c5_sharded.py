"""BASELINE config 5 across N GPUs (run under torchrun): 3840x2160 textured scene, 48-frame camera dolly, 16 spp,
depth 5, denoise + bloom / chromatic aberration / tone map through the batch driver (fr_batch_run), frame f on
rank f mod N (parallel.frames_for_rank), no collective on the data path; finished RGBA8 frames stay in host
memory.  Prints one JSON object on rank 0: aggregate frames/s = frames / slowest rank's wall time.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/tools/c5_sharded.py
"""
import json
import os
import pickle
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
N_FRAMES, W, H = int(os.environ.get("C5_FRAMES", 48)), 3840, 2160
# the procedural 4096^2 textures take ~50 s of numpy: rank 0 makes the scene once, the others read it
cache = "/dev/shm/fredholm_c5_scene.pkl"
t0 = time.time()
if rank == 0:
    s = scenes.textured_scene(tex_res=4096)
    with open(cache + ".tmp", "wb") as f:
        pickle.dump(s, f, protocol=4)
    os.replace(cache + ".tmp", cache)
if world > 1:
    dist.barrier()
if rank != 0:
    with open(cache, "rb") as f:
        s = pickle.load(f)
gen_s = time.time() - t0
L, C = scenes.STANDARD_LIGHTING, scenes.STANDARD_CAMERA
cam = Camera(api.camera_walk(C["origin"], 0.0, 150.0, 0, 0.0), C["fov"], C["F"], C["focus"])
r = Renderer(local)
r.set_scene(s)
r.build_accel()
r.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
r.load_arhosek_sky(L["turbidity"], L["albedo"])
r.set_resolution(W, H)
o0 = np.array(C["origin"], np.float32)
path = np.stack([api.camera_walk(o0 + np.array([0.05 * i, 0.0, -0.08 * i], np.float32), 0.1 * i, 150.0, 0, 0.0)
                 for i in range(N_FRAMES)])
mine = parallel.frames_for_rank(N_FRAMES, rank, world)
kw = dict(camera_path=path, keep_frames=True, denoise=True, use_bloom=True, bloom_threshold=2.0, bloom_sigma=5.0,
          ISO=80.0, chromatic_aberration=1.0, fps=24.0, max_time=1e9, animate=False)
r.batch_run(cam, W, H, 16, 5, 1, first_frame=rank, frame_stride=world, **kw)        # warm-up (allocations)
r.reset_statistics()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
recs, frames, info = r.batch_run(cam, W, H, 16, 5, len(mine), first_frame=rank, frame_stride=world, **kw)
wall = time.perf_counter() - t0
st = r.statistics()
t = torch.tensor([wall, float(st["paths"]), float(st["rays"]), float(info["n_frames"]), float(frames[..., :3].mean())],
                 dtype=torch.float64, device="cuda")
tmax = t.clone()
if world > 1:
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
if rank == 0:
    slowest = float(tmax[0].item())
    out = {"config": "C5 3840x2160 textured (3 x 4096^2 RGBA8), %d tris, %d frames over %d GPUs (frame f on rank f mod N), 16 spp, "
                     "depth 5, denoise + bloom/CA/tone map, frames to host RGBA8" % (s.n_faces, N_FRAMES, world),
           "n_gpus": world, "frames": int(t[3].item()), "frames_rank0": [int(x["frame_idx"]) for x in recs][:8],
           "slowest_rank_wall_s": slowest, "frames_per_s": t[3].item() / slowest,
           "mpaths_per_s_e2e": t[1].item() / slowest / 1e6, "mrays_per_s_e2e": t[2].item() / slowest / 1e6,
           "render_ms_rank0": float(np.mean([x["render_ms"] for x in recs])),
           "mean_rgb_over_ranks": t[4].item() / world, "scene_generation_or_load_s": gen_s}
    print(json.dumps({"c5_sharded": out}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
