#!/usr/bin/env python3
"""Benchmark of the fredholm hot path (BASELINE.json): Mpaths/s and Mrays/s of the
path-tracing core on the 1080p / 1 048 576-triangle Standard-Surface scene
(Hosek sky + directional sun, 64 spp, depth 10).

  python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA core, C ABI)
  torchrun ... bench.py --gpus N ...                       N ranks, sample-sharded
  python bench.py --impl reference ...                     the reference's own integrator
                                                           (host build, oracle/_ref) on
                                                           the box's CPU cores

A "step" is one frame: 64 samples for every pixel of the 1920x1080 image.  With N GPUs every rank renders
its own 64-sample slice [64 r, 64 (r+1)) of a 64 N-sample frame (same scene / BVH replicated, weak scaling) into
sum-accumulators, followed by ONE ncclReduce of the beauty sums to rank 0 and the division by the sample count.
Slice render, reduce and division are one call into the C++ core (fr_render_sharded; the core loads NCCL itself),
enqueued on the renderer's stream and inside the timed region; torch.distributed only launches the ranks,
carries the 128-byte communicator id and takes the max over ranks of the device times.

Prints ONE JSON line (rank 0).  `value` is whole-job Mpaths/s with everything resident in HBM; `e2e` is the same
frame through the host-buffer C-ABI call (clear, render, read the framebuffer back to pinned host memory);
`roofline` is measured live (issue-slot fraction of the traversal kernels from the counting kernels of this run
and the SASS counts of this build; HBM fraction of the queue passes); `strong` is BASELINE config 3 (the same
frame at 4096 spp in total, split over the N GPUs) with its speed-up over one GPU at this run's rate.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(width=1920, height=1080, spp=64, max_depth=10, terrain_res=512, n_spheres=512)
STRONG_SPP = 4096  # BASELINE config 3: the same frame at 4096 spp, split over the N GPUs (strong scaling)


def reference_tiles(width, height, nx=9, ny=5, tw=96, th=90):
    """The reference arm's bounded sample of the frame: nx x ny tiles of tw x th pixels, one centred in each
    cell of a regular grid over the WHOLE image (sky, horizon, spheres, terrain in the frame's own proportions;
    a centre crop over-weights the geometry-dense part).  At 1080p: 45 tiles = 388 800 pixels = 18.75 % of
    the frame."""
    sx, sy = width / WORKLOAD["width"], height / WORKLOAD["height"]
    tw, th = max(1, int(tw * sx)), max(1, int(th * sy))
    tiles = []
    for j in range(ny):
        for i in range(nx):
            cx0, cx1 = width * i // nx, width * (i + 1) // nx
            cy0, cy1 = height * j // ny, height * (j + 1) // ny
            x0 = cx0 + max(0, (cx1 - cx0 - tw) // 2)
            y0 = cy0 + max(0, (cy1 - cy0 - th) // 2)
            tiles.append((x0, y0, min(x0 + tw, width), min(y0 + th, height)))
    return tiles


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=WORKLOAD["spp"])
    ap.add_argument("--width", type=int, default=WORKLOAD["width"])
    ap.add_argument("--height", type=int, default=WORKLOAD["height"])
    ap.add_argument("--max-depth", type=int, default=WORKLOAD["max_depth"])
    ap.add_argument("--small-scene", action="store_true", help="131k-triangle variant (debugging only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--wave-paths", type=int, default=0,
                    help="paths in flight per wave (default: 16 samples x tiled pixels)")
    ap.add_argument("--strong-spp", type=int, default=STRONG_SPP,
                    help="total samples of the strong-scaling sub-record (BASELINE config 3); 0 skips it")
    return ap.parse_args()


def build_scene(args):
    from fredholm_b200 import scenes
    if args.small_scene:
        return scenes.standard_surface_scene(128, 64)
    return scenes.standard_surface_scene(WORKLOAD["terrain_res"], WORKLOAD["n_spheres"])


def camera_for(transform_fn):
    from fredholm_b200 import Camera, scenes
    c = scenes.STANDARD_CAMERA
    return Camera(transform_fn(c["origin"], 0.0, 150.0, 0, 0.0), c["fov"], c["F"], c["focus"])


def workload_name(args, scene):
    return "%dx%d, %d tris Standard-Surface mix + Hosek sky + sun, %d spp, depth %d" % (
        args.width, args.height, scene.n_faces, args.spp, args.max_depth)


# ---------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------
def _reference_oracle(args, scene):
    from oracle import binding as ob
    from fredholm_b200 import scenes
    if not ob.available():
        ob.build()
    L = scenes.STANDARD_LIGHTING
    o = ob.Oracle()
    o.set_scene(scene)
    o.build_accel()
    o.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
    o.load_arhosek_sky(L["turbidity"], L["albedo"])
    o.set_resolution(args.width, args.height)
    return o, camera_for(ob.camera_walk)


def _reference_pass(o, cam, args, tiles, spp, threads):
    """`spp` launches of one sample (canonical mode) over the tiles; returns (seconds, paths, ray counts)."""
    layers = o.new_layers()
    o.init_render_states()
    o.reset_ray_counts()
    secs = 0.0
    for _ in range(spp):
        secs += o.render_tiles(cam, (0, 0, 0), layers, 1, args.max_depth, tiles, n_threads=threads)
    n_paths = sum((x1 - x0) * (y1 - y0) for x0, y0, x1, y1 in tiles) * spp
    return secs, n_paths, o.ray_counts()


def run_reference(args):
    """The reference's own integrator sources compiled for the host (oracle/_ref), all host threads, on a
    bounded sample of the workload: 45 tiles of 96x90 pixels spread over the whole 1080p frame, all 64 samples
    of every sampled pixel."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene = build_scene(args)
    o, cam = _reference_oracle(args, scene)
    cores = os.cpu_count() or 1
    tiles = reference_tiles(args.width, args.height)
    for _ in range(args.warmup):
        _reference_pass(o, cam, args, tiles, max(1, args.spp // 16), cores)   # short: the CPU has no clocks to ramp
    # a step is the tiles at all 64 samples (~8 s on 16 cores); a long run (the driver's --steps 25) takes the first
    # 16 samples of every sampled pixel per step instead -- one whole CMJ pattern -- so that it still ends in minutes
    ref_spp = args.spp if args.steps <= 8 else max(16, args.spp // 4)
    t = 0.0
    n_paths = rays = 0
    for _ in range(args.steps):
        secs, n, rc = _reference_pass(o, cam, args, tiles, ref_spp, cores)
        t += secs
        n_paths += n
        rays += rc["rays"]
    value = n_paths / t / 1e6
    frac = n_paths / args.steps / float(args.width * args.height * ref_spp)
    sample = "%d tiles of %dx%d px on a regular grid over the %dx%d frame (%.2f %% of its pixels), %d of the %d spp: %d paths per step" % (
        len(tiles), tiles[0][2] - tiles[0][0], tiles[0][3] - tiles[0][1], args.width, args.height, 100.0 * frac,
        ref_spp, args.spp, n_paths // args.steps)
    line = {
        "impl": "reference",
        "metric": "Mpaths/s (1080p, 1M tris, Standard Surface + Hosek sky, depth 10)",
        "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, scene), "sample": sample, "sampled_fraction": frac},
        "mrays_per_s": rays / t / 1e6, "rays_per_path": rays / max(n_paths, 1),
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------------------
def cpu_baseline(args, scene):
    """Reference integrator (oracle/_ref) timed on this box's host cores, on the reference arm's tiles:
    `value` is the all-core rate (the quantity the driver's ratio uses), `single_thread` the one-core rate
    BASELINE.md section 2 names as the reported baseline."""
    try:
        from oracle import binding as ob
        if not ob.available():
            return None
        o, cam = _reference_oracle(args, scene)
        cores = os.cpu_count() or 1
        tiles = reference_tiles(args.width, args.height)
        spp_all = max(1, args.spp // 4)                         # 16 spp on all cores: ~3-6 s
        secs, n_paths, rc = _reference_pass(o, cam, args, tiles, spp_all, cores)
        spp_one = 1                                             # 1 spp on one core: ~3-4 s
        secs1, n1, rc1 = _reference_pass(o, cam, args, tiles, spp_one, 1)
        return {"value": n_paths / secs / 1e6, "unit": "Mpaths/s", "cores": cores, "kind": "reference",
                "sample": "%d tiles of %dx%d px over the whole frame, %d spp on %d threads (%.1f s); 1 spp on one thread (%.1f s)"
                          % (len(tiles), tiles[0][2] - tiles[0][0], tiles[0][3] - tiles[0][1], spp_all, cores, secs, secs1),
                "mrays_per_s": rc["rays"] / secs / 1e6, "rays_per_path": rc["rays"] / max(n_paths, 1),
                "single_thread": {"value": n1 / secs1 / 1e6, "unit": "Mpaths/s", "cores": 1,
                                  "mrays_per_s": rc1["rays"] / secs1 / 1e6}}
    except Exception as e:  # the baseline is reported, never required
        return {"value": None, "unit": "Mpaths/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % e}


# algorithmic HBM bytes per unit of every stage: queue entry + record in, result out
ALGO_BYTES = {"trace_closest": 4 + 32 + 16 + 4, "trace_shadow": 48 + 16, "trace_light": 48 + 32,
              "shade": 4 + 16 + 16 + 16 + 3 * 48 + 48 + 48 + 4,
              # beauty-only frame: radiance word zeroed (16) + origin / direction / throughput (48) + queue entry (4);
              # the three first-hit AOV words (48 more) are only written when an AOV layer is bound
              "generate": 16 + 3 * 16 + 4, "film": 16}


def load_json(path):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


def issue_roofline(stage, rays, ms, per_ray, sass, sm_count, clock_mhz):
    """useful warp-instructions / s = rays/s x (N_node c_node + N_tri c_tri) / 32 against the
    chip's issue peak, SMs x 4 schedulers x clock.  N_node / N_tri: the counting kernels, this run; c_node /
    c_tri: SASS instructions of one node / triangle test in the shipped cubin (fredholm_b200/sass_counts.json)."""
    kernel = {"trace_closest": "k_trace_closest", "trace_shadow": "k_trace_shadow", "trace_light": "k_trace_light"}[stage]
    k = (sass or {}).get("kernels", {}).get(kernel)
    if not k or not ms or not rays:
        return None
    n_node, n_tri = per_ray
    inst_per_ray = n_node * k["c_node"] + n_tri * k["c_tri"]
    achieved = rays / (ms / 1e3) * inst_per_ray / 32.0 / 1e9          # G warp-instructions / s, all lanes useful
    peak = sm_count * 4 * clock_mhz * 1e6 / 1e9
    return {"kernel": kernel, "rays": rays, "ms": ms, "grays_per_s": rays / (ms / 1e3) / 1e9,
            "nodes_per_ray": n_node, "tris_per_ray": n_tri, "c_node": k["c_node"], "c_tri": k["c_tri"],
            "thread_inst_per_ray": inst_per_ray, "achieved": achieved, "peak": peak, "frac": achieved / peak}


def run_ours(args):
    from fredholm_b200 import Renderer, api, scenes
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch = None
    if world > 1:
        # torch.distributed is the launcher-side plumbing (rendezvous, barriers, max over ranks); the data path --
        # slice render + ncclReduce of the accumulation buffers -- runs inside the C++ core (fr_render_sharded)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene = build_scene(args)
    L = scenes.STANDARD_LIGHTING
    cam = camera_for(api.camera_walk)
    W, H, spp, depth = args.width, args.height, args.spp, args.max_depth
    n_pixels = W * H

    r = Renderer(local_rank)
    r.set_scene(scene)
    r.build_accel()
    accel = r.accel_info()
    r.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
    r.load_arhosek_sky(L["turbidity"], L["albedo"])
    r.set_resolution(W, H)
    r.set_max_wave_paths(wave_paths(args))
    dev = api.DeviceLayers(W, H, names=("beauty",))
    if world > 1:
        from fredholm_b200 import parallel
        parallel.init_core_communicator(r, dist)    # rank 0's id through the process group, ncclCommInitRank inside the core

    def clear():
        dev.clear()

    def step(total_spp=None):
        """One frame, device resident.  N = 1: render(spp).  N > 1: every rank renders its slice of a
        (spp x N)-sample frame as sums, ONE ncclReduce to rank 0 and the division there -- all enqueued by the
        core on the renderer's stream."""
        r.init_render_states()
        if world > 1:
            r.render_sharded(cam, (0, 0, 0), dev, total_spp or spp * world, depth, root=0)
        else:
            r.render(cam, (0, 0, 0), dev, total_spp or spp, depth)

    def sync_all():
        r.wait()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        clear()
        step()
    sync_all()

    # ---- timed: device-resident frames ----
    r.reset_statistics()
    r.set_stage_timing(True)
    r.stage_times()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    step_ms = []
    for _ in range(args.steps):
        clear()
        sync_all()
        e0 = r.record_event()
        step()
        e1 = r.record_event()
        sync_all()
        # device time from CUDA events on the launching stream (the render kernels and, for N > 1, the
        # ncclReduce the core enqueues on the same stream); max over ranks below
        step_ms.append(api.event_elapsed_ms(e0, e1))
    stats = r.statistics()
    stages = r.stage_times()
    r.set_stage_timing(False)
    wave_state_bytes = r.wave_state_bytes()   # of the timed frames (the 4096-sample sub-record below holds more)
    total_ms = max_over_ranks(float(sum(step_ms)))
    if world > 1:
        cnt = torch.tensor([stats["paths"], stats["rays"], stats["kernel_launches"], stats["rays_skipped"]],
                           dtype=torch.float64, device="cuda")
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        paths_all, rays_all, launches_all, skipped_all = [float(v) for v in cnt.tolist()]
    else:
        paths_all, rays_all, launches_all, skipped_all = (float(stats["paths"]), float(stats["rays"]),
                                                          float(stats["kernel_launches"]), float(stats["rays_skipped"]))

    # ---- timed: end to end through host buffers (every rank, own frame; N=1 is the headline) ----
    host = {"beauty": api.pinned_array((H, W, 4))}
    r.render_frame_host(cam, (0, 0, 0), spp, depth, names=("beauty",), out=host)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r.render_frame_host(cam, (0, 0, 0), spp, depth, names=("beauty",), out=host)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clk = clocks.stop() if rank == 0 else None
    image_mean = float(host["beauty"][..., :3].mean())

    # ---- untimed: one frame through the counting kernels -> nodes / triangles per ray (rank 0 only needs it) ----
    per_ray = {}
    if rank == 0:
        r.set_traversal_counting(True)
        r.reset_statistics()
        clear()
        r.init_render_states()
        r.render(cam, (0, 0, 0), dev, spp, depth)
        r.wait()
        cst, cnt2 = r.statistics(), r.traversal_counters()
        r.set_traversal_counting(False)
        n_vis = cst["rays_shadow"] + (cst["rays_light"] if cnt2["light"][0] == 0 else 0)  # no emitters: MIS rays are visibility rays
        per_ray["trace_closest"] = (cnt2["radiance"][0] / max(cst["rays_radiance"], 1), cnt2["radiance"][1] / max(cst["rays_radiance"], 1))
        per_ray["trace_shadow"] = (cnt2["shadow"][0] / max(n_vis, 1), cnt2["shadow"][1] / max(n_vis, 1))
        per_ray["trace_light"] = (per_ray["trace_shadow"] if cnt2["light"][0] == 0 else
                                  (cnt2["light"][0] / max(cst["rays_light"], 1), cnt2["light"][1] / max(cst["rays_light"], 1)))

    # ---- strong scaling sub-record: BASELINE config 3, --strong-spp samples in total, split over the N GPUs ----
    strong = None
    if args.strong_spp > 0:
        total = args.strong_spp
        clear()
        step(total_spp=min(total, 2 * spp * world))   # warm-up at the real wave size (no allocation in the timed run)
        clear()
        sync_all()
        e0 = r.record_event()
        step(total_spp=total)
        e1 = r.record_event()
        sync_all()
        strong_ms = max_over_ranks(api.event_elapsed_ms(e0, e1))
        strong = {"workload": "BASELINE config 3: the same frame at %d spp in total, sample-sharded over %d GPU(s), one ncclReduce"
                              % (total, world),
                  "spp_total": total, "seconds": strong_ms / 1e3, "mpaths_per_s": n_pixels * total / strong_ms / 1e3}

    if rank != 0:
        if world > 1:
            r.comm_destroy()
            dist.destroy_process_group()
        return

    secs = total_ms / 1e3
    value = paths_all / secs / 1e6
    if strong:
        # how long ONE GPU needs for the same samples at the per-GPU rate of this run's headline
        one_gpu_s = n_pixels * strong["spp_total"] / (value / world * 1e6)
        strong["one_gpu_seconds_at_this_runs_rate"] = one_gpu_s
        strong["speedup_vs_one_gpu"] = one_gpu_s / strong["seconds"]
        strong["efficiency"] = strong["speedup_vs_one_gpu"] / world

    # ---- roofline: measured live.  Traversal and shade are bound by SM issue slots; the queue
    #      passes (generate, film) by HBM.  Per-launch figures of the dominant kernel + the issue-slot fraction of
    #      every traversal stage ----
    peak_hbm, peak_src = measured_peaks()
    attrs = api.device_attributes(local_rank)
    clock_mhz = (clk or {}).get("sm_mhz") or attrs["clock_khz"] / 1e3
    sass = load_json(os.path.join(ROOT, "fredholm_b200", "sass_counts.json"))
    stage_rays = {"trace_closest": stats["rays_radiance"], "trace_shadow": stats["rays_shadow"],
                  "trace_light": stats["rays_light"]}
    issue = {k: issue_roofline(k, stage_rays[k], stages[k][0], per_ray.get(k, (0, 0)), sass, attrs["sm_count"], clock_mhz)
             for k in stage_rays}
    top = max(stage_rays, key=lambda k: stages[k][0])
    ms_top, n_top = stages[top]
    units_per_launch = stage_rays[top] / max(n_top, 1)
    hbm_achieved = (stage_rays[top] * ALGO_BYTES[top] / 1e9) / (ms_top / 1e3) if ms_top > 0 else 0.0
    # DRAM traffic and hardware counters of the same kernel from the committed ncu --set full capture of this
    # build (profiles/r2_ncu_traversal.json, written by tools/ncu_to_json.py; stale if the cubin hash differs)
    ncu = load_json(os.path.join(ROOT, "profiles", "r2_ncu_traversal.json")) or {}
    ncu_k = (ncu.get("kernels") or {}).get("k_" + top) or {}
    traffic = ncu_k.get("dram_bytes_per_ray") * units_per_launch if ncu_k.get("dram_bytes_per_ray") else None
    hbm_passes = {}
    for k, units in (("generate", stats["paths"]), ("film", stats["paths"])):
        ms_k = stages[k][0]
        if ms_k > 0:
            gbs = units * ALGO_BYTES[k] / 1e9 / (ms_k / 1e3)
            hbm_passes["k_" + k] = {"achieved": gbs, "peak": peak_hbm, "frac": gbs / peak_hbm, "unit": "GB/s",
                                    "algorithmic_bytes_per_path": ALGO_BYTES[k], "ms": ms_k}
    it = issue[top] or {}
    roofline = {"kernel": "k_" + top, "bound": "sm_issue",
                "achieved": it.get("achieved"), "peak": it.get("peak"), "unit": "Gwarp-inst/s", "frac": it.get("frac"),
                "traffic": traffic,
                "definition": "useful warp-instructions/s = rays/s x (nodes/ray x c_node + tris/ray x c_tri) / 32 over "
                              "SMs x 4 schedulers x SM clock; nodes/ray and tris/ray from one extra untimed frame of "
                              "the counting kernels, c_node / c_tri = SASS instructions of one node / triangle test "
                              "in this build (fredholm_b200/sass_counts.json)",
                "sm_count": attrs["sm_count"], "sm_clock_mhz": clock_mhz,
                "cubin_sha256": (sass or {}).get("cubin_sha256"),
                "launches": n_top, "avg_launch_ms": ms_top / max(n_top, 1), "units_per_launch": units_per_launch,
                "sm_issue": issue,
                "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": peak_hbm, "unit": "GB/s",
                        "frac": hbm_achieved / peak_hbm, "algorithmic_bytes_per_unit": ALGO_BYTES[top],
                        "peak_source": peak_src,
                        "note": "queue bytes of the dominant kernel against HBM: small by construction, not the binding roof"},
                "hbm_bound_passes": hbm_passes,
                "ncu": {"source": ncu.get("source"), "capture_matches_build": bool(ncu) and ncu.get("cubin_sha256") == (sass or {}).get("cubin_sha256"),
                        "kernel": ncu_k},
                "traffic_source": "profiles/r2_ncu_traversal.json: dram__bytes_read.sum + dram__bytes_write.sum per ray of the "
                                  "ncu --set full capture x rays per launch here",
                "stage_ms": {k: round(v[0], 3) for k, v in stages.items()},
                "stage_share": {k: round(v[0] / max(sum(x[0] for x in stages.values()), 1e-9), 4)
                                for k, v in stages.items()}}
    line = {
        "metric": "Mpaths/s (1080p, 1M tris, Standard Surface + Hosek sky, depth 10)",
        "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, scene), "samples_per_gpu": spp, "parallelism": "sample-sharded x%d" % world,
                   "collective": None if world == 1 else "one ncclReduce of the beauty sums per frame, issued by the C++ core (fr_render_sharded)",
                   "wave_paths": wave_paths(args),
                   "wave_state_gb": round(wave_state_bytes / 1e9, 2),
                   "wave_compaction": "waves of %d samples; the paths alive after 3 bounces finish in a straggler set shared by the frame's waves" % WAVE_SPP,
                   "l2": "per-step working set (path state + queues, %.1f GB) exceeds the 126 MB L2; no explicit flush"
                         % (wave_state_bytes / 1e9),
                   "bvh": {"nodes": accel["n_nodes"], "depth": accel["depth"], "build_ms": round(accel["build_ms"], 2),
                           "bytes": accel["bytes"]}},
        "mrays_per_s": rays_all / secs / 1e6,
        "rays_per_path": rays_all / max(paths_all, 1),
        "ray_accounting": {"traced": rays_all, "zero_contribution_rays_not_traced": skipped_all,
                           "reference_trace_calls": rays_all + skipped_all,
                           "mrays_per_s_reference_accounting": (rays_all + skipped_all) / secs / 1e6,
                           "rays_per_path_reference_accounting": (rays_all + skipped_all) / max(paths_all, 1)},
        "gpu_launches": int(launches_all),
        "clocks": clk,
        "e2e": {"value": paths_all / e2e_s / 1e6 if e2e_s > 0 else None,
                "unit": "Mpaths/s", "h2d_bytes_per_step": 64, "d2h_bytes_per_step": n_pixels * 16,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "note": "fr_render_frame_host: clear + render + framebuffer read-back to pinned host memory"},
        "roofline": roofline,
        "strong": strong,
        "image_mean": image_mean,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, scene)
    else:
        line["cpu_baseline"] = None
    emit(line)
    if world > 1:
        r.comm_destroy()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def slots_per_sample(W, H):
    """8x4 pixel tiles, one path slot per tile lane (fredholm_b200/csrc/wavefront.h)."""
    return ((W + 7) // 8) * ((H + 3) // 4) * 32


WAVE_SPP = 16


def wave_paths(args):
    """Waves of 16 samples unless overridden.  With wave compaction (the stragglers of the four waves of a frame
    finish their late bounces together) that is 12.7 GB of wave state at 2 % below the throughput of the whole
    frame in ONE wave, which holds 35.6 GB (--wave-paths 132710400; profiles/r2k_wave_sweep.jsonl)."""
    return args.wave_paths or slots_per_sample(args.width, args.height) * min(args.spp, WAVE_SPP)


def main():
    args = parse_args()
    # the contract is ONE JSON line on stdout: native libraries (NCCL's version banner) write to
    # fd 1 directly, so fd 1 is pointed at stderr while working and only the line goes to the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
