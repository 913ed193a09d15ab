#!/usr/bin/env python3
"""Benchmark of the fredholm hot path (BASELINE.json): Mpaths/s and Mrays/s of the
path-tracing core on the 1080p / 1 048 576-triangle Standard-Surface scene
(Hosek sky + directional sun, 64 spp, depth 10).

  python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA core, C ABI)
  torchrun ... bench.py --gpus N ...                       N ranks, sample-sharded
  python bench.py --impl reference ...                     the reference's own integrator
                                                           (host build, oracle/_ref) on
                                                           the box's CPU cores

A "step" is one frame: 64 samples for every pixel of the 1920x1080 image.  With N GPUs
every rank renders its own 64-sample slice [64 r, 64 (r+1)) of a 64 N-sample frame (same
scene / BVH replicated, weak scaling) into sum-accumulators, followed by one NCCL reduce
of the beauty sums to rank 0 and the division by the sample count -- the reduce is inside
the timed region.

Prints ONE JSON line (rank 0).  `value` is whole-job Mpaths/s with everything resident in
HBM; `e2e` is the same frame through the host-buffer C-ABI call (clear, render, read the
framebuffer back to pinned host memory).
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
REF_SAMPLE = dict(window=(600, 337, 1320, 742), spp=64)  # 720x405 centre crop of the same frame, all 64 samples (~10 s on 16 cores)


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
                    help="paths in flight per wave (default: the whole frame, spp x tiled pixels)")
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
def run_reference(args):
    """The reference's own integrator sources compiled for the host (oracle/_ref), all
    host threads, on a bounded sample of the workload: a 480x270 window of the 1080p
    frame at 16 spp."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as ob
    if not ob.available():
        ob.build()
    scene = build_scene(args)
    cam = camera_for(ob.camera_walk)
    from fredholm_b200 import scenes
    L = scenes.STANDARD_LIGHTING
    o = ob.Oracle()
    o.set_scene(scene)
    o.build_accel()
    o.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
    o.load_arhosek_sky(L["turbidity"], L["albedo"])
    o.set_resolution(args.width, args.height)
    cores = os.cpu_count() or 1
    x0, y0, x1, y1 = REF_SAMPLE["window"]
    sx, sy = args.width / WORKLOAD["width"], args.height / WORKLOAD["height"]
    win = (int(x0 * sx), int(y0 * sy), int(x1 * sx), int(y1 * sy))
    spp = REF_SAMPLE["spp"]
    n_paths = (win[2] - win[0]) * (win[3] - win[1]) * spp
    layers = o.new_layers()

    def step():
        o.init_render_states()
        secs = 0.0
        for _ in range(spp):  # canonical mode: one sample per launch
            secs += o.render(cam, (0, 0, 0), layers, 1, args.max_depth, window=win, n_threads=cores)
        return secs

    for _ in range(args.warmup):
        step()
    o.reset_ray_counts()
    t = 0.0
    for _ in range(args.steps):
        t += step()
    rays = o.ray_counts()["rays"]
    value = n_paths * args.steps / t / 1e6
    sample = "window %s of the %dx%d frame, %d spp (%d paths per step)" % (win, args.width, args.height, spp, n_paths)
    line = {
        "impl": "reference",
        "metric": "Mpaths/s (1080p, 1M tris, Standard Surface + Hosek sky, depth 10)",
        "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, scene), "sample": sample},
        "mrays_per_s": rays / t / 1e6,
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------------------
def cpu_baseline(args, scene):
    """Reference integrator (oracle/_ref) timed once on this box's host cores."""
    try:
        from oracle import binding as ob
        from fredholm_b200 import scenes
        if not ob.available():
            return None
        L = scenes.STANDARD_LIGHTING
        o = ob.Oracle()
        o.set_scene(scene)
        o.build_accel()
        o.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
        o.load_arhosek_sky(L["turbidity"], L["albedo"])
        o.set_resolution(args.width, args.height)
        cam = camera_for(ob.camera_walk)
        cores = os.cpu_count() or 1
        x0, y0, x1, y1 = REF_SAMPLE["window"]
        sx, sy = args.width / WORKLOAD["width"], args.height / WORKLOAD["height"]
        win = (int(x0 * sx), int(y0 * sy), int(x1 * sx), int(y1 * sy))
        spp = REF_SAMPLE["spp"]
        layers = o.new_layers()
        secs = 0.0
        for _ in range(spp):
            secs += o.render(cam, (0, 0, 0), layers, 1, args.max_depth, window=win, n_threads=cores)
        n_paths = (win[2] - win[0]) * (win[3] - win[1]) * spp
        return {"value": n_paths / secs / 1e6, "unit": "Mpaths/s", "cores": cores, "kind": "reference",
                "sample": "window %s of the frame, %d spp, %.1f s" % (win, spp, secs),
                "mrays_per_s": o.ray_counts()["rays"] / secs / 1e6}
    except Exception as e:  # the baseline is reported, never required
        return {"value": None, "unit": "Mpaths/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % e}


def run_ours(args):
    from fredholm_b200 import Renderer, api, scenes
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch = None
    if world > 1:
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

    if world > 1:
        # torch's current stream := the renderer's stream, so the NCCL reduce is stream-ordered
        # after the render kernels and the CUDA events recorded on that stream bracket both
        torch.cuda.set_stream(torch.cuda.ExternalStream(r.stream(), device=torch.device("cuda", local_rank)))
        beauty = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        layers = {"beauty": beauty.data_ptr()}
        r.set_film_mode("sum")
    else:
        dev = api.DeviceLayers(W, H, names=("beauty",))
        layers = dev

    def clear():
        if world > 1:
            beauty.zero_()
            torch.cuda.synchronize()
        else:
            dev.clear()

    def step():
        """One frame, device resident."""
        r.set_sample_offset(rank * spp)
        r.render(cam, (0, 0, 0), layers, spp, depth)
        if world > 1:
            dist.reduce(beauty, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                r.scale_layers(layers, 1.0 / (spp * world))

    def sync_all():
        r.wait()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

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
        # device time from CUDA events on the launching stream (render kernels and, for N > 1,
        # the NCCL reduce, which torch orders on the same stream); max over ranks below
        step_ms.append(api.event_elapsed_ms(e0, e1))
    stats = r.statistics()
    stages = r.stage_times()
    r.set_stage_timing(False)
    total_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        cnt = torch.tensor([stats["paths"], stats["rays"], stats["kernel_launches"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        paths_all, rays_all, launches_all = [float(v) for v in cnt.tolist()]
    else:
        paths_all, rays_all, launches_all = float(stats["paths"]), float(stats["rays"]), float(stats["kernel_launches"])

    # ---- timed: end to end through host buffers (every rank, own slice; N=1 is the headline) ----
    r.set_film_mode("mean")
    host = {"beauty": api.pinned_array((H, W, 4))}
    r.render_frame_host(cam, (0, 0, 0), spp, depth, names=("beauty",), out=host)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r.render_frame_host(cam, (0, 0, 0), spp, depth, names=("beauty",), out=host)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clk = clocks.stop() if rank == 0 else None
    image_mean = float(host["beauty"][..., :3].mean())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    secs = total_ms / 1e3
    value = paths_all / secs / 1e6
    # ---- roofline of the dominant kernel (per launch, live CUDA-event durations) ----
    peak, peak_src = measured_peaks()
    top = max(("trace_closest", "trace_shadow", "trace_light", "shade"), key=lambda k: stages[k][0])
    per_rank = {"trace_closest": stats["rays_radiance"], "trace_shadow": stats["rays_shadow"],
                "trace_light": stats["rays_light"], "shade": stats["rays_radiance"]}[top]
    # algorithmic HBM bytes per unit (DESIGN.md "Kernels"): queue index + ray record in, result out
    bytes_per_unit = {"trace_closest": 4 + 32 + 16, "trace_shadow": 48 + 16, "trace_light": 48 + 32,
                      "shade": 4 + 32 + 16 + 16 + 32 + 16 + 4 + 3 * 48 + 48}[top]
    ms_top, n_top = stages[top]
    achieved = (per_rank * bytes_per_unit / 1e9) / (ms_top / 1e3) if ms_top > 0 else 0.0
    # DRAM bytes per ray of that kernel from the committed `ncu --set full` capture
    # (profiles/r1k_kernels_ncu_full.txt: dram__bytes_read.sum + dram__bytes_write.sum over the rays of
    # the captured launches -- 33.18 M primary rays; <= 3 x 18.2 M depth-0 visibility rays), scaled to the
    # rays of one launch here
    ncu_dram_bytes_per_unit = {"trace_closest": 68.0, "trace_shadow": 99.0}.get(top)
    ncu_issue = {"trace_closest": (70.4, 20.7), "trace_shadow": (72.9, 19.0)}.get(top, (None, None))
    units_per_launch = per_rank / max(n_top, 1)
    traffic = ncu_dram_bytes_per_unit * units_per_launch if ncu_dram_bytes_per_unit else None
    roofline = {"kernel": "k_" + top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "traffic_source": "ncu --set full capture (profiles/), DRAM bytes per ray x rays per launch",
                "sm_issue": {"issue_slots_busy_pct": ncu_issue[0], "active_lanes_per_instruction": ncu_issue[1],
                             "source": "profiles/r1k_kernels_ncu_full.txt (ncu --set full of the depth-0 launches, 16 spp)"},
                "launches": n_top, "avg_launch_ms": ms_top / max(n_top, 1),
                "units_per_launch": units_per_launch, "algorithmic_bytes_per_unit": bytes_per_unit,
                "note": "traversal is SM-issue / latency bound, not HBM bound (SURVEY.md 8(d)); issue-slot "
                        "utilisation from ncu is in profiles/",
                "stage_ms": {k: round(v[0], 3) for k, v in stages.items()},
                "stage_share": {k: round(v[0] / max(sum(x[0] for x in stages.values()), 1e-9), 4)
                                for k, v in stages.items()}}
    line = {
        "metric": "Mpaths/s (1080p, 1M tris, Standard Surface + Hosek sky, depth 10)",
        "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, scene), "samples_per_gpu": spp, "parallelism": "sample-sharded x%d" % world,
                   "wave_paths": wave_paths(args),
                   "l2": "per-step working set (path state + queues, %.1f GB) exceeds the 126 MB L2; no explicit flush"
                         % (r_state_gb(W, H, spp, args)),
                   "bvh": {"nodes": accel["n_nodes"], "depth": accel["depth"], "build_ms": round(accel["build_ms"], 2),
                           "bytes": accel["bytes"]}},
        "mrays_per_s": rays_all / secs / 1e6,
        "rays_per_path": rays_all / max(paths_all, 1),
        "gpu_launches": int(launches_all),
        "clocks": clk,
        "e2e": {"value": paths_all / e2e_s / 1e6 if e2e_s > 0 else None,
                "unit": "Mpaths/s", "h2d_bytes_per_step": 64, "d2h_bytes_per_step": n_pixels * 16,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "note": "fr_render_frame_host: clear + render + framebuffer read-back to pinned host memory"},
        "roofline": roofline,
        "image_mean": image_mean,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, scene)
    else:
        line["cpu_baseline"] = None
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def slots_per_sample(W, H):
    """8x4 pixel tiles, one path slot per tile lane (fredholm_b200/csrc/wavefront.h)."""
    return ((W + 7) // 8) * ((H + 3) // 4) * 32


def wave_paths(args):
    """All samples of the frame in ONE wave unless overridden: fewer, larger launches (the
    persistent kernels have a fixed tail per launch) at the price of HBM for the wave state."""
    return args.wave_paths or slots_per_sample(args.width, args.height) * args.spp


PATH_STATE_BYTES = 8 * 16 + (2 + 9) * 4 + 4 * 48   # path SoA + queues (integrator.cpp: ensure_capacity)


def r_state_gb(W, H, spp, args):
    slots = slots_per_sample(W, H)
    per_wave = max(1, min(spp, wave_paths(args) // slots))
    return per_wave * slots * PATH_STATE_BYTES / 1e9


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
