#!/usr/bin/env python3
"""ncu --set full capture of the depth-0 traversal launches -> the small JSON bench.py reads for
`roofline.traffic` and `roofline.ncu` (profiles/r2_ncu_traversal.json).

usage: ncu_to_json.py capture.ncu-rep stats.json out.json
stats.json: what tools/profile_run.py --stats-out wrote for the SAME run (ray counts of the captured frame), so
that DRAM bytes can be expressed per ray.  Hardware counters are sums / means over the captured launches of each
kernel (one k_trace_closest launch, three k_trace_shadow launches: sun, sky, MIS-as-visibility rays)."""
import csv, json, os, subprocess, sys

rep, stats_path, out_path = sys.argv[1:4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
stats = json.load(open(stats_path))
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def num(r, key):
    try:
        return float(r[col[key]].replace(",", ""))
    except Exception:
        return float("nan")


def unit(key):
    return rows[1][col[key]]


def to_bytes(v, u):
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


def to_ms(v, u):
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)


kernels = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = "k_trace_closest" if "k_trace_closest" in name else "k_trace_shadow" if "k_trace_shadow" in name else \
            "k_trace_light" if "k_trace_light" in name else None
    if not short:
        continue
    k = kernels.setdefault(short, {"launches": 0, "duration_ms": 0.0, "dram_bytes": 0.0, "warp_inst": 0.0, "thread_inst": 0.0,
                                   "issue_w": 0.0, "l1_w": 0.0, "l2_w": 0.0})
    d = to_ms(num(r, "gpu__time_duration.sum"), unit("gpu__time_duration.sum"))
    k["launches"] += 1
    k["duration_ms"] += d
    k["dram_bytes"] += to_bytes(num(r, "dram__bytes_read.sum"), unit("dram__bytes_read.sum")) + \
        to_bytes(num(r, "dram__bytes_write.sum"), unit("dram__bytes_write.sum"))
    wi = num(r, "smsp__inst_executed.sum")
    k["warp_inst"] += wi
    k["thread_inst"] += wi * num(r, "smsp__thread_inst_executed_per_inst_executed.ratio")
    k["issue_w"] += d * num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    k["l1_w"] += d * num(r, "l1tex__t_sector_hit_rate.pct")
    k["l2_w"] += d * num(r, "lts__t_sector_hit_rate.pct")
rays = {"k_trace_closest": stats["rays_radiance"], "k_trace_shadow": stats["rays_shadow"], "k_trace_light": stats["rays_light"]}
if "k_trace_light" not in kernels:  # no emitters: the MIS rays are traced by k_trace_shadow
    rays["k_trace_shadow"] += stats["rays_light"]
res = {}
for name, k in kernels.items():
    n = max(rays.get(name, 0), 1)
    res[name] = {"launches_captured": k["launches"], "rays_captured": rays.get(name, 0), "duration_ms": k["duration_ms"],
                 "grays_per_s_under_ncu": n / k["duration_ms"] / 1e6,
                 "dram_bytes_per_ray": k["dram_bytes"] / n, "warp_inst_per_ray": k["warp_inst"] / n,
                 "active_lanes_per_instruction": k["thread_inst"] / max(k["warp_inst"], 1),
                 "issue_slots_busy_pct": k["issue_w"] / k["duration_ms"], "l1_hit_pct": k["l1_w"] / k["duration_ms"],
                 "l2_hit_pct": k["l2_w"] / k["duration_ms"]}
sass = json.load(open(os.path.join(ROOT, "fredholm_b200", "sass_counts.json")))
json.dump({"source": "ncu --set full --clock-control none, depth-0 launches of the bench scene at %s spp (tools/ncu_r2.sh)" % stats.get("spp"),
           "cubin_sha256": sass["cubin_sha256"], "kernels": res}, open(out_path, "w"), indent=1)
print(json.dumps(res, indent=1))
