#!/usr/bin/env python3
"""Per-SASS-instruction view of one kernel launch from an .ncu-rep (source page):
prints contiguous regions of similar execution count with instructions executed, average
active threads and sampled stalls -- enough to see where a traversal / shade kernel spends
its issue slots.  Usage: ncu_source.py rep kernel_regex [launch_skip] [--dump]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else "0"
dump = "--dump" in sys.argv
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
c = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
tot = sum(int(r[c["Instructions Executed"]]) for r in body)
tot_s = sum(int(r[c["# Samples"]]) for r in body)
print("# %s  kernel %s launch %s: %d SASS lines, %d warp instrs, %d samples" % (rep, pat, skip, len(body), tot, tot_s))
if dump:
    for i, r in enumerate(body):
        print("%5d %10s %5.1f %6s  %s" % (i, r[c["Instructions Executed"]], float(r[c["Avg. Threads Executed"]] or 0),
                                          r[c["# Samples"]], r[c["Source"]].strip()))
    sys.exit(0)
# group into regions: consecutive instructions whose exec count is within 25%
regions = []
cur = None
for i, r in enumerate(body):
    n = int(r[c["Instructions Executed"]]); th = int(r[c["Thread Instructions Executed"]]); s = int(r[c["# Samples"]])
    if cur and (abs(n - cur["n0"]) <= 0.25 * max(cur["n0"], 1)):
        cur["n"] += n; cur["th"] += th; cur["s"] += s; cur["end"] = i
    else:
        cur = dict(start=i, end=i, n0=n, n=n, th=th, s=s)
        regions.append(cur)
print("%-12s %6s %14s %7s %7s %7s  first instruction" % ("sass lines", "len", "warp instrs", "%instr", "lanes", "%samp"))
for g in regions:
    if g["n"] < 0.004 * tot:
        continue
    print("%5d-%-6d %6d %14d %6.1f%% %7.1f %6.1f%%  %s" % (g["start"], g["end"], g["end"] - g["start"] + 1, g["n"], 100.0 * g["n"] / tot,
                                                    g["th"] / max(g["n"], 1), 100.0 * g["s"] / max(tot_s, 1), body[g["start"]][c["Source"]].strip()[:60]))
