#!/usr/bin/env python3
"""Per CUDA source line view of one kernel from `ncu -i rep --page source --csv --print-source cuda,sass`
output: warp instructions, share, active lanes and stall samples per line, heaviest first.
Usage: ncu_lines.py file.csv [top_n]"""
import csv, os, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = []; path = ""; cols = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": path = os.path.basename(r[1]); continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": cols = {h: i for i, h in enumerate(r)}; continue
    if cols and r[0].strip().isdigit():
        try:
            n = int(r[cols["Instructions Executed"]]); th = int(r[cols["Thread Instructions Executed"]]); s = int(r[cols["# Samples"]])
        except (ValueError, IndexError):
            continue
        out.append((n, th, s, path, int(r[0]), r[1].strip()))
tot = sum(o[0] for o in out) or 1; tots = sum(o[2] for o in out) or 1
print("# %s: %d source lines, %d warp instrs, %d samples" % (sys.argv[1], len(out), tot, tots))
print("%-22s %12s %6s %6s %6s  source" % ("file:line", "warp instrs", "%ins", "lanes", "%samp"))
for n, th, s, p, ln, src in sorted(out, reverse=True)[:top]:
    print("%-22s %12d %5.1f%% %6.1f %5.1f%%  %s" % ("%s:%d" % (p, ln), n, 100.0 * n / tot, th / max(n, 1), 100.0 * s / tots, src[:90]))
