#!/bin/bash
# One GPU-box pass over everything that gets a number in DESIGN.md / README.md (one B200):
#   GPU parity suite, bench.py (both arms), the one-wave operating point, the ncu launch list of the bench command,
#   the five BASELINE configurations, compute-sanitizer.  Outputs: gpurun_out/<tag>_*.
tag=${1:-r2n}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/${tag}_pytest_gpu.log
python bench.py > gpurun_out/${tag}_bench_line.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
python bench.py --wave-paths 132710400 --no-cpu-baseline --strong-spp 0 > gpurun_out/${tag}_bench_onewave.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_bench_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --strong-spp 0 > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 900 python tests/tools/config_runs.py c1 c2 c3 c4 c5 > gpurun_out/${tag}_config_runs.jsonl 2> gpurun_out/${tag}_config.err
{
  for t in memcheck racecheck synccheck; do
    echo "== compute-sanitizer --tool $t python tools/sanitize_run.py"
    timeout 600 compute-sanitizer --tool $t python tools/sanitize_run.py 2>&1 | grep -v "^=========     " | tail -12
  done
} > gpurun_out/${tag}_sanitizer.txt 2>&1
cat gpurun_out/${tag}_pytest_gpu.log
python - <<PY
import json
for n in ("bench_line", "bench_reference", "bench_onewave"):
    try:
        d = json.load(open("gpurun_out/${tag}_%s.json" % n))
        print(n, d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("config") or {}).get("wave_state_gb"), (d.get("roofline") or {}).get("frac"),
              ((d.get("roofline") or {}).get("ncu") or {}).get("capture_matches_build"), (d.get("strong") or {}).get("seconds"))
    except Exception as e:
        print(n, "unreadable:", e)
for l in open("gpurun_out/${tag}_config_runs.jsonl"):
    d = json.loads(l)
    for k, v in d.items():
        print(k, {a: (round(b, 5) if isinstance(b, float) else b) for a, b in v.items() if not isinstance(b, (dict, list))})
PY
grep -n "SUMMARY\|identical\|^==" gpurun_out/${tag}_sanitizer.txt
