#!/bin/bash
# A/B of the 8-wide collapse: largest-area-first (FRD_COLLAPSE=greedy) against the SAH-cost choice (default) at
# several triangle-cost ratios.  Run on a GPU box: bash tools/exp_collapse.sh > gpurun_out/exp_collapse.txt
echo "== greedy"; FRD_COLLAPSE=greedy python tools/stage_bench.py --spp 16 --reps 3 --count 2>&1 | grep -v "^\[bvh\]"
for ct in 0.3 0.6 1.0 1.5; do
  echo "== sah c_tri=$ct"; FRD_SAH_CT=$ct python tools/stage_bench.py --spp 16 --reps 3 --count 2>&1 | grep -v "^\[bvh\]"
done
