#!/bin/bash
# Round 2: ncu --set full of every traversal launch of one 16-spp frame of the bench scene, all 10 bounces
# (final kernels), condensed to profiles/r2_ncu_traversal.json (bench.py roofline.traffic / roofline.ncu)
# and the per-launch text summary.  The .ncu-rep stays on the box.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_trace" -c 60 -f -o /tmp/r2_trav \
  python tools/profile_run.py --spp 16 --depth 10 --stats-out /tmp/r2_stats.json > gpurun_out/r2_ncu_run.log 2>&1
python tools/ncu_summary.py /tmp/r2_trav.ncu-rep > gpurun_out/r2_kernels_ncu_full.txt 2>&1
python tools/ncu_to_json.py /tmp/r2_trav.ncu-rep /tmp/r2_stats.json gpurun_out/r2_ncu_traversal.json > gpurun_out/r2_ncu_to_json.log 2>&1
python tools/ncu_source.py /tmp/r2_trav.ncu-rep k_trace_closest 0 > gpurun_out/r2_trace_closest_source.txt 2>&1
python tools/ncu_source.py /tmp/r2_trav.ncu-rep k_trace_shadow 0 > gpurun_out/r2_trace_shadow_source.txt 2>&1
