#!/bin/bash
# Round 2: ncu --set full of the shade kernels of the first two bounces (benchmark scene, 16 spp): issue slots,
# L1/TEX throughput, load requests and sectors, stall reasons -> gpurun_out/r2_shade_ncu_full.txt
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_shade" -c 12 -f -o /tmp/r2_shade \
  python tools/profile_run.py --spp 16 > gpurun_out/r2_shade_run.log 2>&1
python tools/ncu_summary.py /tmp/r2_shade.ncu-rep > gpurun_out/r2_shade_ncu_full.txt 2>&1
for i in 1 7; do
  ncu -i /tmp/r2_shade.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:k_shade --launch-skip $i --launch-count 1 > /tmp/r2_shade_$i.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/r2_shade_$i.csv 60 > gpurun_out/r2_shade_lines_$i.txt
done
