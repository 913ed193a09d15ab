#!/bin/bash
# ncu --set full of the six depth-0 shade kernels (benchmark scene, 16 spp); per-source-line summaries come back.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_shade" -c 6 -f -o /tmp/shade python tools/profile_run.py --spp 16 > gpurun_out/r1i_shade.log 2>&1
for i in 0 1 2 3 4 5; do
  ncu -i /tmp/shade.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:k_shade --launch-skip $i --launch-count 1 > /tmp/shade_$i.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/shade_$i.csv 70 > gpurun_out/r1i_shade_lines_$i.txt
done
python tools/ncu_source.py /tmp/shade.ncu-rep k_shade 1 > gpurun_out/r1i_shade68_source.txt 2>&1
ls -la gpurun_out | tail -8
