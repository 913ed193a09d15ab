#!/bin/bash
# Round-1i capture: ncu --set full of the first bounce's kernels of the benchmark scene at 16 spp; the .ncu-rep
# (100 MB) stays on the box, only the text summaries come back in gpurun_out/.
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shade|k_generate|k_miss|k_film" -c 24 -f -o /tmp/r1k_full python tools/profile_run.py --spp 16 > gpurun_out/r1k_full.log 2>&1
python tools/ncu_summary.py /tmp/r1k_full.ncu-rep > gpurun_out/r1k_kernels_ncu_full.txt 2>&1
python tools/ncu_source.py /tmp/r1k_full.ncu-rep k_trace_closest 0 > gpurun_out/r1k_trace_closest_source.txt 2>&1
python tools/ncu_source.py /tmp/r1k_full.ncu-rep k_trace_shadow 0 > gpurun_out/r1k_trace_shadow_source.txt 2>&1
python tools/ncu_source.py /tmp/r1k_full.ncu-rep k_shade 1 > gpurun_out/r1k_shade68_source.txt 2>&1
ncu -i /tmp/r1k_full.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:k_shade --launch-skip 1 --launch-count 1 > gpurun_out/r1k_shade68_cuda_sass.csv 2>&1
ncu -i /tmp/r1k_full.ncu-rep --page source --csv --print-source cuda,sass --kernel-name "regex:k_trace_closest" --launch-count 1 > gpurun_out/r1k_closest_cuda_sass.csv 2>&1
ls -la gpurun_out
python tools/ncu_lines.py gpurun_out/r1k_closest_cuda_sass.csv 60 > gpurun_out/r1k_trace_closest_cuda_lines.txt
rm -f gpurun_out/r1k_closest_cuda_sass.csv gpurun_out/r1k_shade68_cuda_sass.csv
