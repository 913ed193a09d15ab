#!/bin/bash
# Round-2c: does a warp of 32 camera rays through ONE pixel (spw 32, refill only when drained) run the
# closest-hit kernel differently from the 8x4-tile warps with dynamic refill?  ncu --set full, depth-0 launch, 32 spp.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_trace_closest|k_trace_shadow" -c 2 -f -o /tmp/r2c_base python tools/profile_run.py --spp 32 --depth 1 > gpurun_out/r2c_base.log 2>&1
FRD_REFILL_LANES_COHERENT=32 ncu --set full --clock-control none --import-source on -k regex:"k_trace_closest|k_trace_shadow" -c 2 -f -o /tmp/r2c_beam python tools/profile_run.py --spp 32 --depth 1 --spw 32 > gpurun_out/r2c_beam.log 2>&1
python tools/ncu_summary.py /tmp/r2c_base.ncu-rep > gpurun_out/r2c_base_ncu.txt 2>&1
python tools/ncu_summary.py /tmp/r2c_beam.ncu-rep > gpurun_out/r2c_beam_ncu.txt 2>&1
python tools/ncu_source.py /tmp/r2c_base.ncu-rep k_trace_closest 0 > gpurun_out/r2c_base_closest_source.txt 2>&1
python tools/ncu_source.py /tmp/r2c_beam.ncu-rep k_trace_closest 0 > gpurun_out/r2c_beam_closest_source.txt 2>&1
