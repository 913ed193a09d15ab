run() { echo "$1"; env $1 python tools/stage_bench.py --spp 16 --reps 4 2>&1 | grep "untimed" ; }
run "FRD_REFILL_LANES=8"
for r in 4 6 12 16; do run "FRD_REFILL_LANES=$r"; done
for t in 2 4 8; do run "FRD_TRI_LANES=$t"; done
for t in 2 4 8; do run "FRD_TRI_LANES_ANY=$t"; done
run "FRD_REFILL_LANES=8"
