run() { echo "$1"; env $1 python tools/stage_bench.py --spp 16 --reps 4 --count 2>&1 | grep "untimed\|nodes/ray" ; }
for r in 2 3 4 5 6 8 4; do run "FRD_PLOC_RADIUS=$r"; done
