run() { echo "$1"; env $1 python tools/stage_bench.py --spp 16 --reps 4 2>&1 | grep "untimed\|accel" | sed 's/.*n_nodes.: \([0-9]*\).*build_ms.: \([0-9.]*\).*/  nodes \1 build_ms \2/' ; }
run "FRD_SAH_CT=1.0"
for c in 0.5 0.7 0.85 1.2 1.5; do run "FRD_SAH_CT=$c"; done
for r in 4 16 32; do run "FRD_PLOC_RADIUS=$r"; done
run "FRD_BVH_BUILDER=lbvh"
run "FRD_COLLAPSE=greedy"
run "FRD_SAH_CT=1.0"
