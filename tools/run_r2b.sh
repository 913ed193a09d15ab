for rc in 8 16 24 32; do
  echo "== FRD_REFILL_LANES_COHERENT=$rc depth 1" 
  FRD_REFILL_LANES_COHERENT=$rc python tools/exp_spw.py --spw 1,8,32 --depth 1 --reps 2
done
echo "== depth 10, refill_coherent 32"
FRD_REFILL_LANES_COHERENT=32 python tools/exp_spw.py --spw 1,32 --depth 10 --reps 2
