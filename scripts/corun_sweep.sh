# sweep of the co-run split (resident same-event : mixed-event warps per SM) on one C5-shape group
# usage: corun_sweep.sh <variant> n:m ...      (0:0 = one kernel after the other)
v=$1; shift
for nm in "$@"; do
  n=${nm%:*}; m=${nm#*:}
  if [ $n = 0 ]; then e="--env HBT_B200_CORUN=0"; else e="--env HBT_B200_CORUN_SAME=$n --env HBT_B200_CORUN_MIXED=$m"; fi
  python scripts/ab_variants.py $v $e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v corun $nm', d.get('fused_ms'), d.get('same_ms'), d.get('mixed_ms'), d.get('accepted_same_per_launch'), d.get('accepted_mixed_per_launch'), d.get('error'))"
done
