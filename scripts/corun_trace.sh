for nm in 12:8 9:12 13:7 6:16; do
  n=${nm%:*}; m=${nm#*:}
  HBT_B200_CORUN_TRACE=1 HBT_B200_CORUN_SAME=$n HBT_B200_CORUN_MIXED=$m python scripts/ab_variants.py --worker base --reps 2 2>&1 | grep -E "corun|fused_ms" | tail -3 | cut -c1-400
done
