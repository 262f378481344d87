# many C5 groups in a row (bench.py --scaling strong, 40 groups): co-run split and lanes in the steady state
for cfg in "12:8:2" "12:8:1" "11:9:2" "10:10:2" "13:7:2" "10:12:2"; do
IFS=: read n m l <<< "$cfg"
HBT_B200_CORUN_SAME=$n HBT_B200_CORUN_MIXED=$m HBT_B200_LANES=$l python bench.py --scaling strong --total-groups 40 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('same:mixed:lanes $cfg', '%.4e'%d['value'], 'ms per group', round(d['ms_per_step']/40,3))"
done
