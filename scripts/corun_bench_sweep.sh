# co-run split on the other bench shapes: resident pairs/s of bench.py --config c3 / c4 for a few (same : mixed) splits
for c in c3 c4; do for nm in 12:8 10:10 14:6 9:12; do
n=${nm%:*}; m=${nm#*:}
HBT_B200_CORUN_SAME=$n HBT_B200_CORUN_MIXED=$m python bench.py --config $c --no-cpu-baseline --no-e2e --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$c corun $nm', '%.4e'%d['value'], round(d['ms_per_step'],2))"
done; done
