for c in c2 c3; do for k in 4 8 12 16; do
HBT_B200_COALESCE_HOST=$k python bench.py --config $c --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$c coalesce_host=$k value', '%.3e'%d['value'], 'e2e', '%.3e'%d['e2e']['value'], 'e2e ms/step', round(d['e2e']['ms_per_step'],2), 'ms/step', round(d['ms_per_step'],2))"
done; done
