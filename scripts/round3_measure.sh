set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r03u_t_all.log 2>&1; tail -1 gpurun_out/r03u_t_all.log
python bench.py > gpurun_out/r03u_bench_c5.json 2> gpurun_out/r03u_bench_c5.err
for c in c2 c3 c4; do python bench.py --config $c --no-cpu-baseline > gpurun_out/r03u_bench_$c.json 2> gpurun_out/r03u_bench_$c.err; done
python bench.py --scaling strong --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r03u_bench_c5_strong.json 2> gpurun_out/r03u_bench_c5_strong.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r03u_bench_reference.json 2> gpurun_out/r03u_bench_reference.err
for f in c5 c2 c3 c4 c5_strong reference; do python -c "
import json
d=json.loads(open('gpurun_out/r03u_bench_$f.json').read().strip().splitlines()[-1])
print('$f', '%.4e'%d['value'], round(d.get('ms_per_step',0),2), 'e2e', '%.4e'%(d.get('e2e') or {}).get('value',0), 'frac', (d.get('roofline') or {}).get('frac'), 'launches', d.get('gpu_launches'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))"; done
compute-sanitizer --tool memcheck python scripts/sanitize_round3.py > gpurun_out/r03u_san_mem.log 2>&1; tail -2 gpurun_out/r03u_san_mem.log
compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_round3.py > gpurun_out/r03u_san_race.log 2>&1; tail -2 gpurun_out/r03u_san_race.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03u_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r03u_ncu_bench.log 2>&1
HBT_B200_CORUN=0 ncu --set full --clock-control none --import-source on -k regex:hbt_pairs -s 2 -c 2 -o gpurun_out/r03u_sep python scripts/profile_step.py --launches 2 > gpurun_out/r03u_ncu.log 2>&1; tail -1 gpurun_out/r03u_ncu.log
