set -x
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r02_t51.log 2>&1; tail -2 gpurun_out/r02_t51.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err; tail -c 200 gpurun_out/r02k_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02k_bench_n1.json').read().strip().splitlines()[-1])
print('cfg4 N=1', d['ms_per_step'], d['phases_ms_one_agent'], 'e2e/value', d['e2e']['value']/d['value'], d['roofline']['frac'], d['step_level']['frac'], d['clocks'])
print(d['rooflines']['gradient'])
for k,v in d['other_workloads'].items(): print(k, v['ms_per_step'])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:grad_projected -c 1 -o gpurun_out/r02_grad_cfg4_tmap -f python tools/profile_step.py --reps 1 > gpurun_out/r02_ncu_grad_tmap.log 2>&1; tail -2 gpurun_out/r02_ncu_grad_tmap.log
