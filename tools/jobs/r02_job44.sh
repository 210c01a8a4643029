python -m pytest tests -m gpu -q > gpurun_out/r02_t44.log 2>&1; tail -2 gpurun_out/r02_t44.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err; tail -c 300 gpurun_out/r02j_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02j_bench_n1.json').read().strip().splitlines()[-1])
print('cfg4 N=1', d['ms_per_step'], d['phases_ms_one_agent'], 'e2e/value', d['e2e']['value']/d['value'], d['roofline']['frac'], d['step_level']['frac'], d['clocks'])
for k,v in d['other_workloads'].items(): print(k, v['ms_per_step'])
PY
