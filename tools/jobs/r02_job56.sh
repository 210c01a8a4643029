timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/r02_t56.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r02_t56.log
timeout 150 python bench.py --steps 10 --warmup 3 --also "cfg5,cfg3" --no-cpu-baseline > gpurun_out/r02l_bench_n1_short.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02l_bench_n1_short.json').read().strip().splitlines()[-1])
print('cfg4 N=1', d['ms_per_step'], d['phases_ms_one_agent'], d['roofline']['frac'], d['step_level']['frac'], d['e2e']['value']/d['value'], d['final_z_head'])
for k,v in d['other_workloads'].items(): print(k, v['ms_per_step'])
"
