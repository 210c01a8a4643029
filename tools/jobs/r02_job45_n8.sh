set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02k_bench_n8.json 2> gpurun_out/r02k_bench_n8.err
tail -c 400 gpurun_out/r02k_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02k_bench_n8.json').read().strip().splitlines()[-1])
print('cfg4 N=8', d['ms_per_step'], 'e2e/value', d['e2e']['value']/d['value'], d['clocks'], d['final_z_head'])
for k,v in d['other_workloads'].items(): print(k, v['ms_per_step'])
PY
