python -m pytest tests -m gpu -q > gpurun_out/r02_t38.log 2>&1; tail -5 gpurun_out/r02_t38.log
for w in cfg3 cfg2 cfg1; do
  python bench.py --workload $w --steps 30 --warmup 5 --also "" --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w N=1', d['ms_per_step'], d.get('phases_ms_one_agent'), 'e2e', d['e2e']['value']/d['value'])"
done
