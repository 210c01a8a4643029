python -m pytest tests -m gpu -q -x > gpurun_out/r02_t49.log 2>&1; tail -3 gpurun_out/r02_t49.log
for v in 0 1 0 1; do
  if [ $v = 1 ]; then export DQGP_FID_NO_TMAP=1; else unset DQGP_FID_NO_TMAP; fi
  python bench.py --workload cfg3 --steps 30 --warmup 5 --also "" --no-cpu-baseline --skip-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 N=1 no_tmap=$v', d['ms_per_step'], d['phases_ms_one_agent']['gradient'], d['final_nll_rank0'][0], d['final_z_head'])"
done
