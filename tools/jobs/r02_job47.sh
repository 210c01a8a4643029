for v in 0 1 0 1; do
  if [ $v = 1 ]; then export DQGP_GEMM_NO_PRELOAD=1; else unset DQGP_GEMM_NO_PRELOAD; fi
  python bench.py --steps 12 --warmup 4 --also "" --no-cpu-baseline --skip-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4 N=1 no_preload=$v', d['ms_per_step'], d['phases_ms_one_agent']['factor'], d['final_nll_rank0'][0])"
done
