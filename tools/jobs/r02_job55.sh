DQGP_GEMM_TMAP=1 timeout 100 python tools/factor_breakdown.py 2>&1 | tail -2
timeout 100 python tools/factor_breakdown.py 2>&1 | tail -2
for v in 1 0; do
  if [ $v = 1 ]; then export DQGP_GEMM_TMAP=1; else unset DQGP_GEMM_TMAP; fi
  timeout 150 python bench.py --steps 8 --warmup 3 --also "" --no-cpu-baseline --skip-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4 N=1 gemm_tmap=$v', d['ms_per_step'], d['phases_ms_one_agent']['factor'], d['final_nll_rank0'][0], d['final_z_head'])"
done
