timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_guards.py -m gpu -q -x -k "gradient or fullsize or agent_step or grad" > gpurun_out/r02_t50.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_t50.log
for v in 0 1 0; do
  if [ $v = 1 ]; then export DQGP_GRAD_NO_TMAP=1; else unset DQGP_GRAD_NO_TMAP; fi
  timeout 200 python bench.py --steps 10 --warmup 4 --also "" --no-cpu-baseline --skip-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4 N=1 no_tmap=$v', d['ms_per_step'], d['phases_ms_one_agent']['gradient'], d['final_nll_rank0'][0], d['final_z_head'])"
done
