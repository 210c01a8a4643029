python -m pytest tests -m gpu -q -x > gpurun_out/r02_t41.log 2>&1; tail -3 gpurun_out/r02_t41.log
OBS=1,2 python tools/potrf_latency.py 1024 2048 8192 2>&1 | tail -6
python tools/factor_breakdown.py 2>&1 | tail -2
