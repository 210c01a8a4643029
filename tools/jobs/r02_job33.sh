set -x
OBS=4 DQGP_POTRF_TRACE=1 python tools/potrf_latency.py 8192 > gpurun_out/r02_potrf_trace.txt 2>&1
tail -50 gpurun_out/r02_potrf_trace.txt
python tools/factor_breakdown.py > gpurun_out/r02_factor_breakdown.txt 2>&1; tail -12 gpurun_out/r02_factor_breakdown.txt
