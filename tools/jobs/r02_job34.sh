set -x
OBS=1,2,4,8,16 python tools/potrf_latency.py 8192 4096 > gpurun_out/r02_potrf_obs.txt 2>&1
cat gpurun_out/r02_potrf_obs.txt
python tools/gemm_bench.py > gpurun_out/r02_gemm_bench.txt 2>&1; cat gpurun_out/r02_gemm_bench.txt
