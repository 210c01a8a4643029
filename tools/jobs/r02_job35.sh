set -x
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "dgemm or potrf or lean or lookahead or lu_" > gpurun_out/r02_t35.log 2>&1; tail -3 gpurun_out/r02_t35.log
python tools/gemm_bench.py > gpurun_out/r02_gemm_bench_preload.txt 2>&1; grep "K=128\|K=256" gpurun_out/r02_gemm_bench_preload.txt
DQGP_GEMM_NO_PRELOAD=1 python tools/gemm_bench.py 2>&1 | grep "K=128\|K=256"
python tools/factor_breakdown.py 2>&1 | tail -2
DQGP_GEMM_NO_PRELOAD=1 python tools/factor_breakdown.py 2>&1 | tail -2
