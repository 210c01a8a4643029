DQGP_GEMM_TMAP=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "dgemm or potrf or lookahead or lean" > gpurun_out/r02_t54.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_t54.log
DQGP_GEMM_TMAP=1 timeout 100 python tools/gemm_bench.py 2>&1 | grep "K=8192\|K=128\|K=256" | head -9
timeout 100 python tools/gemm_bench.py 2>&1 | grep "K=8192\|K=128\|K=256" | head -9
