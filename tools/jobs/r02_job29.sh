set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "dgemm or potrf or lookahead" 2>&1 | tail -2
for i in 1 2; do
timeout 120 python tools/factor_breakdown.py 8192 2>&1 | tail -1
DQGP_GEMM_NO_PREFETCH=1 timeout 120 python tools/factor_breakdown.py 8192 2>&1 | tail -1
done
timeout 100 python tools/gemm_bench.py 2>&1 | grep -E "K=128|K=256" 
DQGP_GEMM_NO_PREFETCH=1 timeout 100 python tools/gemm_bench.py 2>&1 | grep -E "K=128|K=256"
