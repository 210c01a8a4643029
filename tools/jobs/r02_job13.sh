set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t13.log
tail -4 gpurun_out/r02_t13.log
timeout 120 python tools/gemm_bench.py > gpurun_out/r02_gemm_bench_bulk.log 2>&1; cat gpurun_out/r02_gemm_bench_bulk.log
DQGP_GEMM_NO_BULK=1 timeout 120 python tools/gemm_bench.py | head -3
for i in 1 2; do
timeout 120 python tools/profile_step.py | grep -E "factor|step"
DQGP_GEMM_NO_BULK=1 timeout 120 python tools/profile_step.py | grep -E "factor|step"
done
timeout 120 python tools/profile_step.py --n 2048 --encoding hubregtsen --kernel fidelity --q 5 --layers 2 --d 2 | grep -E "gradient|gram|step"
DQGP_FID_NO_BULK=1 timeout 120 python tools/profile_step.py --n 2048 --encoding hubregtsen --kernel fidelity --q 5 --layers 2 --d 2 | grep -E "gradient|gram|step"
