set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t7.log
tail -4 gpurun_out/r02_t7.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_n1_cfg5.json 2> gpurun_out/r02c_bench_n1_cfg5.err
python tools/gemm_bench.py > gpurun_out/r02_gemm_bench.log 2>&1
cat gpurun_out/r02_gemm_bench.log
ncu --set full --clock-control none --import-source on -k regex:gemm_group -c 1 -o gpurun_out/r02_gemm_8192 -f python tools/gemm_bench.py > gpurun_out/r02_ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:grad_projected -c 1 -o gpurun_out/r02_grad_cfg4 -f python tools/profile_step.py --reps 1 > gpurun_out/r02_ncu_grad_cfg4.log 2>&1
