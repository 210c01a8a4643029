set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t2.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err
python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_n1_cfg5.json 2> gpurun_out/r02a_bench_n1_cfg5.err
python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_n1_cfg3.json 2> gpurun_out/r02a_bench_n1_cfg3.err
tail -5 gpurun_out/r02_t2.log
