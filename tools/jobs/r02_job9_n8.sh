set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02c_bench_n8.json 2> gpurun_out/r02c_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --workload cfg5 > gpurun_out/r02c_bench_n8_cfg5.json 2> gpurun_out/r02c_bench_n8_cfg5.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --workload cfg3 > gpurun_out/r02c_bench_n8_cfg3.json 2> gpurun_out/r02c_bench_n8_cfg3.err
python -m pytest tests -m gpu -x -q -k "second_device" > gpurun_out/r02_t9.log 2>&1; tail -3 gpurun_out/r02_t9.log
tail -c 600 gpurun_out/r02c_bench_n8.err
