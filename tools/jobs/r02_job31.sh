set -x
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02_t31.log 2>&1; tail -25 gpurun_out/r02_t31.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02d_ref.json 2> gpurun_out/r02d_ref.err; tail -c 400 gpurun_out/r02d_ref.json; tail -c 400 gpurun_out/r02d_ref.err
