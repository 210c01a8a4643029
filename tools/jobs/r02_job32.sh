set -x
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_t32.log 2>&1; tail -40 gpurun_out/r02_t32.log
