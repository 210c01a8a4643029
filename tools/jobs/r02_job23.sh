set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t23.log
tail -5 gpurun_out/r02_t23.log
