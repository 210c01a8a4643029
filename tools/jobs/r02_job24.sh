set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t24.log
tail -5 gpurun_out/r02_t24.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
