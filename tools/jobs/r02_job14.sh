set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "gradient or fullsize or medium_size or duplicate" > gpurun_out/r02_t14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t14.log
tail -3 gpurun_out/r02_t14.log
for i in 1 2; do
timeout 120 python tools/profile_step.py | grep -E "gradient"
DQGP_GRAD_POLL_ALL=1 timeout 120 python tools/profile_step.py | grep -E "gradient"
done
