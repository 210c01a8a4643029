set -x
python -m pytest tests -m gpu -x -q -k "gradient or fullsize or train_and_update or medium_size or duplicate or trajectory" > gpurun_out/r02_t12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t12.log
tail -4 gpurun_out/r02_t12.log
for i in 1 2; do
python tools/profile_step.py | grep gradient
DQGP_GRAD_NO_P28=1 python tools/profile_step.py | grep gradient
done
