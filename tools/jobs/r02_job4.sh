set -x
python -m pytest tests -m gpu -x -q -k "lu or dgemm_general or predict or honoured or repairs or walks or train_and_update" > gpurun_out/r02_t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t4.log
tail -25 gpurun_out/r02_t4.log
