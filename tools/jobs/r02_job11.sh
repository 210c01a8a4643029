set -x
python -m pytest tests -m gpu -x -q -k "gradient or fullsize or train_and_update or medium_size or duplicate" > gpurun_out/r02_t11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t11.log
tail -4 gpurun_out/r02_t11.log
for i in 1 2; do
python tools/profile_step.py | grep gradient
DQGP_GRAD_NO_BULK=1 python tools/profile_step.py | grep gradient
done
python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern | grep gradient
DQGP_GRAD_NO_BULK=1 python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern | grep gradient
ncu --set full --clock-control none --import-source on -k regex:grad_projected -c 1 -o gpurun_out/r02_grad_cfg4_bulk -f python tools/profile_step.py --reps 1 > gpurun_out/r02_ncu_grad_bulk.log 2>&1
