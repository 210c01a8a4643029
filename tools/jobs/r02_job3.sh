set -x
python -m pytest tests -m gpu -x -q -k "gradient or outer or fullsize or gram" > gpurun_out/r02_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t3.log
python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern > gpurun_out/r02_step_cfg5_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:grad_projected -c 1 -o gpurun_out/r02_grad_matern_fast -f python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern --reps 1 > gpurun_out/r02_ncu_grad2.log 2>&1
tail -5 gpurun_out/r02_t3.log; cat gpurun_out/r02_step_cfg5_b.log
