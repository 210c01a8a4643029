set -x
python -m pytest tests -m gpu -x -q -k "not cfg5" > gpurun_out/r02_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t1.log
python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern > gpurun_out/r02_step_cfg5.log 2>&1
python tools/profile_step.py > gpurun_out/r02_step_cfg4.log 2>&1
python tools/profile_step.py --n 2048 --encoding hubregtsen --kernel fidelity --q 5 --layers 2 --d 2 > gpurun_out/r02_step_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:statevec -c 1 -o gpurun_out/r02_sv_q10 -f python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern --reps 1 > gpurun_out/r02_ncu_sv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:grad_projected -c 1 -o gpurun_out/r02_grad_matern -f python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern --reps 1 > gpurun_out/r02_ncu_grad.log 2>&1
tail -3 gpurun_out/r02_t1.log; cat gpurun_out/r02_step_cfg5.log
