set -x
python -m pytest tests -m gpu -x -q -k "shared_prefix or features or fullsize or medium_size" > gpurun_out/r02_t8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t8.log
tail -4 gpurun_out/r02_t8.log
DQGP_SV_FORCE_LC2=1 python tools/profile_step.py > gpurun_out/r02_step_cfg4_e_lc2.log 2>&1
python tools/profile_step.py > gpurun_out/r02_step_cfg4_e.log 2>&1
python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern > gpurun_out/r02_step_cfg5_e.log 2>&1
grep statevector gpurun_out/r02_step_cfg*_e*.log
DQGP_SV_FORCE_LC2=1 ncu --set full --clock-control none --import-source on -k regex:statevec -c 1 -o gpurun_out/r02_sv_q8_lc2c -f python tools/profile_step.py --reps 1 > gpurun_out/r02_ncu_sv5.log 2>&1
