set -x
python -m pytest tests -m gpu -x -q -k "shared_prefix or features or fullsize or medium_size or trajectory or jacobian" > gpurun_out/r02_t5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t5.log
tail -15 gpurun_out/r02_t5.log
python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern > gpurun_out/r02_step_cfg5_c.log 2>&1
DQGP_SV_NO_PAIR=1 python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern > gpurun_out/r02_step_cfg5_c_nopair.log 2>&1
python tools/profile_step.py > gpurun_out/r02_step_cfg4_c.log 2>&1
DQGP_SV_NO_PAIR=1 python tools/profile_step.py > gpurun_out/r02_step_cfg4_c_nopair.log 2>&1
grep statevector gpurun_out/r02_step_cfg*_c*.log
ncu --set full --clock-control none --import-source on -k regex:statevec -c 1 -o gpurun_out/r02_sv_q10_lc2 -f python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern --reps 1 > gpurun_out/r02_ncu_sv2.log 2>&1
