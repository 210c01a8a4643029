set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "shared_prefix or features or fullsize or medium_size or jacobian" > gpurun_out/r02_t20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t20.log
tail -12 gpurun_out/r02_t20.log
for i in 1 2; do
timeout 120 python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern | grep statevector
DQGP_SV_NO_MAPPED=1 timeout 120 python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern | grep statevector
done
DQGP_SV_FORCE_LC2=1 timeout 120 python tools/profile_step.py | grep statevector
timeout 120 python tools/profile_step.py | grep statevector
timeout 300 ncu --set full --clock-control none --import-source on -k regex:statevec -c 1 -o gpurun_out/r02_sv_q10_lc3 -f python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern --reps 1 > gpurun_out/r02_ncu_sv6.log 2>&1
