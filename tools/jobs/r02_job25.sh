set -x
for i in 1 2; do
timeout 120 python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern | grep statevector
DQGP_SV_UNPAIRED=1 timeout 120 python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern | grep statevector
done
DQGP_SV_UNPAIRED=1 timeout 300 python -m pytest tests -m gpu -x -q -k "shared_prefix and (kyriienko-10 or kyriienko-11 or yz_cx-9 or kyriienko-12)" 2>&1 | tail -3
DQGP_SV_UNPAIRED=1 timeout 300 ncu --set full --clock-control none -k regex:statevec -c 1 -o gpurun_out/r02_sv_q10_unpaired -f python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern --reps 1 > gpurun_out/r02_ncu_sv8.log 2>&1
