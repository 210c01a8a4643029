set -x
for i in 1 2; do
timeout 120 python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern | grep statevector
timeout 120 python tools/profile_step.py | grep statevector
DQGP_SV_PAIRED_SMALL=1 timeout 120 python tools/profile_step.py | grep statevector
done
DQGP_SV_PAIRED_SMALL=1 timeout 300 python -m pytest tests -m gpu -x -q -k "shared_prefix" 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q -k "shared_prefix or fullsize or medium" 2>&1 | tail -3
