set -x
for i in 1 2 3; do
timeout 120 python tools/profile_step.py | grep statevector
done
timeout 120 python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern | grep statevector
