set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "gradient or fullsize or train_and_update or medium_size or duplicate or trajectory or smoke" 2>&1 | tail -3
for i in 1 2 3; do timeout 120 python tools/profile_step.py | grep -E "gradient"; done
timeout 120 python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern | grep gradient
