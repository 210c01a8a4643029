set -x
python -m pytest tests -m gpu -x -q -k "shared_prefix or features or fullsize or medium_size" > gpurun_out/r02_t6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t6.log
tail -4 gpurun_out/r02_t6.log
python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern > gpurun_out/r02_step_cfg5_d.log 2>&1
python tools/profile_step.py > gpurun_out/r02_step_cfg4_d.log 2>&1
DQGP_SV_NO_LC2=1 python tools/profile_step.py > gpurun_out/r02_step_cfg4_d_old.log 2>&1
grep statevector gpurun_out/r02_step_cfg*_d*.log
python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r02b_cfg3_conn8.json 2>/dev/null
CUDA_DEVICE_MAX_CONNECTIONS=32 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r02b_cfg3_conn32.json 2>/dev/null
CUDA_DEVICE_MAX_CONNECTIONS=32 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r02b_cfg4_conn32.json 2>/dev/null
python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e --graph off > gpurun_out/r02b_cfg3_nograph.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b_*.json')):
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1]); print(f, l["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
ncu --set full --clock-control none --import-source on -k regex:statevec -c 1 -o gpurun_out/r02_sv_q10_lc2b -f python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern --reps 1 > gpurun_out/r02_ncu_sv3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:statevec -c 1 -o gpurun_out/r02_sv_q8_lc2b -f python tools/profile_step.py --reps 1 > gpurun_out/r02_ncu_sv4.log 2>&1
