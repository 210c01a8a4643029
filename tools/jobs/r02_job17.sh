set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t17.log
tail -3 gpurun_out/r02_t17.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err
python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_n1_cfg5.json 2> gpurun_out/r02e_bench_n1_cfg5.err
python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 10 > gpurun_out/r02e_bench_n1_cfg3.json 2> gpurun_out/r02e_bench_n1_cfg3.err
python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 10 > gpurun_out/r02e_bench_n1_cfg1.json 2> gpurun_out/r02e_bench_n1_cfg1.err
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02e_bench_reference.json 2> gpurun_out/r02e_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02e_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --graph off > gpurun_out/r02e_ncu_bench.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02e_*.json')):
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "%.2f ms"%l["ms_per_step"], "value %.3e e2e %.3e"%(l["value"], l["e2e"]["value"]), l.get("phases_ms_one_agent"), (l.get("roofline") or {}).get("frac"), (l.get("step_level") or {}).get("frac"))
    except Exception as e: print(f, "ERR", e)
PY
