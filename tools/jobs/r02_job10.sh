set -x
python -m pytest tests -m gpu -x -q -k "agent or train or predict or trajectory or admm" > gpurun_out/r02_t10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t10.log
tail -5 gpurun_out/r02_t10.log
python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 10 > gpurun_out/r02d_bench_n1_cfg3.json 2> gpurun_out/r02d_bench_n1_cfg3.err
DQGP_AGENT_GRAPH=0 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 10 > gpurun_out/r02d_bench_n1_cfg3_nograph.json 2> /dev/null
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 4 > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02d_*.json')):
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "%.2f ms"%l["ms_per_step"], "value %.3e e2e %.3e ratio %.3f"%(l["value"], l["e2e"]["value"], l["e2e"]["value"]/l["value"]))
    except Exception as e: print(f, "ERR", e)
PY
