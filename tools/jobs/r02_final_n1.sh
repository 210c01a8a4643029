set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tfinal.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tfinal.log
tail -4 gpurun_out/r02_tfinal.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02h_bench_reference.json 2> gpurun_out/r02h_bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02h_bench_n1.json').read().strip().splitlines()[-1])
print("%.2f ms"%l["ms_per_step"], "value %.3e e2e %.3e"%(l["value"], l["e2e"]["value"]), l["phases_ms_one_agent"], l["step_level"]["frac"], l["roofline"]["frac"], l["gpu_launches"])
for k,v in l.get("other_workloads",{}).items(): print(k, round(v["ms_per_step"],3), {a:round(b,3) for a,b in v["phases_ms_one_agent"].items()}, v["roofline"]["kernel"], round(v["roofline"]["frac"],3))
r=json.loads(open('gpurun_out/r02h_bench_reference.json').read().strip().splitlines()[-1])
print("reference", r["value"], r["steps"], r.get("extrapolated"), r["ms_per_step"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r02h_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --graph off --also "" > gpurun_out/r02h_ncu_bench.log 2>&1
tail -2 gpurun_out/r02h_ncu_bench.log | cut -c1-300
