set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t21.log
tail -5 gpurun_out/r02_t21.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python tools/profile_step.py | grep -E "statevector|^step"
timeout 120 python tools/profile_step.py --encoding kyriienko --q 10 --layers 4 --d 6 --outer-kernel matern | grep -E "statevector|^step"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:statevec -c 1 -o gpurun_out/r02_sv_q8_lc3 -f python tools/profile_step.py --reps 1 > gpurun_out/r02_ncu_sv7.log 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02g_bench_n1.json').read().strip().splitlines()[-1])
print("%.2f ms"%l["ms_per_step"], "value %.3e e2e %.3e"%(l["value"], l["e2e"]["value"]), l["phases_ms_one_agent"], l["step_level"]["frac"])
for k,v in l.get("other_workloads",{}).items(): print(k, v["ms_per_step"], v["phases_ms_one_agent"], v["roofline"]["kernel"], v["roofline"]["frac"])
PY
