set -x
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02f_bench_n1.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02f_bench_n1.json').read().strip().splitlines()[-1])
print("%.2f ms"%l["ms_per_step"], "value %.3e e2e %.3e"%(l["value"], l["e2e"]["value"]), l["gpu_launches"], l["clocks"])
for k,v in l.get("other_workloads",{}).items(): print(k, v["ms_per_step"], v["phases_ms_one_agent"], v["roofline"]["kernel"], v["roofline"]["frac"], v["rooflines_frac"])
PY
which compute-sanitizer
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q -k "features_empty or ragged or honoured or lu_solve or second_device or walks or shared_prefix_simulation_matches_per_set_kernels[kyriienko-10 or medium_size" > gpurun_out/r02_sanitizer.log 2>&1; echo "sanitizer rc=$?"
tail -15 gpurun_out/r02_sanitizer.log
