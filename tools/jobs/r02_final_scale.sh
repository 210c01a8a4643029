set -x
for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02h_bench_n$N.json 2> gpurun_out/r02h_bench_n$N.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02h_bench_n[248].json')):
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "%.2f ms"%l["ms_per_step"], "value %.3e e2e %.3e ratio %.3f"%(l["value"], l["e2e"]["value"], l["e2e"]["value"]/l["value"]), l["final_z_head"], {k:round(v["ms_per_step"],2) for k,v in l.get("other_workloads",{}).items()}, round(l["step_level"]["frac"],3))
    except Exception as e: print(f, "ERR", e)
PY
