for ob in 1 2 4; do
  python bench.py --workload cfg3 --steps 30 --warmup 5 --also "" --no-cpu-baseline --skip-e2e --outer-blocks $ob 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 N=1 ob=$ob', d['ms_per_step'], d.get('phases_ms_one_agent'))"
done
for ob in 1 2 4; do
  python bench.py --workload cfg2 --steps 50 --warmup 5 --also "" --no-cpu-baseline --skip-e2e --outer-blocks $ob 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 N=1 ob=$ob', d['ms_per_step'], d.get('phases_ms_one_agent'))"
done
