OBS=1 python tools/potrf_latency.py 1024 2>&1 | tail -12
