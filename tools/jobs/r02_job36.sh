OBS=1,2,4 python tools/potrf_latency.py 1024 2048 3072 4096 6144 2>&1 | tee gpurun_out/r02_potrf_obs_small.txt
