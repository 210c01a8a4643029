python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "potrf or lean or lookahead" > gpurun_out/r02_t42.log 2>&1; tail -2 gpurun_out/r02_t42.log
OBS=1,2 python tools/potrf_latency.py 1024 2048 8192 2>&1 | tail -6
