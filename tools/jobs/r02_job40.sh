nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_latency tools/dmma_latency.cu && /tmp/dmma_latency
