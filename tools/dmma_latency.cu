// Dependent-issue latency of DMMA.8x8x4 / DFMA / LDS on one warp (and on 8 warps of one CTA): cycles per instruction for 1, 2, 4, 8
// independent accumulator chains.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_latency tools/dmma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void dmma_chain(double* out, long long* cycles, int iters) {
    double c[CHAINS][2];
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    const long long t1 = clock64();
    double s = 0.0;
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int CHAINS>
__global__ void dfma_chain(double* out, long long* cycles, int iters) {
    double c[CHAINS];
    for (int i = 0; i < CHAINS; ++i) c[i] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) c[i] = fma(c[i], a, b);
    }
    const long long t1 = clock64();
    double s = 0.0;
    for (int i = 0; i < CHAINS; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
__global__ void rsqrt_chain(double* out, long long* cycles, int iters) {
    double x = 1.5 + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) x = rsqrt(x) + 1.25;
    const long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Cold-start cost of the DMMA unit: a burst of 8 independent DMMAs after the warp spent `gap` cycles without issuing one
// (GAP_KIND 0: spinning on clock64; 1: a dependent DFMA chain, i.e. the FP64 pipe stays busy; 2: a dependent FFMA chain).
template <int GAP_KIND>
__global__ void dmma_after_gap(double* out, long long* cycles, int gap, int reps) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double f = a; float ff = (float)a;
    long long burst = 0;
    for (int r = 0; r < reps; ++r) {
        const long long g0 = clock64();
        if (GAP_KIND == 0) { while (clock64() - g0 < gap) { } }
        else if (GAP_KIND == 1) { while (clock64() - g0 < gap) { f = fma(f, b, a); } }
        else { while (clock64() - g0 < gap) { ff = fmaf(ff, 0.999f, 1.0f); } }
        __syncwarp();
        const long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += c[i][0];
        if (s == 1.2345e300) f += 1.0;         // consume: the clock read below waits for the last DMMA
        const long long t1 = clock64();
        if (r > 0) burst += t1 - t0;
    }
    out[threadIdx.x] = c[0][0] + c[7][1] + f + ff;
    if (threadIdx.x == 0) cycles[0] = burst / (reps - 1);
}

template <int GAP_KIND>
static void run_gap(const char* what) {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 32);
    for (int gap : {0, 100, 300, 1000, 2000, 5000, 20000}) {
        dmma_after_gap<GAP_KIND><<<1, 32>>>(out, cyc, gap, 50);
        long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("burst of 8 DMMA after a %5d-cycle gap (%s): %lld cycles\n", gap, what, h);
    }
    cudaFree(out); cudaFree(cyc);
}

template <typename K>
static void run(const char* name, K kern, int chains, int threads) {
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 1024); cudaMalloc(&cyc, sizeof(long long) * 4);
    const int iters = 2000;
    kern<<<1, threads>>>(out, cyc, iters);
    kern<<<1, threads>>>(out, cyc, iters);
    long long h = 0;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-6s chains=%d warps=%d: %.1f cycles per instruction per warp (%.1f per round)\n", name, chains, threads / 32, double(h) / iters / chains,
           double(h) / iters);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {32, 128, 256}) {
        run("DMMA", dmma_chain<1>, 1, threads); run("DMMA", dmma_chain<2>, 2, threads); run("DMMA", dmma_chain<4>, 4, threads);
        run("DMMA", dmma_chain<8>, 8, threads); run("DMMA", dmma_chain<16>, 16, threads);
    }
    for (int threads : {32, 128}) {
        run("DFMA", dfma_chain<1>, 1, threads); run("DFMA", dfma_chain<4>, 4, threads); run("DFMA", dfma_chain<16>, 16, threads);
    }
    {
        double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 32);
        rsqrt_chain<<<1, 32>>>(out, cyc, 2000); rsqrt_chain<<<1, 32>>>(out, cyc, 2000);
        long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("rsqrt(double)+add dependent chain: %.1f cycles per step\n", double(h) / 2000);
    }
    run_gap<0>("idle warp");
    run_gap<1>("DFMA chain in the gap");
    run_gap<2>("FFMA chain in the gap");
    return 0;
}
