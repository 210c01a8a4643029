// FP64 pipe microbenchmark for B200 (sm_100a): measures the achievable DFMA (vector) and
// DMMA (mma.sync f64 tensor) rates, whether they overlap, and fp64 exp() throughput.
// These are the roofline denominators for the fp64 kernels of this repo (MEASURED_PEAKS.json
// only carries HBM and bf16). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int NCH>
__global__ void k_dfma(double* out, int iters, double x) {
    double acc[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) acc[i] = threadIdx.x * 1e-3 + i;
    double m = x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) acc[i] = fma(acc[i], m, 1e-9);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NCH>
__global__ void k_dmma884(double* out, int iters, double x) {
    double c0[NCH], c1[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { c0[i] = 0; c1[i] = 0; }
    double a = x + threadIdx.x * 1e-6, b = x - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NCH>
__global__ void k_dmma16816(double* out, int iters, double x) {
    double c[NCH][4];
#pragma unroll
    for (int i = 0; i < NCH; i++) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0; }
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = x + i * 1e-3 + threadIdx.x * 1e-6;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = x - i * 1e-3 - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NCH>
__global__ void k_dmma1688(double* out, int iters, double x) {
    double c[NCH][4];
#pragma unroll
    for (int i = 0; i < NCH; i++) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0; }
    double a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = x + i * 1e-3 + threadIdx.x * 1e-6;
#pragma unroll
    for (int i = 0; i < 2; i++) b[i] = x - i * 1e-3 - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed in the same warp: NM DMMA884 + ND DFMA per inner iteration
template <int NM, int ND>
__global__ void k_mixed(double* out, int iters, double x) {
    double c0[NM], c1[NM], acc[ND];
#pragma unroll
    for (int i = 0; i < NM; i++) { c0[i] = 0; c1[i] = 0; }
#pragma unroll
    for (int i = 0; i < ND; i++) acc[i] = threadIdx.x * 1e-3 + i;
    double a = x + threadIdx.x * 1e-6, b = x - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NM; i++) dmma884(c0[i], c1[i], a, b);
#pragma unroll
        for (int i = 0; i < ND; i++) acc[i] = fma(acc[i], x, 1e-9);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; i++) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < ND; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed across warps: even warps DMMA, odd warps DFMA
__global__ void k_mixed_warps(double* out, int iters, double x) {
    int w = threadIdx.x >> 5;
    double s = 0;
    if (w & 1) {
        double acc[8];
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = threadIdx.x * 1e-3 + i;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], x, 1e-9);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) s += acc[i];
    } else {
        double c0[8], c1[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { c0[i] = 0; c1[i] = 0; }
        double a = x + threadIdx.x * 1e-6, b = x - threadIdx.x * 1e-6;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) dmma884(c0[i], c1[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) s += c0[i] + c1[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double* out, int iters, double x) {
    double acc[4];
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i] = -(threadIdx.x * 1e-2 + i) * x;
    double s = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) { s += exp(acc[i]); acc[i] -= 1e-7; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_sqrt(double* out, int iters, double x) {
    double acc[4];
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i] = (threadIdx.x * 1e-2 + i + 1) * x;
    double s = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) { s += sqrt(acc[i]); acc[i] += 1e-7; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_sincos(double* out, int iters, double x) {
    double acc[4];
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i] = (threadIdx.x * 1e-2 + i + 1) * x;
    double s = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) { double sn, cs; sincos(acc[i], &sn, &cs); s += sn * cs; acc[i] += 1e-7; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F launch, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    int threads = 256, blocks = sms * 8;
    double* out; CK(cudaMalloc(&out, sizeof(double) * threads * blocks));
    int iters = 20000;
    double x = 0.999999;
    double nthreads = (double)threads * blocks, nwarps = nthreads / 32;
    printf("{\"gpu\": \"%s\", \"sms\": %d,\n", prop.name, sms);
    {
        float ms = time_it([&] { k_dfma<8><<<blocks, threads>>>(out, iters, x); });
        printf(" \"dfma_tflops\": %.3f,\n", 2.0 * nthreads * 8 * iters / (ms * 1e-3) / 1e12);
    }
    {
        float ms = time_it([&] { k_dmma884<8><<<blocks, threads>>>(out, iters, x); });
        printf(" \"dmma_m8n8k4_tflops\": %.3f,\n", 2.0 * 256 * nwarps * 8 * iters / (ms * 1e-3) / 1e12);
    }
    {
        float ms = time_it([&] { k_dmma1688<4><<<blocks, threads>>>(out, iters, x); });
        printf(" \"dmma_m16n8k8_tflops\": %.3f,\n", 2.0 * 1024 * nwarps * 4 * iters / (ms * 1e-3) / 1e12);
    }
    {
        float ms = time_it([&] { k_dmma16816<4><<<blocks, threads>>>(out, iters, x); });
        printf(" \"dmma_m16n8k16_tflops\": %.3f,\n", 2.0 * 2048 * nwarps * 4 * iters / (ms * 1e-3) / 1e12);
    }
    {
        // same-warp mix: 4 DMMA884 (4*256 MAC/warp) + 16 DFMA (16*32 MAC/warp) per iteration
        float ms = time_it([&] { k_mixed<4, 16><<<blocks, threads>>>(out, iters, x); });
        double fl_m = 2.0 * 256 * nwarps * 4 * iters, fl_d = 2.0 * nthreads * 16 * iters;
        printf(" \"mixed_same_warp\": {\"ms\": %.3f, \"dmma_tflops\": %.3f, \"dfma_tflops\": %.3f, \"sum_tflops\": %.3f},\n", ms,
               fl_m / (ms * 1e-3) / 1e12, fl_d / (ms * 1e-3) / 1e12, (fl_m + fl_d) / (ms * 1e-3) / 1e12);
    }
    {
        float ms = time_it([&] { k_mixed_warps<<<blocks, threads>>>(out, iters, x); });
        double fl_m = 2.0 * 256 * (nwarps / 2) * 8 * iters, fl_d = 2.0 * (nthreads / 2) * 16 * iters;
        printf(" \"mixed_across_warps\": {\"ms\": %.3f, \"dmma_tflops\": %.3f, \"dfma_tflops\": %.3f, \"sum_tflops\": %.3f},\n", ms,
               fl_m / (ms * 1e-3) / 1e12, fl_d / (ms * 1e-3) / 1e12, (fl_m + fl_d) / (ms * 1e-3) / 1e12);
    }
    {
        int it2 = 2000;
        float ms = time_it([&] { k_exp<<<blocks, threads>>>(out, it2, x); });
        printf(" \"exp_gops\": %.3f,\n", nthreads * 4 * it2 / (ms * 1e-3) / 1e9);
        ms = time_it([&] { k_sqrt<<<blocks, threads>>>(out, it2, x); });
        printf(" \"sqrt_gops\": %.3f,\n", nthreads * 4 * it2 / (ms * 1e-3) / 1e9);
        ms = time_it([&] { k_sincos<<<blocks, threads>>>(out, it2, x); });
        printf(" \"sincos_gops\": %.3f,\n", nthreads * 4 * it2 / (ms * 1e-3) / 1e9);
    }
    printf(" \"clock_khz\": %d}\n", prop.clockRate);
    return 0;
}
