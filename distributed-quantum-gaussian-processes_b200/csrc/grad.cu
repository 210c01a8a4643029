// Fused central-difference gradient of the GP negative log marginal likelihood.
//   grad_i = 1/2 * sum_jk (A^-1 - alpha alpha^T)_jk * (K(p + h e_i) - K(p - h e_i))_kj / (2h)
// Replaces agent_riemannian.py:270-275 (the (P,n,n) dK tensor, 64 GB/agent at config 5) and :431-436 (the
// Python loop of n^2 elementwise products): no shifted Gram is ever written.  A CTA owns one 64x64 tile
// (lower tiles only; off-diagonal tiles weigh 2 by symmetry), keeps its tile of B = A^-1 - alpha alpha^T in
// registers, and loops over the 2P shifted feature sets staged through shared memory, evaluating the outer
// kernel and contracting on the fly.  Partial sums go to partial[tile][i]; a second kernel adds them in a
// fixed order, so the result is bit-reproducible run to run (the reference rounds it to 4 decimals).
// Bound: FP64 pipe (3m + outer-kernel flops per entry); HBM traffic is one read of A^-1.
#include "pairwise.cuh"

namespace dqgp {

__device__ __forceinline__ void tile_from_index(int local, int& bi, int& bj) {
    bi = int((sqrt(8.0 * local + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= local) ++bi;
    while (bi * (bi + 1) / 2 > local) --bi;
    bj = local - bi * (bi + 1) / 2;
}

__device__ __forceinline__ void load_bracket(const double* __restrict__ Ainv, int ld, const double* __restrict__ alpha, int n,
                                             int row0, int col0, int ty, int tx, double weight, double (&b)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = row0 + ty + 16 * i;
        const double ar = (r < n) ? alpha[r] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = col0 + 2 * tx + 32 * (j >> 1) + (j & 1);
            b[i][j] = (r < n && c < n) ? weight * (Ainv[(size_t)r * ld + c] - ar * alpha[c]) : 0.0;
        }
    }
}

__device__ __forceinline__ void block_commit(double acc, double* s_red, double* dst) {
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < PW_THREADS / 32; ++w) t += s_red[w];
        *dst = t;
    }
}

template <int OUTER>
__global__ void __launch_bounds__(PW_THREADS) grad_projected_kernel(const double* __restrict__ Ainv, int ld,
                                                                    const double* __restrict__ alpha,
                                                                    const double* __restrict__ F, int n, int m, int P,
                                                                    OuterHyp hyp, double* __restrict__ partial) {
    __shared__ __align__(16) double FrT[PW_MAX_M * PW_PITCH];
    __shared__ __align__(16) double FcT[PW_MAX_M * PW_PITCH];
    __shared__ double s_red[PW_THREADS / 32];
    int bi, bj;
    tile_from_index(blockIdx.x, bi, bj);
    const int row0 = bi * PW_TILE, col0 = bj * PW_TILE;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double b[4][4];
    load_bracket(Ainv, ld, alpha, n, row0, col0, ty, tx, bi == bj ? 1.0 : 2.0, b);
    const size_t set_stride = (size_t)n * m;
    for (int i = 0; i < P; ++i) {
        double acc = 0.0;
#pragma unroll 1
        for (int sg = 0; sg < 2; ++sg) {
            const double* Fs = F + (size_t)(1 + 2 * i + sg) * set_stride;
            __syncthreads();
            stage_features_T(FrT, Fs, row0, n, m);
            stage_features_T(FcT, Fs, col0, n, m);
            __syncthreads();
            double d2[4][4];
            micro_sqdist(FrT, FcT, m, ty, tx, d2);
            double part = 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) part = fma(b[r][c], outer_eval<OUTER>(d2[r][c], hyp), part);
            acc += sg ? -part : part;
        }
        block_commit(acc, s_red, partial + (size_t)blockIdx.x * P + i);
    }
}

constexpr int FID_KC_G = 16;
constexpr int FID_PITCH_G = 65;

__global__ void __launch_bounds__(PW_THREADS) grad_fidelity_kernel(const double* __restrict__ Ainv, int ld,
                                                                   const double* __restrict__ alpha,
                                                                   const double2* __restrict__ Psi, int n, int dim, int P,
                                                                   double* __restrict__ partial) {
    __shared__ __align__(16) double2 ArT[FID_KC_G * FID_PITCH_G];
    __shared__ __align__(16) double2 AcT[FID_KC_G * FID_PITCH_G];
    __shared__ double s_red[PW_THREADS / 32];
    int bi, bj;
    tile_from_index(blockIdx.x, bi, bj);
    const int row0 = bi * PW_TILE, col0 = bj * PW_TILE;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double b[4][4];
    load_bracket(Ainv, ld, alpha, n, row0, col0, ty, tx, bi == bj ? 1.0 : 2.0, b);
    const size_t set_stride = (size_t)n * dim;
    const int vr = min(PW_TILE, n - row0), vc = min(PW_TILE, n - col0);
    for (int i = 0; i < P; ++i) {
        double acc = 0.0;
#pragma unroll 1
        for (int sg = 0; sg < 2; ++sg) {
            const double2* Ps = Psi + (size_t)(1 + 2 * i + sg) * set_stride;
            double re[4][4], im[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) re[r][c] = im[r][c] = 0.0;
            for (int k0 = 0; k0 < dim; k0 += FID_KC_G) {
                __syncthreads();
                for (int e = threadIdx.x; e < PW_TILE * FID_KC_G; e += PW_THREADS) {
                    const int r = e / FID_KC_G, kk = e % FID_KC_G;
                    const bool kin = (k0 + kk) < dim;
                    ArT[kk * FID_PITCH_G + r] = (r < vr && kin) ? Ps[(size_t)(row0 + r) * dim + k0 + kk] : make_double2(0.0, 0.0);
                    AcT[kk * FID_PITCH_G + r] = (r < vc && kin) ? Ps[(size_t)(col0 + r) * dim + k0 + kk] : make_double2(0.0, 0.0);
                }
                __syncthreads();
#pragma unroll 4
                for (int kk = 0; kk < FID_KC_G; ++kk) {
                    double2 a[4], bb[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) a[r] = ArT[kk * FID_PITCH_G + ty + 16 * r];
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        bb[2 * jj] = AcT[kk * FID_PITCH_G + 2 * tx + 32 * jj];
                        bb[2 * jj + 1] = AcT[kk * FID_PITCH_G + 2 * tx + 32 * jj + 1];
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            re[r][c] = fma(a[r].x, bb[c].x, re[r][c]);
                            re[r][c] = fma(a[r].y, bb[c].y, re[r][c]);
                            im[r][c] = fma(a[r].y, bb[c].x, im[r][c]);
                            im[r][c] = fma(-a[r].x, bb[c].y, im[r][c]);
                        }
                }
            }
            double part = 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) part = fma(b[r][c], re[r][c] * re[r][c] + im[r][c] * im[r][c], part);
            acc += sg ? -part : part;
        }
        block_commit(acc, s_red, partial + (size_t)blockIdx.x * P + i);
    }
}

// grad[i] = scale * sum_tiles partial[tile][i], fixed summation order (strided per thread, then a fixed tree)
__global__ void __launch_bounds__(256) grad_reduce_kernel(const double* __restrict__ partial, int n_tiles, int P, double scale,
                                                          double* __restrict__ grad) {
    __shared__ double s[256];
    const int i = blockIdx.x;
    double acc = 0.0;
    for (int t = threadIdx.x; t < n_tiles; t += 256) acc += partial[(size_t)t * P + i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) grad[i] = scale * s[0];
}

static inline int grad_tiles(int n) {
    const int t = (n + PW_TILE - 1) / PW_TILE;
    return t * (t + 1) / 2;
}

}  // namespace dqgp

extern "C" {

size_t dqgp_grad_workspace_bytes(int n, int P) {
    if (n <= 0 || P <= 0) return 0;
    return sizeof(double) * (size_t)dqgp::grad_tiles(n) * (size_t)P;
}

int dqgp_grad_projected(int outer, const double* h_hyp, const double* d_Ainv, int ld, const double* d_alpha, const double* d_F,
                        int n, int m, int P, double h, double* d_grad, void* d_work, void* stream) {
    using namespace dqgp;
    DQGP_REQUIRE(d_Ainv && d_alpha && d_F && d_grad && d_work, "dqgp_grad_projected: NULL argument");
    DQGP_REQUIRE(n >= 1 && ld >= n && P >= 1 && m >= 1 && m <= PW_MAX_M && h != 0.0, "dqgp_grad_projected: bad shape / shift");
    OuterHyp hyp;
    if (make_outer_hyp(outer, h_hyp, &hyp)) return -1;
    const int tiles = grad_tiles(n);
    cudaStream_t st = as_stream(stream);
    double* partial = static_cast<double*>(d_work);
    switch (outer) {
        case DQGP_OUTER_GAUSSIAN: grad_projected_kernel<DQGP_OUTER_GAUSSIAN><<<tiles, PW_THREADS, 0, st>>>(d_Ainv, ld, d_alpha, d_F, n, m, P, hyp, partial); break;
        case DQGP_OUTER_MATERN15: grad_projected_kernel<DQGP_OUTER_MATERN15><<<tiles, PW_THREADS, 0, st>>>(d_Ainv, ld, d_alpha, d_F, n, m, P, hyp, partial); break;
        default: grad_projected_kernel<DQGP_OUTER_EXPSINE2><<<tiles, PW_THREADS, 0, st>>>(d_Ainv, ld, d_alpha, d_F, n, m, P, hyp, partial); break;
    }
    DQGP_LAUNCH_CHECK("grad_projected_kernel");
    grad_reduce_kernel<<<P, 256, 0, st>>>(partial, tiles, P, 0.5 / (2.0 * h), d_grad);
    DQGP_LAUNCH_CHECK("grad_reduce_kernel");
    return 0;
}

int dqgp_grad_fidelity(const double* d_Ainv, int ld, const double* d_alpha, const double* d_Psi, int n, int dim, int P, double h,
                       double* d_grad, void* d_work, void* stream) {
    using namespace dqgp;
    DQGP_REQUIRE(d_Ainv && d_alpha && d_Psi && d_grad && d_work, "dqgp_grad_fidelity: NULL argument");
    DQGP_REQUIRE(n >= 1 && ld >= n && P >= 1 && dim >= 1 && h != 0.0, "dqgp_grad_fidelity: bad shape / shift");
    const int tiles = grad_tiles(n);
    cudaStream_t st = as_stream(stream);
    double* partial = static_cast<double*>(d_work);
    grad_fidelity_kernel<<<tiles, PW_THREADS, 0, st>>>(d_Ainv, ld, d_alpha, reinterpret_cast<const double2*>(d_Psi), n, dim, P, partial);
    DQGP_LAUNCH_CHECK("grad_fidelity_kernel");
    grad_reduce_kernel<<<P, 256, 0, st>>>(partial, tiles, P, 0.5 / (2.0 * h), d_grad);
    DQGP_LAUNCH_CHECK("grad_reduce_kernel");
    return 0;
}
}
