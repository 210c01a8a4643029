// Grouped fp64 GEMM, 128x128 tile per CTA, 8 warps (2x4) each owning 64x32 = 8x4 DMMA 8x8 blocks.
// Operands stream HBM/L2 -> shared memory through a 3-stage cp.async (LDGSTS) ring, 16-byte copies;
// shared-memory pitches (20 / 132 doubles) make every DMMA fragment load bank-conflict-free.
// A launch covers a *group* of independent tasks (device-resident table) so the many small products of
// the recursive triangular inverse fill the machine in one launch.  Roofline: FP64 pipe (DMMA and DFMA
// share it on B200: 37.1 TFLOP/s measured, profiles/r01_fp64_peak.json).
#include "gemm64.cuh"

namespace dqgp {

__device__ __forceinline__ void gm_load_operand(double* sm, const double* __restrict__ g, int ld, int row0, int k0, int k_contig) {
    // 128 rows x 16 k doubles = 1024 16-byte chunks, 4 per thread
    if (k_contig) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = threadIdx.x + GM_THREADS * i;
            const int row = c >> 3, kc = (c & 7) * 2;
            cp_async16(sm + row * GM_PITCH_K + kc, g + (size_t)(row0 + row) * ld + k0 + kc);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = threadIdx.x + GM_THREADS * i;
            const int k = c >> 6, mc = (c & 63) * 2;
            cp_async16(sm + k * GM_PITCH_M + mc, g + (size_t)(k0 + k) * ld + row0 + mc);
        }
    }
}

__global__ void __launch_bounds__(GM_THREADS, 1) gemm_group_kernel(const GemmTask* __restrict__ tasks, int n_tasks) {
    extern __shared__ __align__(16) double gm_smem[];
    // locate the task that owns this tile (tables are short: <= a few hundred entries)
    int ti = 0;
    {
        int lo = 0, hi = n_tasks - 1;
        const int tile = blockIdx.x;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (tasks[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
        }
        ti = lo;
    }
    const GemmTask T = tasks[ti];
    const int local = blockIdx.x - T.tile_begin;
    int tm, tn;
    if (T.lower_tiles) {
        tm = int((sqrt(8.0 * local + 1.0) - 1.0) * 0.5);
        while ((tm + 1) * (tm + 2) / 2 <= local) ++tm;
        while (tm * (tm + 1) / 2 > local) --tm;
        tn = local - tm * (tm + 1) / 2;
    } else {
        const int tiles_n = T.N / GM_BN;
        tm = local / tiles_n;
        tn = local - tm * tiles_n;
    }
    const int m0 = tm * GM_BM, n0 = tn * GM_BN;
    int kb = 0, ke = T.K;
    if (T.krule == GM_KRULE_A_LOWER) ke = min(T.K, m0 + GM_BM);
    else if (T.krule == GM_KRULE_B_LOWER) kb = n0;
    else if (T.krule == GM_KRULE_LAUUM) kb = max(m0, n0);
    const int n_chunks = (ke - kb) / GM_KC;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp >> 2, wn = warp & 3;     // 2 x 4 warps
    const int g = lane >> 2, t = lane & 3;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto stage_a = [&](int s) { return gm_smem + (size_t)s * 2 * GM_OPERAND_DOUBLES; };
    auto stage_b = [&](int s) { return gm_smem + (size_t)s * 2 * GM_OPERAND_DOUBLES + GM_OPERAND_DOUBLES; };

#pragma unroll
    for (int s = 0; s < GM_STAGES - 1; ++s) {
        if (s < n_chunks) {
            gm_load_operand(stage_a(s), T.A, T.lda, m0, kb + s * GM_KC, T.a_k_contig);
            gm_load_operand(stage_b(s), T.B, T.ldb, n0, kb + s * GM_KC, T.b_k_contig);
        }
        cp_async_commit();
    }
    for (int ch = 0; ch < n_chunks; ++ch) {
        cp_async_wait<GM_STAGES - 2>();
        __syncthreads();
        {
            const int nx = ch + GM_STAGES - 1;
            if (nx < n_chunks) {
                const int s = nx % GM_STAGES;
                gm_load_operand(stage_a(s), T.A, T.lda, m0, kb + nx * GM_KC, T.a_k_contig);
                gm_load_operand(stage_b(s), T.B, T.ldb, n0, kb + nx * GM_KC, T.b_k_contig);
            }
            cp_async_commit();
        }
        const double* As = stage_a(ch % GM_STAGES);
        const double* Bs = stage_b(ch % GM_STAGES);
#pragma unroll
        for (int kk = 0; kk < GM_KC / 4; ++kk) {
            double a[8], b[4];
            if (T.a_k_contig) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = As[(wm * 64 + i * 8 + g) * GM_PITCH_K + kk * 4 + t];
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = As[(kk * 4 + t) * GM_PITCH_M + wm * 64 + i * 8 + g];
            }
            if (T.b_k_contig) {
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Bs[(wn * 32 + j * 8 + g) * GM_PITCH_K + kk * 4 + t];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Bs[(kk * 4 + t) * GM_PITCH_M + wn * 32 + j * 8 + g];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: each lane owns 2 adjacent doubles per 8x8 block -> 16-byte accesses
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + wm * 64 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + wn * 32 + j * 8 + 2 * t;
            double2* dst = reinterpret_cast<double2*>(T.C + (size_t)r * T.ldc + c);
            double2 v = make_double2(T.alpha * acc[i][j][0], T.alpha * acc[i][j][1]);
            if (T.beta != 0.0) {
                const double2 old = *dst;
                v.x = fma(T.beta, old.x, v.x);
                v.y = fma(T.beta, old.y, v.y);
            }
            *dst = v;
        }
    }
}

int gemm_init() {
    static bool done[64] = {false};
    int dev = 0;
    DQGP_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && done[dev]) return 0;
    DQGP_CUDA(cudaFuncSetAttribute(gemm_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GM_SMEM_BYTES));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return 0;
}

int launch_gemm_group(const GemmTask* d_tasks, int n_tasks, int total_tiles, cudaStream_t st) {
    if (n_tasks <= 0 || total_tiles <= 0) return 0;
    int rc = gemm_init();
    if (rc) return rc;
    gemm_group_kernel<<<total_tiles, GM_THREADS, GM_SMEM_BYTES, st>>>(d_tasks, n_tasks);
    DQGP_LAUNCH_CHECK("gemm_group_kernel");
    return 0;
}

}  // namespace dqgp
