// Grouped fp64 tile GEMM on the DMMA tensor path (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4), the one
// contraction engine behind the blocked Cholesky, triangular inverse and inverse product (chol.cu).
#pragma once
#include "common.cuh"

namespace dqgp {

constexpr int GM_BM = 128, GM_BN = 128, GM_KC = 16, GM_STAGES = 3, GM_THREADS = 256;
constexpr int GM_PITCH_K = GM_KC + 4;    // operand stored [row][k]  (k contiguous in HBM)
constexpr int GM_PITCH_M = GM_BM + 4;    // operand stored [k][row]  (row contiguous in HBM)
constexpr int GM_OPERAND_DOUBLES = (GM_BM * GM_PITCH_K > GM_KC * GM_PITCH_M) ? GM_BM * GM_PITCH_K : GM_KC * GM_PITCH_M;
constexpr size_t GM_SMEM_BYTES = size_t(GM_STAGES) * 2 * GM_OPERAND_DOUBLES * sizeof(double);

enum { GM_KRULE_ALL = 0, GM_KRULE_A_LOWER = 1, GM_KRULE_B_LOWER = 2, GM_KRULE_LAUUM = 3 };

// C(MxN) = alpha * A(MxK) * B(KxN) + beta * C.  M, N multiples of 128; K multiple of 16.
struct GemmTask {
    const double* A;   // a_rowmajor_k ? A[m*lda + k] : A[k*lda + m]
    const double* B;   // b_rowmajor_k ? B[n*ldb + k] : B[k*ldb + n]
    double* C;         // C[m*ldc + n]
    int M, N, K;
    int lda, ldb, ldc;
    int a_k_contig, b_k_contig;
    int lower_tiles;   // only tiles with m0 >= n0 (square outputs)
    int krule;         // restrict the contraction range per tile when an operand is lower-triangular
    double alpha, beta;
    int tile_begin;    // first global tile id of this task within its launch group
    int tiles;         // number of tiles of this task
};

int launch_gemm_group(const GemmTask* d_tasks, int n_tasks, int total_tiles, cudaStream_t st);
int gemm_init();   // raises the dynamic shared memory limit once per process/device

static inline int gemm_task_tiles(const GemmTask& t) {
    const int tm = t.M / GM_BM, tn = t.N / GM_BN;
    return t.lower_tiles ? tm * (tm + 1) / 2 : tm * tn;
}

}  // namespace dqgp
