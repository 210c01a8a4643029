// Blocked fp64 Cholesky + explicit SPD inverse + solve on one GPU (sm_100a).
// Replaces the LAPACK sequence of RiemannianAgent.train_and_update (reference agent_riemannian.py:410-418:
// cholesky, 4x general solve incl. an explicit inverse against eye(n); :442 slogdet) with
//   potrf   right-looking, nb = 128: [leaf: factor + invert the diagonal block] -> [panel = panel * inv(Lkk)^T]
//           -> [trailing -= panel panel^T]           (trailing update on DMMA, n^3/3 flops)
//   trtri   W = L^-1 by recursive halving: W21 = -W22 (L21 W11), all products of one level in ONE grouped
//           launch (2 launches per level, log2(n/128) levels, n^3/3 flops on DMMA)
//   lauum   A^-1 = W^T W, lower tiles only, contraction range clipped to the triangular support (n^3/3)
//   solve   alpha = W^T (W y);  logdet = 2 sum log Lii (accumulated by the leaves)
// The matrix is padded to a multiple of 128 with an identity block, so no kernel has edge cases.
#include <cmath>
#include "gemm64.cuh"

namespace dqgp {
constexpr int NB = 128;
}  // namespace dqgp

struct dqgp_solver {
    int n, np, ld, nblk, device;
    double *A, *W, *T;            // factor (in place), L^-1, scratch / A^-1
    double *y_pad, *w, *partial;  // padded rhs, W y, column partial sums
    double *strip, *V;            // prediction: padded 128-row strip of K(test,train), V = W strip^T
    dqgp::GemmTask* d_tasks;
    // launch groups: [first task, task count, tiles]
    struct Group { int first, count, tiles; };
    std::vector<Group> trsm, inner;                   // per 128-column step: panel solve, update inside the outer panel
    std::vector<Group> syrk_next, syrk_rest;          // per outer panel (512 columns): next panel's columns / the bulk
    int ob;                                           // outer panel width in 128-blocks (1 = single-level)
    cudaStream_t helper;                              // HIGH-priority stream carrying the critical path (leaf, panel solve, next column)
    cudaEvent_t ev_fork, ev_join;
    std::vector<cudaEvent_t> ev_trsm, ev_rest;
    std::vector<Group> tri_t, tri_w;       // per trtri level
    Group lauum, quad;
    size_t bytes;
};

namespace dqgp {

// ---- leaf: Cholesky of a 128x128 diagonal block and its triangular inverse, one CTA, 256 threads -----------
// One shared 128x130 array M holds both results: the strict upper triangle keeps L transposed
// (M[k][i] = L[i][k], k < i) and the strict lower triangle receives W = L^-1 (M[k][j] = W[k][j], j < k);
// the diagonals live in s_diag / s_rdiag.
// Phase 1 is a left-looking Cholesky in panels of 8 columns.  Row i is owned by the thread pair (i, i+128): both
// run the update loop over half of the k range each (per k: 1 conflict-free LDS + 4 broadcast LDS.128 for 8 DFMA),
// the helper hands its partial sums over through shared memory, and the primary thread factors the 8x8 diagonal
// block redundantly in registers (no cross-lane chain) and forward-substitutes its row.
// Phase 2 inverts L column by column in panels of 8 rows; column j is owned by two ADJACENT lanes that split the k
// range and combine with one shuffle, so warps never wait for each other (no block barrier in the whole phase).
// Measured with clock64 (round 1): v3 (128 threads) spent 32K cycles in the phase-1 k-loops, 50K in the per-panel
// serial part, 51K in the phase-2 k-loops of the slowest thread, 19K in its serial part, 9K storing W = 164K cycles.
constexpr int LEAF_THREADS = 256;
constexpr int LP = 130;
constexpr size_t LEAF_SMEM_V2 = sizeof(double) * (NB * LP + 2 * NB + 8 * NB);

__global__ void __launch_bounds__(LEAF_THREADS, 1) potrf_leaf_kernel(double* __restrict__ A, int ld, double* __restrict__ W,
                                                                      int blk, double* logdet, int* info, int n_real) {
    extern __shared__ __align__(16) double leaf_smem[];
    double* M = leaf_smem;
    double* s_diag = M + NB * LP;
    double* s_rdiag = s_diag + NB;
    double* s_part = s_rdiag + NB;            // [8][128] partial sums of the helper threads
    __shared__ double s_red[LEAF_THREADS / 32];
    __shared__ double s_blk[64];
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = tid & (NB - 1);             // row owned in phase 1
    const bool helper = tid >= NB;
    double* Ablk = A + (size_t)blk * NB * ld + (size_t)blk * NB;
    double* Wblk = W + (size_t)blk * NB * ld + (size_t)blk * NB;
    if (tid == 0) s_bad = 0;
    __syncthreads();

    // ---------------- phase 1: Cholesky ----------------
    double a[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = 0.0;
    if (!helper) {
        const double2* src = reinterpret_cast<const double2*>(Ablk + (size_t)i * ld);
#pragma unroll
        for (int c = 0; c < 4; ++c) { const double2 v = src[c]; a[2 * c] = v.x; a[2 * c + 1] = v.y; }
    }
    for (int j0 = 0; j0 < NB; j0 += 8) {
        double nxt[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) nxt[c] = 0.0;
        if (!helper && j0 + 8 < NB && i >= j0 + 8) {   // prefetch the next panel's entries of this row
            const double2* src = reinterpret_cast<const double2*>(Ablk + (size_t)i * ld + j0 + 8);
#pragma unroll
            for (int c = 0; c < 4; ++c) { const double2 v = src[c]; nxt[2 * c] = v.x; nxt[2 * c + 1] = v.y; }
        }
        if (i >= j0) {
            const int kh = (j0 >> 1) & ~3;
            const int kb = helper ? kh : 0, ke = helper ? j0 : kh;
#pragma unroll 4
            for (int k = kb; k < ke; ++k) {
                const double lik = M[k * LP + i];
                const double2* row = reinterpret_cast<const double2*>(&M[k * LP + j0]);
                const double2 v0 = row[0], v1 = row[1], v2 = row[2], v3 = row[3];
                a[0] = fma(-lik, v0.x, a[0]); a[1] = fma(-lik, v0.y, a[1]);
                a[2] = fma(-lik, v1.x, a[2]); a[3] = fma(-lik, v1.y, a[3]);
                a[4] = fma(-lik, v2.x, a[4]); a[5] = fma(-lik, v2.y, a[5]);
                a[6] = fma(-lik, v3.x, a[6]); a[7] = fma(-lik, v3.y, a[7]);
            }
            if (helper) {
#pragma unroll
                for (int c = 0; c < 8; ++c) s_part[c * NB + i] = a[c];
            }
        }
        __syncthreads();
        if (!helper && i >= j0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) a[c] += s_part[c * NB + i];
            if (i < j0 + 8) {
#pragma unroll
                for (int c = 0; c < 8; ++c) s_blk[(i - j0) * 8 + c] = a[c];
            }
        }
        __syncthreads();
        if (!helper && i >= j0) {
            // every primary thread factors the 8x8 diagonal block redundantly in registers, then substitutes its row
            double l[8][8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) l[r][c] = s_blk[r * 8 + c];
            double rd[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const double piv = l[c][c];
                if (!(piv > 0.0) && i == j0 && s_bad == 0) s_bad = j0 + c + 1;
                rd[c] = rsqrt(piv);
                l[c][c] = piv * rd[c];
#pragma unroll
                for (int r = c + 1; r < 8; ++r) l[r][c] *= rd[c];
#pragma unroll
                for (int c2 = c + 1; c2 < 8; ++c2)
#pragma unroll
                    for (int r = c2; r < 8; ++r) l[r][c2] = fma(-l[r][c], l[c2][c], l[r][c2]);
            }
            if (i < j0 + 8) {
                const int r0 = i - j0;
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (r == r0) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) a[c] = (c <= r) ? l[r][c] : 0.0;
                        s_diag[i] = l[r][r];
                        s_rdiag[i] = rd[r];
                    }
            } else {
                // column-oriented substitution: as soon as x_k is known it is removed from all later columns
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const double x = a[c] * rd[c];
                    a[c] = x;
#pragma unroll
                    for (int c2 = c + 1; c2 < 8; ++c2) a[c2] = fma(-x, l[c2][c], a[c2]);
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (i > j0 + c) M[(j0 + c) * LP + i] = a[c];
            double2* dst = reinterpret_cast<double2*>(Ablk + (size_t)i * ld + j0);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                dst[c] = make_double2((i >= j0 + 2 * c) ? a[2 * c] : 0.0, (i >= j0 + 2 * c + 1) ? a[2 * c + 1] : 0.0);
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 8; ++c) a[c] = nxt[c];     // helpers restart from zero
    }
    {
        double v = (tid < NB) ? log(s_diag[tid]) : 0.0;
        v = warp_sum(v);
        if (lane == 0) s_red[warp] = v;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < LEAF_THREADS / 32; ++w) tot += s_red[w];
            *logdet += 2.0 * tot;
            if (s_bad && *info == 0 && blk * NB + s_bad <= n_real) *info = blk * NB + s_bad;
        }
    }

    // ---------------- phase 2: W = L^-1; column j is owned by lanes (2*(j%16), 2*(j%16)+1) of warp j/16 ----------------
    const int j = warp * 16 + (lane >> 1);
    const int h = lane & 1;
    for (int i0 = 0; i0 < NB; i0 += 8) {
        if (warp * 16 > i0 + 7) continue;         // whole warp is right of this row panel (warp-uniform)
        double acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = 0.0;
        if (j < i0) {
            // k = j .. i0-1 (W[j][j] = 1/L[j][j], W[k][j] = M[k][j] for k > j); the two lanes take alternate k
            for (int k = j + h; k < i0; k += 2) {
                const double w = (k == j) ? s_rdiag[j] : M[k * LP + j];
                const double2* row = reinterpret_cast<const double2*>(&M[k * LP + i0]);
                const double2 v0 = row[0], v1 = row[1], v2 = row[2], v3 = row[3];
                acc[0] = fma(w, v0.x, acc[0]); acc[1] = fma(w, v0.y, acc[1]);
                acc[2] = fma(w, v1.x, acc[2]); acc[3] = fma(w, v1.y, acc[3]);
                acc[4] = fma(w, v2.x, acc[4]); acc[5] = fma(w, v2.y, acc[5]);
                acc[6] = fma(w, v3.x, acc[6]); acc[7] = fma(w, v3.y, acc[7]);
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 1);
        if (h == 0 && j <= i0 + 7) {
            double wv[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int row = i0 + r;
                double w = 0.0;
                if (row == j) w = s_rdiag[j];
                else if (row > j) {
                    double sres = acc[r];
#pragma unroll
                    for (int r2 = 0; r2 < r; ++r2)
                        if (i0 + r2 >= j) sres = fma(M[(i0 + r2) * LP + row], wv[r2], sres);
                    w = -sres * s_rdiag[row];
                    M[row * LP + j] = w;
                }
                wv[r] = w;
            }
        }
        __syncwarp();
    }
    __syncthreads();
    for (int e = tid; e < NB * NB / 2; e += LEAF_THREADS) {
        const int r = e >> 6, c = (e & 63) * 2;
        double2 v;
        v.x = (c < r) ? M[r * LP + c] : (c == r ? s_rdiag[r] : 0.0);
        v.y = (c + 1 < r) ? M[r * LP + c + 1] : (c + 1 == r ? s_rdiag[r] : 0.0);
        *reinterpret_cast<double2*>(Wblk + (size_t)r * ld + c) = v;
    }
}

// ---- padding: rows/cols >= n become the identity ---------------------------------------------------------
__global__ void pad_identity_kernel(double* A, int n, int np, int ld) {
    const int r = blockIdx.x;   // np rows
    for (int c = threadIdx.x; c < np; c += blockDim.x) {
        if (r >= n || c >= n) A[(size_t)r * ld + c] = (r == c) ? 1.0 : 0.0;
    }
}
__global__ void pad_vector_kernel(const double* y, int n, int np, double* out, double* logdet, int* info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) out[i] = (y != nullptr && i < n) ? y[i] : 0.0;
    if (i == 0) { *logdet = 0.0; *info = 0; }
}

// ---- copy the strictly-lower 128x128 blocks of src into dst (panels of L: T -> A) ---------------------------
__global__ void copy_lower_blocks_kernel(const double* __restrict__ src, double* __restrict__ dst, int ld) {
    const int rb = blockIdx.y + 1, cb = blockIdx.x;
    if (cb >= rb) return;
    const double2* s2 = reinterpret_cast<const double2*>(src + (size_t)rb * NB * ld + (size_t)cb * NB);
    double2* d2 = reinterpret_cast<double2*>(dst + (size_t)rb * NB * ld + (size_t)cb * NB);
    for (int e = threadIdx.x; e < NB * NB / 2; e += blockDim.x) {
        const int r = e >> 6, c = e & 63;
        d2[(size_t)r * (ld / 2) + c] = s2[(size_t)r * (ld / 2) + c];
    }
}

// ---- w = W y (lower-triangular, one warp per row) --------------------------------------------------------
__global__ void trmv_lower_kernel(const double* __restrict__ W, int np, int ld, const double* __restrict__ y, double* __restrict__ w) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= np) return;
    double acc = 0.0;
    for (int k = lane; k <= row; k += 32) acc = fma(W[(size_t)row * ld + k], y[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) w[row] = acc;
}
// ---- alpha = W^T w: per (row block, column block) partial sums, then a fixed-order reduction --------------
__global__ void trmv_lower_T_partial_kernel(const double* __restrict__ W, int ld, const double* __restrict__ w, double* __restrict__ partial, int np) {
    const int rb = blockIdx.y, cb = blockIdx.x;
    const int col = cb * NB + threadIdx.x;
    double acc = 0.0;
    if (rb >= cb) {
        const int r0 = rb * NB;
        for (int r = 0; r < NB; ++r) acc = fma(W[(size_t)(r0 + r) * ld + col], w[r0 + r], acc);   // zero above the diagonal
    }
    partial[(size_t)rb * np + col] = acc;
}
__global__ void reduce_partial_kernel(const double* __restrict__ partial, int nblk, int np, int n, double* __restrict__ out) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= n) return;
    double acc = 0.0;
    for (int rb = 0; rb < nblk; ++rb) acc += partial[(size_t)rb * np + col];
    out[col] = acc;
}
// ---- mirror the lower triangle of a padded square into the upper one (tiled transpose through smem) --------
__global__ void symmetrize_lower_kernel(double* A, int ld) {
    __shared__ double tile[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) tile[r][tx] = A[(size_t)(bi * 32 + r) * ld + bj * 32 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int gr = bj * 32 + r, gc = bi * 32 + tx;
        if (bi != bj || gc > gr) A[(size_t)gr * ld + gc] = tile[tx][r];
    }
}
__global__ void colsumsq_kernel(const double* __restrict__ V, int rows, int ld, int ncols, double* __restrict__ out) {
    // out[c] = sum_r V[r][c]^2, fixed order; one thread per column, coalesced across threads
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    double acc = 0.0;
    for (int r = 0; r < rows; ++r) { const double v = V[(size_t)r * ld + c]; acc = fma(v, v, acc); }
    out[c] = acc;
}
__global__ void copy_pad_rows_kernel(const double* __restrict__ B, int nb, int n, int ldb, double* __restrict__ out, int nbp, int np, int ldo) {
    const int r = blockIdx.x;
    for (int c = threadIdx.x; c < np; c += blockDim.x) out[(size_t)r * ldo + c] = (r < nb && c < n) ? B[(size_t)r * ldb + c] : 0.0;
}

static GemmTask make_task(const double* A, const double* B, double* C, int M, int N, int K, int ld, int a_k, int b_k, int lower,
                          int krule, double alpha, double beta) {
    GemmTask t;
    t.A = A; t.B = B; t.C = C; t.M = M; t.N = N; t.K = K; t.lda = t.ldb = t.ldc = ld;
    t.a_k_contig = a_k; t.b_k_contig = b_k; t.lower_tiles = lower; t.krule = krule; t.alpha = alpha; t.beta = beta;
    t.tile_begin = 0; t.tiles = gemm_task_tiles(t);
    return t;
}

}  // namespace dqgp

extern "C" {

int dqgp_solver_create(int n, dqgp_solver** out) { return dqgp_solver_create_ex(n, 0, out); }

int dqgp_solver_create_ex(int n, int outer_blocks, dqgp_solver** out) {
    using namespace dqgp;
    DQGP_REQUIRE(out != nullptr, "dqgp_solver_create: out is NULL");
    *out = nullptr;
    DQGP_REQUIRE(n >= 1 && n <= (1 << 17), "dqgp_solver_create: n = %d outside [1, 131072]", n);
    dqgp_solver* s = new dqgp_solver();
    s->n = n;
    s->ob = outer_blocks > 0 ? (outer_blocks > 16 ? 16 : outer_blocks) : 4;
    s->nblk = (n + NB - 1) / NB;
    s->np = s->nblk * NB;
    s->ld = s->np;
    s->A = s->W = s->T = s->y_pad = s->w = s->partial = s->strip = s->V = nullptr;
    s->d_tasks = nullptr;
    s->helper = nullptr; s->ev_fork = nullptr; s->ev_join = nullptr;
    cudaError_t e = cudaGetDevice(&s->device);
    const size_t mat = sizeof(double) * (size_t)s->np * s->ld;
    s->bytes = 3 * mat;
    if (e == cudaSuccess) e = cudaMalloc(&s->A, mat);
    if (e == cudaSuccess) e = cudaMalloc(&s->W, mat);
    if (e == cudaSuccess) e = cudaMalloc(&s->T, mat);
    if (e == cudaSuccess) e = cudaMalloc(&s->y_pad, sizeof(double) * s->np);
    if (e == cudaSuccess) e = cudaMalloc(&s->w, sizeof(double) * s->np);
    if (e == cudaSuccess) e = cudaMalloc(&s->partial, sizeof(double) * (size_t)s->nblk * s->np);
    if (e == cudaSuccess) e = cudaMalloc(&s->strip, sizeof(double) * (size_t)NB * s->ld);
    if (e == cudaSuccess) e = cudaMalloc(&s->V, sizeof(double) * (size_t)s->np * NB);
    if (e != cudaSuccess) { dqgp_solver_destroy(s); return cuda_fail(e, "dqgp_solver_create (allocation; this library has no CPU fallback)"); }

    std::vector<GemmTask> tasks;
    auto push_group = [&](std::vector<GemmTask>& grp) {
        dqgp_solver::Group g;
        g.first = (int)tasks.size(); g.count = (int)grp.size(); g.tiles = 0;
        for (auto& t : grp) { t.tile_begin = g.tiles; g.tiles += t.tiles; tasks.push_back(t); }
        grp.clear();
        return g;
    };
    const int ld = s->ld, np = s->np, nblk = s->nblk;
    auto at = [&](double* base, int rb, int cb) { return base + (size_t)rb * NB * ld + (size_t)cb * NB; };
    std::vector<GemmTask> grp;
    // potrf: two-level blocking.  Outer panels of OB = 4 block columns (512); inside a panel the 128-wide steps
    // update only the panel's own columns (K = 128, little work); the trailing matrix gets ONE rank-512 update per
    // outer panel (K = 512: the GEMM runs near its long-K efficiency instead of the 60% of K = 128 updates).
    const int OB = s->ob;
    for (int k = 0; k + 1 < nblk; ++k) {
        const int rest = np - (k + 1) * NB;
        // panel solve out of place (A -> T): two 64-column tiles share the same input rows, so in place would race;
        // the strictly-lower blocks of L are copied back T -> A once, after the last step
        grp.push_back(make_task(at(s->A, k + 1, k), at(s->W, k, k), at(s->T, k + 1, k), rest, NB, NB, ld, 1, 1, 0, GM_KRULE_ALL, 1.0, 0.0));
        s->trsm.push_back(push_group(grp));
        const int pend = std::min(((k / OB) + 1) * OB, nblk);     // first block column after this outer panel
        for (int c = k + 1; c < pend; ++c)                         // block column c of the panel, rows c..end
            grp.push_back(make_task(at(s->T, c, k), at(s->T, c, k), at(s->A, c, c), np - c * NB, NB, NB, ld, 1, 1, 0, GM_KRULE_ALL, -1.0, 1.0));
        s->inner.push_back(push_group(grp));
    }
    for (int p0 = 0; p0 < nblk; p0 += OB) {
        const int w = std::min(OB, nblk - p0);                    // panel width in blocks
        const int c0 = p0 + w;                                    // first trailing block column
        const int wn = std::min(OB, nblk - c0);                   // width of the next panel
        if (wn > 0)   // columns of the next outer panel, all rows below: on the critical path
            grp.push_back(make_task(at(s->T, c0, p0), at(s->T, c0, p0), at(s->A, c0, c0), np - c0 * NB, wn * NB, w * NB, ld, 1, 1, 0, GM_KRULE_ALL, -1.0, 1.0));
        s->syrk_next.push_back(push_group(grp));
        const int c1 = c0 + wn;
        if (wn > 0 && c1 < nblk)   // everything further right: the bulk, overlapped with the next panel's factorisation
            grp.push_back(make_task(at(s->T, c1, p0), at(s->T, c1, p0), at(s->A, c1, c1), np - c1 * NB, np - c1 * NB, w * NB, ld, 1, 1, 1, GM_KRULE_ALL, -1.0, 1.0));
        s->syrk_rest.push_back(push_group(grp));
    }
    // trtri levels: spans of `span` blocks are already inverted; join neighbours pairwise
    for (int span = 1; span < nblk; span *= 2) {
        std::vector<GemmTask> gt, gw;
        for (int o = 0; o + span < nblk; o += 2 * span) {
            const int h1 = span, h2 = std::min(span, nblk - (o + span));
            // T21 = L21 * W11   (W11 lower-triangular as the K x N operand)
            gt.push_back(make_task(at(s->A, o + h1, o), at(s->W, o, o), at(s->T, o + h1, o), h2 * NB, h1 * NB, h1 * NB, ld, 1, 0, 0, GM_KRULE_B_LOWER, 1.0, 0.0));
            // W21 = -W22 * T21  (W22 lower-triangular as the M x K operand)
            gw.push_back(make_task(at(s->W, o + h1, o + h1), at(s->T, o + h1, o), at(s->W, o + h1, o), h2 * NB, h1 * NB, h2 * NB, ld, 1, 0, 0, GM_KRULE_A_LOWER, -1.0, 0.0));
        }
        s->tri_t.push_back(push_group(gt));
        s->tri_w.push_back(push_group(gw));
    }
    // lauum: Ainv = W^T W  (both operands row-contiguous: A[k][m] = W[k][m])
    grp.push_back(make_task(s->W, s->W, s->T, np, np, np, ld, 0, 0, 1, GM_KRULE_LAUUM, 1.0, 0.0));
    s->lauum = push_group(grp);
    // prediction: V[m][c] = sum_k W[m][k] strip[c][k]  (W lower-triangular, both operands k-contiguous)
    {
        GemmTask t = make_task(s->W, s->strip, s->V, np, NB, np, ld, 1, 1, 0, GM_KRULE_A_LOWER, 1.0, 0.0);
        t.ldc = NB;
        grp.push_back(t);
        s->quad = push_group(grp);
    }

    e = cudaMalloc(&s->d_tasks, sizeof(GemmTask) * tasks.size());
    if (e == cudaSuccess) e = cudaMemcpy(s->d_tasks, tasks.data(), sizeof(GemmTask) * tasks.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        e = cudaStreamCreateWithPriority(&s->helper, cudaStreamNonBlocking, hi);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming);
    for (int k = 0; k < (nblk + s->ob - 1) / s->ob && e == cudaSuccess; ++k) {
        cudaEvent_t a, b;
        e = cudaEventCreateWithFlags(&a, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b, cudaEventDisableTiming);
        if (e == cudaSuccess) { s->ev_trsm.push_back(a); s->ev_rest.push_back(b); }
    }
    if (e == cudaSuccess) e = cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LEAF_SMEM_V2);
    if (e != cudaSuccess) { dqgp_solver_destroy(s); return cuda_fail(e, "dqgp_solver_create (task table)"); }
    int rc = gemm_init();
    if (rc) { dqgp_solver_destroy(s); return rc; }
    *out = s;
    return 0;
}

void dqgp_solver_destroy(dqgp_solver* s) {
    if (!s) return;
    for (auto ev : s->ev_trsm) cudaEventDestroy(ev);
    for (auto ev : s->ev_rest) cudaEventDestroy(ev);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    if (s->helper) cudaStreamDestroy(s->helper);
    cudaFree(s->A); cudaFree(s->W); cudaFree(s->T); cudaFree(s->y_pad); cudaFree(s->w); cudaFree(s->partial); cudaFree(s->strip); cudaFree(s->V); cudaFree(s->d_tasks);
    delete s;
}
int dqgp_solver_n(const dqgp_solver* s) { return s ? s->n : -1; }
int dqgp_solver_ld(const dqgp_solver* s) { return s ? s->ld : -1; }
double* dqgp_solver_matrix(dqgp_solver* s) { return s ? s->A : nullptr; }
double* dqgp_solver_inverse(dqgp_solver* s) { return s ? s->T : nullptr; }
double* dqgp_solver_factor(dqgp_solver* s) { return s ? s->A : nullptr; }
size_t dqgp_solver_bytes(const dqgp_solver* s) { return s ? s->bytes : 0; }

int dqgp_potrf_solve_inv(dqgp_solver* s, const double* d_y, double* d_alpha, double* d_logdet, int* d_info, int want_inverse,
                         void* stream) {
    using namespace dqgp;
    const bool factor_only = want_inverse < 0;
    DQGP_REQUIRE(s && d_logdet && d_info && (factor_only || (d_y && d_alpha)), "dqgp_potrf_solve_inv: NULL argument");
    cudaStream_t st = as_stream(stream);
    const int np = s->np, ld = s->ld, nblk = s->nblk;
    pad_vector_kernel<<<(np + 255) / 256, 256, 0, st>>>(factor_only ? nullptr : d_y, s->n, np, s->y_pad, d_logdet, d_info);
    if (np != s->n) pad_identity_kernel<<<np, 256, 0, st>>>(s->A, s->n, np, ld);
    DQGP_LAUNCH_CHECK("pad kernels");
    // Two-level right-looking Cholesky with look-ahead.  The critical path (leaves, panel solves, updates inside the
    // current 512-column outer panel, and the rank-512 update of the NEXT panel's columns) runs on the solver's
    // HIGH-priority stream; the bulk rank-512 update of everything further right stays on the caller's stream and
    // overlaps the next panel's factorisation.  Priority matters: a leaf CTA needs 133 KB of shared
    // memory and only fits beside ONE resident GEMM CTA, so it must win the slot a retiring GEMM CTA frees.
    cudaStream_t crit = s->helper;
    DQGP_CUDA(cudaEventRecord(s->ev_fork, st));
    DQGP_CUDA(cudaStreamWaitEvent(crit, s->ev_fork, 0));
    int last_rest = -1;
    const int OB = s->ob;
    for (int k = 0; k < nblk; ++k) {
        potrf_leaf_kernel<<<1, LEAF_THREADS, LEAF_SMEM_V2, crit>>>(s->A, ld, s->W, k, d_logdet, d_info, s->n);
        DQGP_LAUNCH_CHECK("potrf_leaf_kernel");
        if (k + 1 >= nblk) break;
        int rc = launch_gemm_group(s->d_tasks + s->trsm[k].first, s->trsm[k].count, s->trsm[k].tiles, crit);
        if (rc) return rc;
        rc = launch_gemm_group(s->d_tasks + s->inner[k].first, s->inner[k].count, s->inner[k].tiles, crit);
        if (rc) return rc;
        if ((k + 1) % OB == 0) {                       // an outer panel is complete: rank-(OB*128) trailing update
            const int p = k / OB;
            const bool has_rest = s->syrk_rest[p].tiles > 0;
            if (has_rest) DQGP_CUDA(cudaEventRecord(s->ev_trsm[p], crit));
            if (last_rest >= 0) DQGP_CUDA(cudaStreamWaitEvent(crit, s->ev_rest[last_rest], 0));   // earlier bulk updates hit these columns
            rc = launch_gemm_group(s->d_tasks + s->syrk_next[p].first, s->syrk_next[p].count, s->syrk_next[p].tiles, crit);
            if (rc) return rc;
            if (has_rest) {
                DQGP_CUDA(cudaStreamWaitEvent(st, s->ev_trsm[p], 0));
                rc = launch_gemm_group(s->d_tasks + s->syrk_rest[p].first, s->syrk_rest[p].count, s->syrk_rest[p].tiles, st);
                if (rc) return rc;
                DQGP_CUDA(cudaEventRecord(s->ev_rest[p], st));
                last_rest = p;
            }
        }
    }
    DQGP_CUDA(cudaEventRecord(s->ev_join, crit));
    DQGP_CUDA(cudaStreamWaitEvent(st, s->ev_join, 0));
    if (nblk > 1) {
        copy_lower_blocks_kernel<<<dim3(nblk - 1, nblk - 1), 256, 0, st>>>(s->T, s->A, ld);
        DQGP_LAUNCH_CHECK("copy_lower_blocks_kernel");
    }
    if (factor_only) return 0;      // L is in place (dqgp_solver_factor); its diagonal-block inverses are in W
    for (size_t l = 0; l < s->tri_t.size(); ++l) {
        int rc = launch_gemm_group(s->d_tasks + s->tri_t[l].first, s->tri_t[l].count, s->tri_t[l].tiles, st);
        if (rc) return rc;
        rc = launch_gemm_group(s->d_tasks + s->tri_w[l].first, s->tri_w[l].count, s->tri_w[l].tiles, st);
        if (rc) return rc;
    }
    // alpha = W^T (W y)
    trmv_lower_kernel<<<(np + 7) / 8, 256, 0, st>>>(s->W, np, ld, s->y_pad, s->w);
    trmv_lower_T_partial_kernel<<<dim3(nblk, nblk), NB, 0, st>>>(s->W, ld, s->w, s->partial, np);
    reduce_partial_kernel<<<(s->n + 255) / 256, 256, 0, st>>>(s->partial, nblk, np, s->n, d_alpha);
    DQGP_LAUNCH_CHECK("solve kernels");
    if (want_inverse) {
        int rc = launch_gemm_group(s->d_tasks + s->lauum.first, s->lauum.count, s->lauum.tiles, st);
        if (rc) return rc;
        if (want_inverse > 1) {
            symmetrize_lower_kernel<<<dim3(np / 32, np / 32), 256, 0, st>>>(s->T, ld);
            DQGP_LAUNCH_CHECK("symmetrize_lower_kernel");
        }
    }
    return 0;
}

int dqgp_solver_apply_factor(dqgp_solver* s, const double* d_x, double* d_y, void* stream) {
    // y = L x with the Cholesky factor held by the solver (sampling from N(0, K): main.py:272-274)
    using namespace dqgp;
    DQGP_REQUIRE(s && d_x && d_y, "dqgp_solver_apply_factor: NULL argument");
    cudaStream_t st = as_stream(stream);
    pad_vector_kernel<<<(s->np + 255) / 256, 256, 0, st>>>(d_x, s->n, s->np, s->y_pad, s->w, reinterpret_cast<int*>(s->w + 1));
    trmv_lower_kernel<<<(s->np + 7) / 8, 256, 0, st>>>(s->A, s->np, s->ld, s->y_pad, s->partial);
    DQGP_LAUNCH_CHECK("apply_factor kernels");
    DQGP_CUDA(cudaMemcpyAsync(d_y, s->partial, sizeof(double) * s->n, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int dqgp_solver_quadform_rows(dqgp_solver* s, const double* d_B, int nb, int ldb, double* d_out, void* stream) {
    // d_out[i] = || L^-1 b_i ||^2 for every row b_i of B (nb, n): V = W B^T in strips of 128 rows of B on the
    // DMMA GEMM, then fixed-order column sums of squares (main.py:1462-1463).
    using namespace dqgp;
    DQGP_REQUIRE(s && d_B && d_out && nb >= 0 && ldb >= s->n, "dqgp_solver_quadform_rows: bad arguments");
    cudaStream_t st = as_stream(stream);
    for (int r0 = 0; r0 < nb; r0 += NB) {
        const int cnt = std::min(NB, nb - r0);
        copy_pad_rows_kernel<<<NB, 256, 0, st>>>(d_B + (size_t)r0 * ldb, cnt, s->n, ldb, s->strip, NB, s->np, s->ld);
        int rc = launch_gemm_group(s->d_tasks + s->quad.first, s->quad.count, s->quad.tiles, st);
        if (rc) return rc;
        colsumsq_kernel<<<1, NB, 0, st>>>(s->V, s->n, NB, cnt, d_out + r0);
        DQGP_LAUNCH_CHECK("quadform kernels");
    }
    return 0;
}

}  // extern "C"
