// Batched statevector simulation of the encoding circuits + Pauli-XYZ features (sm_100a).
//
// Replaces what squlearn does behind q_kernel.evaluate for every sample and parameter set
// (reference call site agent_riemannian.py:118; S = 2P+1 sets built at :245-256).
//
// Mapping: one warp owns one state (2^q complex128 amplitudes in shared memory, never in HBM); for
// q <= 5 a warp is split into 32/2^(q-1) lane groups, one state per group, so no lane idles.  Each lane
// owns 2^(q-1)/group amplitude pairs per gate; gates are applied in place with only __syncwarp between
// them (no block barrier).  All sin/cos of a state's gate angles are computed once, cooperatively, into
// a per-state table (a gate's angle is shared by all 2^(q-1) pairs), and arccos(x) once per sample.
// The epilogue reduces <X_k>,<Y_k>,<Z_k> with group-local shuffles, or streams the state to HBM for the
// fidelity kernel.  Work is a warp-granular grid-stride loop over S*n states on a grid sized to the SMs.
#include "common.cuh"

namespace dqgp {

__device__ __forceinline__ int insert_zero_bit(int k, int t) {
    const int lo = k & ((1 << t) - 1);
    return ((k >> t) << (t + 1)) | lo;
}

template <int Q>
struct SvGeom {
    static constexpr int DIM = 1 << Q;
    static constexpr int PAIRS = (Q == 0) ? 1 : (DIM / 2);
    static constexpr int GROUP = PAIRS >= 32 ? 32 : (PAIRS < 1 ? 1 : PAIRS);  // lanes per state
    static constexpr int SPW = 32 / GROUP;                                     // states per warp
    static constexpr int PPL = PAIRS / GROUP;                                  // pairs per lane
};

template <int Q, bool WANT_STATES>
__global__ void __launch_bounds__(128) statevec_kernel(const dqgp_gate* __restrict__ gates, int n_gates, int d, int P,
                                                        int uses_acos, const double* __restrict__ X, int n,
                                                        const double* __restrict__ Pm, int S, double* __restrict__ out) {
    using G = SvGeom<Q>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // block layout: [gate program][per warp: SPW x (DIM double2 | n_gates double2 trig | d double acos)]
    dqgp_gate* s_gates = reinterpret_cast<dqgp_gate*>(smem_raw);
    const int warps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t gate_bytes = (sizeof(dqgp_gate) * n_gates + 15) & ~size_t(15);
    const size_t state_bytes = sizeof(double2) * G::DIM + sizeof(double2) * n_gates + ((sizeof(double) * d + 15) & ~size_t(15));
    for (int i = threadIdx.x; i < n_gates; i += blockDim.x) s_gates[i] = gates[i];
    __syncthreads();

    const int sub = lane / G::GROUP;   // which state of this warp
    const int lig = lane % G::GROUP;   // lane in group
    unsigned char* my = smem_raw + gate_bytes + state_bytes * (size_t(warp) * G::SPW + sub);
    double2* amp = reinterpret_cast<double2*>(my);
    double2* trig = amp + G::DIM;
    double* acx = reinterpret_cast<double*>(trig + n_gates);

    const long long total = (long long)S * n;
    const long long n_groups = (total + G::SPW - 1) / G::SPW;
    const int m = 3 * Q;
    for (long long grp = (long long)blockIdx.x * warps + warp; grp < n_groups; grp += (long long)gridDim.x * warps) {
        long long st = grp * G::SPW + sub;
        const bool live = st < total;
        if (!live) st = total - 1;
        const int s = int(st / n), j = int(st % n);
        const double* x = X + (size_t)j * d;
        const double* p = Pm + (size_t)s * P;

        if (uses_acos) {
            for (int f = lig; f < d; f += G::GROUP) acx[f] = acos(x[f]);
            __syncwarp();
        }
        for (int g = lig; g < n_gates; g += G::GROUP) {
            const dqgp_gate gt = s_gates[g];
            if (gt.form == DQGP_A_NONE) continue;   // H, CX carry no angle
            double ang = 0.0;
            switch (gt.form) {
                case DQGP_A_P: ang = p[gt.pidx]; break;
                case DQGP_A_X: ang = x[gt.fidx]; break;
                case DQGP_A_P_PLUS_CX: ang = p[gt.pidx] + gt.coef * x[gt.fidx]; break;
                case DQGP_A_P_TIMES_ACOS: ang = p[gt.pidx] * acx[gt.fidx]; break;
                case DQGP_A_C_TIMES_ACOS: ang = gt.coef * acx[gt.fidx]; break;
                default: break;
            }
            double sn, cs;
            sincos(0.5 * ang, &sn, &cs);
            trig[g] = make_double2(cs, sn);
        }
        for (int i = lig; i < G::DIM; i += G::GROUP) amp[i] = make_double2(i == 0 ? 1.0 : 0.0, 0.0);
        __syncwarp();

        for (int g = 0; g < n_gates; ++g) {
            const dqgp_gate gt = s_gates[g];
            const double2 cs = trig[g];
            const double c = cs.x, sn = cs.y;
            const int t = (gt.kind >= DQGP_G_CX) ? gt.q1 : gt.q0;
            const int bit = 1 << t;
            const int cbit = (gt.kind >= DQGP_G_CX) ? (1 << gt.q0) : 0;
#pragma unroll
            for (int r = 0; r < G::PPL; ++r) {
                const int k = lig + G::GROUP * r;
                const int i0 = insert_zero_bit(k, t), i1 = i0 | bit;
                if (cbit && !(i0 & cbit)) continue;
                const double2 a = amp[i0], b = amp[i1];
                double2 na, nb;
                switch (gt.kind) {
                    case DQGP_G_H: {
                        const double r2 = 0.70710678118654752440;
                        na = make_double2((a.x + b.x) * r2, (a.y + b.y) * r2);
                        nb = make_double2((a.x - b.x) * r2, (a.y - b.y) * r2);
                        break;
                    }
                    case DQGP_G_RX:
                        na = make_double2(c * a.x + sn * b.y, c * a.y - sn * b.x);
                        nb = make_double2(sn * a.y + c * b.x, c * b.y - sn * a.x);
                        break;
                    case DQGP_G_RY:
                        na = make_double2(c * a.x - sn * b.x, c * a.y - sn * b.y);
                        nb = make_double2(sn * a.x + c * b.x, sn * a.y + c * b.y);
                        break;
                    case DQGP_G_RZ:
                    case DQGP_G_CRZ:
                        na = make_double2(c * a.x + sn * a.y, c * a.y - sn * a.x);
                        nb = make_double2(c * b.x - sn * b.y, c * b.y + sn * b.x);
                        break;
                    default:  // CX
                        na = b;
                        nb = a;
                        break;
                }
                amp[i0] = na;
                amp[i1] = nb;
            }
            __syncwarp();
        }

        if (WANT_STATES) {
            if (live) {
                double2* dst = reinterpret_cast<double2*>(out) + (size_t)st * G::DIM;
                for (int i = lig; i < G::DIM; i += G::GROUP) dst[i] = amp[i];
            }
        } else {
            double* dst = out + (size_t)st * m;
#pragma unroll 1
            for (int k = 0; k < Q; ++k) {
                double fx = 0.0, fy = 0.0, fz = 0.0;
#pragma unroll
                for (int r = 0; r < G::PPL; ++r) {
                    const int kk = lig + G::GROUP * r;
                    const int i0 = insert_zero_bit(kk, k), i1 = i0 | (1 << k);
                    const double2 a = amp[i0], b = amp[i1];
                    fx += a.x * b.x + a.y * b.y;
                    fy += a.x * b.y - a.y * b.x;
                    fz += (a.x * a.x + a.y * a.y) - (b.x * b.x + b.y * b.y);
                }
#pragma unroll
                for (int o = G::GROUP / 2; o > 0; o >>= 1) {
                    fx += __shfl_xor_sync(0xffffffffu, fx, o);
                    fy += __shfl_xor_sync(0xffffffffu, fy, o);
                    fz += __shfl_xor_sync(0xffffffffu, fz, o);
                }
                if (live && lig == 0) {
                    dst[k] = 2.0 * fx;
                    dst[Q + k] = 2.0 * fy;
                    dst[2 * Q + k] = fz;
                }
            }
        }
        __syncwarp();
    }
}

template <int Q, bool WANT_STATES>
static int launch_sv(const dqgp_circuit* c, const double* X, int n, const double* Pm, int S, double* out, cudaStream_t st) {
    using G = SvGeom<Q>;
    const int n_gates = (int)c->gates.size();
    const size_t gate_bytes = (sizeof(dqgp_gate) * n_gates + 15) & ~size_t(15);
    const size_t state_bytes = sizeof(double2) * G::DIM + sizeof(double2) * n_gates + ((sizeof(double) * c->d + 15) & ~size_t(15));
    int warps = 4;
    while (warps > 1 && gate_bytes + state_bytes * G::SPW * warps > 100 * 1024) warps >>= 1;
    const size_t smem = gate_bytes + state_bytes * G::SPW * warps;
    DQGP_REQUIRE(smem <= 227 * 1024, "statevector kernel needs %zu bytes of shared memory (q=%d, %d gates)", smem, Q, n_gates);
    auto kern = statevec_kernel<Q, WANT_STATES>;
    DQGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    DQGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const long long total = (long long)S * n;
    const long long n_groups = (total + G::SPW - 1) / G::SPW;
    long long blocks = (n_groups + warps - 1) / warps;
    const long long cap = (long long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) return 0;
    kern<<<(unsigned)blocks, warps * 32, smem, st>>>(c->d_gates, n_gates, c->d, c->P, c->uses_acos ? 1 : 0, X, n, Pm, S, out);
    DQGP_LAUNCH_CHECK("statevec_kernel");
    return 0;
}

template <bool WANT_STATES>
static int dispatch_sv(const dqgp_circuit* c, const double* X, int n, const double* Pm, int S, double* out, void* stream) {
    DQGP_REQUIRE(n >= 0 && S >= 0, "statevector: negative size");
    if (n == 0 || S == 0) return 0;      // empty input: nothing to do (pointers may be NULL)
    DQGP_REQUIRE(c && X && Pm && out, "statevector: NULL argument");
    int rc = circuit_on_device(c);
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    switch (c->q) {
#define DQGP_SV_CASE(QQ) case QQ: return launch_sv<QQ, WANT_STATES>(c, X, n, Pm, S, out, st);
        DQGP_SV_CASE(1) DQGP_SV_CASE(2) DQGP_SV_CASE(3) DQGP_SV_CASE(4) DQGP_SV_CASE(5) DQGP_SV_CASE(6)
        DQGP_SV_CASE(7) DQGP_SV_CASE(8) DQGP_SV_CASE(9) DQGP_SV_CASE(10) DQGP_SV_CASE(11) DQGP_SV_CASE(12)
#undef DQGP_SV_CASE
    }
    set_error("statevector: unsupported qubit count %d", c->q);
    return -1;
}

}  // namespace dqgp

extern "C" {
int dqgp_features(const dqgp_circuit* c, const double* d_X, int n, const double* d_Pm, int S, double* d_F, void* stream) {
    return dqgp::dispatch_sv<false>(c, d_X, n, d_Pm, S, d_F, stream);
}
int dqgp_states(const dqgp_circuit* c, const double* d_X, int n, const double* d_Pm, int S, double* d_Psi, void* stream) {
    return dqgp::dispatch_sv<true>(c, d_X, n, d_Pm, S, d_Psi, stream);
}
}
