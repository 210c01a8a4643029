// Gate programs of the four in-scope encoding circuits, built on the host and kept resident on the
// device.  Restates squlearn 0.9.1's ChebyshevPQC / HubregtsenEncodingCircuit / YZ_CX_EncodingCircuit
// (ctor sites: reference main.py:68-83, agent_riemannian.py:51-66) with default options; the Kyriienko
// program is this project's own definition because the reference's call has no upstream behaviour
// (SURVEY Q13).  An independent restatement lives in oracle/circuits.py; tests compare the two.
#include <algorithm>
#include <mutex>
#include <string>
#include "common.cuh"

namespace dqgp {

static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return -2;
}
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!cached[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}

static int count_parameters(int enc, int q, int layers) {
    switch (enc) {
        case DQGP_CHEBYSHEV: return 2 * q + layers * q + layers * (q > 2 ? q : (q == 2 ? 1 : 0));
        case DQGP_HUBREGTSEN: return layers * q + (q > 2 ? layers * q : 0);
        case DQGP_YZ_CX: return 2 * q * layers;
        case DQGP_KYRIIENKO: return 3 * q * layers;
    }
    return -1;
}

static dqgp_gate mk(int kind, int q0, int q1 = -1, int form = DQGP_A_NONE, int pidx = -1, int fidx = -1, double coef = 1.0) {
    dqgp_gate g;
    g.kind = kind; g.q0 = q0; g.q1 = q1; g.form = form; g.pidx = pidx; g.fidx = fidx; g.coef = coef;
    return g;
}

static void build(dqgp_circuit& c) {
    const int q = c.q, d = c.d, L = c.layers, P = c.P;
    int next = 0;
    auto take = [&]() { int k = P > 0 ? next % P : 0; ++next; return k; };
    auto& G = c.gates;
    switch (c.encoding) {
        case DQGP_CHEBYSHEV: {
            // basis change, L x [RX(p*arccos x) ; CRZ ring: even pairs then odd pairs incl. closing pair], basis change
            for (int i = 0; i < q; ++i) G.push_back(mk(DQGP_G_RY, i, -1, DQGP_A_P, take()));
            // feature index: a running offset over all layers (squlearn 0.9.x, recalled); equals i % d when q % d == 0 or L == 1
            for (int l = 0; l < L; ++l) {
                for (int i = 0; i < q; ++i) G.push_back(mk(DQGP_G_RX, i, -1, DQGP_A_P_TIMES_ACOS, take(), (l * q + i) % d));
                if (q >= 2)   // range(0, q + closed - 1, 2) with closed = 1
                    for (int i = 0; i < q; i += 2) G.push_back(mk(DQGP_G_CRZ, i, (i + 1) % q, DQGP_A_P, take()));
                if (q > 2)
                    for (int i = 1; i < q; i += 2) G.push_back(mk(DQGP_G_CRZ, i, (i + 1) % q, DQGP_A_P, take()));
            }
            for (int i = 0; i < q; ++i) G.push_back(mk(DQGP_G_RY, i, -1, DQGP_A_P, take()));
            c.uses_acos = true;
            break;
        }
        case DQGP_HUBREGTSEN: {
            for (int i = 0; i < q; ++i) G.push_back(mk(DQGP_G_H, i));
            const int loops = (d + q - 1) / q;
            for (int l = 0; l < L; ++l) {
                for (int i = 0; i < loops * q; ++i)
                    G.push_back(mk(((i / q) % 2 == 0) ? DQGP_G_RZ : DQGP_G_RX, i % q, -1, DQGP_A_X, -1, i % d));
                for (int i = 0; i < q; ++i) G.push_back(mk(DQGP_G_RY, i, -1, DQGP_A_P, take()));
                if (q > 2)
                    for (int i = 0; i < q; ++i) G.push_back(mk(DQGP_G_CRZ, i, (i + 1) % q, DQGP_A_P, take()));
            }
            break;
        }
        case DQGP_YZ_CX: {
            for (int l = 0; l < L; ++l) {
                for (int i = 0; i < q; ++i) {
                    G.push_back(mk(DQGP_G_RY, i, -1, DQGP_A_P_PLUS_CX, take(), (l * q + i) % d, 1.0));   // running feature offset
                    G.push_back(mk(DQGP_G_RZ, i, -1, DQGP_A_P_PLUS_CX, take(), (l * q + i) % d, 1.0));
                }
                for (int i = (l % 2 == 0 ? 0 : 1); i < q - 1; i += 2) G.push_back(mk(DQGP_G_CX, i, i + 1));
            }
            break;
        }
        case DQGP_KYRIIENKO: {
            for (int l = 0; l < L; ++l) {
                for (int i = 0; i < q; ++i) G.push_back(mk(DQGP_G_RY, i, -1, DQGP_A_C_TIMES_ACOS, -1, i % d, 2.0 * (i + 1)));
                for (int i = 0; i < q; ++i) {
                    G.push_back(mk(DQGP_G_RZ, i, -1, DQGP_A_P, take()));
                    G.push_back(mk(DQGP_G_RX, i, -1, DQGP_A_P, take()));
                    G.push_back(mk(DQGP_G_RZ, i, -1, DQGP_A_P, take()));
                }
                for (int i = 0; i < q - 1; i += 2) G.push_back(mk(DQGP_G_CX, i, i + 1));
                for (int i = 1; i < q - 1; i += 2) G.push_back(mk(DQGP_G_CX, i, i + 1));
            }
            c.uses_acos = true;
            break;
        }
    }
}

// Greedy regrouping of the gate list into register-blocked passes.  A gate joins the current pass if its target
// is (or can still become) one of the <= 3 block qubits and none of its qubits is touched by a gate that was
// skipped earlier in program order (gates on disjoint qubits commute, so pulling it forward is exact).
static void build_plan(dqgp_circuit& c) {
    const int n = (int)c.gates.size();
    std::vector<char> done(n, 0);
    std::vector<std::vector<int>> fused;   // constituent gates of every SV_U2, in application order
    int remaining = n;
    const int bmax = c.q < 3 ? c.q : 3;
    while (remaining > 0) {
        SvPass pass;
        pass.nq = 0; pass.q[0] = pass.q[1] = pass.q[2] = -1;
        pass.op_begin = (int)c.ops.size();
        unsigned blocked = 0;                  // qubits of skipped gates
        std::vector<int> members;
        for (int g = 0; g < n; ++g) {
            if (done[g]) continue;
            const dqgp_gate& gt = c.gates[g];
            const bool two = gt.kind >= DQGP_G_CX;
            const int tgt = two ? gt.q1 : gt.q0;
            const unsigned qs = (1u << tgt) | (two ? (1u << gt.q0) : 0u);
            bool in_block = false;
            for (int k = 0; k < pass.nq; ++k) in_block |= (pass.q[k] == tgt);
            if ((qs & blocked) == 0 && (in_block || pass.nq < bmax)) {
                if (!in_block) pass.q[pass.nq++] = tgt;
                members.push_back(g);
                done[g] = 1;
                --remaining;
            } else {
                blocked |= qs;
            }
        }
        // sort block qubits ascending and express the ops in local bits, fusing runs of 1-qubit gates per qubit
        for (int a = 0; a < pass.nq; ++a)
            for (int b = a + 1; b < pass.nq; ++b)
                if (pass.q[b] < pass.q[a]) { int t = pass.q[a]; pass.q[a] = pass.q[b]; pass.q[b] = t; }
        auto local_of = [&](int qubit) { for (int k = 0; k < pass.nq; ++k) if (pass.q[k] == qubit) return k; return -1; };
        int open_mat[3] = {-1, -1, -1};          // per block qubit: U2 that can still absorb gates
        for (int g : members) {
            const dqgp_gate& gt = c.gates[g];
            const bool two = gt.kind >= DQGP_G_CX;
            if (!two) {
                const int lb = local_of(gt.q0);
                if (open_mat[lb] < 0) {
                    SvOp op;
                    op.kind = SV_U2; op.lbit = (int8_t)lb; op.cloc = -1; op.cq = -1; op.pad = 0;
                    op.idx = (int16_t)fused.size();
                    open_mat[lb] = (int)fused.size();
                    fused.emplace_back();
                    c.ops.push_back(op);
                }
                fused[open_mat[lb]].push_back(g);
            } else {
                SvOp op;
                op.kind = (int8_t)(gt.kind == DQGP_G_CX ? SV_CX : SV_CRZ);
                op.lbit = (int8_t)local_of(gt.q1);
                op.cloc = (int8_t)local_of(gt.q0);
                op.cq = (int8_t)(op.cloc < 0 ? gt.q0 : -1);
                op.idx = (int16_t)g;
                op.pad = 0;
                c.ops.push_back(op);
                open_mat[op.lbit] = -1;                       // later gates on the target must stay behind this op
                if (op.cloc >= 0) open_mat[op.cloc] = -1;     // ... and so must gates on the control
            }
        }
        pass.op_end = (int)c.ops.size();
        pass.lead_end = pass.op_begin;
        while (pass.lead_end < pass.op_end && c.ops[pass.lead_end].kind == SV_CX && c.ops[pass.lead_end].cq >= 0) ++pass.lead_end;
        pass.trail_begin = pass.op_end;
        while (pass.trail_begin > pass.lead_end && c.ops[pass.trail_begin - 1].kind == SV_CX && c.ops[pass.trail_begin - 1].cq >= 0) --pass.trail_begin;
        c.passes.push_back(pass);
    }
    for (auto& f : fused) {
        SvMat m;
        m.g_begin = (int)c.mat_gates.size();
        for (int g : f) c.mat_gates.push_back(g);
        m.g_end = (int)c.mat_gates.size();
        c.mats.push_back(m);
    }
    // where does each parameter enter?  (gate -> fused matrix / CRZ op -> pass)
    const int P = c.P;
    std::vector<int> uses(P, 0), gate_mat(n, -1), gate_pass(n, -1);
    c.par_gate.assign(P, -1);
    for (int g = 0; g < n; ++g)
        if (c.gates[g].pidx >= 0) { ++uses[c.gates[g].pidx]; c.par_gate[c.gates[g].pidx] = g; }
    c.shareable = P > 0;
    for (int i = 0; i < P; ++i) c.shareable = c.shareable && uses[i] == 1;
    for (int ip = 0; ip < (int)c.passes.size(); ++ip)
        for (int o = c.passes[ip].op_begin; o < c.passes[ip].op_end; ++o) {
            const SvOp& op = c.ops[o];
            if (op.kind == SV_U2) {
                for (int e = c.mats[op.idx].g_begin; e < c.mats[op.idx].g_end; ++e) { gate_mat[c.mat_gates[e]] = op.idx; gate_pass[c.mat_gates[e]] = ip; }
            } else {
                gate_pass[op.idx] = ip;
            }
        }
    c.par_mat.assign(P, -1);
    c.pass_par_begin.assign(c.passes.size() + 1, 0);
    if (c.shareable) {
        for (int i = 0; i < P; ++i) c.par_mat[i] = gate_mat[c.par_gate[i]];
        for (int ip = 0; ip < (int)c.passes.size(); ++ip) {
            c.pass_par_begin[ip] = (int)c.pass_params.size();
            for (int i = 0; i < P; ++i)
                if (gate_pass[c.par_gate[i]] == ip) c.pass_params.push_back(i);
        }
        c.pass_par_begin[c.passes.size()] = (int)c.pass_params.size();
    }
}

// CX-free plan (SvPass3): list scheduling over the circuit's events (fused 1-qubit runs, CX).  A CX whose predecessors are done is
// applied to the map at once; a pass takes the (up to three) earliest ready fused unitaries, on distinct qubits; no CX lies between the
// unitaries of a pass, so they all see one map.
static void build_plan3(dqgp_circuit& c) {
    c.has_plan3 = false;
    c.n_passes3 = 0;
    const int n = (int)c.gates.size(), q = c.q, P = c.P;
    if (!c.shareable || P <= 0) return;
    for (const auto& g : c.gates)
        if (g.kind == DQGP_G_CRZ) return;
    struct Event { bool cx; int a, b; int mat; };
    std::vector<Event> ev;
    std::vector<int> open_run(q, -1), gate_mat(n, -1);
    for (int g = 0; g < n; ++g) {
        const dqgp_gate& gt = c.gates[g];
        if (gt.kind == DQGP_G_CX) {
            open_run[gt.q0] = open_run[gt.q1] = -1;
            ev.push_back({true, gt.q0, gt.q1, -1});
        } else {
            if (open_run[gt.q0] < 0) {
                open_run[gt.q0] = (int)ev.size();
                ev.push_back({false, gt.q0, -1, (int)c.mats3.size()});
                SvMat m; m.g_begin = m.g_end = 0;
                c.mats3.push_back(m);
            }
            gate_mat[g] = ev[open_run[gt.q0]].mat;
        }
    }
    // constituent gates of every fused matrix, in application order
    for (size_t f = 0; f < c.mats3.size(); ++f) {
        c.mats3[f].g_begin = (int)c.mat_gates3.size();
        for (int g = 0; g < n; ++g)
            if (gate_mat[g] == (int)f) c.mat_gates3.push_back(g);
        c.mats3[f].g_end = (int)c.mat_gates3.size();
    }
    const int ne = (int)ev.size();
    std::vector<char> done(ne, 0);
    auto ready = [&](int e) {
        const unsigned qs = (1u << ev[e].a) | (ev[e].cx ? (1u << ev[e].b) : 0u);
        for (int f = 0; f < e; ++f)
            if (!done[f] && (qs & ((1u << ev[f].a) | (ev[f].cx ? (1u << ev[f].b) : 0u)))) return false;
        return true;
    };
    std::vector<int> mask(q);
    for (int k = 0; k < q; ++k) mask[k] = 1 << k;
    auto fill_rest = [&](SvPass3& ps, const std::vector<int>& blk) {
        int r = 0;
        for (int k = 0; k < q; ++k) {
            bool in = false;
            for (int b : blk) in |= (b == k);
            if (!in) ps.rest[r++] = mask[k];
        }
        for (; r < MAX_QUBITS; ++r) ps.rest[r] = 0;
        ps.nq = (int)blk.size();
        for (int l = 0; l < 3; ++l) ps.bm[l] = l < (int)blk.size() ? mask[blk[l]] : 0;
    };
    std::vector<int> mat_pass(c.mats3.size(), -1);
    int left = ne;
    while (left > 0) {
        for (bool progress = true; progress;) {
            progress = false;
            for (int e = 0; e < ne; ++e)
                if (!done[e] && ev[e].cx && ready(e)) { mask[ev[e].a] ^= mask[ev[e].b]; done[e] = 1; --left; progress = true; }
        }
        std::vector<int> blk, members;
        for (int e = 0; e < ne && (int)blk.size() < 3; ++e)
            if (!done[e] && !ev[e].cx && ready(e)) { blk.push_back(ev[e].a); members.push_back(e); }
        if (members.empty()) break;
        SvPass3 ps;
        fill_rest(ps, blk);
        ps.op_begin = (int)c.ops3.size();
        for (size_t l = 0; l < members.size(); ++l) {
            SvOp op;
            op.kind = SV_U2; op.lbit = (int8_t)l; op.cloc = -1; op.cq = -1; op.pad = 0; op.idx = (int16_t)ev[members[l]].mat;
            c.ops3.push_back(op);
            mat_pass[ev[members[l]].mat] = (int)c.passes3.size();
            done[members[l]] = 1; --left;
        }
        ps.op_end = (int)c.ops3.size();
        c.passes3.push_back(ps);
    }
    if (left != 0) { c.passes3.clear(); c.ops3.clear(); c.mats3.clear(); c.mat_gates3.clear(); return; }   // cannot happen for a valid circuit
    c.n_passes3 = (int)c.passes3.size();
    // epilogue groups under the final map: qubits {0,1,2}, {3,4,5}, ... and the remainder
    for (int k0 = 0; k0 < q; k0 += 3) {
        std::vector<int> blk;
        for (int k = k0; k < std::min(q, k0 + 3); ++k) blk.push_back(k);
        SvPass3 ps;
        fill_rest(ps, blk);
        ps.op_begin = ps.op_end = 0;
        c.passes3.push_back(ps);
    }
    // where each parameter enters: its rotation's fused matrix and that matrix's pass
    c.par_mat3.assign(P, -1);
    for (int i = 0; i < P; ++i) c.par_mat3[i] = gate_mat[c.par_gate[i]];
    c.pass_par_begin3.assign(c.n_passes3 + 1, 0);
    for (int ip = 0; ip < c.n_passes3; ++ip) {
        c.pass_par_begin3[ip] = (int)c.pass_params3.size();
        for (int i = 0; i < P; ++i)
            if (c.par_mat3[i] >= 0 && mat_pass[c.par_mat3[i]] == ip) c.pass_params3.push_back(i);
    }
    c.pass_par_begin3[c.n_passes3] = (int)c.pass_params3.size();
    c.has_plan3 = (int)c.pass_params3.size() == P;
}

static std::mutex g_upload_mutex;
int circuit_on_device(const dqgp_circuit* cc) {
    dqgp_circuit* c = const_cast<dqgp_circuit*>(cc);
    std::lock_guard<std::mutex> lock(g_upload_mutex);
    int dev = -1;
    DQGP_CUDA(cudaGetDevice(&dev));
    if (c->d_gates && c->device == dev) return 0;
    DQGP_REQUIRE(c->d_gates == nullptr, "circuit handle was created for device %d but used on device %d", c->device, dev);
    DQGP_CUDA(cudaMalloc(&c->d_gates, sizeof(dqgp_gate) * c->gates.size()));
    DQGP_CUDA(cudaMemcpy(c->d_gates, c->gates.data(), sizeof(dqgp_gate) * c->gates.size(), cudaMemcpyHostToDevice));
    DQGP_CUDA(cudaMalloc(&c->d_passes, sizeof(SvPass) * c->passes.size()));
    DQGP_CUDA(cudaMemcpy(c->d_passes, c->passes.data(), sizeof(SvPass) * c->passes.size(), cudaMemcpyHostToDevice));
    // the kernels stage n_gates op slots (#ops <= #gates): size the buffers accordingly
    DQGP_CUDA(cudaMalloc(&c->d_ops, sizeof(SvOp) * std::max(c->ops.size(), c->gates.size())));
    DQGP_CUDA(cudaMemcpy(c->d_ops, c->ops.data(), sizeof(SvOp) * c->ops.size(), cudaMemcpyHostToDevice));
    DQGP_CUDA(cudaMalloc(&c->d_mats, sizeof(SvMat) * (c->mats.size() + 1)));
    DQGP_CUDA(cudaMemcpy(c->d_mats, c->mats.data(), sizeof(SvMat) * c->mats.size(), cudaMemcpyHostToDevice));
    DQGP_CUDA(cudaMalloc(&c->d_mat_gates, sizeof(int) * (c->mat_gates.size() + 1)));
    DQGP_CUDA(cudaMemcpy(c->d_mat_gates, c->mat_gates.data(), sizeof(int) * c->mat_gates.size(), cudaMemcpyHostToDevice));
    if (c->shareable) {
        std::vector<int> pack;
        pack.insert(pack.end(), c->par_gate.begin(), c->par_gate.end());
        pack.insert(pack.end(), c->par_mat.begin(), c->par_mat.end());
        pack.insert(pack.end(), c->pass_par_begin.begin(), c->pass_par_begin.end());
        pack.insert(pack.end(), c->pass_params.begin(), c->pass_params.end());
        DQGP_CUDA(cudaMalloc(&c->d_share, sizeof(int) * pack.size()));
        DQGP_CUDA(cudaMemcpy(c->d_share, pack.data(), sizeof(int) * pack.size(), cudaMemcpyHostToDevice));
    }
    if (c->has_plan3) {
        DQGP_CUDA(cudaMalloc(&c->d_passes3, sizeof(SvPass3) * c->passes3.size()));
        DQGP_CUDA(cudaMemcpy(c->d_passes3, c->passes3.data(), sizeof(SvPass3) * c->passes3.size(), cudaMemcpyHostToDevice));
        DQGP_CUDA(cudaMalloc(&c->d_ops3, sizeof(SvOp) * std::max(c->ops3.size(), c->gates.size())));
        DQGP_CUDA(cudaMemcpy(c->d_ops3, c->ops3.data(), sizeof(SvOp) * c->ops3.size(), cudaMemcpyHostToDevice));
        DQGP_CUDA(cudaMalloc(&c->d_mats3, sizeof(SvMat) * c->mats3.size()));
        DQGP_CUDA(cudaMemcpy(c->d_mats3, c->mats3.data(), sizeof(SvMat) * c->mats3.size(), cudaMemcpyHostToDevice));
        DQGP_CUDA(cudaMalloc(&c->d_mat_gates3, sizeof(int) * c->mat_gates3.size()));
        DQGP_CUDA(cudaMemcpy(c->d_mat_gates3, c->mat_gates3.data(), sizeof(int) * c->mat_gates3.size(), cudaMemcpyHostToDevice));
        std::vector<int> pack;
        pack.insert(pack.end(), c->par_gate.begin(), c->par_gate.end());
        pack.insert(pack.end(), c->par_mat3.begin(), c->par_mat3.end());
        pack.insert(pack.end(), c->pass_par_begin3.begin(), c->pass_par_begin3.end());
        pack.insert(pack.end(), c->pass_params3.begin(), c->pass_params3.end());
        DQGP_CUDA(cudaMalloc(&c->d_share3, sizeof(int) * pack.size()));
        DQGP_CUDA(cudaMemcpy(c->d_share3, pack.data(), sizeof(int) * pack.size(), cudaMemcpyHostToDevice));
    }
    c->device = dev;
    return 0;
}

}  // namespace dqgp

extern "C" {

int dqgp_version(void) { return DQGP_VERSION; }
const char* dqgp_last_error(void) { return dqgp::g_err.c_str(); }

int dqgp_circuit_create(int encoding, int num_qubits, int num_features, int num_layers, dqgp_circuit** out) {
    DQGP_REQUIRE(out != nullptr, "dqgp_circuit_create: out is NULL");
    *out = nullptr;
    DQGP_REQUIRE(encoding >= DQGP_CHEBYSHEV && encoding <= DQGP_KYRIIENKO, "Unknown encoding type: %d", encoding);
    DQGP_REQUIRE(num_qubits >= 1 && num_qubits <= dqgp::MAX_QUBITS, "num_qubits must be in [1,%d], got %d", dqgp::MAX_QUBITS, num_qubits);
    DQGP_REQUIRE(num_features >= 1 && num_features <= 64, "num_features must be in [1,64], got %d", num_features);
    DQGP_REQUIRE(num_layers >= 1 && num_layers <= 64, "num_layers must be in [1,64], got %d", num_layers);
    dqgp_circuit* c = new dqgp_circuit();
    c->encoding = encoding; c->q = num_qubits; c->d = num_features; c->layers = num_layers;
    c->P = dqgp::count_parameters(encoding, num_qubits, num_layers);
    c->uses_acos = false; c->d_gates = nullptr; c->d_passes = nullptr; c->d_ops = nullptr; c->d_mats = nullptr; c->d_mat_gates = nullptr; c->d_share = nullptr; c->shareable = false;
    c->d_passes3 = nullptr; c->d_ops3 = nullptr; c->d_mats3 = nullptr; c->d_mat_gates3 = nullptr; c->d_share3 = nullptr;
    dqgp::build(*c);
    dqgp::build_plan(*c);
    dqgp::build_plan3(*c);
    c->device = -1;
    *out = c;   // the device copy of the program is made on first use (dqgp::circuit_on_device)
    return 0;
}
void dqgp_circuit_destroy(dqgp_circuit* c) {
    if (!c) return;
    if (c->d_gates) cudaFree(c->d_gates);
    if (c->d_passes) cudaFree(c->d_passes);
    if (c->d_ops) cudaFree(c->d_ops);
    if (c->d_mats) cudaFree(c->d_mats);
    if (c->d_mat_gates) cudaFree(c->d_mat_gates);
    if (c->d_share) cudaFree(c->d_share);
    if (c->d_passes3) cudaFree(c->d_passes3);
    if (c->d_ops3) cudaFree(c->d_ops3);
    if (c->d_mats3) cudaFree(c->d_mats3);
    if (c->d_mat_gates3) cudaFree(c->d_mat_gates3);
    if (c->d_share3) cudaFree(c->d_share3);
    delete c;
}
int dqgp_circuit_num_parameters(const dqgp_circuit* c) { return c ? c->P : -1; }
int dqgp_circuit_num_gates(const dqgp_circuit* c) { return c ? (int)c->gates.size() : -1; }
int dqgp_circuit_num_passes(const dqgp_circuit* c) { return c ? (int)c->passes.size() : -1; }
int dqgp_circuit_num_passes_cx_free(const dqgp_circuit* c) { return c ? (c->has_plan3 ? c->n_passes3 : 0) : -1; }
int dqgp_circuit_num_fused_ops(const dqgp_circuit* c) { return c ? (int)c->ops.size() : -1; }
long long dqgp_circuit_shifted_u2_applications(const dqgp_circuit* c) {
    if (!c) return -1;
    // fused 2x2 unitary applications (one per amplitude pair sweep of a state) that ONE sample costs in dqgp_features_shifted /
    // dqgp_states_shifted: the base circuit twice (final base state + the pass-by-pass advance), then for every parameter the
    // passes from its own pass to the end - once for a rotation parameter (both signs from one fork), twice for a CRZ parameter.
    // A CRZ / CX op counts 3/8 of a 2x2 unitary (6 of 16 flops per pair) / nothing.  Without prefix sharing: (2P + 1) circuits.
    if (c->has_plan3) {      // the CX-free plan (what dqgp_features_shifted runs for circuits without CRZ): fused unitaries only
        const int np3 = c->n_passes3;
        double all3 = 0.0, total3 = 0.0;
        std::vector<double> cost3(np3, 0.0);
        for (int ip = 0; ip < np3; ++ip) { cost3[ip] = c->passes3[ip].op_end - c->passes3[ip].op_begin; all3 += cost3[ip]; }
        total3 = 2.0 * all3;
        for (int ip = 0; ip < np3; ++ip) {
            double suffix = 0.0;
            for (int k = ip; k < np3; ++k) suffix += cost3[k];
            total3 += (c->pass_par_begin3[ip + 1] - c->pass_par_begin3[ip]) * suffix;
        }
        return (long long)(total3 + 0.5);
    }
    const int np = (int)c->passes.size();
    std::vector<double> pass_cost(np, 0.0);
    for (int ip = 0; ip < np; ++ip)
        for (int o = c->passes[ip].op_begin; o < c->passes[ip].op_end; ++o)
            pass_cost[ip] += c->ops[o].kind == SV_U2 ? 1.0 : (c->ops[o].kind == SV_CRZ ? 0.375 : 0.0);
    double all = 0.0;
    for (double v : pass_cost) all += v;
    if (!c->shareable) return (long long)((2.0 * c->P + 1.0) * all + 0.5);
    double total = 2.0 * all;
    for (int ip = 0; ip < np; ++ip) {
        double suffix = 0.0;
        for (int k = ip; k < np; ++k) suffix += pass_cost[k];
        for (int e = c->pass_par_begin[ip]; e < c->pass_par_begin[ip + 1]; ++e)
            total += (c->par_mat[c->pass_params[e]] >= 0 ? 1.0 : 2.0) * suffix;
    }
    return (long long)(total + 0.5);
}
int dqgp_circuit_describe(const dqgp_circuit* c, dqgp_gate* h_out, int capacity) {
    DQGP_REQUIRE(c && h_out, "dqgp_circuit_describe: NULL argument");
    int n = (int)c->gates.size();
    DQGP_REQUIRE(capacity >= n, "dqgp_circuit_describe: capacity %d < %d gates", capacity, n);
    for (int i = 0; i < n; ++i) h_out[i] = c->gates[i];
    return n;
}

}  // extern "C"
