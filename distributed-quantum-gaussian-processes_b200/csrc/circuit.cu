// Gate programs of the four in-scope encoding circuits, built on the host and kept resident on the
// device.  Restates squlearn 0.9.1's ChebyshevPQC / HubregtsenEncodingCircuit / YZ_CX_EncodingCircuit
// (ctor sites: reference main.py:68-83, agent_riemannian.py:51-66) with default options; the Kyriienko
// program is this project's own definition because the reference's call has no upstream behaviour
// (SURVEY Q13).  An independent restatement lives in oracle/circuits.py; tests compare the two.
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
            for (int l = 0; l < L; ++l) {
                for (int i = 0; i < q; ++i) G.push_back(mk(DQGP_G_RX, i, -1, DQGP_A_P_TIMES_ACOS, take(), i % d));
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
                    G.push_back(mk(DQGP_G_RY, i, -1, DQGP_A_P_PLUS_CX, take(), i % d, 1.0));
                    G.push_back(mk(DQGP_G_RZ, i, -1, DQGP_A_P_PLUS_CX, take(), i % d, 1.0));
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
    c->uses_acos = false; c->d_gates = nullptr;
    dqgp::build(*c);
    c->device = -1;
    *out = c;   // the device copy of the program is made on first use (dqgp::circuit_on_device)
    return 0;
}
void dqgp_circuit_destroy(dqgp_circuit* c) {
    if (!c) return;
    if (c->d_gates) cudaFree(c->d_gates);
    delete c;
}
int dqgp_circuit_num_parameters(const dqgp_circuit* c) { return c ? c->P : -1; }
int dqgp_circuit_num_gates(const dqgp_circuit* c) { return c ? (int)c->gates.size() : -1; }
int dqgp_circuit_describe(const dqgp_circuit* c, dqgp_gate* h_out, int capacity) {
    DQGP_REQUIRE(c && h_out, "dqgp_circuit_describe: NULL argument");
    int n = (int)c->gates.size();
    DQGP_REQUIRE(capacity >= n, "dqgp_circuit_describe: capacity %d < %d gates", capacity, n);
    for (int i = 0; i < n; ++i) h_out[i] = c->gates[i];
    return n;
}

}  // extern "C"
