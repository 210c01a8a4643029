"""Oracle results at the REAL shard sizes of BASELINE.json configs[2..4] (one agent's shard each), committed as
fixtures because they cost minutes of host CPU (a q = 10 simulation of 8192 samples takes ~1 minute per parameter set
in NumPy) — the GPU tests compare against them instead of re-running the oracle on the GPU box.

    python tests/golden/make_fullsize_golden.py [cfg3] [cfg4] [cfg5]

Everything here is the oracle (oracle/agent_step.py restating agent_riemannian.py:209-277, :410-486): the unshifted Gram
(sampled rows x sampled columns are stored, the full matrix is 537 MB), its NLL terms through LAPACK, the central-difference
gradient for the listed parameters through materialised shifted Grams, and the local update that follows from it.
"""
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import agent_step, circuits, driver, qkernels, torus  # noqa: E402

H = np.pi / 8
CASES = {
    # name: encoding, kernel_type, q, layers, d, n, training outer kernels to record, number of gradient parameters
    "cfg3": ("hubregtsen", "fidelity", 5, 2, 2, 2048, ("gaussian",), None),
    "cfg4": ("yz_cx", "projected", 8, 3, 4, 8192, ("gaussian",), None),
    "cfg5": ("kyriienko", "projected", 10, 4, 6, 8192, ("matern", "gaussian"), 12),
}
_G = {}      # shared with the forked workers: x, kernel config, brackets per outer kernel


def _gram_from(enc, ktype, outer, x, p):
    if ktype == "fidelity":
        s = enc.states(x, p)
        ov = s @ s.conj().T
        return ov.real ** 2 + ov.imag ** 2
    f = enc.features(x, p)
    return {o: qkernels.outer_kernel_matrix(o, f, f) for o in outer}


def _grad_job(i):
    """0.5 * sum(bracket o dK_i^T) for parameter i, dK_i = (K(p + h e_i) - K(p - h e_i)) / 2h (agent_riemannian.py:275, :431-437)."""
    enc, ktype, outers, x, sets = _G["enc"], _G["ktype"], _G["outers"], _G["x"], _G["sets"]
    up = _gram_from(enc, ktype, outers, x, sets[1 + 2 * i])
    dn = _gram_from(enc, ktype, outers, x, sets[2 + 2 * i])
    out = []
    for o in outers:
        dk = ((up - dn) if ktype == "fidelity" else (up[o] - dn[o])) / (2.0 * H)
        out.append(0.5 * np.sum(_G["bracket"][o] * dk.T))
    return i, out


def make(name):
    encoding, ktype, q, layers, d, n, outers, n_grad = CASES[name]
    t0 = time.time()
    x, y = driver.synthetic_dataset(n, d, encoding)
    P = circuits.num_parameters(encoding, q, layers)
    rs = np.random.RandomState(42)
    z, psi = np.round(rs.rand(P), 4), np.round(rs.rand(P), 4)
    sets = agent_step.shifted_parameter_sets(z, H)
    enc = qkernels.EncodingCircuit(encoding, q, d, layers)
    idx = np.arange(P) if n_grad is None else np.unique(np.linspace(0, P - 1, n_grad).round().astype(int))
    pick = np.random.default_rng(1)
    rows, cols = np.sort(pick.choice(n, 256, replace=False)), np.sort(pick.choice(n, 512, replace=False))
    k0 = _gram_from(enc, ktype, outers, x, sets[0])
    out = dict(encoding=encoding, kernel_type=ktype, q=q, layers=layers, d=d, n=n, z=z, psi=psi, h=H, noise_std=0.1, rho=100.0,
               L=100.0, grad_index=idx, rows=rows, cols=cols, outers=np.array(outers))
    brackets = {}
    for o in outers:
        k = k0 if ktype == "fidelity" else k0[o]
        grad0, comp, _, alpha, c_inv = agent_step.gp_terms(k, np.zeros((0, n, n)), y, 0.1, want_cond=False)
        brackets[o] = c_inv - np.outer(alpha, alpha)
        out.update({f"K_{o}": k[np.ix_(rows, cols)], f"Kdiag_{o}": np.diag(k).copy(), f"alpha_{o}": alpha,
                    f"nll_{o}": comp["total"], f"log_det_term_{o}": comp["log_det_term"],
                    f"quadratic_term_{o}": comp["quadratic_term"], f"constant_term_{o}": comp["constant_term"]})
        print(name, o, "K + LAPACK terms done, nll =", comp["total"], f"({time.time() - t0:.0f} s)", flush=True)
        del c_inv
    del k0
    _G.update(enc=enc, ktype=ktype, outers=outers, x=x, sets=sets, bracket=brackets)
    workers = min(len(idx), int(os.environ.get("DQGP_GOLDEN_WORKERS", "6")))
    with ProcessPoolExecutor(max_workers=workers) as pool:       # fork: the workers see _G without pickling the brackets
        res = dict(pool.map(_grad_job, [int(i) for i in idx]))
    for k_o, o in enumerate(outers):
        g = np.array([res[int(i)][k_o] for i in idx])
        full = np.zeros(P)
        full[idx] = g
        theta, psi_new = agent_step.local_update(torus.wrap(z), np.round(full, 4), psi, 100.0, 100.0)
        out.update({f"grad_{o}": g, f"theta_{o}": theta[idx], f"psi_out_{o}": psi_new[idx]})
    np.savez_compressed(os.path.join(HERE, f"fullsize_{name}.npz"), **out)
    print(name, "written", f"({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    for case in (sys.argv[1:] or list(CASES)):
        make(case)
