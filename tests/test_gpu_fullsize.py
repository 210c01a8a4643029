"""GPU parity at the REAL shard sizes of BASELINE.json configs[2], [3] and [4] (one agent's shard: n = 2048 / 8192 / 8192)
against the ORACLE: unshifted Gram entries on sampled rows x columns (1e-10 relative), NLL and its components (1e-8), the
fused central-difference gradient (1e-8) and the local ADMM update (same 1e-4 grid point).  The oracle side costs minutes
of NumPy per case (q = 10: one minute per parameter set), so it is committed as fixtures: tests/golden/fullsize_cfg*.npz,
written by tests/golden/make_fullsize_golden.py (pure oracle: oracle/agent_step.py = agent_riemannian.py:209-277, :410-486)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def d():
    import dqgp_b200
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box")
    return dqgp_b200


CASES = [("cfg3", "gaussian"), ("cfg4", "gaussian"), ("cfg5", "matern"), ("cfg5", "gaussian")]


@pytest.mark.parametrize("cfg,outer", CASES)
def test_full_size_shard_matches_oracle(d, cfg, outer):
    path = os.path.join(GOLDEN, f"fullsize_{cfg}.npz")
    if not os.path.exists(path):
        pytest.fail(f"{path} missing: run tests/golden/make_fullsize_golden.py {cfg}")
    g = load_golden(f"fullsize_{cfg}.npz")
    enc, ktype, q, layers, dd, n = str(g["encoding"]), str(g["kernel_type"]), int(g["q"]), int(g["layers"]), int(g["d"]), int(g["n"])
    x, y = d.synthetic_dataset(n, dd, enc)
    # cfg5: the north-star reading (training Gram honours Matern) and the reference's Q1 behaviour (Gaussian) are both recorded
    eng = d.AgentEngine(x, y, encoding_type=enc, kernel_type=ktype, num_qubits=q, num_layers=layers, noise_std=float(g["noise_std"]),
                        rho=float(g["rho"]), L=float(g["L"]), outer_kernel=outer, training_ignores_outer_kernel=False,
                        shift_value=float(g["h"]))
    z, psi, idx = g["z"], g["psi"], g["grad_index"]
    assert z.size == eng.P
    dz, dpsi = d.kernels.dev_f64(z), d.kernels.dev_f64(psi)
    eng.simulate(dz); eng.gram()
    rows, cols = torch.from_numpy(g["rows"]).cuda(), torch.from_numpy(g["cols"]).cuda()
    A = eng.solver.matrix()
    lo, hi = torch.minimum(rows[:, None], cols[None, :]), torch.maximum(rows[:, None], cols[None, :])
    K = A[hi, lo].cpu().numpy()                     # the training Gram fills the lower tiles (all the factorisation reads)
    K[g["rows"][:, None] == g["cols"][None, :]] -= float(g["noise_std"]) ** 2
    ref = g[f"K_{outer}"]
    # the stored diagonal is K_jj + sigma^2, so the subtraction above leaves the rounding of 1.01 - 0.01 there (and a fidelity
    # diagonal is |<psi|psi>|^2 = 1 to a few ulp on both sides): compare those absolutely
    diag = g["rows"][:, None] == g["cols"][None, :]
    assert np.max(np.abs(K - ref)[diag], initial=0.0) < 1e-13
    rel = np.abs(K - ref) / np.maximum(np.abs(ref), 1e-300)
    assert np.max(rel[~diag]) < 1e-10, f"K differs from the oracle by {np.max(rel[~diag]):.2e} relative"
    th, ps = torch.empty(eng.P, dtype=torch.float64, device="cuda"), torch.empty(eng.P, dtype=torch.float64, device="cuda")
    eng.factor(); eng.gradient(); eng.update(dpsi, th, ps)
    torch.cuda.synchronize(); eng.check_info()
    nll = eng.d_nll.cpu().numpy()
    for k, name in enumerate(("log_det_term", "quadratic_term", "constant_term")):
        r = float(g[f"{name}_{outer}"])
        assert abs(nll[k] - r) < 1e-8 * max(1.0, abs(r)), (name, nll[k], r)
    r = float(g[f"nll_{outer}"])
    assert abs(nll[3] - r) < 1e-8 * max(1.0, abs(r)), (nll[3], r)
    alpha = eng.d_alpha.cpu().numpy()
    assert np.max(np.abs(alpha - g[f"alpha_{outer}"])) < 1e-8 * np.abs(g[f"alpha_{outer}"]).max()
    grad, gref = eng.d_grad.cpu().numpy()[idx], g[f"grad_{outer}"]
    assert np.max(np.abs(grad - gref)) < 1e-8 * max(1.0, np.abs(gref).max()), np.max(np.abs(grad - gref))
    tie_tol = 1e4 * 1e-8 * max(1.0, np.abs(gref).max())                               # the gradient tolerance, in grid units
    near_tie = np.abs(np.abs(gref * 1e4 - np.floor(gref * 1e4)) - 0.5) < tie_tol      # rounding cliffs
    assert np.max(np.abs(th.cpu().numpy()[idx] - g[f"theta_{outer}"])[~near_tie], initial=0.0) < 1e-12
    assert np.max(np.abs(ps.cpu().numpy()[idx] - g[f"psi_out_{outer}"])[~near_tie], initial=0.0) < 1e-9
