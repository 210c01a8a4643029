"""The oracle against the golden vectors produced by the REAL reference code (tests/golden/make_golden.py):
riemannian_optimizer.py imported unmodified; RiemannianAgent.train_and_update / main.predict_quantum_gp /
main.main executed unmodified over the squlearn stand-in; scikit-learn's kernel classes."""
import json
import os

import numpy as np
import pytest

from conftest import AGENT_CASES, GOLDEN, load_golden
from oracle import agent_step, driver, qkernels, torus


def test_torus_and_admm_formulas_bit_exact():
    g = load_golden("torus.npz")
    for c in range(4):
        theta, psi, grad, rho = g[f"c{c}_theta"], g[f"c{c}_psi"], g[f"c{c}_grad"], float(g[f"c{c}_rho"])
        assert np.array_equal(torus.update_z(theta, psi, rho), g[f"c{c}_z"])
        assert np.array_equal(torus.circular_mean(theta), g[f"c{c}_circmean"])
        zr = torus.wrap(np.round(g[f"c{c}_z"], 4))
        th = torus.update_theta(zr, grad, psi[0], rho, 100.0)
        assert np.array_equal(th, g[f"c{c}_theta_new"])
        assert np.array_equal(torus.update_psi(psi[0], th, zr, rho), g[f"c{c}_psi_new"])
        assert np.array_equal(torus.wrap(theta[0] * 3 - 2), g[f"c{c}_wrap"])
        assert np.array_equal(torus.wrap(theta[1] - theta[0]), g[f"c{c}_logmap"])
        assert torus.torus_distance(theta[0], theta[1]) == float(g[f"c{c}_dist"])


def test_outer_kernels_equal_scikit_learn():
    g = load_golden("outer_kernels.npz")
    for name in qkernels.OUTER_KERNELS:
        assert np.array_equal(qkernels.outer_kernel_matrix(name, g["F"], g["G"]), g[name])
    # and live against the installed scikit-learn (reference pins 1.7.0, image has 1.9.0)
    from sklearn.gaussian_process.kernels import RBF, ExpSineSquared, Matern
    assert np.array_equal(RBF(length_scale=1 / np.sqrt(2.0))(g["F"], g["G"]), g["gaussian"])
    assert np.array_equal(Matern(length_scale=1.0, nu=1.5)(g["F"], g["G"]), g["matern"])
    assert np.array_equal(ExpSineSquared()(g["F"], g["G"]), g["expsinesquared"])
    # closed forms quoted in DESIGN.md
    from scipy.spatial.distance import cdist
    dist = cdist(g["F"], g["G"])
    assert np.allclose(g["gaussian"], np.exp(-dist ** 2), rtol=1e-14)
    assert np.allclose(g["matern"], (1 + np.sqrt(3) * dist) * np.exp(-np.sqrt(3) * dist), rtol=1e-14)
    assert np.allclose(g["expsinesquared"], np.exp(-2 * np.sin(np.pi * dist) ** 2), rtol=1e-14)


@pytest.mark.parametrize("case", AGENT_CASES)
def test_agent_step_equals_reference_agent(case):
    g = load_golden(f"agent_step_{case}.npz")
    cfg = agent_step.KernelConfig(str(g["encoding"]), str(g["kernel_type"]), int(g["q"]), int(g["layers"]), str(g["outer_kernel"]))
    r = agent_step.train_and_update(cfg, g["X"], g["Y"], g["z"], g["psi"], float(g["noise_std"]), float(g["rho"]), float(g["L"]),
                                    float(g["h"]), keep_k=True)
    assert np.array_equal(r.theta, g["theta_out"]) and np.array_equal(r.psi, g["psi_out"])
    assert r.nll == float(g["nll"]) and r.cond == float(g["cond"])
    assert r.components["log_det_term"] == float(g["log_det_term"])
    assert r.components["quadratic_term"] == float(g["quadratic_term"])
    assert np.array_equal(r.K, g["K"])
    mean, var = driver.predict(cfg, g["X"], g["Y"], g["X_test"], np.mod(g["z"], np.pi), 0.1)
    assert np.array_equal(mean, g["pred_mean"]) and np.array_equal(var, g["pred_var"])


def test_first_iteration_of_reference_main_trajectory():
    with open(os.path.join(GOLDEN, "trajectory_cfg1.json")) as f:
        rec = json.load(f)
    data = load_golden("trajectory_cfg1_data.npz")
    it = rec["iterations"][0]
    cfg = agent_step.KernelConfig("chebyshev", "projected", 3, 1, "matern")
    a = 3                                               # smallest shard keeps the CPU suite fast
    r = agent_step.train_and_update(cfg, data[f"X_{a}"], data[f"Y_{a}"], np.array(it["z"]), np.array(it["psi_in"][a]), 0.1,
                                    100.0, 100.0, workers=None, want_cond=False)
    assert np.array_equal(r.theta, np.array(it["theta_out"][a]))
    assert np.array_equal(r.psi, np.array(it["psi_out"][a]))
    assert r.nll == it["nll"][a]
    # consensus of iteration 2 from the recorded outputs of iteration 1 (main.py:2523)
    z2 = np.round(torus.update_z(np.array(it["theta_out"]), np.array(it["psi_out"]), 100.0), 4)
    assert np.array_equal(z2, np.array(rec["iterations"][1]["z"]))


@pytest.mark.parametrize("name,late", [("trajectory_cfg1", 29), ("trajectory_cfg2", 9), ("trajectory_cfg2_srtm", 7), ("trajectory_cfg3s", 7), ("trajectory_cfg4s", 5),
                                       ("trajectory_cfg5s", 5)])
def test_whole_reference_main_trajectory_consensus_chain(name, late):
    """Every z of the recorded main.main() runs (30 iterations of configs[0]; 10 of the configs[1]-shaped run; 8 / 6 / 6 of the
    reduced configs[2], [3], [4]) follows bit for bit from the previous iteration's recorded theta / psi (main.py:2523), and the
    oracle's agent step reproduces a LATE iteration's outputs of the last shard exactly (the state has left the initial random
    grid points by then)."""
    with open(os.path.join(GOLDEN, f"{name}.json")) as f:
        rec = json.load(f)
    its = rec["iterations"]
    assert len(its) == late + 1
    for k in range(1, len(its)):
        z = np.round(torus.update_z(np.array(its[k - 1]["theta_out"]), np.array(its[k - 1]["psi_out"]), 100.0), 4)
        assert np.array_equal(z, np.array(its[k]["z"])), k
        assert its[k]["psi_in"] == its[k - 1]["psi_out"]
    data = load_golden(f"{name}_data.npz")
    it, a = its[late], rec["n_agents"] - 1
    argv = rec["argv"].split()
    get = lambda flag, default=None: argv[argv.index(flag) + 1] if flag in argv else default
    cfg = agent_step.KernelConfig(get("--encoding"), get("--kernel-type"), int(get("--num-qubits")), int(get("--num-layers")),
                                  get("--outer-kernel", "gaussian"))
    r = agent_step.train_and_update(cfg, data[f"X_{a}"], data[f"Y_{a}"], np.array(it["z"]), np.array(it["psi_in"][a]), 0.1,
                                    100.0, 100.0, workers=None, want_cond=False)
    assert np.array_equal(r.theta, np.array(it["theta_out"][a])) and np.array_equal(r.psi, np.array(it["psi_out"][a]))
    assert r.nll == it["nll"][a]


def test_q1_training_ignores_outer_kernel_switch():
    g = load_golden("agent_step_cheb_proj_matern_q3.npz")
    kw = dict(encoding_type="chebyshev", kernel_type="projected", num_qubits=3, num_layers=1, outer_kernel="matern")
    quirk = agent_step.KernelConfig(**kw)
    honest = agent_step.KernelConfig(training_ignores_outer_kernel=False, **kw)
    assert quirk.training_outer_kernel() == "gaussian" and honest.training_outer_kernel() == "matern"
    k1 = quirk.make(2, training=True); k2 = honest.make(2, training=True)
    for k in (k1, k2):
        k.assign_parameters(np.mod(g["z"], np.pi))
    assert np.array_equal(k1.evaluate(g["X"], g["X"]), g["K"])
    assert not np.allclose(k2.evaluate(g["X"], g["X"]), g["K"])


def test_finite_difference_is_central_difference_not_shift_rule():
    """Q3: dK_i = (K+ - K-)/(2h) with h = pi/8; sanity against a tight numerical derivative (loose tolerance)."""
    rng = np.random.default_rng(0)
    cfg = agent_step.KernelConfig("yz_cx", "projected", 2, 1, "gaussian")
    x = rng.uniform(-1, 1, (8, 2)); p = rng.uniform(0.5, 2.5, 4)
    _, dk = agent_step.kernel_and_derivatives(cfg, x, p, np.pi / 8)
    _, dk_small = agent_step.kernel_and_derivatives(cfg, x, p, 1e-5)
    assert 1e-6 < np.abs(dk - dk_small).max() < 0.3
