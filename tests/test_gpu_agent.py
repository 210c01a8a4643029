"""GPU parity, agent / trajectory level: the drop-in Python surface against golden vectors produced by the
REAL reference code (tests/golden/make_golden.py) and against the oracle at larger sizes."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import AGENT_CASES, GOLDEN, load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def d():
    import dqgp_b200
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box")
    return dqgp_b200


def _agent(d, g, **kw):
    return d.RiemannianAgent("golden", g["X"], g["Y"], int(g["q"]), float(g["noise_std"]), float(g["rho"]), float(g["L"]),
                             q_kernel=None, use_parameter_shift=True, num_workers=2, shift_value=float(g["h"]),
                             num_layers=int(g["layers"]), encoding_type=str(g["encoding"]), kernel_type=str(g["kernel_type"]),
                             measurement="XYZ", outer_kernel=str(g["outer_kernel"]), **kw)


@pytest.mark.parametrize("case", AGENT_CASES)
def test_train_and_update_matches_reference(d, case):
    """RiemannianAgent.train_and_update == the real reference agent's output on the same inputs:
    theta_i, psi_i on the same 1e-4 grid point, NLL and its components to 1e-8."""
    g = load_golden(f"agent_step_{case}.npz")
    theta, psi, nll, cond, comp = _agent(d, g).train_and_update(g["z"], g["psi"])
    assert np.max(np.abs(theta - g["theta_out"])) < 1e-12
    assert np.max(np.abs(psi - g["psi_out"])) < 1e-9
    assert abs(nll - float(g["nll"])) < 1e-8 * max(1.0, abs(float(g["nll"])))
    assert set(comp) == {"log_det_term", "quadratic_term", "constant_term", "total"}
    assert np.isnan(cond)


def test_condition_number_diagnostic(d):
    g = load_golden("agent_step_yzcx_fid_q2.npz")
    _, _, _, cond, _ = _agent(d, g, compute_condition_number=True).train_and_update(g["z"], g["psi"])
    ref = float(g["cond"])
    assert cond > 0 and abs(np.log10(cond) - np.log10(ref)) < 1.0     # SVD of a numerically singular Gram: order of magnitude


def test_process_agent_training_tuple_interface(d):
    g = load_golden("agent_step_yzcx_proj_gauss_q4.npz")
    tup = ("agent_1", g["X"], g["Y"], int(g["q"]), 0.1, 100.0, 100.0, g["z"], g["psi"], True, int(g["d"]), int(g["layers"]), None,
           float(g["h"]), "yz_cx", "projected", "XYZ", 0.015, "gradient_descent", 0.9, "gaussian", {"gamma": 1.0}, None)
    theta, psi, nll, cond, comp = d.process_agent_training(tup)
    assert np.max(np.abs(theta - g["theta_out"])) < 1e-12 and np.max(np.abs(psi - g["psi_out"])) < 1e-9


def test_train_agents_concurrent_equals_sequential(d):
    """dqgp_b200.train_agents (one CUDA stream per agent, async submit/collect) returns exactly what sequential
    train_and_update calls return — the GPU counterpart of the reference's executor.map fan-out (main.py:2530-2542)."""
    cases = ["yzcx_proj_gauss_q4", "cheb_proj_gauss_q4", "hub_fid_q5"]
    gs = [load_golden(f"agent_step_{c}.npz") for c in cases[:1]] * 3
    agents = [_agent(d, g) for g in gs]
    for i, a in enumerate(agents):
        a.agent_id = f"agent_{i}"                      # distinct engines -> really concurrent
    z, psis = gs[0]["z"], [gs[0]["psi"], gs[0]["psi"] * 0.5, gs[0]["psi"] * 0.25]
    par = d.train_agents(agents, z, psis)
    seq = [a.train_and_update(z, p) for a, p in zip(agents, psis)]
    for (t1, p1, n1, _, _), (t2, p2, n2, _, _) in zip(par, seq):
        assert np.array_equal(t1, t2) and np.array_equal(p1, p2) and n1 == n2
    assert np.max(np.abs(par[0][0] - gs[0]["theta_out"])) < 1e-12


@pytest.mark.parametrize("outer_blocks", [1, 2, 4, -1])
def test_cholesky_panel_width_does_not_change_results(d, outer_blocks):
    from oracle import driver
    x, y = driver.synthetic_dataset(1100, 4, "yz_cx")
    res = []
    for ob in (0, outer_blocks):
        eng = d.AgentEngine(x, y, encoding_type="yz_cx", kernel_type="projected", num_qubits=4, num_layers=2, noise_std=0.1,
                            rho=100.0, L=100.0, cholesky_outer_blocks=ob)
        z = d.kernels.dev_f64(np.round(np.random.RandomState(1).rand(eng.P), 4))
        eng.simulate(z); eng.gram(); eng.factor(); eng.gradient()
        torch.cuda.synchronize()
        eng.check_info()
        res.append((eng.d_grad.cpu().numpy(), eng.d_nll.cpu().numpy()))
    assert np.max(np.abs(res[0][0] - res[1][0])) < 1e-9 * max(1.0, np.abs(res[0][0]).max())
    assert abs(res[0][1][3] - res[1][1][3]) < 1e-9 * abs(res[0][1][3])


def test_cuda_graph_replay_equals_eager_iterations(d):
    """AdmmEngine.capture()/replay() (whole iteration, all agent streams and the solver's look-ahead streams in one CUDA
    graph) walks exactly the same trajectory as eager iteration() calls."""
    from oracle import driver
    x, y = driver.synthetic_dataset(4 * 150, 2, "chebyshev")
    shards = [(x[a * 150:(a + 1) * 150], y[a * 150:(a + 1) * 150]) for a in range(4)]
    rs = np.random.RandomState(42)
    theta0, psi0 = np.round(rs.rand(4, 12), 4), np.round(rs.rand(4, 12), 4)
    kw = dict(rho=100.0, L=100.0, encoding_type="chebyshev", kernel_type="projected", num_qubits=3, num_layers=1, noise_std=0.1)
    eager = d.AdmmEngine(shards, theta0, psi0, **kw)
    for _ in range(4):
        eager.iteration()
    graph = d.AdmmEngine(shards, theta0, psi0, **kw)
    graph.capture()                      # its warm-up iteration is rolled back: capture() does not advance the run
    for _ in range(4):
        graph.replay()
    for a, b in zip(eager.state(), graph.state()):
        assert np.array_equal(a, b)


def test_wrong_parameter_count_raises(d):
    g = load_golden("agent_step_yzcx_fid_q2.npz")
    with pytest.raises(ValueError):
        _agent(d, g).train_and_update(g["z"][:-1], g["psi"][:-1])
    with pytest.raises(ValueError):
        d.create_quantum_kernel(3, 2, 1, True, "no_such_circuit", "fidelity")
    with pytest.raises(ValueError):
        d.create_quantum_kernel(3, 2, 1, True, "yz_cx", "no_such_kernel")


TRAJECTORIES = ["trajectory_cfg1", "trajectory_cfg2", "trajectory_cfg2_srtm", "trajectory_cfg3s", "trajectory_cfg4s", "trajectory_cfg5s"]


def _trajectory_config(rec):
    """Engine keyword arguments from the argv the real main.main() was run with (recorded in the golden file)."""
    a = rec["argv"].split()
    get = lambda flag, default=None: a[a.index(flag) + 1] if flag in a else default
    return dict(encoding_type=get("--encoding"), kernel_type=get("--kernel-type"), num_qubits=int(get("--num-qubits")),
                num_layers=int(get("--num-layers")), noise_std=0.1, outer_kernel=get("--outer-kernel", "gaussian"))


@pytest.mark.parametrize("name", TRAJECTORIES)
def test_trajectory_matches_reference_main(d, name):
    """BASELINE.json configs[0] (30 iterations), a configs[1]-shaped run (chebyshev q=4, 3 layers, P = 32, 10 iterations; the SRTM
    tile is absent, so main.py's synthetic branch), configs[1] itself on main.py's SRTM branch over a synthetic tile (8 iterations;
    tests/test_data_plumbing.py rebuilds its shards from the tile) and reduced-size runs of configs[2], [3], [4] (hubregtsen fidelity q=5, 8 agents;
    yz_cx projected-Gaussian q=6, 8 agents; kyriienko projected-Matern q=4, 4 agents): replay the ADMM trajectory the real
    main.main() produced (regional partitions of unequal size, matern flag -> Gaussian training Grams (Q1), rho = L = 100) through
    AdmmEngine on the device - every iteration's z, theta on the same 1e-4 grid point, psi, per-agent NLL to 1e-8."""
    with open(os.path.join(GOLDEN, f"{name}.json")) as f:
        rec = json.load(f)
    data = load_golden(f"{name}_data.npz")
    A = rec["n_agents"]
    shards = [(data[f"X_{a}"], data[f"Y_{a}"]) for a in range(A)]
    it0 = rec["iterations"][0]
    P = len(it0["z"])
    psi0 = np.array(it0["psi_in"])
    # theta before the first z-update is not recorded; drive iteration 1 from its recorded z instead
    eng = d.AdmmEngine(shards, np.zeros((A, P)), psi0, rho=100.0, L=100.0, **_trajectory_config(rec))
    for k, it in enumerate(rec["iterations"]):
        if k == 0:
            eng.z.copy_(torch.tensor(it["z"], dtype=torch.float64, device="cuda"))
            for i, ag in enumerate(eng.agents):
                ag.step(eng.z, eng.psi[i], eng.local_theta[i], eng.local_psi[i])
            eng.theta.copy_(eng.local_theta); eng.psi.copy_(eng.local_psi)
        else:
            eng.iteration()
        z, theta, psi, nll = eng.state()
        assert np.max(np.abs(z - np.array(it["z"]))) < 1e-12, f"z differs at iteration {k + 1}"
        assert np.max(np.abs(theta - np.array(it["theta_out"]))) < 1e-12, f"theta differs at iteration {k + 1}"
        assert np.max(np.abs(psi - np.array(it["psi_out"]))) < 1e-9, f"psi differs at iteration {k + 1}"
        ref_nll = np.array(it["nll"])
        assert np.max(np.abs(nll - ref_nll) / np.maximum(1.0, np.abs(ref_nll))) < 1e-8
    assert len(rec["iterations"]) >= 6


@pytest.mark.parametrize("case", AGENT_CASES)
def test_predict_matches_reference(d, case):
    g = load_golden(f"agent_step_{case}.npz")
    # hub_proj_ess_q3: ExpSineSquared on 9 features is indefinite (eigenvalues -2.5 .. 15.8): np.linalg.cholesky raises in the
    # reference and its except-branch inverts with np.linalg.inv (main.py:1479-1486).  The device path walks the same rung
    # (partial-pivoting LU, csrc/lu.cu) and must reproduce the reference's numbers like any other case.
    d.predict_quantum_gp.used_lu_fallback = False
    mean, var, *_ = d.predict_quantum_gp(g["X"], g["Y"], g["X_test"], np.mod(g["z"], np.pi), int(g["q"]), int(g["layers"]), 0.1,
                                         True, str(g["encoding"]), str(g["kernel_type"]), "XYZ", str(g["outer_kernel"]),
                                         Y_test=g["Y_test"])
    assert np.max(np.abs(mean - g["pred_mean"])) < 1e-8 * max(1.0, np.abs(g["pred_mean"]).max())
    assert np.max(np.abs(var - g["pred_var"])) < 1e-8
    from oracle import driver
    ref_nlpd = driver.nlpd(g["Y_test"], g["pred_mean"], g["pred_var"])
    assert abs(d.predict_quantum_gp.last_nlpd - ref_nlpd) < 1e-8 * max(10.0, abs(ref_nlpd))
    assert abs(d.nlpd(g["Y_test"], mean, var) - d.predict_quantum_gp.last_nlpd) < 1e-10 * max(1.0, abs(d.predict_quantum_gp.last_nlpd))
    assert d.predict_quantum_gp.used_lu_fallback == (case == "hub_proj_ess_q3")


@pytest.mark.parametrize("case", ["cheb_proj_matern_q3", "hub_fid_q5", "yzcx_proj_gauss_q4"])
def test_lean_predict_matches_reference(d, case):
    """The lean prediction path (one square, substitution solves, in-place K(test,train)) against the reference's
    predict_quantum_gp outputs, and against the three-square path."""
    g = load_golden(f"agent_step_{case}.npz")
    args = (g["X"], g["Y"], g["X_test"], np.mod(g["z"], np.pi), int(g["q"]), int(g["layers"]), 0.1, True, str(g["encoding"]),
            str(g["kernel_type"]), "XYZ", str(g["outer_kernel"]))
    mean, var, *_ = d.predict_quantum_gp(*args, Y_test=g["Y_test"], lean=True)
    nl = d.predict_quantum_gp.last_nlpd
    assert np.max(np.abs(mean - g["pred_mean"])) < 1e-8 * max(1.0, np.abs(g["pred_mean"]).max())
    assert np.max(np.abs(var - g["pred_var"])) < 1e-8
    mean2, var2, *_ = d.predict_quantum_gp(*args, Y_test=g["Y_test"], lean=False)
    assert np.max(np.abs(mean - mean2)) < 1e-10 and np.max(np.abs(var - var2)) < 1e-10
    assert abs(nl - d.predict_quantum_gp.last_nlpd) < 1e-9


@pytest.mark.parametrize("lean", [False, True])
def test_generate_quantum_gp_data_matches_reference_main(d, lean):
    """the dataset main.main() generated for configs[0] (seed 42, data seed 7), regenerated on the GPU
    with the same host RNG stream, split like main.py:2356: X identical, Y equal to the Cholesky-sampling accuracy;
    with the three-square workspace and with the in-place (lean) factorisation used at N >= 65536."""
    from sklearn.model_selection import train_test_split
    data = load_golden("trajectory_cfg1_data.npz")
    X, Y, truth = d.generate_quantum_gp_data(1000, 2, 3, 1, (-2.0, 2.0), 0.1, True, None, "chebyshev", "projected", "XYZ", "matern",
                                             {"length_scale": 1.0, "nu": 1.5}, None, data_seed=7, param_seed=42, lean=lean)
    Xtr, Xte, Ytr, Yte = train_test_split(X, Y, test_size=0.1, random_state=42, shuffle=True)
    assert np.array_equal(Xtr, data["X_train"])
    assert np.max(np.abs(Ytr - data["Y_train"])) < 1e-6 * max(1.0, np.abs(data["Y_train"]).max())
    assert truth.shape == (12,) and np.all((truth >= 0) & (truth <= np.pi))


@pytest.mark.parametrize("name", TRAJECTORIES)
def test_run_admm_driver_matches_reference_main(d, name):
    """The whole driver loop (z-update, agents, collect/round, per-iteration 5-fold CV NLPD, stop at max_iter with the
    best-CV consensus) against what the real main.main() did to ITS stop (max_iter 30 for BASELINE configs[0], 10 for the
    configs[1]-shaped run): every iteration's z / theta / CV NLPD, and the FINAL consensus parameters main() ends with
    (the best-CV z, main.py:2777-2780)."""
    with open(os.path.join(GOLDEN, f"{name}.json")) as f:
        rec = json.load(f)
    data = load_golden(f"{name}_data.npz")
    A = rec["n_agents"]
    shards = [(data[f"X_{a}"], data[f"Y_{a}"]) for a in range(A)]
    # the state before iteration 1 is not recorded: start from iteration 1's outputs and replay iterations 2..end
    it1 = rec["iterations"][0]
    n_it = len(rec["iterations"]) - 1
    out = d.run_admm(shards, rho=100.0, L=100.0, max_iter=n_it, theta0=np.array(it1["theta_out"]), psi0=np.array(it1["psi_out"]),
                     cv_data=(data["X_train"], data["Y_train"]), cv_folds=5, seed=42 + 1, **_trajectory_config(rec))
    assert out["iterations"] == n_it and out["stop_reason"] == "max_iter"
    for k, h in enumerate(out["history"]):
        ref_it, ref_cv = rec["iterations"][k + 1], rec["cv"][k + 1]
        assert np.max(np.abs(h["z"] - np.array(ref_it["z"]))) < 1e-12
        assert np.max(np.abs(h["theta"] - np.array(ref_it["theta_out"]))) < 1e-12
        assert ref_cv["random_seed"] == 42 + k + 2
        assert abs(h["cv"]["mean_nlpd"] - ref_cv["mean_nlpd"]) < 1e-7
    # main()'s final consensus = the z of its best-CV iteration (strict improvement, first wins), over ALL its iterations
    scores = [c["mean_nlpd"] for c in rec["cv"]]
    best = int(np.argmin(scores))
    if best >= 1:                                   # (iteration 1 itself is not replayed here)
        assert np.max(np.abs(out["z"] - np.array(rec["iterations"][best]["z"]))) < 1e-8
        assert np.array_equal(out["z"], out["history"][best - 1]["z"])
    else:
        best2 = 1 + int(np.argmin(scores[1:]))
        assert np.array_equal(out["z"], out["history"][best2 - 1]["z"])


def test_cv_nlpd_matches_reference_main(d):
    with open(os.path.join(GOLDEN, "trajectory_cfg1.json")) as f:
        rec = json.load(f)
    data = load_golden("trajectory_cfg1_data.npz")
    cv = rec["cv"][0]
    out = d.k_fold_cross_validation_consensus(data["X_train"], data["Y_train"], np.array(cv["params"]), 3, 1, 0.1, k_folds=5,
                                              encoding_type="chebyshev", kernel_type="projected", outer_kernel="matern",
                                              random_seed=cv["random_seed"])
    assert np.max(np.abs(np.array(out["fold_nlpds"]) - np.array(cv["fold_nlpds"]))) < 1e-7
    assert abs(out["mean_nlpd"] - cv["mean_nlpd"]) < 1e-7


@pytest.mark.parametrize("enc,ktype,q,layers,dd,n", [("yz_cx", "projected", 8, 3, 4, 700), ("hubregtsen", "fidelity", 5, 2, 2, 520),
                                                   ("kyriienko", "projected", 6, 2, 3, 400), ("chebyshev", "fidelity", 4, 2, 2, 330)])
def test_agent_step_medium_size_against_oracle(d, enc, ktype, q, layers, dd, n):
    from oracle import agent_step, circuits, driver
    x, y = driver.synthetic_dataset(n, dd, enc)
    P = circuits.num_parameters(enc, q, layers)
    rs = np.random.RandomState(42)
    z, psi = np.round(rs.rand(P), 4), np.round(rs.rand(P), 4)
    cfg = agent_step.KernelConfig(enc, ktype, q, layers, "gaussian")
    ref = agent_step.train_and_update(cfg, x, y, z, psi, 0.1, 100.0, 100.0, workers=None, want_cond=False)
    ag = d.RiemannianAgent("m", x, y, q, 0.1, 100.0, 100.0, use_parameter_shift=True, num_layers=layers, encoding_type=enc,
                           kernel_type=ktype)
    theta, psi_new, nll, _, _ = ag.train_and_update(z, psi)
    near_tie = np.abs(np.abs(ref.grad * 1e4 - np.floor(ref.grad * 1e4)) - 0.5) < 1e-5      # rounding cliffs
    assert np.max(np.abs(ag.last_gradient - ref.grad)) < 1e-8 * max(1.0, np.abs(ref.grad).max())
    assert np.max(np.abs(theta - ref.theta)[~near_tie]) < 1e-12
    assert abs(nll - ref.nll) < 1e-8 * max(1.0, abs(ref.nll))


def test_nan_inputs_propagate_and_are_reported(d):
    """arccos of |x| > 1 (chebyshev / kyriienko without clipping): NaNs propagate through features and K as in
    the reference; dataset generation reports them like main.py:248-249."""
    x = np.array([[0.1, 0.2], [1.5, 0.3], [0.4, -0.2]])
    qk = d.create_quantum_kernel(3, 2, 1, True, "kyriienko", "projected")
    qk.assign_parameters(np.full(qk.encoding_circuit.num_parameters, 0.3))
    K = qk.evaluate(x, x)
    assert np.isnan(K[1]).all() and np.isnan(K[:, 1]).all() and np.isfinite(K[0, 2]) and K[0, 0] == 1.0
    with pytest.raises(ValueError, match="NaN"):
        d.generate_quantum_gp_data(20, 2, 3, 1, (-2.0, 2.0), 0.1, True, None, "kyriienko", "projected", data_seed=1)


def test_duplicate_samples_match_oracle(d):
    """Collisions: repeated inputs give identical Gram rows (exact ones off the diagonal); K + sigma^2 I stays SPD."""
    from oracle import agent_step, driver
    x, y = driver.synthetic_dataset(90, 3, "hubregtsen")
    x[10:20] = x[0:10]
    x[50] = x[51]
    cfg = agent_step.KernelConfig("hubregtsen", "projected", 4, 2, "gaussian")
    rs = np.random.RandomState(3)
    z, psi = np.round(rs.rand(16), 4), np.round(rs.rand(16), 4)
    ref = agent_step.train_and_update(cfg, x, y, z, psi, 0.1, 100.0, 100.0, want_cond=False, keep_k=True)
    ag = d.RiemannianAgent("dup", x, y, 4, 0.1, 100.0, 100.0, use_parameter_shift=True, num_layers=2, encoding_type="hubregtsen",
                           kernel_type="projected")
    theta, psi_new, nll, _, _ = ag.train_and_update(z, psi)
    assert ref.K[0, 10] == 1.0
    assert np.max(np.abs(ag.last_gradient - ref.grad)) < 1e-8 * max(1.0, np.abs(ref.grad).max())
    assert np.max(np.abs(theta - ref.theta)) < 1e-12 and abs(nll - ref.nll) < 1e-8 * max(1.0, abs(ref.nll))


@pytest.mark.parametrize("enc,ktype,q,layers,dd,n,outer", [("hubregtsen", "fidelity", 5, 2, 2, 2048, "gaussian"),
                                                          ("kyriienko", "projected", 10, 4, 6, 8192, "matern")])
def test_full_size_properties_other_configs(d, enc, ktype, q, layers, dd, n, outer):
    """BASELINE.json configs[2] and configs[4] shard sizes: A * A^-1 = I, alpha solves the system, the fused gradient is
    reproducible bit for bit, and agrees with a materialised torch fp64 contraction for one parameter."""
    x, y = d.synthetic_dataset(n, dd, enc)
    eng = d.AgentEngine(x, y, encoding_type=enc, kernel_type=ktype, num_qubits=q, num_layers=layers, noise_std=0.1, rho=100.0,
                        L=100.0, outer_kernel=outer, training_ignores_outer_kernel=False)
    z = d.kernels.dev_f64(np.round(np.random.RandomState(42).rand(eng.P), 4))
    eng.simulate(z); eng.gram()
    K = eng.solver.matrix().clone()
    K = torch.tril(K) + torch.tril(K, -1).T
    assert torch.allclose(torch.diagonal(K), torch.full((n,), 1.01, dtype=torch.float64, device="cuda"), rtol=0, atol=1e-13)
    eng.factor(); eng.gradient()
    torch.cuda.synchronize(); eng.check_info()
    Ainv = torch.tril(eng.solver.inverse()); Ainv = Ainv + Ainv.T - torch.diag(torch.diagonal(Ainv))
    assert (K @ Ainv - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max().item() < 1e-8
    yy = torch.from_numpy(y).cuda()
    assert ((K @ eng.d_alpha - yy).abs().max() / yy.abs().max()).item() < 1e-9
    g1 = eng.d_grad.clone()
    eng.gradient(); torch.cuda.synchronize()
    assert torch.equal(g1, eng.d_grad)
    B = Ainv - torch.outer(eng.d_alpha, eng.d_alpha)
    i = eng.P // 2
    def gram(s):
        f = eng.d_feat[s]
        if ktype == "fidelity":
            c = torch.view_as_complex(f.reshape(n, -1, 2).contiguous())
            return (c @ c.conj().T).abs() ** 2
        dist = torch.cdist(f, f)
        k = dist * 3 ** 0.5
        return (1 + k) * torch.exp(-k)
    dK = (gram(1 + 2 * i) - gram(2 + 2 * i)) / (2 * eng.h)
    ref = 0.5 * (B * dK).sum().item()
    assert abs(g1[i].item() - ref) < 1e-6 * max(1.0, abs(ref))


def test_full_size_properties_config4_shard(d):
    """BASELINE.json configs[3] shard size (n = 8192, yz_cx q=8 L=3 projected-gaussian): size-independent
    properties — K symmetric with unit diagonal in [0,1]; A * A^-1 = I; alpha solves A alpha = y; the fused
    gradient is linear in B (doubling alpha^T alpha part) and reproducible bit-for-bit run to run."""
    n, q, layers, dd = 8192, 8, 3, 4
    x, y = d.synthetic_dataset(n, dd, "yz_cx")
    eng = d.AgentEngine(x, y, encoding_type="yz_cx", kernel_type="projected", num_qubits=q, num_layers=layers, noise_std=0.1,
                        rho=100.0, L=100.0)
    z = d.kernels.dev_f64(np.round(np.random.RandomState(42).rand(eng.P), 4))
    eng.simulate(z); eng.gram()
    K = eng.solver.matrix().clone()          # the training Gram fills the lower tiles only (all the factorisation reads)
    K = torch.tril(K) + torch.tril(K, -1).T
    f0 = eng.d_feat[0]
    full = torch.empty((n, n), dtype=torch.float64, device="cuda")
    hyp = (C.c_double * 1)(1.0)
    assert d.load().dqgp_gram_projected(0, hyp, f0.data_ptr(), n, f0.data_ptr(), n, 3 * q, full.data_ptr(), n, 1,
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    assert torch.equal(full, full.T)         # mirrored mode: exactly symmetric
    assert torch.equal(torch.tril(full, -1), torch.tril(K, -1))
    assert torch.all(torch.diagonal(K) == 1.0 + 0.1 ** 2)
    off = K - torch.diag(torch.diagonal(K))
    assert off.min() >= 0.0 and off.max() <= 1.0
    eng.factor(); eng.gradient()
    torch.cuda.synchronize()
    eng.check_info()
    Ainv = torch.tril(eng.solver.inverse()); Ainv = Ainv + Ainv.T - torch.diag(torch.diagonal(Ainv))
    R = K @ Ainv - torch.eye(n, dtype=torch.float64, device="cuda")
    assert R.abs().max().item() < 1e-8
    yy = torch.from_numpy(y).cuda()
    assert ((K @ eng.d_alpha - yy).abs().max() / yy.abs().max()).item() < 1e-9
    g1 = eng.d_grad.clone()
    eng.gradient(); torch.cuda.synchronize()
    assert torch.equal(g1, eng.d_grad)                                   # deterministic reduction
    # independent check of the fused gradient for two parameters with materialised Grams (torch fp64)
    B = Ainv - torch.outer(eng.d_alpha, eng.d_alpha)
    F = eng.d_feat
    for i in (0, eng.P - 1):
        def gram(s):
            f = F[s]
            d2 = torch.cdist(f, f).pow(2)
            return torch.exp(-d2)
        dK = (gram(1 + 2 * i) - gram(2 + 2 * i)) / (2 * eng.h)
        ref = 0.5 * (B * dK).sum().item()
        assert abs(g1[i].item() - ref) < 1e-6 * max(1.0, abs(ref))


@pytest.mark.parametrize("enc,ktype,q,dd,layers,n", [("yz_cx", "projected", 4, 2, 2, 150), ("kyriienko", "projected", 3, 2, 1, 200),
                                                     ("chebyshev", "projected", 3, 2, 1, 180), ("hubregtsen", "fidelity", 5, 2, 2, 160),
                                                     ("yz_cx", "fidelity", 2, 2, 1, 120), ("chebyshev", "fidelity", 3, 2, 1, 100)])
def test_analytic_gradient_is_the_derivative_of_the_oracle_nll(d, enc, ktype, q, dd, layers, n):
    """AgentEngine(gradient="analytic") (opt-in): the gradient equals the derivative of the NLL computed by
    the ORACLE (NumPy kernel + LAPACK), by central differences of the NLL itself; and the reference's h = pi/8 rule, which the
    default mode reproduces, is visibly a different number."""
    from oracle import qkernels
    x, y = d.synthetic_dataset(n, dd, enc)
    z = np.round(np.random.RandomState(5).uniform(0.2, 2.9, d.EncodingCircuit(enc, q, dd, layers).num_parameters), 4)
    grads = {}
    for mode in ("analytic", "central_difference"):
        eng = d.AgentEngine(x, y, encoding_type=enc, kernel_type=ktype, num_qubits=q, num_layers=layers, noise_std=0.1, rho=100.0,
                            L=100.0, gradient=mode)
        dz = d.kernels.dev_f64(z)
        eng.simulate(dz); eng.gram(); eng.factor(); eng.gradient()
        torch.cuda.synchronize(); eng.check_info()
        grads[mode] = eng.d_grad.cpu().numpy().copy()
        nll_gpu = float(eng.d_nll[3].item())

    def nll(p):
        qk = qkernels.create_quantum_kernel(q, dd, layers, enc, ktype, "XYZ", "gaussian")
        qk.assign_parameters(p)
        a = qk.evaluate(x, x) + 0.01 * np.eye(n)
        L = np.linalg.cholesky(a)
        al = np.linalg.solve(a, y)
        return np.log(np.diag(L)).sum() + 0.5 * y @ al + 0.5 * n * np.log(2 * np.pi)

    assert abs(nll(z) - nll_gpu) < 1e-8 * abs(nll_gpu)
    h = 1e-5
    for i in range(0, len(z), max(1, len(z) // 5)):
        e = np.zeros_like(z); e[i] = h
        ref = (nll(z + e) - nll(z - e)) / (2 * h)
        assert abs(grads["analytic"][i] - ref) < 2e-5 * max(1.0, abs(ref)), (i, grads["analytic"][i], ref)
    assert np.max(np.abs(grads["analytic"] - grads["central_difference"])) > 1e-3 * np.max(np.abs(grads["analytic"]))


def test_riemannian_agent_analytic_mode_through_the_reference_api(d):
    """RiemannianAgent(..., gradient="analytic").train_and_update: same NLL as the default mode, gradient equal to the engine's
    analytic gradient, and a local update that follows from it by the reference's formulas."""
    from oracle import agent_step, torus
    x, y = d.synthetic_dataset(140, 2, "yz_cx")
    rs = np.random.RandomState(11)
    P = d.EncodingCircuit("yz_cx", 4, 2, 2).num_parameters
    z, psi = np.round(rs.uniform(0.2, 2.9, P), 4), np.round(rs.rand(P), 4)
    out = {}
    for mode in ("central_difference", "analytic"):
        ag = d.RiemannianAgent(f"m_{mode}", x, y, 4, 0.1, 100.0, 100.0, use_parameter_shift=True, num_layers=2, encoding_type="yz_cx",
                               kernel_type="projected", gradient=mode)
        theta, psi_new, nll, _, _ = ag.train_and_update(z, psi)
        out[mode] = (theta, psi_new, nll, ag.last_gradient.copy())
    assert abs(out["analytic"][2] - out["central_difference"][2]) < 1e-9 * abs(out["analytic"][2])
    g4 = np.round(out["analytic"][3], 4)
    th_ref, ps_ref = agent_step.local_update(torus.wrap(z), g4, psi, 100.0, 100.0)
    assert np.max(np.abs(out["analytic"][0] - th_ref)) < 1e-12 and np.max(np.abs(out["analytic"][1] - ps_ref)) < 1e-9
    assert np.max(np.abs(out["analytic"][3] - out["central_difference"][3])) > 1e-3


def test_agent_step_walks_the_lu_rung_when_cholesky_fails(d):
    """use_parameter_shift=False with a projected kernel is the reference's `_manual_projected_kernel_derivatives_riemannian`
    branch (agent_riemannian.py:397-400): central differences through self.q_kernel, i.e. with the REAL outer kernel.  With
    ExpSineSquared the Gram is indefinite, np.linalg.cholesky raises and the reference refactors with scipy's LU (:419-425).
    RiemannianAgent must do the same on the device and land on the oracle's theta / psi."""
    from oracle import agent_step
    x, y = d.synthetic_dataset(97, 4, "hubregtsen", seed=5)
    cfg = agent_step.KernelConfig("hubregtsen", "projected", 3, 2, "expsinesquared", training_ignores_outer_kernel=False)
    rs = np.random.RandomState(8)
    z, psi = np.round(rs.rand(12) * np.pi, 4), np.round(rs.rand(12), 4)
    K, dK = agent_step.kernel_and_derivatives(cfg, x, np.mod(z, np.pi), np.pi / 8)
    assert np.linalg.eigvalsh(K + 0.01 * np.eye(97)).min() < -0.1                # really indefinite
    ref = agent_step.train_and_update(cfg, x, y, z, psi, 0.1, 100.0, 100.0, want_cond=False)
    ag = d.RiemannianAgent("lu", x, y, 3, 0.1, 100.0, 100.0, use_parameter_shift=False, num_layers=2, encoding_type="hubregtsen",
                           kernel_type="projected", outer_kernel="expsinesquared")
    theta, psi_new, nll, _, comp = ag.train_and_update(z, psi)
    assert ag.used_lu_fallback
    assert np.max(np.abs(ag.last_gradient - ref.grad)) < 1e-8 * max(1.0, np.abs(ref.grad).max())
    assert np.max(np.abs(theta - ref.theta)) < 1e-12 and np.max(np.abs(psi_new - ref.psi)) < 1e-9
    assert abs(comp["quadratic_term"] - ref.components["quadratic_term"]) < 1e-8 * max(1.0, abs(ref.components["quadratic_term"]))
    if np.isfinite(ref.nll):
        assert abs(nll - ref.nll) < 1e-8 * max(1.0, abs(ref.nll))
    else:
        assert not np.isfinite(nll)                # negative determinant: the reference takes the log of a negative number


def test_admm_engine_repairs_failed_agents_with_the_lu_rung(d):
    """AdmmEngine.iteration + repair_failed_agents on an indefinite training kernel (ExpSineSquared honoured) equals the
    oracle's iteration, which walks the reference's LU rung inside gp_terms."""
    from oracle import agent_step, driver
    x, y = d.synthetic_dataset(2 * 80, 4, "hubregtsen", seed=3)
    shards = [(x[:80], y[:80]), (x[80:], y[80:])]
    cfg = agent_step.KernelConfig("hubregtsen", "projected", 3, 1, "expsinesquared", training_ignores_outer_kernel=False)
    rs = np.random.RandomState(2)
    theta0, psi0 = np.round(rs.rand(2, 6), 4), np.round(rs.rand(2, 6), 4)
    eng = d.AdmmEngine(shards, theta0, psi0, rho=100.0, L=100.0, encoding_type="hubregtsen", kernel_type="projected", num_qubits=3,
                       num_layers=1, noise_std=0.1, outer_kernel="expsinesquared", training_ignores_outer_kernel=False)
    th, ps = theta0, psi0
    for _ in range(2):
        eng.iteration()
        assert eng.repair_failed_agents() == 2
        z_ref, th, ps, _ = driver.admm_iteration(cfg, shards, th, ps, 0.1, 100.0, 100.0)
        z, theta, psi, _ = eng.state()
        assert np.max(np.abs(z - z_ref)) < 1e-12 and np.max(np.abs(theta - th)) < 1e-12 and np.max(np.abs(psi - ps)) < 1e-9


def test_agent_step_on_a_second_device_after_the_first(d):
    """Kernel attributes (dynamic shared memory above 48 KB) are per DEVICE: one process that touches two GPUs must be able to
    run the fused gradient / fidelity kernels on both (round-1 advisor finding: a process-wide `static bool` guard)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oracle import agent_step
    x, y = d.synthetic_dataset(130, 3, "yz_cx")
    rs = np.random.RandomState(1)
    res = {}
    for kt in ("projected", "fidelity"):
        P = d.EncodingCircuit("yz_cx", 4, 3, 2).num_parameters
        z, psi = np.round(rs.rand(P), 4), np.round(rs.rand(P), 4)
        ref = agent_step.train_and_update(agent_step.KernelConfig("yz_cx", kt, 4, 2), x, y, z, psi, 0.1, 100.0, 100.0, want_cond=False)
        for dev in (0, 1):
            with torch.cuda.device(dev):
                eng = d.AgentEngine(x, y, encoding_type="yz_cx", kernel_type=kt, num_qubits=4, num_layers=2, noise_std=0.1, rho=100.0, L=100.0)
                dz = d.kernels.dev_f64(z, device=f"cuda:{dev}")
                eng.simulate(dz); eng.gram(); eng.factor(); eng.gradient()
                torch.cuda.synchronize()
                g = eng.d_grad.cpu().numpy()
            assert np.max(np.abs(g - ref.grad)) < 1e-8 * max(1.0, np.abs(ref.grad).max()), (kt, dev)
    torch.cuda.set_device(0)


def test_predict_propagates_nan_like_numpy(d):
    """A chebyshev test point outside [-1, 1] gives a NaN row of K(test, train); np.maximum (main.py:1466) propagates the NaN into
    the predictive variance, and so must the device epilogue (CUDA's fmax alone would return the 1e-10 floor instead)."""
    g = load_golden("agent_step_cheb_proj_matern_q3.npz")
    xt = g["X_test"].copy()
    xt[3, 0] = 1.5                                   # arccos(1.5) = NaN
    mean, var, *_ = d.predict_quantum_gp(g["X"], g["Y"], xt, np.mod(g["z"], np.pi), int(g["q"]), int(g["layers"]), 0.1, True,
                                         str(g["encoding"]), str(g["kernel_type"]), "XYZ", str(g["outer_kernel"]))
    assert np.isnan(mean[3]) and np.isnan(var[3])
    ok = np.arange(len(mean)) != 3
    assert np.max(np.abs(mean[ok] - g["pred_mean"][ok])) < 1e-8 * max(1.0, np.abs(g["pred_mean"]).max())
    assert np.max(np.abs(var[ok] - g["pred_var"][ok])) < 1e-8
