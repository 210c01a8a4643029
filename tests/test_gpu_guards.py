"""Out-of-bounds WRITE check without compute-sanitizer (it is closed on the GPU pool): every device buffer the host layer
allocates for the kernels is placed inside a larger allocation whose 512-byte margins hold a sentinel; after the calls (ragged
sizes on purpose: nothing is a multiple of a tile) the margins must be untouched.  Results are still checked against the
oracle, so a kernel cannot pass by writing nothing.  Reads past the end and writes further than 512 bytes away are not seen."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

GUARD_BYTES = 512
SENTINEL = {torch.float64: -7.25e77, torch.int32: 0x5A5A5A5A}


@pytest.fixture(scope="module")
def d():
    import dqgp_b200
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box")
    return dqgp_b200


class Guards:
    def __init__(self):
        self.buffers = []          # (whole allocation, guard elements, payload elements, shape)

    def allocate(self, real_full, shape, dtype, device, fill=None):
        numel = int(np.prod(shape)) if len(shape) else 1
        g = GUARD_BYTES // torch.empty((), dtype=dtype).element_size()
        whole = real_full((numel + 2 * g,), SENTINEL[dtype], dtype=dtype, device=device)
        view = whole[g:g + numel].view(shape)
        if fill is not None:
            view.fill_(fill)
        self.buffers.append((whole, g, numel, tuple(shape)))
        return view

    def check(self):
        torch.cuda.synchronize()
        assert len(self.buffers) > 0
        for whole, g, numel, shape in self.buffers:
            s = SENTINEL[whole.dtype]
            assert bool((whole[:g] == s).all()), f"write BEFORE a buffer of shape {shape}"
            assert bool((whole[g + numel:] == s).all()), f"write AFTER a buffer of shape {shape}"
        return len(self.buffers)


@pytest.fixture
def guards(monkeypatch):
    real_empty, real_zeros, real_full = torch.empty, torch.zeros, torch.full
    gd = Guards()

    def shape_of(args):
        if len(args) == 1 and isinstance(args[0], (tuple, list, torch.Size)):
            return tuple(int(v) for v in args[0])
        return tuple(int(v) for v in args)

    def wrap(real, fill):
        def alloc(*args, **kw):
            dev, dtype = kw.get("device"), kw.get("dtype", torch.float32)
            on_gpu = dev is not None and torch.device(dev).type == "cuda"
            if not on_gpu or dtype not in SENTINEL or set(kw) - {"device", "dtype"}:
                return real(*args, **kw)
            return gd.allocate(real_full, shape_of(args), dtype, dev, fill)
        return alloc

    monkeypatch.setattr(torch, "empty", wrap(real_empty, None))
    monkeypatch.setattr(torch, "zeros", wrap(real_zeros, 0))
    return gd


def test_guard_fixture_sees_a_stray_write(guards):
    """The checker itself: a write one element past the end is reported."""
    buf = torch.empty((3, 5), dtype=torch.float64, device="cuda")
    assert buf.shape == (3, 5) and guards.check() == 1
    whole, g, numel, _ = guards.buffers[0]
    whole[g + numel] = 0.0
    with pytest.raises(AssertionError, match="write AFTER"):
        guards.check()


@pytest.mark.parametrize("enc,q,dd,layers", [("chebyshev", 3, 2, 1), ("hubregtsen", 5, 2, 2), ("yz_cx", 8, 4, 3), ("kyriienko", 10, 6, 2),
                                            ("yz_cx", 11, 3, 1)])
def test_simulator_outputs_stay_inside_their_buffers(d, guards, enc, q, dd, layers):
    from oracle import circuits, statevector
    rng = np.random.default_rng(q)
    n, S = 37, 3
    lo, hi = (-0.99, 0.99) if enc in ("chebyshev", "kyriienko") else (-2, 2)
    x = rng.uniform(lo, hi, (n, dd))
    pm = rng.uniform(0, np.pi, (S, circuits.num_parameters(enc, q, layers)))
    ec = d.EncodingCircuit(enc, q, dd, layers)
    dx, dpm = d.kernels.dev_f64(x), d.kernels.dev_f64(pm)
    F, Psi = ec.features(dx, dpm), ec.states(dx, dpm)
    assert guards.check() >= 2
    ref = statevector.simulate(circuits.build_circuit(enc, q, dd, layers), q, x, pm[S - 1])
    assert np.abs(F[S - 1].cpu().numpy() - statevector.pauli_features(ref, q)).max() < 1e-12
    got = Psi[S - 1].cpu().numpy()
    assert np.abs(got[..., 0] + 1j * got[..., 1] - ref).max() < 1e-12


@pytest.mark.parametrize("enc,ktype,q,layers,dd,n,outer,honour", [("chebyshev", "projected", 3, 1, 2, 97, "gaussian", False),
                                                                 ("hubregtsen", "fidelity", 5, 2, 2, 113, "gaussian", False),
                                                                 ("kyriienko", "projected", 4, 2, 6, 131, "matern", True),
                                                                 ("yz_cx", "projected", 9, 1, 3, 201, "gaussian", False),
                                                                 ("yz_cx", "projected", 6, 2, 4, 259, "expsinesquared", True)])
def test_agent_step_stays_inside_its_buffers(d, guards, enc, ktype, q, layers, dd, n, outer, honour):
    """Simulation of the 2P+1 sets, Gram, factorisation (Cholesky or the LU rung), fused gradient, NLL, local update - with the
    engine's feature / parameter-set / gradient / workspace buffers fenced."""
    from oracle import agent_step, circuits, driver
    x, y = driver.synthetic_dataset(n, dd, enc)
    P = circuits.num_parameters(enc, q, layers)
    rs = np.random.RandomState(5)
    z, psi = np.round(rs.rand(P), 4), np.round(rs.rand(P), 4)
    cfg = agent_step.KernelConfig(enc, ktype, q, layers, outer, training_ignores_outer_kernel=not honour)
    ref = agent_step.train_and_update(cfg, x, y, z, psi, 0.1, 100.0, 100.0, workers=1, want_cond=False)
    ag = d.RiemannianAgent("g", x, y, q, 0.1, 100.0, 100.0, use_parameter_shift=True, num_layers=layers, encoding_type=enc,
                           kernel_type=ktype, outer_kernel=outer, training_ignores_outer_kernel=not honour)
    for _ in range(3):                                           # eager, graph capture, graph replay
        theta, psi_new, nll, _, _ = ag.train_and_update(z, psi)
    assert guards.check() >= 6
    near_tie = np.abs(np.abs(ref.grad * 1e4 - np.floor(ref.grad * 1e4)) - 0.5) < 1e-5
    assert np.max(np.abs(ag.last_gradient - ref.grad)) < 1e-8 * max(1.0, np.abs(ref.grad).max())
    assert np.max(np.abs(theta - ref.theta)[~near_tie]) < 1e-12
    assert (np.isnan(nll) and np.isnan(ref.nll)) or abs(nll - ref.nll) < 1e-8 * max(1.0, abs(ref.nll))


def test_rectangular_grams_and_prediction_stay_inside_their_buffers(d, guards):
    from oracle import agent_step, driver
    x, y = driver.synthetic_dataset(150, 2, "hubregtsen")
    xt = x[:23] + 0.01
    for ktype, outer in (("projected", "matern"), ("fidelity", "gaussian")):
        qk = d.create_quantum_kernel(4, 2, 2, True, "hubregtsen", ktype, outer_kernel=outer)
        p = np.round(np.random.RandomState(1).rand(qk.encoding_circuit.num_parameters), 4)
        qk.assign_parameters(p)
        K = qk.evaluate(xt[:13], x[:29])
        cfg = agent_step.KernelConfig("hubregtsen", ktype, 4, 2, outer, training_ignores_outer_kernel=False)
        ref = cfg.make(2, training=False)
        ref._parameters = p
        assert np.max(np.abs(K - ref.evaluate(xt[:13], x[:29]))) < 1e-10
        mean, var = d.predict_quantum_gp(x, y, xt, p, 4, 2, 0.1, True, "hubregtsen", ktype, "XYZ", outer)[:2]
        rm, rv = driver.predict(cfg, x, y, xt, p, 0.1)[:2]
        assert np.max(np.abs(mean - rm)) < 1e-8 * max(1.0, np.abs(rm).max()) and np.max(np.abs(var - rv)) < 1e-8
    assert guards.check() >= 6


def test_admm_iterations_stay_inside_their_buffers(d, guards):
    """Three agents with unequal shards on one GPU: consensus, per-agent streams, the (A, 2, P) row table."""
    from oracle import driver
    x, y = driver.synthetic_dataset(333, 3, "yz_cx")
    shards = [(x[:101], y[:101]), (x[101:230], y[101:230]), (x[230:], y[230:])]
    kw = dict(encoding_type="yz_cx", kernel_type="projected", num_qubits=4, num_layers=2, noise_std=0.1)
    P = d.EncodingCircuit("yz_cx", 4, 3, 2).num_parameters
    rs = np.random.RandomState(42)
    th0, ps0 = np.round(rs.rand(3, P), 4), np.round(rs.rand(3, P), 4)
    eng = d.AdmmEngine(shards, th0, ps0, rho=100.0, L=100.0, **kw)
    for _ in range(3):
        eng.iteration()
    z, theta, psi, nll = eng.state()
    assert guards.check() >= 10
    from oracle import agent_step
    cfg = agent_step.KernelConfig("yz_cx", "projected", 4, 2, "gaussian")
    t, p = th0.copy(), ps0.copy()
    for _ in range(3):
        zr, t, p, res = driver.admm_iteration(cfg, shards, t, p, 0.1, 100.0, 100.0)
    nl = np.array([r.nll for r in res])
    assert np.array_equal(z, zr) and np.max(np.abs(theta - t)) < 1e-12 and np.max(np.abs(nll - nl)) < 1e-8 * np.abs(nl).max()
