"""CPU-side checks: the C-ABI library loads and exports every symbol include/dqgp.h declares (no compute
calls), the library's gate programs equal the oracle's, the host mirror equals the reference's formulas,
and the product fails loudly without a GPU instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


@pytest.fixture(scope="module")
def d():
    import __graft_entry__
    __graft_entry__.build()
    import dqgp_b200
    return dqgp_b200


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dqgp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dqgp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(d):
    names = _declared_symbols()
    assert len(names) >= 30
    lib = ctypes.CDLL(d._lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dqgp.h but not exported by libdqgp.so"
    assert sorted(d._lib.SIGNATURES) == names, "ctypes signature table and header drifted"
    assert d.load().dqgp_version() == 100


def test_library_is_sm100a_dmma_code(d):
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", d._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "DMMA.8x8x4" in sass            # fp64 tensor path in the factorisation
    assert "LDGSTS" in sass                # per-thread async staging (the GEMM's fallback ring, small-tile kernel)
    assert "UTMALDG.3D" in sass            # tensor-map (TMA) staging of the fused gradients' feature / state tiles
    assert "UTMALDG.2D" in sass            # ... and of the GEMM's operand tiles
    assert "UBLKCP" in sass                # per-row bulk copies: their fallback, and the Gram kernels' staging
    assert "SYNCS.PHASECHK.TRANS64.TRYWAIT" in sass      # ... completed on mbarriers


@pytest.mark.parametrize("enc,q,dd,layers", [("chebyshev", 3, 2, 1), ("chebyshev", 4, 2, 3), ("chebyshev", 2, 1, 2),
                                            ("hubregtsen", 5, 2, 2), ("hubregtsen", 3, 7, 2), ("hubregtsen", 2, 2, 1),
                                            ("yz_cx", 8, 4, 3), ("yz_cx", 5, 2, 4), ("yz_cx", 1, 1, 1),
                                            ("kyriienko", 10, 6, 4), ("kyriienko", 3, 2, 1)])
def test_library_gate_program_equals_oracle(d, enc, q, dd, layers):
    from oracle import circuits
    kinds = {"h": 0, "rx": 1, "ry": 2, "rz": 3, "cx": 4, "crz": 5}
    forms = {"": 0, "p": 1, "x": 2, "p+cx": 3, "p*acos": 4, "c*acos": 5}
    ec = d.EncodingCircuit(enc, q, dd, layers)       # host-side program; no device needed
    got = ec.describe()
    ref = circuits.build_circuit(enc, q, dd, layers)
    assert ec.num_parameters == circuits.num_parameters(enc, q, layers)
    assert len(got) == len(ref)
    for (kind, q0, q1, form, pidx, fidx, coef), g in zip(got, ref):
        assert (kind, q0, q1, form, pidx, fidx) == (kinds[g.name], g.q0, g.q1, forms[g.form], g.pidx, g.fidx)
        if g.form in ("p+cx", "c*acos"):
            assert coef == g.coef


def test_circuit_argument_errors(d):
    with pytest.raises(ValueError, match="Unknown encoding type"):
        d.EncodingCircuit("layered", 3, 2, 1)
    with pytest.raises(d.DqgpError):
        d.EncodingCircuit("yz_cx", 40, 2, 1)


def test_host_mirror_equals_reference_formulas(d):
    g = load_golden("torus.npz")
    for c in range(4):
        theta, psi, grad, rho = g[f"c{c}_theta"], g[f"c{c}_psi"], g[f"c{c}_grad"], float(g[f"c{c}_rho"])
        man, opt, admm = d.create_riemannian_framework(theta.shape[1], rho=rho)
        assert np.array_equal(admm.update_z(theta, psi), g[f"c{c}_z"])
        assert np.array_equal(d.circular_mean(theta), g[f"c{c}_circmean"])
        zr = man.wrap_to_manifold(np.round(g[f"c{c}_z"], 4))
        th = admm.update_theta(zr, grad, psi[0], 100.0, opt)
        assert np.array_equal(th, g[f"c{c}_theta_new"])
        assert np.array_equal(admm.update_psi(psi[0], th, zr), g[f"c{c}_psi_new"])
        assert np.array_equal(man.log_map(theta[0], theta[1]), g[f"c{c}_logmap"])
        assert man.distance(theta[0], theta[1]) == float(g[f"c{c}_dist"])
    with pytest.raises(NotImplementedError):
        opt.step(theta[0], grad)


def test_no_cpu_fallback(d):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(d.DqgpError, match="no CPU fallback"):
        d.create_quantum_kernel(3, 2, 1, True, "yz_cx", "projected")
    x = np.zeros((4, 2))
    with pytest.raises(d.DqgpError):
        d.RiemannianAgent("a", x, np.zeros(4), 3, 0.1, 100, 100, encoding_type="yz_cx", kernel_type="projected",
                          num_layers=1).train_and_update(np.zeros(6), np.zeros(6))
    with pytest.raises(d.DqgpError):
        d.predict_quantum_gp(x, np.zeros(4), x, np.zeros(6), 3, 1, 0.1, encoding_type="yz_cx", kernel_type="projected")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "distributed-quantum-gaussian-processes_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{fn} imports the oracle"


def test_synthetic_dataset_identical_to_oracle(d):
    from oracle import driver
    for enc in ("yz_cx", "chebyshev"):
        a, b = d.synthetic_dataset(50, 3, enc), driver.synthetic_dataset(50, 3, enc)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def _python_plan(enc, q, dd, layers):
    """Independent Python port of the planner's greedy pass grouping (csrc/circuit.cu: build_plan): gate indices per pass."""
    from oracle import circuits
    gates = circuits.build_circuit(enc, q, dd, layers)
    done, passes, bmax = [False] * len(gates), [], min(q, 3)
    while not all(done):
        blk, blocked, members = [], 0, []
        for g, gt in enumerate(gates):
            if done[g]:
                continue
            two = gt.name in ("cx", "crz")
            tgt = gt.q1 if two else gt.q0
            qs = (1 << tgt) | ((1 << gt.q0) if two else 0)
            if (qs & blocked) == 0 and (tgt in blk or len(blk) < bmax):
                if tgt not in blk:
                    blk.append(tgt)
                members.append(g)
                done[g] = True
            else:
                blocked |= qs
        passes.append((sorted(blk), members))
    return gates, passes


@pytest.mark.parametrize("enc,q,dd,layers", [("kyriienko", 10, 6, 4), ("yz_cx", 8, 4, 3), ("chebyshev", 4, 2, 3), ("hubregtsen", 5, 2, 2)])
def test_plan_statistics_match_an_independent_port(d, enc, q, dd, layers):
    """Pass count, fused-op count and the executed-work counter of bench.py's statevector roofline
    (dqgp_circuit_shifted_u2_applications) against a Python port of the planner: host-side, no device needed."""
    ec = d.EncodingCircuit(enc, q, dd, layers)
    gates, passes = _python_plan(enc, q, dd, layers)
    lib = d.load()
    assert lib.dqgp_circuit_num_passes(ec.handle) == len(passes)
    # fused ops per pass: runs of 1-qubit gates on one qubit collapse into one 2x2 unitary until a 2-qubit gate touches the qubit
    cost, par_pass, uses, n_u2 = [], {}, {}, 0
    for ip, (blk, members) in enumerate(passes):
        open_q, u2, crz = set(), 0, 0
        for g in members:
            gt = gates[g]
            if gt.name in ("cx", "crz"):
                open_q.discard(gt.q1); open_q.discard(gt.q0)
                crz += gt.name == "crz"
            elif gt.q0 not in open_q:
                open_q.add(gt.q0); u2 += 1
            if gt.pidx >= 0:
                par_pass[gt.pidx] = (ip, gt.name == "crz")
                uses[gt.pidx] = uses.get(gt.pidx, 0) + 1
        cost.append(u2 + 0.375 * crz)
        n_u2 += u2
    n_ops = sum(sum(1 for g in m if gates[g].name in ("cx", "crz")) for _, m in passes) + n_u2
    assert lib.dqgp_circuit_num_fused_ops(ec.handle) == n_ops
    assert all(v == 1 for v in uses.values())
    total = 2.0 * sum(cost) + sum((2.0 if is_crz else 1.0) * sum(cost[ip:]) for ip, is_crz in par_pass.values())
    ref3 = _python_cx_free_plan(enc, q, dd, layers)
    if ref3 is not None:
        # circuits without CRZ run the CX-free plan: the counter follows THAT plan (one fused unitary per qubit of a pass)
        passes3 = ref3[0]
        gates3 = gates
        # pass of every parameter: replay the event construction to find which run (and so which pass) owns the parameter's gate
        run_of_gate, open_run, runs = {}, {}, []
        for g, gt in enumerate(gates3):
            if gt.name == "cx":
                open_run.pop(gt.q0, None); open_run.pop(gt.q1, None)
            else:
                if gt.q0 not in open_run:
                    open_run[gt.q0] = len(runs); runs.append(gt.q0)
                run_of_gate[g] = open_run[gt.q0]
        # runs are consumed per qubit in order: the k-th run of qubit a is the k-th time a appears in the passes
        seen, run_pass, counters = {}, {}, {}
        for ip, blk in enumerate(passes3):
            for a in blk:
                counters[a] = counters.get(a, 0) + 1
                seen[(a, counters[a])] = ip
        occ = {}
        for r, a in enumerate(runs):
            occ[a] = occ.get(a, 0) + 1
            run_pass[r] = seen[(a, occ[a])]
        cost3 = [len(blk) for blk in passes3]
        total = 2.0 * sum(cost3)
        for g, gt in enumerate(gates3):
            if gt.pidx >= 0:
                total += sum(cost3[run_pass[run_of_gate[g]]:])
    assert lib.dqgp_circuit_shifted_u2_applications(ec.handle) == int(total + 0.5)


def test_use_parameter_shift_selects_the_reference_branch(d):
    """agent_riemannian.py:383-404: True -> the (2P+1)-job workers (Gaussian training Grams, Q1); False + projected -> central
    differences through the agent's own kernel (REAL outer kernel); False + fidelity -> analytic derivatives.  Host logic only."""
    x, y = np.zeros((4, 2)), np.zeros(4)
    mk = lambda **kw: d.RiemannianAgent("a", x, y, 3, 0.1, 100.0, 100.0, num_layers=1, encoding_type="yz_cx", outer_kernel="matern", **kw)
    a = mk(use_parameter_shift=True, kernel_type="projected")
    assert a.training_ignores_outer_kernel is True and a.gradient == "central_difference"
    a = mk(use_parameter_shift=False, kernel_type="projected")
    assert a.training_ignores_outer_kernel is False and a.gradient == "central_difference"
    a = mk(use_parameter_shift=False, kernel_type="fidelity")
    assert a.gradient == "analytic"
    a = mk(use_parameter_shift=False, kernel_type="fidelity", gradient="central_difference", training_ignores_outer_kernel=True)
    assert a.gradient == "central_difference" and a.training_ignores_outer_kernel is True


def _python_cx_free_plan(enc, q, dd, layers):
    """Independent port of the CX-free planner (csrc/circuit.cu: build_plan3): list scheduling over fused 1-qubit runs and CX gates,
    a ready CX is absorbed into the index map at once (control's mask ^= target's mask), a pass takes the earliest <= 3 ready runs."""
    from oracle import circuits
    gates = circuits.build_circuit(enc, q, dd, layers)
    events, open_run = [], {}
    for gt in gates:
        if gt.name == "crz":
            return None
        if gt.name == "cx":
            open_run.pop(gt.q0, None); open_run.pop(gt.q1, None)
            events.append(("cx", gt.q0, gt.q1))
        elif gt.q0 not in open_run:
            open_run[gt.q0] = len(events)
            events.append(("u2", gt.q0, -1))
    done, mask, passes = [False] * len(events), [1 << k for k in range(q)], []
    qubits = lambda e: {events[e][1]} | ({events[e][2]} if events[e][0] == "cx" else set())
    ready = lambda e: all(done[f] or not (qubits(e) & qubits(f)) for f in range(e))
    while not all(done):
        progress = True
        while progress:
            progress = False
            for e, ev in enumerate(events):
                if not done[e] and ev[0] == "cx" and ready(e):
                    mask[ev[1]] ^= mask[ev[2]]; done[e] = True; progress = True
        blk = [e for e, ev in enumerate(events) if not done[e] and ev[0] == "u2" and ready(e)][:3]
        if not blk:
            break
        for e in blk:
            done[e] = True
        passes.append([events[e][1] for e in blk])
    return passes, mask


@pytest.mark.parametrize("enc,q,dd,layers,expect", [("kyriienko", 10, 6, 4, 14), ("yz_cx", 8, 4, 3, 8), ("yz_cx", 11, 5, 1, 4),
                                                   ("kyriienko", 12, 6, 4, 16), ("chebyshev", 4, 2, 3, 0), ("hubregtsen", 5, 2, 2, 0)])
def test_cx_free_plan_pass_count(d, enc, q, dd, layers, expect):
    ec = d.EncodingCircuit(enc, q, dd, layers)
    got = d.load().dqgp_circuit_num_passes_cx_free(ec.handle)
    assert got == expect
    ref = _python_cx_free_plan(enc, q, dd, layers)
    assert got == (0 if ref is None else len(ref[0]))
    if ref is not None:
        # the final index map is a bijection: its column masks are linearly independent over GF(2)
        basis = []
        for m in ref[1]:
            for b in basis:
                m = min(m, m ^ b)
            assert m != 0
            basis.append(m)
