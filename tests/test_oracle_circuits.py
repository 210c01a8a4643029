"""Pins for the part of the oracle that could NOT be checked against squlearn (parity unpinned there):
two independent simulators, analytic known answers, invariants."""
import numpy as np
import pytest

from oracle import circuits, qkernels, statevector

CASES = [("chebyshev", 3, 2, 1), ("chebyshev", 4, 2, 3), ("chebyshev", 2, 1, 2), ("hubregtsen", 5, 2, 2),
         ("hubregtsen", 3, 5, 1), ("hubregtsen", 2, 2, 1), ("yz_cx", 4, 4, 3), ("yz_cx", 5, 2, 2), ("yz_cx", 1, 1, 2),
         ("kyriienko", 4, 6, 2), ("kyriienko", 3, 2, 1)]


@pytest.mark.parametrize("enc,q,d,layers", CASES)
def test_gate_simulator_matches_dense_unitaries(enc, q, d, layers):
    rng = np.random.default_rng(q + 10 * d)
    gates = circuits.build_circuit(enc, q, d, layers)
    P = circuits.num_parameters(enc, q, layers)
    x = rng.uniform(-0.95, 0.95, (6, d))
    p = rng.uniform(0, np.pi, P)
    a = statevector.simulate(gates, q, x, p)
    b = statevector.simulate_dense(gates, q, x, p)
    assert np.abs(a - b).max() < 1e-13
    assert np.abs(np.linalg.norm(a, axis=1) - 1).max() < 1e-13
    assert np.abs(statevector.pauli_features(a, q) - statevector.pauli_features_dense(b, q)).max() < 1e-13


def test_parameter_counts_match_survey_appendix_c():
    assert circuits.num_parameters("chebyshev", 3, 1) == 12 and len(circuits.build_circuit("chebyshev", 3, 2, 1)) == 12
    assert circuits.num_parameters("chebyshev", 4, 3) == 32
    assert circuits.num_parameters("hubregtsen", 5, 2) == 20 and len(circuits.build_circuit("hubregtsen", 5, 2, 2)) == 35
    assert circuits.num_parameters("yz_cx", 8, 3) == 48 and len(circuits.build_circuit("yz_cx", 8, 4, 3)) == 59
    assert circuits.num_parameters("kyriienko", 10, 4) == 120 and len(circuits.build_circuit("kyriienko", 10, 6, 4)) == 196
    for enc, q, d, layers in CASES:        # every parameter is used exactly once
        used = sorted(g.pidx for g in circuits.build_circuit(enc, q, d, layers) if g.pidx >= 0)
        assert used == list(range(circuits.num_parameters(enc, q, layers)))


def test_known_answer_yz_cx_one_qubit():
    """RZ(p1+x) RY(p0+x)|0>: <Z> = cos a, <X> = sin a cos b, <Y> = sin a sin b."""
    gates = circuits.build_circuit("yz_cx", 1, 1, 1)
    x = np.array([[0.3], [-1.1]])
    p = np.array([0.7, 1.9])
    f = statevector.pauli_features(statevector.simulate(gates, 1, x, p), 1)
    a, b = p[0] + x[:, 0], p[1] + x[:, 0]
    assert np.allclose(f[:, 0], np.sin(a) * np.cos(b), atol=1e-15)
    assert np.allclose(f[:, 1], np.sin(a) * np.sin(b), atol=1e-15)
    assert np.allclose(f[:, 2], np.cos(a), atol=1e-15)


def test_known_answer_chebyshev_one_qubit():
    """q=1: RY(p0) RX(p1 acos x) RY(p2) with no entangler; p0 = p2 = 0 gives <Z> = cos(p1 acos x) = T_{p1}(x)."""
    gates = circuits.build_circuit("chebyshev", 1, 1, 1)
    assert [g.name for g in gates] == ["ry", "rx", "ry"]
    x = np.array([[0.2], [0.9], [-0.5]])
    f = statevector.pauli_features(statevector.simulate(gates, 1, x, np.array([0.0, 2.0, 0.0])), 1)
    assert np.allclose(f[:, 2], 2 * x[:, 0] ** 2 - 1, atol=1e-15)         # Chebyshev T_2
    assert np.allclose(f[:, 1], -np.sin(2 * np.arccos(x[:, 0])), atol=1e-15)


def test_known_answer_hubregtsen_h_then_rz():
    """q=1, d=1: H, RZ(x), RY(p): after H RZ(x) the Bloch vector is (cos x, sin x, 0)."""
    gates = circuits.build_circuit("hubregtsen", 1, 1, 1)
    x = np.array([[0.4], [2.0]])
    f = statevector.pauli_features(statevector.simulate(gates, 1, x, np.array([0.0])), 1)
    assert np.allclose(f, np.stack([np.cos(x[:, 0]), np.sin(x[:, 0]), 0 * x[:, 0]], 1), atol=1e-15)


def test_known_answer_two_qubit_cx_and_crz():
    # YZ_CX q=2, one layer: product state then CX(0,1); with RY(pi) on qubit 0 -> |1>, CX flips qubit 1
    gates = circuits.build_circuit("yz_cx", 2, 1, 1)
    psi = statevector.simulate(gates, 2, np.zeros((1, 1)), np.array([np.pi, 0.0, 0.0, 0.0]))
    assert np.isclose(abs(psi[0, 3]), 1.0)                                  # |11>
    # CRZ acts only when the control is |1>: chebyshev q=2 with all-zero rotations leaves |00> untouched
    gates = circuits.build_circuit("chebyshev", 2, 1, 1)
    P = circuits.num_parameters("chebyshev", 2, 1)
    p = np.zeros(P); p[4] = 1.234                                           # the single CRZ angle
    psi = statevector.simulate(gates, 2, np.full((1, 1), 0.5), p)
    assert np.isclose(psi[0, 0], 1.0)


@pytest.mark.parametrize("ktype,outer", [("fidelity", "gaussian"), ("projected", "gaussian"), ("projected", "matern"),
                                         ("projected", "expsinesquared")])
def test_kernel_invariants(ktype, outer):
    rng = np.random.default_rng(2)
    k = qkernels.create_quantum_kernel(4, 3, 2, "hubregtsen", ktype, "XYZ", outer)
    k.assign_parameters(rng.uniform(0, np.pi, k.encoding_circuit.num_parameters))
    x = rng.uniform(-2, 2, (40, 3))
    K = k.evaluate(x, x)
    assert np.allclose(K, K.T, atol=1e-14)
    assert np.allclose(np.diag(K), 1.0, atol=1e-13)
    assert K.min() >= -1e-15 and K.max() <= 1 + 1e-13
    if outer != "expsinesquared":      # ExpSineSquared of a Euclidean distance is not PSD for m > 1 features
        assert np.linalg.eigvalsh(K).min() > -1e-12
    if ktype == "fidelity":
        s = k.encoding_circuit.states(x, k.parameters)
        assert np.allclose(K, np.abs(s @ s.conj().T) ** 2, atol=1e-14)


def test_projected_kernel_invariant_under_qubit_relabelling():
    rng = np.random.default_rng(3)
    q, d, layers = 3, 3, 1
    gates = circuits.build_circuit("yz_cx", q, d, layers)
    p = rng.uniform(0, np.pi, circuits.num_parameters("yz_cx", q, layers))
    x = rng.uniform(-1, 1, (9, d))
    f = statevector.pauli_features(statevector.simulate(gates, q, x, p), q)
    perm = np.array([2, 0, 1])
    fp = np.concatenate([f[:, perm], f[:, q + perm], f[:, 2 * q + perm]], axis=1)
    for outer in qkernels.OUTER_KERNELS:
        assert np.allclose(qkernels.outer_kernel_matrix(outer, f, f), qkernels.outer_kernel_matrix(outer, fp, fp), atol=1e-15)
