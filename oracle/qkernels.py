"""Oracle stand-ins for the squlearn objects the reference builds in ``create_quantum_kernel``
(``main.py:43-145``) and in the shifted-kernel worker (``agent_riemannian.py:33-123``).

They expose exactly the surface the reference touches: ``evaluate(x, y)``, ``assign_parameters(p)``, a
writable ``_parameters``, ``num_parameters`` (``None`` before first use for the projected kernel, see
``main.py:198-199``), ``encoding_circuit.num_parameters`` and ``executor``.  [UPSTREAM-RECALLED] for the
squlearn semantics; the outer kernels restate scikit-learn 1.9.0's arithmetic
(``sklearn/gaussian_process/kernels.py`` ``RBF.__call__``, ``Matern.__call__``, ``ExpSineSquared.__call__``)
and are checked against the installed classes in ``tests/test_oracle_kernels.py``.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.spatial.distance import cdist

from . import circuits, statevector

OUTER_KERNELS = ("gaussian", "matern", "expsinesquared")


class EncodingCircuit:
    def __init__(self, encoding: str, num_qubits: int, num_features: int, num_layers: int):
        self.encoding = encoding
        self.num_qubits = int(num_qubits)
        self.num_features = int(num_features)
        self.num_layers = int(num_layers)
        self.gates = circuits.build_circuit(encoding, self.num_qubits, self.num_features, self.num_layers)
        self.num_parameters = circuits.num_parameters(encoding, self.num_qubits, self.num_layers)

    def states(self, x, p):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x.reshape(-1, self.num_features)
        return statevector.simulate(self.gates, self.num_qubits, x, np.asarray(p, dtype=np.float64))

    def features(self, x, p):
        return statevector.pauli_features(self.states(x, p), self.num_qubits)


class Executor:
    """Name holder only: both squlearn executors are exact, shot-free simulators."""

    def __init__(self, name: str = "statevector_simulator"):
        self.name = name

    def __repr__(self):
        return f"Executor({self.name})"


def outer_kernel_matrix(name: str, f: np.ndarray, g: np.ndarray, gamma: float = 1.0,
                        length_scale: float = 1.0, nu: float = 1.5, periodicity: float = 1.0) -> np.ndarray:
    """Outer kernel on feature matrices f (n1, m), g (n2, m) with scikit-learn's operation order.
    Defaults are the only values the reference can reach (Q2): gamma=1; Matern l=1, nu=1.5; ESS l=1, p=1."""
    if name == "gaussian":
        ls = 1.0 / math.sqrt(2.0 * gamma)      # squlearn wraps sklearn RBF(length_scale=1/sqrt(2 gamma))
        d2 = cdist(f / ls, g / ls, metric="sqeuclidean")
        return np.exp(-0.5 * d2)
    if name == "matern":
        if nu != 1.5:
            raise NotImplementedError("only nu=1.5 is reachable from the reference (Q2)")
        d = cdist(f / length_scale, g / length_scale, metric="euclidean")
        k = d * math.sqrt(3)
        return (1.0 + k) * np.exp(-k)
    if name == "expsinesquared":
        d = cdist(f, g, metric="euclidean")
        arg = np.pi * d / periodicity
        s = np.sin(arg)
        return np.exp(-2 * (s / length_scale) ** 2)
    raise ValueError(f"outer kernel {name!r} is outside the hot path (SURVEY §2 #13)")


class _KernelBase:
    def __init__(self, encoding_circuit: EncodingCircuit, executor=None, parameter_seed=0):
        self.encoding_circuit = encoding_circuit
        self.executor = executor if executor is not None else Executor()
        rng = np.random.default_rng(parameter_seed)
        # squlearn seeds *initial* parameters; the reference always overwrites them before evaluating.
        self._parameters = rng.uniform(-np.pi, np.pi, encoding_circuit.num_parameters)

    def assign_parameters(self, p):
        p = np.asarray(p, dtype=np.float64)
        if p.shape != (self.encoding_circuit.num_parameters,):
            raise ValueError(f"expected {self.encoding_circuit.num_parameters} parameters, got {p.shape}")
        self._parameters = p.copy()

    @property
    def parameters(self):
        return self._parameters


class FidelityKernel(_KernelBase):
    """K[j,k] = |<psi(y_k;p)|psi(x_j;p)>|^2  (use_expectation / evaluate_duplicates="all" do not change
    the exact-simulation value; ``main.py:118-124``)."""

    def __init__(self, encoding_circuit, executor=None, parameter_seed=0, use_expectation=True,
                 evaluate_duplicates="all", **_):
        super().__init__(encoding_circuit, executor, parameter_seed)
        self.num_parameters = encoding_circuit.num_parameters

    def evaluate(self, x, y=None):
        sx = self.encoding_circuit.states(x, self._parameters)
        sy = sx if y is None else self.encoding_circuit.states(y, self._parameters)
        ov = sx @ sy.conj().T
        return ov.real ** 2 + ov.imag ** 2


class ProjectedQuantumKernel(_KernelBase):
    """Outer kernel on the XYZ Pauli-expectation features (``main.py:130-137``)."""

    def __init__(self, encoding_circuit, measurement="XYZ", outer_kernel="gaussian", executor=None,
                 parameter_seed=0, regularization=None, **outer_kernel_params):
        super().__init__(encoding_circuit, executor, parameter_seed)
        if measurement != "XYZ":
            raise NotImplementedError("only measurement='XYZ' is on the hot path")
        if regularization is not None:
            raise NotImplementedError("regularization is outside the hot path (SURVEY §2 #13)")
        self.outer_kernel = str(outer_kernel).lower()
        self.outer_kernel_params = dict(outer_kernel_params)
        self.num_parameters = None          # squlearn quirk mirrored: unknown until first evaluation

    def evaluate(self, x, y=None):
        fx = self.encoding_circuit.features(x, self._parameters)
        fy = fx if y is None else self.encoding_circuit.features(y, self._parameters)
        self.num_parameters = self.encoding_circuit.num_parameters
        return outer_kernel_matrix(self.outer_kernel, fx, fy, **self.outer_kernel_params)


def create_quantum_kernel(num_qubits, num_features=1, num_layers=2, encoding_type="yz_cx",
                          kernel_type="fidelity", measurement="XYZ", outer_kernel="gaussian"):
    """Factory with the argument meaning of ``main.create_quantum_kernel`` (``main.py:43-145``); note Q2:
    outer-kernel hyper-parameters are never forwarded by the reference, so none are accepted here."""
    if encoding_type not in circuits.ENCODINGS:
        raise ValueError(f"Unknown encoding type: {encoding_type}")
    enc = EncodingCircuit(encoding_type, num_qubits, num_features, num_layers)
    if kernel_type == "fidelity":
        return FidelityKernel(enc)
    if kernel_type == "projected":
        return ProjectedQuantumKernel(enc, measurement=measurement, outer_kernel=outer_kernel)
    raise ValueError(f"Unknown kernel type: {kernel_type}")
