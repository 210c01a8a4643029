"""Satisfy ``import squlearn...`` (and matplotlib) with oracle-backed stand-ins so that the REAL
reference modules under /root/reference can be imported and executed *unmodified* in this container.

Used only by ``tests/golden/make_golden.py`` (golden-vector generation) — never by the product, the GPU
tests or the bench (``/root/reference`` does not exist on the GPU box).  With this in place,
``agent_riemannian.RiemannianAgent.train_and_update``, ``main.predict_quantum_gp`` and ``main.main`` run
their own code line by line; only the arithmetic *behind* the squlearn boundary comes from ``oracle``.
"""
from __future__ import annotations

import sys
import types
from unittest import mock

from . import qkernels


def _circuit_class(encoding):
    class _Circuit(qkernels.EncodingCircuit):
        def __init__(self, num_qubits, num_features=1, num_layers=1, **kwargs):
            if kwargs:
                raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
            super().__init__(encoding, num_qubits, num_features, num_layers)
    _Circuit.__name__ = encoding
    return _Circuit


def _unsupported(name):
    class _Nope:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is outside the hot path (SURVEY §2 #12)")
    _Nope.__name__ = name
    return _Nope


def install():
    if "squlearn" in sys.modules and getattr(sys.modules["squlearn"], "__dqgp_fake__", False):
        return
    root = types.ModuleType("squlearn")
    root.__dqgp_fake__ = True
    enc = types.ModuleType("squlearn.encoding_circuit")
    enc.ChebyshevPQC = _circuit_class("chebyshev")
    enc.YZ_CX_EncodingCircuit = _circuit_class("yz_cx")
    enc.HubregtsenEncodingCircuit = _circuit_class("hubregtsen")
    enc.KyriienkoEncodingCircuit = _circuit_class("kyriienko")   # our definition, see oracle.circuits
    for other in ("MultiControlEncodingCircuit", "LayeredEncodingCircuit", "RandomEncodingCircuit",
                  "HighDimEncodingCircuit"):
        setattr(enc, other, _unsupported(other))
    ker = types.ModuleType("squlearn.kernel")
    ker.FidelityKernel = qkernels.FidelityKernel
    ker.ProjectedQuantumKernel = qkernels.ProjectedQuantumKernel
    util = types.ModuleType("squlearn.util")
    util.Executor = qkernels.Executor
    root.encoding_circuit, root.kernel, root.util = enc, ker, util
    sys.modules.update({"squlearn": root, "squlearn.encoding_circuit": enc, "squlearn.kernel": ker,
                        "squlearn.util": util})
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
