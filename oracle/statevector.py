"""Exact statevector simulation of the oracle gate programs + Pauli-XYZ features.

Two independent implementations:
  * ``simulate``        gate-by-gate application on a batch of states, vectorised over samples;
  * ``simulate_dense``  per sample, the full 2^q x 2^q unitary built from Kronecker products
                        (slow, obviously right) — used by tests to pin ``simulate``.
Conventions as in ``oracle.circuits`` (Qiskit little-endian).  Replaces what squlearn's
``Executor("statevector_simulator")`` / ``Executor("pennylane")`` do behind
``q_kernel.evaluate`` (call site ``agent_riemannian.py:118``).  [UPSTREAM-RECALLED]
"""
from __future__ import annotations

from typing import List

import numpy as np

from .circuits import Gate

_SQ = 1.0 / np.sqrt(2.0)


def _u2(name: str, theta: np.ndarray) -> np.ndarray:
    """(n, 2, 2) complex matrices of a 1-qubit rotation for a vector of angles."""
    c = np.cos(theta / 2.0)
    s = np.sin(theta / 2.0)
    u = np.zeros(theta.shape + (2, 2), dtype=np.complex128)
    if name == "rx":
        u[..., 0, 0] = c
        u[..., 0, 1] = -1j * s
        u[..., 1, 0] = -1j * s
        u[..., 1, 1] = c
    elif name == "ry":
        u[..., 0, 0] = c
        u[..., 0, 1] = -s
        u[..., 1, 0] = s
        u[..., 1, 1] = c
    elif name == "rz":
        u[..., 0, 0] = c - 1j * s
        u[..., 1, 1] = c + 1j * s
    else:
        raise ValueError(name)
    return u


def simulate(gates: List[Gate], q: int, x: np.ndarray, p: np.ndarray) -> np.ndarray:
    """States U(x_j; p)|0..0> for every row x_j.  Returns (n, 2^q) complex128."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    n = x.shape[0]
    dim = 1 << q
    psi = np.zeros((n, dim), dtype=np.complex128)
    psi[:, 0] = 1.0
    for g in gates:
        lo = 1 << g.q0
        if g.name == "h":
            v = psi.reshape(n, dim // (2 * lo), 2, lo)
            a, b = v[:, :, 0, :].copy(), v[:, :, 1, :].copy()
            v[:, :, 0, :] = (a + b) * _SQ
            v[:, :, 1, :] = (a - b) * _SQ
        elif g.name in ("rx", "ry", "rz"):
            u = _u2(g.name, g.angle(x, p))
            v = psi.reshape(n, dim // (2 * lo), 2, lo)
            a, b = v[:, :, 0, :].copy(), v[:, :, 1, :].copy()
            v[:, :, 0, :] = u[:, 0, 0, None, None] * a + u[:, 0, 1, None, None] * b
            v[:, :, 1, :] = u[:, 1, 0, None, None] * a + u[:, 1, 1, None, None] * b
        elif g.name in ("cx", "crz"):
            idx = np.arange(dim)
            cbit = (idx >> g.q0) & 1
            tbit = (idx >> g.q1) & 1
            if g.name == "cx":
                src = np.where(cbit == 1, idx ^ (1 << g.q1), idx)
                psi = psi[:, src]
            else:
                th = g.angle(x, p)[:, None] / 2.0
                phase = np.where(tbit[None, :] == 1, np.exp(1j * th), np.exp(-1j * th))
                psi = np.where(cbit[None, :] == 1, psi * phase, psi)
        else:
            raise ValueError(g.name)
    return psi


# ---- independent dense implementation -------------------------------------------------------------

_I2 = np.eye(2, dtype=np.complex128)
_P0 = np.array([[1, 0], [0, 0]], dtype=np.complex128)
_P1 = np.array([[0, 0], [0, 1]], dtype=np.complex128)
_PAULI = {
    "X": np.array([[0, 1], [1, 0]], dtype=np.complex128),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    "Z": np.array([[1, 0], [0, -1]], dtype=np.complex128),
}
_H = np.array([[1, 1], [1, -1]], dtype=np.complex128) * _SQ


def _embed(ops: dict, q: int) -> np.ndarray:
    """kron over qubits q-1 .. 0 (qubit 0 is the right-most / least-significant factor)."""
    out = np.array([[1.0 + 0j]])
    for k in range(q - 1, -1, -1):
        out = np.kron(out, ops.get(k, _I2))
    return out


def _rot_dense(name: str, theta: float) -> np.ndarray:
    from scipy.linalg import expm

    return expm(-0.5j * theta * _PAULI[name[1].upper()])


def unitary_dense(gates: List[Gate], q: int, x_row: np.ndarray, p: np.ndarray) -> np.ndarray:
    u = np.eye(1 << q, dtype=np.complex128)
    xr = np.asarray(x_row, dtype=np.float64)[None, :]
    for g in gates:
        if g.name == "h":
            m = _embed({g.q0: _H}, q)
        elif g.name in ("rx", "ry", "rz"):
            m = _embed({g.q0: _rot_dense(g.name, float(g.angle(xr, p)[0]))}, q)
        elif g.name == "cx":
            m = _embed({g.q0: _P0}, q) + _embed({g.q0: _P1, g.q1: _PAULI["X"]}, q)
        elif g.name == "crz":
            m = _embed({g.q0: _P0}, q) + _embed({g.q0: _P1, g.q1: _rot_dense("rz", float(g.angle(xr, p)[0]))}, q)
        else:
            raise ValueError(g.name)
        u = m @ u
    return u


def simulate_dense(gates: List[Gate], q: int, x: np.ndarray, p: np.ndarray) -> np.ndarray:
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    return np.stack([unitary_dense(gates, q, row, p)[:, 0] for row in x])


def pauli_features_dense(psi: np.ndarray, q: int) -> np.ndarray:
    out = np.zeros((psi.shape[0], 3 * q))
    for a, name in enumerate("XYZ"):
        for k in range(q):
            op = _embed({k: _PAULI[name]}, q)
            out[:, a * q + k] = np.real(np.einsum("ni,ij,nj->n", psi.conj(), op, psi))
    return out


# ---- features -----------------------------------------------------------------------------------------

def pauli_features(psi: np.ndarray, q: int) -> np.ndarray:
    """[<X_0>..<X_{q-1}>, <Y_0>.., <Z_0>..] for every state row (measurement="XYZ" ordering of
    squlearn's ProjectedQuantumKernel, ``main.py:130-137``).  Returns (n, 3q) float64."""
    n, dim = psi.shape
    f = np.zeros((n, 3 * q))
    for k in range(q):
        lo = 1 << k
        v = psi.reshape(n, dim // (2 * lo), 2, lo)
        a, b = v[:, :, 0, :], v[:, :, 1, :]
        cross = np.sum(np.conj(a) * b, axis=(1, 2))
        f[:, k] = 2.0 * cross.real
        f[:, q + k] = 2.0 * cross.imag
        f[:, 2 * q + k] = np.sum(np.abs(a) ** 2, axis=(1, 2)) - np.sum(np.abs(b) ** 2, axis=(1, 2))
    return f
