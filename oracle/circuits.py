"""Gate programs of the four in-scope squlearn 0.9.1 encoding circuits  [UPSTREAM-RECALLED].

The reference only *constructs* these (``main.py:68-83``, ``agent_riemannian.py:51-66,331-338``) with
``num_qubits, num_features, num_layers`` and all other options at their defaults; the arithmetic lives
in squlearn, which is not available here.  Each builder below restates the published circuit as a flat
list of ``Gate`` records.  Conventions (Qiskit): little-endian (qubit 0 = least-significant index bit),
``RX(t)=exp(-i t X/2)``, ``RY``, ``RZ`` likewise, ``CRZ(t; c, t)=|0><0|(x)I + |1><1|(x)RZ(t)``, start |0..0>.
Parameter and feature indices wrap (``parameters[ioff % P]``, ``features[k % d]``).  ChebyshevPQC and YZ_CX carry a RUNNING
feature offset across layers (``features[feature_offset % d]``, one step per encoded qubit, never reset) as recalled from
squlearn 0.9.x; Hubregtsen restarts at feature 0 in every layer.  The two rules coincide whenever ``q % d == 0`` or there
is one layer — every BASELINE.json config and every golden fixture — so this choice is invisible to them.

Angle forms (``Gate.form``):
  "p"      angle = p[pidx]
  "x"      angle = x[fidx]
  "p+cx"   angle = p[pidx] + coef * x[fidx]
  "p*acos" angle = p[pidx] * arccos(x[fidx])
  "c*acos" angle = coef * arccos(x[fidx])
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

ENCODINGS = ("chebyshev", "hubregtsen", "yz_cx", "kyriienko")


@dataclass(frozen=True)
class Gate:
    name: str            # "h" | "rx" | "ry" | "rz" | "cx" | "crz"
    q0: int              # target (1-qubit gates) / control (cx, crz)
    q1: int = -1         # target of cx / crz
    form: str = ""       # angle form, "" for h / cx
    pidx: int = -1
    fidx: int = -1
    coef: float = 1.0

    def angle(self, x: np.ndarray, p: np.ndarray) -> np.ndarray:
        """Angle for every sample row of ``x`` (n, d) under parameter vector ``p`` (P,)."""
        n = x.shape[0]
        if self.form == "p":
            return np.full(n, p[self.pidx])
        if self.form == "x":
            return x[:, self.fidx].copy()
        if self.form == "p+cx":
            return p[self.pidx] + self.coef * x[:, self.fidx]
        if self.form == "p*acos":
            return p[self.pidx] * np.arccos(x[:, self.fidx])
        if self.form == "c*acos":
            return self.coef * np.arccos(x[:, self.fidx])
        raise ValueError(f"gate {self.name} has no angle")


def num_parameters(encoding: str, q: int, layers: int) -> int:
    if encoding == "chebyshev":
        # squlearn ChebyshevPQC.num_parameters (closed=True): 2q basis-change + q per layer for the
        # feature map + entanglers per layer (q if q>2, 1 if q==2, 0 if q==1).
        ent = q if q > 2 else (1 if q == 2 else 0)
        return 2 * q + layers * q + layers * ent
    if encoding == "hubregtsen":
        return layers * q + (layers * q if q > 2 else 0)
    if encoding == "yz_cx":
        return 2 * q * layers
    if encoding == "kyriienko":
        return 3 * q * layers
    raise ValueError(f"Unknown encoding type: {encoding}")


def build_circuit(encoding: str, q: int, d: int, layers: int) -> List[Gate]:
    P = num_parameters(encoding, q, layers)
    gates: List[Gate] = []
    ioff = 0

    def take() -> int:
        nonlocal ioff
        k = ioff % P
        ioff += 1
        return k

    if encoding == "chebyshev":
        # ChebyshevPQC(closed=True, entangling_gate="crz", nonlinearity="arccos"); alpha only bounds
        # the initial parameters, which the reference overwrites (agent_riemannian.py:114).
        for i in range(q):
            gates.append(Gate("ry", i, form="p", pidx=take()))
        foff = 0
        for _ in range(layers):
            for i in range(q):
                gates.append(Gate("rx", i, form="p*acos", pidx=take(), fidx=foff % d))
                foff += 1
            closed = 1
            for i in range(0, q + closed - 1, 2):
                if q >= 2:
                    gates.append(Gate("crz", i, (i + 1) % q, form="p", pidx=take()))
            if q > 2:
                for i in range(1, q + closed - 1, 2):
                    gates.append(Gate("crz", i, (i + 1) % q, form="p", pidx=take()))
        for i in range(q):
            gates.append(Gate("ry", i, form="p", pidx=take()))
    elif encoding == "hubregtsen":
        # HubregtsenEncodingCircuit(closed=True, final_encoding=False)
        for i in range(q):
            gates.append(Gate("h", i))
        loops = int(np.ceil(d / q))
        for _ in range(layers):
            for i in range(loops * q):
                name = "rz" if (i // q) % 2 == 0 else "rx"
                gates.append(Gate(name, i % q, form="x", fidx=i % d))
            for i in range(q):
                gates.append(Gate("ry", i, form="p", pidx=take()))
            if q > 2:
                for i in range(q):  # closed ring
                    gates.append(Gate("crz", i, (i + 1) % q, form="p", pidx=take()))
    elif encoding == "yz_cx":
        # YZ_CX_EncodingCircuit(c=1.0)
        foff = 0
        for layer in range(layers):
            for i in range(q):
                gates.append(Gate("ry", i, form="p+cx", pidx=take(), fidx=foff % d, coef=1.0))
                gates.append(Gate("rz", i, form="p+cx", pidx=take(), fidx=foff % d, coef=1.0))
                foff += 1
            start = 0 if layer % 2 == 0 else 1
            for i in range(start, q - 1, 2):
                gates.append(Gate("cx", i, i + 1))
    elif encoding == "kyriienko":
        # OUR definition: squlearn 0.9.1's KyriienkoEncodingCircuit does not accept
        # ``num_layers`` so the reference call has no upstream behaviour.  L x [Chebyshev-tower RY
        # encoding; HEA block RZ RX RZ per qubit; CX ladder on even pairs then odd pairs].
        for _ in range(layers):
            for i in range(q):
                gates.append(Gate("ry", i, form="c*acos", fidx=i % d, coef=2.0 * (i + 1)))
            for i in range(q):
                gates.append(Gate("rz", i, form="p", pidx=take()))
                gates.append(Gate("rx", i, form="p", pidx=take()))
                gates.append(Gate("rz", i, form="p", pidx=take()))
            for i in range(0, q - 1, 2):
                gates.append(Gate("cx", i, i + 1))
            for i in range(1, q - 1, 2):
                gates.append(Gate("cx", i, i + 1))
    else:
        raise ValueError(f"Unknown encoding type: {encoding}")
    assert ioff == P or P == 0, (encoding, q, layers, ioff, P)
    return gates
