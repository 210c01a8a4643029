"""Oracle restatement of one agent's step: the (2P+1)-evaluation Gram + central-difference tensor
(``agent_riemannian.py:209-277``, worker ``:33-123``) and the linear algebra / gradient / NLL / local
ADMM update of ``RiemannianAgent.train_and_update`` (``agent_riemannian.py:376-491``).

Pinned against the real ``RiemannianAgent`` (run unmodified over ``oracle.fake_squlearn``) by the golden
vectors in ``tests/golden/agent_step_*.npz``.
"""
from __future__ import annotations

from concurrent.futures import ProcessPoolExecutor
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import qkernels, torus


@dataclass
class KernelConfig:
    """What the reference ships to its shifted-kernel workers (dict built at agent_riemannian.py:231-239)."""
    encoding_type: str = "yz_cx"
    kernel_type: str = "fidelity"
    num_qubits: int = 4
    num_layers: int = 2
    outer_kernel: str = "gaussian"
    # Q1: the worker dict carries no 'outer_kernel' key, so training Grams always use the Gaussian
    # outer kernel whatever --outer-kernel says.  True reproduces that; False honours `outer_kernel`.
    training_ignores_outer_kernel: bool = True

    def training_outer_kernel(self) -> str:
        return "gaussian" if self.training_ignores_outer_kernel else self.outer_kernel

    def make(self, num_features: int, training: bool):
        ok = self.training_outer_kernel() if training else self.outer_kernel
        return qkernels.create_quantum_kernel(self.num_qubits, num_features, self.num_layers,
                                              self.encoding_type, self.kernel_type, "XYZ", ok)


def shifted_parameter_sets(p, h, period=torus.PERIOD) -> np.ndarray:
    """(2P+1, P): [p, p+h e_0, p-h e_0, p+h e_1, ...], each wrapped mod period
    (agent_riemannian.py:219 wraps p, :245-256 builds the list, the worker wraps again at :41)."""
    p = torus.wrap(np.asarray(p, dtype=np.float64), period)
    rows = [p.copy()]
    for i in range(p.size):
        up = p.copy()
        up[i] += h
        dn = p.copy()
        dn[i] -= h
        rows += [up, dn]
    return torus.wrap(np.stack(rows), period)


def _gram_job(args):
    cfg, x, params = args
    k = cfg.make(x.shape[1], training=True)
    k._parameters = params
    return k.evaluate(x, x)


def _noop(_):
    return None


def process_pool(workers: Optional[int]):
    """The reference's process pool of shifted-kernel workers (agent_riemannian.py:261), started with "spawn" instead of the
    platform's fork: with OpenBLAS 0.3.30 a threaded LAPACK call in the PARENT after a fork pool can deadlock (seen with
    np.linalg.solve at n = 113).  The workers are started before the pool is returned, so a timed region around the map
    excludes the interpreter start-up that fork would not pay."""
    import multiprocessing
    pool = ProcessPoolExecutor(max_workers=workers, mp_context=multiprocessing.get_context("spawn"))
    list(pool.map(_noop, range(pool._max_workers)))
    return pool


def kernel_and_derivatives(cfg: KernelConfig, x, p, h, workers: Optional[int] = 1, subset=None):
    """K (n,n) and dK (P,n,n) by central differences, dK_i = (K(p+h e_i) - K(p-h e_i)) / (2h)
    (agent_riemannian.py:270-275, Q3).  ``workers`` > 1 mirrors the reference's nested process pool."""
    sets = shifted_parameter_sets(p, h)
    if subset is not None:
        sets = sets[list(subset)]
    jobs = [(cfg, x, s) for s in sets]
    if workers is not None and workers == 1:
        grams = [_gram_job(j) for j in jobs]
    else:
        with process_pool(workers) as pool:
            grams = list(pool.map(_gram_job, jobs))
    if subset is not None:
        return grams
    n_p = (len(grams) - 1) // 2
    dk = np.zeros((n_p, x.shape[0], x.shape[0]))
    for i in range(n_p):
        dk[i] = (grams[1 + 2 * i] - grams[2 + 2 * i]) / (2.0 * h)
    return grams[0], dk


@dataclass
class StepResult:
    theta: np.ndarray
    psi: np.ndarray
    nll: float
    cond: float
    components: dict
    grad: np.ndarray = field(default=None)          # unrounded dL/dtheta, for diagnostics
    K: np.ndarray = field(default=None)


def gp_terms(c, dk, y, noise_std, want_cond=True):
    """agent_riemannian.py:410-460: noise, cond(C), Cholesky, alpha, explicit inverse, gradient, NLL."""
    n = c.shape[0]
    c_noise = c + noise_std ** 2 * np.eye(n)
    cond = float(np.linalg.cond(c)) if want_cond else float("nan")
    try:
        chol = np.linalg.cholesky(c_noise)
        alpha = np.linalg.solve(chol.T, np.linalg.solve(chol, y))
        c_inv = np.linalg.solve(chol.T, np.linalg.solve(chol, np.eye(n)))
    except np.linalg.LinAlgError:
        try:
            from scipy.linalg import lu_factor, lu_solve
            lu = lu_factor(c_noise)
            alpha = lu_solve(lu, y)
            c_inv = lu_solve(lu, np.eye(n))
        except np.linalg.LinAlgError:
            c_inv = np.linalg.pinv(c_noise)
            alpha = c_inv @ y
    bracket = c_inv - np.outer(alpha, alpha)
    grad = 0.5 * np.array([np.sum(bracket * dk[i].T) for i in range(dk.shape[0])])
    sign, log_det = np.linalg.slogdet(c_noise)
    if sign <= 0:
        log_det = np.log(np.linalg.det(c_noise + 1e-8 * np.eye(n)))
    comp = {
        "log_det_term": float(0.5 * log_det),
        "quadratic_term": float(0.5 * (y.T @ alpha)),
        "constant_term": float(0.5 * len(y) * np.log(2 * np.pi)),
    }
    comp["total"] = float(comp["log_det_term"] + comp["quadratic_term"] + comp["constant_term"])
    return grad, comp, cond, alpha, c_inv


def local_update(z_wrapped, grad4, psi, rho, lipschitz):
    """agent_riemannian.py:479-486: theta from z, psi from the UNROUNDED theta, then both to 4 dp."""
    theta = torus.update_theta(z_wrapped, grad4, psi, rho, lipschitz)
    psi_new = torus.update_psi(psi, theta, z_wrapped, rho)
    return np.round(theta, 4), np.round(psi_new, 4)


def train_and_update(cfg: KernelConfig, x, y, z, psi, noise_std, rho, lipschitz, h=np.pi / 8,
                     workers: Optional[int] = 1, want_cond=True, keep_k=False) -> StepResult:
    z_m = torus.wrap(np.asarray(z, dtype=np.float64))
    c, dk = kernel_and_derivatives(cfg, x, z_m, h, workers)
    grad, comp, cond, _, _ = gp_terms(c, dk, y, noise_std, want_cond)
    grad4 = np.round(grad, 4)
    theta, psi_new = local_update(z_m, grad4, np.asarray(psi, dtype=np.float64), rho, lipschitz)
    return StepResult(theta, psi_new, comp["total"], cond, comp, grad, c if keep_k else None)
