"""Host-side torus manifold and ADMM update formulas with the reference's names and semantics
(``riemannian_optimizer.py``: ``circular_mean`` :26-51, ``TorusManifold`` :53-146, ``RiemannianADMM``
:285-399, ``create_riemannian_framework`` :402-428).  These are O(P) NumPy operations of the driver
layer; the device-resident equivalents used inside the GPU iteration are
``dqgp_admm_local`` / ``dqgp_admm_consensus`` (csrc/admm.cu).  ``RiemannianOptimizer`` is dead code in the
reference (never stepped, Q8); a stub keeps the factory's 3-tuple shape.
"""
from __future__ import annotations

import numpy as np


def circular_mean(angles, period=np.pi):
    """Per-column direction of the summed unit vectors, mapped back to [0, period)."""
    phase = 2 * np.pi * np.asarray(angles) / period
    s = np.sum(np.sin(phase), axis=0)
    c = np.sum(np.cos(phase), axis=0)
    return np.mod(np.arctan2(s, c) * period / (2 * np.pi), period)


class TorusManifold:
    """S^1 x ... x S^1 with every coordinate of period ``period`` (pi in the reference, Q6)."""

    def __init__(self, dimension, period=np.pi):
        self.dim = dimension
        self.period = period
        self.name = f"Torus S^1 x ... x S^1 ({dimension}D, period={period:.3f})"

    def wrap_to_manifold(self, x):
        return np.mod(x, self.period)

    def random_point(self):
        return np.random.uniform(0, self.period, self.dim)

    def distance(self, x, y):
        d = np.asarray(x) - np.asarray(y)
        return np.linalg.norm(np.mod(d + self.period / 2, self.period) - self.period / 2)

    def exp_map(self, x, v):
        return self.wrap_to_manifold(x + v)

    def log_map(self, x, y):
        # the reference returns (y - x) mod period in [0, period), not the signed shortest arc (Q7)
        return self.wrap_to_manifold(y - x)

    def retraction(self, x, v):
        return self.exp_map(x, v)

    def vector_transport(self, x, v, d):
        return v

    def riemannian_gradient(self, x, euclidean_grad):
        return euclidean_grad


class RiemannianOptimizer:
    """Placeholder for the reference's optimizer object: it is constructed but never stepped there
    (Q8), so only the constructor signature is kept."""

    def __init__(self, manifold, learning_rate=0.015, method="gradient_descent", beta=0.9,
                 gradient_clip_norm=1.0, max_step_size=0.08):
        self.manifold, self.lr, self.method, self.beta = manifold, learning_rate, method, beta
        self.gradient_clip_norm, self.max_step_size = gradient_clip_norm, max_step_size

    def step(self, x, grad):
        raise NotImplementedError("RiemannianOptimizer.step is unreachable in the reference (SURVEY Q8) "
                                  "and outside the hot path")


class RiemannianADMM:
    def __init__(self, manifold, rho=1.0):
        self.manifold = manifold
        self.rho = rho
        self.iteration = 0

    def update_z(self, theta, psi):
        return circular_mean(theta + psi / self.rho, period=self.manifold.period)

    def update_theta(self, z, grad, psi, L, optimizer=None):
        return self.manifold.retraction(z, -(grad + psi) / (self.rho + L))

    def update_psi(self, psi, theta, z):
        return psi + self.rho * self.manifold.log_map(z, theta)

    def compute_primal_residual(self, theta, z):
        return np.linalg.norm([self.manifold.distance(t, z) for t in theta])

    def compute_dual_residual(self, z_new, z_old):
        return self.manifold.distance(z_new, z_old)


def create_riemannian_framework(num_parameters, learning_rate=0.01, rho=1.0, method="gradient_descent",
                                gradient_clip_norm=1.0, max_step_size=0.1):
    manifold = TorusManifold(num_parameters)
    optimizer = RiemannianOptimizer(manifold, learning_rate, method, gradient_clip_norm=gradient_clip_norm,
                                    max_step_size=max_step_size)
    return manifold, optimizer, RiemannianADMM(manifold, rho)
