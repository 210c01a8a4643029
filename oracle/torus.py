"""Oracle restatement of the torus manifold and the ADMM update formulas
(``riemannian_optimizer.py:26-51`` circular mean, ``:73-129`` manifold maps, ``:302-368`` ADMM).
Checked bit-for-bit against the real module (imported unmodified) by ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import numpy as np

PERIOD = np.pi    # every parameter lives on a circle of period pi (Q6; riemannian_optimizer.py:61)


def wrap(x, period=PERIOD):
    """riemannian_optimizer.py:73-83."""
    return np.mod(x, period)


def circular_mean(angles, period=PERIOD):
    """riemannian_optimizer.py:26-51: per column, direction of the summed unit vectors."""
    ang = 2 * np.pi * np.asarray(angles) / period
    c = np.sum(np.cos(ang), axis=0)
    s = np.sum(np.sin(ang), axis=0)
    return np.mod(np.arctan2(s, c) * period / (2 * np.pi), period)


def torus_distance(x, y, period=PERIOD):
    """riemannian_optimizer.py:89-105 (shortest-arc norm; used only for residual reporting)."""
    d = np.asarray(x) - np.asarray(y)
    return np.linalg.norm(np.mod(d + period / 2, period) - period / 2)


def update_z(theta, psi, rho, period=PERIOD):
    """riemannian_optimizer.py:302-322."""
    return circular_mean(theta + psi / rho, period=period)


def update_theta(z, grad, psi, rho, lipschitz, period=PERIOD):
    """riemannian_optimizer.py:324-348: one closed-form step from z, retracted (= wrapped)."""
    step = -(grad + psi) / (rho + lipschitz)
    return wrap(z + step, period)


def update_psi(psi, theta, z, rho, period=PERIOD):
    """riemannian_optimizer.py:350-368: log_map(z, theta) = (theta - z) mod period, NOT the signed arc (Q7)."""
    return psi + rho * wrap(theta - z, period)
