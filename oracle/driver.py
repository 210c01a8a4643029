"""Oracle restatement of the ADMM consensus loop (``main.py:2403-2555``: init, z-update, fan-out,
collect/round; stop rule ``:2719``), GP prediction (``main.py:1399-1466``) and NLPD (``:1546-1552``).
Excludes the per-iteration 5-fold CV (Q14), plotting and printing.
"""
from __future__ import annotations

import numpy as np

from . import agent_step, torus


def initial_state(n_agents, n_params, rho, rng=None):
    """main.py:2407-2408, 2460: theta, psi ~ round(U[0,1), 4) from the *global legacy* NumPy stream
    (Q12) unless ``rng`` (a RandomState) is given; z from the consensus update, rounded."""
    rand = np.random.rand if rng is None else rng.rand
    theta = np.round(rand(n_agents, n_params), 4)
    psi = np.round(rand(n_agents, n_params), 4)
    z = np.round(torus.update_z(theta, psi, rho), 4)
    return theta, psi, z


def admm_iteration(cfg, shards, theta, psi, noise_std, rho, lipschitz, h=np.pi / 8, workers=1, want_cond=False):
    """One pass of the ``while True`` body, main.py:2507-2555.  ``shards`` = [(X_a, Y_a)], ``lipschitz`` a
    scalar or per-agent list.  Returns (z, theta, psi, [StepResult])."""
    z = np.round(torus.update_z(theta, psi, rho), 4)
    theta = theta.copy()
    psi = psi.copy()
    out = []
    for a, (xa, ya) in enumerate(shards):
        la = lipschitz[a] if np.ndim(lipschitz) else lipschitz
        r = agent_step.train_and_update(cfg, xa, ya, z, psi[a], noise_std, rho, la, h, workers, want_cond)
        theta[a] = np.round(r.theta, 4)
        psi[a] = np.round(r.psi, 4)
        out.append(r)
    return z, theta, psi, out


def converged(z, theta, tol=1e-6):
    """main.py:2719: every agent within tol of the consensus (Euclidean norm)."""
    return all(np.linalg.norm(z - t) < tol for t in theta)


def predict(cfg, x_train, y_train, x_test, params, noise_std):
    """main.py:1399-1466: posterior mean / variance with K + sigma^2 I + 1e-6 I; note the prediction path
    honours the configured outer kernel (unlike training, Q1)."""
    k = cfg.make(x_train.shape[1], training=False)
    k.assign_parameters(np.asarray(params, dtype=np.float64))
    k_tt = k.evaluate(x_train, x_train)
    k_st = k.evaluate(x_test, x_train)
    k_ss = k.evaluate(x_test, x_test)
    a = k_tt + (noise_std ** 2) * np.eye(k_tt.shape[0])
    a += 1e-6 * np.eye(k_tt.shape[0])
    try:
        chol = np.linalg.cholesky(a)
        alpha = np.linalg.solve(chol, y_train)
        alpha = np.linalg.solve(chol.T, alpha)
        mean = k_st @ alpha
        v = np.linalg.solve(chol, k_st.T)
        var = np.maximum(np.diag(k_ss) - np.sum(v ** 2, axis=0), 1e-10)
    except np.linalg.LinAlgError:
        # main.py:1470-1486: direct inversion when the Cholesky fails (e.g. ExpSineSquared of a Euclidean
        # distance is not positive semi-definite in more than one dimension)
        a_inv = np.linalg.inv(a)
        alpha = a_inv @ y_train
        mean = k_st @ alpha
        var = np.maximum(np.diag(k_ss - k_st @ a_inv @ k_st.T), 1e-10)
    return mean, var


def nlpd(y_true, mean, var):
    """main.py:1546-1552 (same formula at :1657-1662)."""
    var = np.maximum(var, 1e-10)
    r = y_true - mean
    return float(np.mean(0.5 * np.log(2 * np.pi) + 0.5 * np.log(var) + 0.5 * (r ** 2 / var)))


def synthetic_dataset(n, d, encoding, seed=0):
    """SURVEY §8(d) / BASELINE.md §4 synthetic inputs (identical for GPU and CPU runs)."""
    rng = np.random.default_rng(seed)
    lo, hi = (-0.99, 0.99) if encoding in ("chebyshev", "kyriienko") else (-2.0, 2.0)
    x = rng.uniform(lo, hi, (n, d))
    y = np.sin(x.sum(axis=1)) + 0.1 * rng.standard_normal(n)
    return x, y
