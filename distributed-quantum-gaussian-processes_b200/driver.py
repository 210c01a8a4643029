"""ADMM consensus driver with the control flow of the reference's training loop (``main.py:2403-2784``): initial z,
then per iteration z-update -> every agent's step -> collect/round -> optional k-fold CV of the consensus parameters
(NLPD, ``main.py:2650-2666``) -> stop on consensus (``main.py:2719,2768``), CV patience (``:2771``) or ``max_iter``
(``:2777``).  Printing, plotting and the ground-truth analysis of the reference are not reproduced.

The agents run on this process's GPU through ``AdmmEngine`` (device-resident theta, psi, z; with several ranks each
process passes its own shards and the shared initial state).  Host RNG: pass ``theta0`` / ``psi0`` explicitly, or let
them be drawn from NumPy's legacy global stream exactly where the reference draws them (``main.py:2407-2408``, Q12).
"""
from __future__ import annotations

import numpy as np

from .engine import AdmmEngine
from .predict import k_fold_cross_validation_consensus


def run_admm(shards, *, encoding_type, kernel_type, num_qubits, num_layers, noise_std=0.1, rho=100.0, L=100.0,
             outer_kernel="gaussian", shift_value=np.pi / 8, max_iter=100, tolerance=1e-6, theta0=None, psi0=None,
             n_agents_total=None, cv_data=None, cv_folds=5, cv_patience=50, seed=42, training_ignores_outer_kernel=True,
             process_group=None, rank=0, world_size=1, callback=None, gradient="central_difference"):
    """Returns a dict: z (final consensus; the best-CV z on early stop / max_iter as in the reference), iterations,
    history (per iteration: z, theta, psi, nll per local agent, cv), stop_reason.
    ``cv_data=(X_train, Y_train)`` enables the per-iteration CV of main.py:2650 (seed + iteration as fold seed).
    ``gradient="analytic"`` switches the agents to the exact NLL derivative (opt-in, not the reference's trajectories)."""
    import torch

    A = n_agents_total if n_agents_total is not None else len(shards) * world_size
    from .kernels import EncodingCircuit
    d = np.asarray(shards[0][0]).reshape(len(shards[0][0]), -1).shape[1]
    P = EncodingCircuit(encoding_type, num_qubits, d, num_layers).num_parameters
    if theta0 is None:
        theta0 = np.round(np.random.rand(A, P), 4)                       # main.py:2407
    if psi0 is None:
        psi0 = np.round(np.random.rand(A, P), 4)                         # main.py:2408
    eng = AdmmEngine(shards, np.asarray(theta0, dtype=np.float64), np.asarray(psi0, dtype=np.float64), rho=rho, L=L,
                     process_group=process_group, rank=rank, world_size=world_size, encoding_type=encoding_type,
                     kernel_type=kernel_type, num_qubits=num_qubits, num_layers=num_layers, noise_std=noise_std,
                     outer_kernel=outer_kernel, shift_value=shift_value, training_ignores_outer_kernel=training_ignores_outer_kernel,
                     gradient=gradient)
    history, z_best_cv, cv_best, patience = [], None, float("inf"), 0
    it, reason = 0, "max_iter"
    while True:
        it += 1
        eng.iteration()
        eng.repair_failed_agents()          # Cholesky failures walk the reference's LU ladder (agent_riemannian.py:419-425)
        z, theta, psi, nll = eng.state()
        rec = {"iteration": it, "z": z.copy(), "theta": theta.copy(), "psi": psi.copy(), "nll": nll.copy(), "cv": None}
        if cv_data is not None:
            cv = k_fold_cross_validation_consensus(cv_data[0], cv_data[1], z, num_qubits, num_layers, noise_std, k_folds=cv_folds,
                                                   encoding_type=encoding_type, kernel_type=kernel_type, outer_kernel=outer_kernel,
                                                   random_seed=seed + it, process_group=process_group, rank=rank,
                                                   world_size=world_size)     # folds dealt to the ranks, scores gathered
            rec["cv"] = cv
            if cv["mean_nlpd"] < cv_best:
                cv_best, z_best_cv, patience = cv["mean_nlpd"], z.copy(), 0
            else:
                patience += 1
        history.append(rec)
        if callback is not None:
            callback(rec)
        norms = np.linalg.norm(z - theta, axis=1)                        # main.py:2719-2720
        if np.all(norms < tolerance):
            reason = "consensus"
            break
        if cv_data is not None and patience >= cv_patience:
            reason, z = "cv_patience", z_best_cv.copy()
            break
        if it >= max_iter:
            reason = "max_iter"
            if z_best_cv is not None:
                z = z_best_cv.copy()
            break
    torch.cuda.synchronize()
    return {"z": z, "iterations": it, "history": history, "stop_reason": reason, "best_cv_nlpd": cv_best, "z_best_cv": z_best_cv}
