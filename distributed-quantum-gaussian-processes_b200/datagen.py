"""Quantum-GP dataset generation with the reference's signature and RNG stream
(``main.generate_quantum_gp_data``, main.py:161-292): ground-truth parameters from ``param_seed``, inputs from
``data_seed``, Y = chol(K + 1e-6 I) z + noise.  SURVEY §8(f) row 2 ("next"): the N x N Gram and its Cholesky run on
the GPU (K is written straight into the solver; only the factor is computed), all random draws stay on the host
with NumPy's legacy global stream so the data are the reference's for the same seeds.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import _lib
from ._lib import check
from .engine import Solver
from .kernels import create_quantum_kernel, dev_f64, stream_ptr


def generate_quantum_gp_data(num_samples, input_dim, num_qubits, num_layers=2, data_range=(-2.0, 2.0), noise_std=0.1,
                             use_parameter_shift=True, kernel_params=None, encoding_type="yz_cx", kernel_type="fidelity",
                             measurement="XYZ", outer_kernel="gaussian", outer_kernel_params=None, regularization=None,
                             data_seed=None, param_seed=42, lean=None):
    """-> (X, Y, ground_truth_params).  ``lean`` (None = decide from free HBM): factor the N x N Gram in place in one padded
    square (N = 131072 is 137 GB) instead of the three-square workspace."""
    if input_dim < 1 or input_dim > 6:
        raise ValueError(f"Input dimension must be between 1 and 6, got {input_dim}")
    lib = _lib.load()
    qk = create_quantum_kernel(num_qubits, input_dim, num_layers, use_parameter_shift, encoding_type, kernel_type, measurement,
                               outer_kernel, outer_kernel_params, regularization)
    n_par = qk.num_parameters if qk.num_parameters is not None else qk.encoding_circuit.num_parameters
    if kernel_params is not None:
        kernel_params = np.asarray(kernel_params, dtype=np.float64)
        if len(kernel_params) != n_par:
            raise ValueError(f"Expected {n_par} parameters, got {len(kernel_params)}")
        qk.assign_parameters(kernel_params)
        truth = np.round(kernel_params.copy(), 4)
    else:
        np.random.seed(param_seed)                                     # main.py:210-212
        truth = np.round(np.random.uniform(0, np.pi, n_par), 4)
        qk.assign_parameters(truth)
    if data_seed is None:
        data_seed = int(time.time() * 1000) % 2 ** 32                  # main.py:216-218 (not reproducible by design)
    np.random.seed(data_seed)
    X = np.random.uniform(data_range[0], data_range[1], size=(num_samples, input_dim))
    if encoding_type == "chebyshev":
        X = np.clip(X, -0.99, 0.99)                                    # main.py:225-236
    n = num_samples
    d_x = dev_f64(X)
    d_p = dev_f64(np.asarray(qk.parameters, dtype=np.float64).reshape(1, -1))
    if lean is None:                                                   # three padded squares do not fit: factor in place
        npad = -(-n // 128) * 128
        lean = 3 * 8 * npad * npad > 0.9 * torch.cuda.mem_get_info()[0]
    solver = Solver(n, lean=bool(lean))
    st = stream_ptr()
    qk.evaluate_device(d_x, d_x, d_p, same=2, out=solver.matrix(), ld=solver.ld)     # lower tiles: all the factorisation reads
    # a NaN feature (arccos outside [-1, 1]) makes the whole row and column of that sample NaN, diagonal included
    if bool(torch.isnan(torch.diagonal(solver.matrix())).any()):
        raise ValueError("Kernel matrix contains NaN or infinite values")   # main.py:248-249 (arccos outside [-1, 1])
    check(lib.dqgp_add_diagonal(solver.matrix_ptr, n, solver.ld, 1e-6, st), "add diagonal")
    logdet = torch.zeros(1, dtype=torch.float64, device=d_x.device)
    info = torch.zeros(1, dtype=torch.int32, device=d_x.device)
    check(lib.dqgp_potrf_solve_inv(solver.handle, None, None, logdet.data_ptr(), info.data_ptr(), -1, st), "potrf")
    z = np.random.normal(0, 1, n)                                      # main.py:273
    d_z = dev_f64(z)
    d_y = torch.empty(n, dtype=torch.float64, device=d_x.device)
    check(lib.dqgp_solver_apply_factor(solver.handle, d_z.data_ptr(), d_y.data_ptr(), st), "apply factor")
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError("Cholesky of K + 1e-6 I failed (the reference falls back to an eigendecomposition, "
                                    "main.py:279-287; not on the GPU path)")
    Y = d_y.cpu().numpy()
    Y += np.random.normal(0, noise_std, n)                             # main.py:277
    return X, Y, truth
