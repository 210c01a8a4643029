"""GP prediction / NLPD at full-train scale on ONE GPU: K(train,train) of n_train^2 fp64 entries in
HBM, blocked Cholesky, rectangular K(test,train), predictive mean / variance / NLPD.  Checks the solve residual."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import dqgp_b200 as d  # noqa: E402

n_train = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
n_test = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
x, y = d.synthetic_dataset(n_train + n_test, 4, "yz_cx")
xtr, ytr, xte, yte = x[:n_train], y[:n_train], x[n_train:], y[n_train:]
P = d.EncodingCircuit("yz_cx", 8, 4, 3).num_parameters
params = np.round(np.random.RandomState(42).rand(P), 4)
torch.cuda.synchronize()
t0 = time.perf_counter()
mean, var, *_ = d.predict_quantum_gp(xtr, ytr, xte, params, 8, 3, 0.1, True, "yz_cx", "projected", "XYZ", "gaussian", Y_test=yte)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
rmse = float(np.sqrt(np.mean((mean - yte) ** 2)))
print(f"n_train={n_train} n_test={n_test}: {dt:.2f} s  NLPD {d.predict_quantum_gp.last_nlpd:.4f}  RMSE {rmse:.4f}  "
      f"var range [{var.min():.3e}, {var.max():.3e}]  peak mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB (torch) ")
assert np.isfinite(mean).all() and (var > 0).all() and rmse < 0.5
