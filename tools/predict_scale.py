"""GP prediction / NLPD at full-train scale on ONE GPU: K(train,train) of n_train^2 fp64 entries in
HBM, blocked Cholesky, rectangular K(test,train), predictive mean / variance / NLPD.
    python tools/predict_scale.py n_train n_test [cfg4|cfg5] [auto|lean|full] [--residual]
--residual re-factors and checks (K + sigma^2 I) alpha = y on 256 random rows."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import dqgp_b200 as d  # noqa: E402

n_train = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
n_test = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfg = sys.argv[3] if len(sys.argv) > 3 else "cfg4"          # cfg4: yz_cx q=8 L=3 d=4 gaussian; cfg5: kyriienko q=10 L=4 d=6 matern
lean = {"lean": True, "full": False}.get(sys.argv[4] if len(sys.argv) > 4 else "auto")
enc, q, layers, dim, outer = ("yz_cx", 8, 3, 4, "gaussian") if cfg == "cfg4" else ("kyriienko", 10, 4, 6, "matern")
x, y = d.synthetic_dataset(n_train + n_test, dim, enc)
xtr, ytr, xte, yte = x[:n_train], y[:n_train], x[n_train:], y[n_train:]
P = d.EncodingCircuit(enc, q, dim, layers).num_parameters
params = np.round(np.random.RandomState(42).rand(P), 4)
torch.cuda.synchronize()
t0 = time.perf_counter()
mean, var, *_ = d.predict_quantum_gp(xtr, ytr, xte, params, q, layers, 0.1, True, enc, "projected", "XYZ", outer, Y_test=yte, lean=lean)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"[{cfg} lean={lean}] ", end="")
rmse = float(np.sqrt(np.mean((mean - yte) ** 2)))
print(f"n_train={n_train} n_test={n_test}: {dt:.2f} s  NLPD {d.predict_quantum_gp.last_nlpd:.4f}  RMSE {rmse:.4f}  "
      f"var range [{var.min():.3e}, {var.max():.3e}]  peak mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB (torch) ")
assert np.isfinite(mean).all() and (var > 0).all() and rmse < (0.5 if cfg == "cfg4" else 1.0)

if "--residual" in sys.argv:
    # size-independent check of the factorisation + substitution at this scale: for 256 random training rows,
    # (K + sigma^2 I + 1e-6 I)[rows, :] alpha == y[rows], with the rows of K recomputed by a rectangular Gram
    import ctypes as C
    from dqgp_b200.engine import Solver
    from dqgp_b200.kernels import create_quantum_kernel, dev_f64, stream_ptr
    lib = d.load()
    del mean, var
    torch.cuda.empty_cache()
    qk = create_quantum_kernel(q, dim, layers, True, enc, "projected", "XYZ", outer)
    qk.assign_parameters(params)
    d_x, d_p, d_y = dev_f64(xtr), dev_f64(params.reshape(1, -1)), dev_f64(ytr)
    s = Solver(n_train, lean=True if lean is None else lean)
    qk.evaluate_device(d_x, d_x, d_p, same=2, out=s.matrix(), ld=s.ld)
    st = stream_ptr()
    lib.dqgp_add_diagonal(s.matrix_ptr, n_train, s.ld, 0.1 ** 2, st)
    lib.dqgp_add_diagonal(s.matrix_ptr, n_train, s.ld, 1e-6, st)
    alpha = torch.empty(n_train, dtype=torch.float64, device="cuda")
    logdet = torch.zeros(1, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    t0 = time.perf_counter()
    assert lib.dqgp_potrf_solve_inv(s.handle, d_y.data_ptr(), alpha.data_ptr(), logdet.data_ptr(), info.data_ptr(), 0, st) == 0
    torch.cuda.synchronize()
    t1 = time.perf_counter() - t0
    assert info.item() == 0
    rows = np.sort(np.random.default_rng(1).choice(n_train, 256, replace=False))
    k_rows = qk.evaluate_device(dev_f64(xtr[rows]), d_x, d_p, same=False)
    r = k_rows @ alpha + (0.1 ** 2 + 1e-6) * alpha[rows] - d_y[rows]
    res = float(r.abs().max() / d_y.abs().max())
    npad = -(-n_train // 128) * 128
    print(f"factor + solve {t1:.2f} s ({npad ** 3 / 3 / t1 / 1e12:.1f} TFLOP/s on n^3/3)  logdet {logdet.item():.6f}  "
          f"max |A alpha - y| / max|y| over 256 rows = {res:.2e}  solver bytes {s.bytes / 1e9:.1f} GB")
    assert res < 1e-9
