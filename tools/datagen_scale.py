"""Quantum-GP dataset generation (main.py:161-292: full N x N Gram + 1e-6 I, Cholesky, Y = L z + noise) at full scale on ONE
GPU with the in-place (lean) factorisation.    python tools/datagen_scale.py N [cfg4|cfg5]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dqgp_b200 as d  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg4"
enc, q, layers, dim, outer, rng = ("yz_cx", 8, 3, 4, "gaussian", (-2.0, 2.0)) if cfg == "cfg4" else ("kyriienko", 10, 4, 6, "matern", (-0.99, 0.99))
torch.cuda.synchronize()
t0 = time.perf_counter()
try:
    X, Y, truth = d.generate_quantum_gp_data(N, dim, q, layers, rng, 0.1, True, None, enc, "projected", "XYZ", outer, None, None,
                                             data_seed=7, param_seed=42)
    dt = time.perf_counter() - t0
    print(f"[{cfg}] N={N}: {dt:.2f} s  Y mean {Y.mean():+.4f} std {Y.std():.4f}  finite {np.isfinite(Y).all()}  "
          f"free HBM after {torch.cuda.mem_get_info()[0] / 1e9:.0f} GB")
except np.linalg.LinAlgError as exc:
    print(f"[{cfg}] N={N}: {time.perf_counter() - t0:.2f} s  {exc}")
