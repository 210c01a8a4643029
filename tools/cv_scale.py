"""5-fold CV of the consensus parameters (main.py:1490-1596, run once per ADMM iteration by the reference, main.py:2650) on the
FULL training set on one GPU.    python tools/cv_scale.py N_train [cfg4|cfg5]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dqgp_b200 as d  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 58982
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg4"
enc, q, layers, dim, outer = ("yz_cx", 8, 3, 4, "gaussian") if cfg == "cfg4" else ("kyriienko", 10, 4, 6, "matern")
x, y = d.synthetic_dataset(n, dim, enc)
P = d.EncodingCircuit(enc, q, dim, layers).num_parameters
z = np.round(np.random.RandomState(42).rand(P), 4)
torch.cuda.synchronize()
t0 = time.perf_counter()
cv = d.k_fold_cross_validation_consensus(x, y, z, q, layers, 0.1, k_folds=5, encoding_type=enc, kernel_type="projected", outer_kernel=outer)
torch.cuda.synchronize()
print(f"[{cfg}] 5-fold CV on {n} training samples: {time.perf_counter() - t0:.1f} s  mean NLPD {cv['mean_nlpd']:.4f} +- {cv['std_nlpd']:.4f}  "
      f"RMSE {cv['mean_rmse']:.4f}  valid folds {cv['valid_folds']}/{cv['total_folds']}")
assert cv["valid_folds"] == 5
