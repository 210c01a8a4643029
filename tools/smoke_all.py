"""Small end-to-end exercise of every kernel (all four circuits, both kernel types, three outer kernels, prediction),
sized for `compute-sanitizer --tool memcheck|racecheck python tools/smoke_all.py`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dqgp_b200 as d  # noqa: E402

rs = np.random.RandomState(0)
for enc, ktype, outer, q, layers, dd, n in [("chebyshev", "projected", "matern", 3, 1, 2, 150), ("hubregtsen", "fidelity", "gaussian", 5, 2, 2, 130),
                                            ("yz_cx", "projected", "gaussian", 8, 2, 4, 200), ("kyriienko", "projected", "expsinesquared", 9, 1, 3, 140),
                                            ("yz_cx", "fidelity", "gaussian", 2, 2, 1, 70)]:
    x, y = d.synthetic_dataset(n, dd, enc)
    P = d.EncodingCircuit(enc, q, dd, layers).num_parameters
    z, psi = np.round(rs.rand(P), 4), np.round(rs.rand(P), 4)
    ag = d.RiemannianAgent("s", x, y, q, 0.1, 100.0, 100.0, use_parameter_shift=True, num_layers=layers, encoding_type=enc,
                           kernel_type=ktype, outer_kernel=outer, training_ignores_outer_kernel=False)
    try:
        th, ps, nll, _, _ = ag.train_and_update(z, psi)
    except np.linalg.LinAlgError as e:      # ExpSineSquared Grams are indefinite: every kernel still ran
        nll = float('nan')
        print(enc, ktype, outer, 'train:', str(e)[:50])
    xt, yt = d.synthetic_dataset(33, dd, enc, seed=5)
    try:
        mean, var, *_ = d.predict_quantum_gp(x, y, xt, np.mod(z, np.pi), q, layers, 0.1, True, enc, ktype, "XYZ",
                                             "gaussian" if outer == "expsinesquared" else outer, Y_test=yt)
        print(enc, ktype, outer, "nll %.6f" % nll, "nlpd %.6f" % d.predict_quantum_gp.last_nlpd)
    except RuntimeError as e:
        print(enc, ktype, outer, "nll %.6f" % nll, "predict:", str(e)[:60])
print("smoke_all ok")
