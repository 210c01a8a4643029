"""One agent's step of the bench workload (default cfg4 shard: n=8192, yz_cx q=8 L=3, projected-gaussian), run
`--reps` times on cuda:0.  Short on purpose: this is the command profiled under ncu (launch list / --set full)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dqgp_b200 as d  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--encoding", default="yz_cx")
ap.add_argument("--kernel", default="projected")
ap.add_argument("--q", type=int, default=8)
ap.add_argument("--layers", type=int, default=3)
ap.add_argument("--d", type=int, default=4)
ap.add_argument("--outer", type=int, default=0)
ap.add_argument("--gradient", default="central_difference", choices=["central_difference", "analytic"])
ap.add_argument("--outer-kernel", default="gaussian", help="training outer kernel (honoured: cfg5 = matern)")
a = ap.parse_args()
x, y = d.synthetic_dataset(a.n, a.d, a.encoding)
eng = d.AgentEngine(x, y, encoding_type=a.encoding, kernel_type=a.kernel, num_qubits=a.q, num_layers=a.layers, noise_std=0.1,
                    rho=100.0, L=100.0, cholesky_outer_blocks=a.outer, gradient=a.gradient, outer_kernel=a.outer_kernel,
                    training_ignores_outer_kernel=False)
rs = np.random.RandomState(42)
z = d.kernels.dev_f64(np.round(rs.rand(eng.P), 4))
psi = d.kernels.dev_f64(np.round(rs.rand(eng.P), 4))
out = torch.empty((2, eng.P), dtype=torch.float64, device="cuda")
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.step(z, psi, out[0], out[1])
    e1.record()
    torch.cuda.synchronize()
    print(f"step {e0.elapsed_time(e1):.2f} ms  nll {eng.d_nll[3].item():.6f}")
eng.check_info()


def timed(name, fn, reps=3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"  {name:12s} {e0.elapsed_time(e1) / reps:8.3f} ms")
    return e0.elapsed_time(e1) / reps


timed("statevector", lambda: eng.simulate(z))
g = timed("gram", eng.gram)
timed("factor", lambda: (eng.gram(), eng.factor()))
print(f"    (factor includes one gram: subtract {g:.3f})")
timed("gradient", eng.gradient)
