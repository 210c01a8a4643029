"""Time the stages of dqgp_potrf_solve_inv on an SPD Gram of size n: factor only / + triangular inverse + solve / + inverse product."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dqgp_b200 as d  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lib = d.load()
x, y = d.synthetic_dataset(n, 4, "yz_cx")
for ob in (1, 4):
    eng = d.AgentEngine(x, y, encoding_type="yz_cx", kernel_type="projected", num_qubits=8, num_layers=3, noise_std=0.1, rho=100.0,
                        L=100.0, cholesky_outer_blocks=ob)
    z = d.kernels.dev_f64(np.round(np.random.RandomState(42).rand(eng.P), 4))
    eng.simulate(z, 0, 1)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = {}
    for name, mode in (("potrf", -1), ("potrf+trtri+solve", 0), ("full", 1)):
        ts = []
        for rep in range(4):
            eng.gram()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.dqgp_potrf_solve_inv(eng.solver.handle, eng.d_Y.data_ptr(), eng.d_alpha.data_ptr(), eng.d_logdet.data_ptr(),
                                          eng.d_info.data_ptr(), mode, st)
            e1.record(); torch.cuda.synchronize()
            assert rc == 0
            ts.append(e0.elapsed_time(e1))
        res[name] = min(ts[1:])
    gf = float(((n + 127) // 128 * 128)) ** 3 / 3 / 1e9
    print(f"n={n} outer_blocks={ob}: potrf {res['potrf']:.2f} ms ({gf / res['potrf']:.1f} TF)  trtri+solve {res['potrf+trtri+solve'] - res['potrf']:.2f} ms "
          f"({gf / (res['potrf+trtri+solve'] - res['potrf']):.1f} TF)  lauum {res['full'] - res['potrf+trtri+solve']:.2f} ms "
          f"({gf / (res['full'] - res['potrf+trtri+solve']):.1f} TF)  total {res['full']:.2f} ms ({3 * gf / res['full']:.1f} TF)")
