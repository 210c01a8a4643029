"""Cholesky-only timing of dqgp_potrf_solve_inv (factor only) on a random SPD matrix for several n: us per 128-column step."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dqgp_b200 as d  # noqa: E402
from dqgp_b200.engine import Solver  # noqa: E402

lib = d.load()
for n in [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192]:
    for ob in [int(o) for o in os.environ.get("OBS", "1,4").split(",")]:
        s = Solver(n, ob)
        g = torch.Generator(device="cuda").manual_seed(1)
        B = torch.randn((n, 64), dtype=torch.float64, device="cuda", generator=g)
        K = B @ B.T / 64 + torch.eye(n, dtype=torch.float64, device="cuda")
        logdet = torch.zeros(1, dtype=torch.float64, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        ts = []
        for rep in range(4):
            s.matrix().copy_(K)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.dqgp_potrf_solve_inv(s.handle, None, None, logdet.data_ptr(), info.data_ptr(), -1, st)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0 and info.item() == 0
            ts.append(e0.elapsed_time(e1))
        nblk = (n + 127) // 128
        print(f"n={n} ob={ob}: potrf {min(ts[1:]):.3f} ms = {1e3 * min(ts[1:]) / nblk:.1f} us per step  logdet {logdet.item():.6f}", flush=True)
