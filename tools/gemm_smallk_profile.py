"""ncu target: three launches of the DMMA GEMM as a rank-K update C -= A B^T (8192 x 8192, beta = 1), K from argv (default 128)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dqgp_b200 as d  # noqa: E402

lib = d.load()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 128
M = N = 8192
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
a = torch.randn((M, K), dtype=torch.float64, device="cuda")
b = torch.randn((N, K), dtype=torch.float64, device="cuda")
c = torch.zeros((M, N), dtype=torch.float64, device="cuda")
for _ in range(3):
    assert lib.dqgp_dgemm(1, 1, M, N, K, -1.0, a.data_ptr(), K, b.data_ptr(), K, 1.0, c.data_ptr(), N, st) == 0
torch.cuda.synchronize()
print("done", float(c[0, 0]))
