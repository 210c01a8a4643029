"""Build libdqgp.so (hand-written sm_100a CUDA + the C ABI of include/dqgp.h) in-tree with nvcc.

    python distributed-quantum-gaussian-processes_b200/build.py [--force]

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libdqgp.so")
# statevec.cu is compiled once per group of entry points (-DDQGP_SV_PART=k): its twelve qubit counts x kernel variants take
# 2.6 minutes in one translation unit
SOURCES = ["circuit.cu", ("statevec.cu", 1), ("statevec.cu", 2), ("statevec.cu", 3), ("statevec.cu", 4), ("statevec.cu", 5),
           "gram.cu", "gemm64.cu", "chol.cu", "lu.cu", "grad.cu", "fid.cu", "admm.cu", "api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _deps_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "dqgp.h")]
    return max(os.path.getmtime(f) for f in files)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    headers_mtime = max([os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith(".cuh")] +
                        [os.path.getmtime(os.path.join(HERE, "..", "include", "dqgp.h"))])
    if os.environ.get("DQGP_NVCC_EXTRA"):
        force = True

    def compile_one(src):
        part = []
        if isinstance(src, tuple):
            src, k = src
            obj = os.path.join(OBJ, src.replace(".cu", f"_p{k}.o"))
            part = [f"-DDQGP_SV_PART={k}"]
        else:
            obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        # incremental: an object is rebuilt when its source or any header (csrc/*.cuh, include/dqgp.h) is newer
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(os.path.join(CSRC, src)), headers_mtime):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *part, *os.environ.get("DQGP_NVCC_EXTRA", "").split(), "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr.strip():
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(os.cpu_count() or 8, len(SOURCES))) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
