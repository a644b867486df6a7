"""Build recipe for libplangen_b200.so (sm_100a only, in-tree, no torch dependency)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libplangen_b200.so")
SOURCES = ["engine.cu"]
HEADERS = ["common.cuh", "gemm.cuh", "lm_kernels.cuh", "attn_tma.cuh", "attn_v4.cuh", "attn_v5.cuh", "step_kernel.cuh", "sample.cuh", "text_decode.cuh", "attn_prefill_tc.cuh", "vq_kernels.cuh",
           os.path.join("..", "..", "include", "plangen_b200.h")]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
           "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcudart"]
    cmd[1:1] = os.environ.get("PG_NVCC_FLAGS", "").split()
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libplangen_b200.so")
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
