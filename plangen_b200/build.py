"""Build recipe for libplangen_b200.so (sm_100a only, in-tree, no torch dependency)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libplangen_b200.so")
SOURCES = ["engine.cu"]
HEADERS = ["common.cuh", "gemm.cuh", "gemm_sk.cuh", "gemm_tc2.cuh", "lm_kernels.cuh", "attn_tma.cuh", "attn_v5.cuh", "sample.cuh", "text_decode.cuh", "attn_prefill_tc.cuh", "vq_kernels.cuh", "vit_kernels.cuh",
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
    # one builder at a time (torchrun starts N ranks that may all find the library stale): exclusive lock, re-check,
    # compile into a private file and publish it with an atomic rename so no rank ever maps a half-written library
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIB
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            tmp = f"{LIB}.tmp{os.getpid()}"
            cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                   "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                   "-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcudart"]
            cmd[1:1] = os.environ.get("PG_NVCC_FLAGS", "").split()
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libplangen_b200.so")
            os.replace(tmp, LIB)
            if verbose:
                print(r.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
