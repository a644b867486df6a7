"""Time the CTA-pair contraction (gemm_tc2.cuh) through pg_test_gemm on the prefill / SigLIP shapes for several ring depths.
Usage: python tools/tc2_sweep.py [stages ...]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import janus_oracle as O
from plangen_b200 import _lib
from plangen_b200.config import Dims
from plangen_b200.engine import FastJanus

sd = O.init_state_dict(O.TINY, seed=0, with_vq=False)
eng = FastJanus(sd, Dims.from_any(O.TINY), mode="bf16", max_batch=2, max_prompt=32, with_vq=False)
shapes = [("prefill qkv", 6228, 6144, 2048), ("prefill o", 6228, 2048, 2048), ("prefill gate|up", 6228, 11264, 2048),
          ("prefill down", 6228, 2048, 5632), ("siglip qkv", 73728, 3072, 1024), ("siglip fc1", 73728, 4096, 1024),
          ("siglip fc2", 73728, 1024, 4096)]
stages = [int(a) for a in sys.argv[1:]] or [3, 4]
st = torch.cuda.current_stream().cuda_stream
for name, M, N, K in shapes:
    X = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(1, M, N, device="cuda", dtype=torch.float32)
    line = f"{name:16s} M={M} N={N} K={K}:"
    for s in stages:
        eng.set_option("tc2_stages", s)
        def run():
            _lib.check(eng._lib.pg_test_gemm(eng._h, 1, 1, C.c_void_p(X.data_ptr()), C.c_void_p(W.data_ptr()), M, N, K, 1,
                                             C.c_void_p(out.data_ptr()), C.c_void_p(st)))
        for _ in range(3): run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): run()
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) * 100
        line += f"  s{s}: {us:7.1f} us {2.0 * M * N * K / us / 1e6:6.0f} TF/s"
    print(line, flush=True)
    del X, W, out
