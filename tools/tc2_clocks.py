"""SM clock / power while the CTA-pair contraction runs back to back for ~4 s on one prefill shape and one SigLIP shape:
is its throughput a clock (power cap) effect?  Prints TFLOP/s per 0.5 s window next to the nvidia-smi samples."""
import ctypes as C, os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import janus_oracle as O
from plangen_b200 import _lib
from plangen_b200.config import Dims
from plangen_b200.engine import FastJanus
eng = FastJanus(O.init_state_dict(O.TINY, seed=0, with_vq=False), Dims.from_any(O.TINY), mode="bf16", max_batch=2, max_prompt=32, with_vq=False)
Q = "clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown"
rows = []
p = subprocess.Popen(["nvidia-smi", "-i", "0", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
def rd():
    for line in p.stdout: rows.append((time.perf_counter(), line.strip()))
threading.Thread(target=rd, daemon=True).start()
st = torch.cuda.current_stream().cuda_stream
for name, M, N, K in [("prefill gate|up", 6228, 11264, 2048), ("siglip fc1", 73728, 4096, 1024)]:
    X = torch.randn(M, K, device="cuda").to(torch.bfloat16); W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(1, M, N, device="cuda", dtype=torch.float32)
    def run():
        _lib.check(eng._lib.pg_test_gemm(eng._h, 1, 1, C.c_void_p(X.data_ptr()), C.c_void_p(W.data_ptr()), M, N, K, 1, C.c_void_p(out.data_ptr()), C.c_void_p(st)))
    for _ in range(3): run()
    torch.cuda.synchronize(); time.sleep(1.0)
    for w in range(8):
        n = max(1, int(0.5 / (2.0 * M * N * K / 1.1e15)))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        for _ in range(n): run()
        b.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
        smp = [r for t, r in rows if t0 <= t <= t1]
        print(f"{name} window {w}: {2.0 * M * N * K * n / (a.elapsed_time(b) * 1e-3) / 1e12:7.0f} TF/s   smi[{len(smp)}] {smp[len(smp) // 2] if smp else ''}", flush=True)
    del X, W, out
p.terminate()
