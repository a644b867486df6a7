"""Throughput of the mmu front-end tower (SigLIP-L/16-384 + aligner) at BASELINE configs[3]'s batch: images/s and the
fraction of the tensor-pipe peak (algorithmic FLOPs: 2 x params x patches + attention)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plangen_b200 import JANUS_1P3B, synthetic
from plangen_b200.config import Dims
from plangen_b200.engine import FastJanus
B = int(os.environ.get("PG_B", "128")); dev = torch.device("cuda", 0)
d = Dims(**{**JANUS_1P3B.__dict__, "L": 2, "name": "janus-1.3b-2layer"})
sd = synthetic.random_state_dict(d, dev, seed=0, with_vq=False, with_vision=True)
opts = {k[4:].lower(): int(v) for k, v in os.environ.items() if k.startswith("OPT_")}
eng = FastJanus(sd, d, mode="bf16", max_batch=2, max_prompt=64, with_vq=False, max_images=B, options=opts)
del sd
img = torch.rand(B, 3, d.sig_image, d.sig_image, device=dev) * 2 - 1
import ctypes as C
from plangen_b200 import _lib
st = torch.cuda.current_stream(dev); sp = C.c_void_p(st.cuda_stream)
def run():
    _lib.check(eng._lib.pg_vision_features(eng._h, C.c_void_p(img.data_ptr()), B, None, sp))
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = int(os.environ.get("PG_ITERS", "5"))
e0.record(st)
for _ in range(iters): run()
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
W, L, NP, F, D = d.sig_width, d.sig_layers, d.sig_patches, d.sig_mlp, d.D
lin = L * (4 * W * W + 2 * W * F) + 3 * d.sig_patch ** 2 * W + W * D + D * D
flop_img = 2 * lin * NP + L * 4 * NP * NP * W
peaks = {}
try: peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception: pass
peak = peaks.get("bf16_tflops_sustained") or 1369.6
tf = flop_img * B / (ms / 1e3) / 1e12
print(json.dumps({"images": B, "ms": ms, "images_per_s": B / (ms / 1e3), "gflop_per_image": flop_img / 1e9, "tflops": tf,
                  "frac_of_sustained_bf16_peak": tf / peak, "peak_tflops": peak}))
