"""In-kernel timeline of the two resid+RMSNorm kernels of one layer inside the decode graph, next to the begin/end of
their neighbours (O / gate|up, down / QKV).  Stamps: entry, past the dependency wait, loads+sums done (warp 0),
block reduction done, stores issued."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from plangen_b200 import JANUS_1P3B, synthetic
from plangen_b200.engine import FastJanus
B = 16; dims = JANUS_1P3B; dev = torch.device("cuda", 0)
sd = synthetic.random_state_dict(dims, dev, seed=0, with_vq=False)
opts = {k: int(v) for k, v in (kv.split("=") for kv in os.environ.get("PG_OPTS", "").split(",") if kv)}
eng = FastJanus(sd, dims, mode="bf16", max_batch=B, max_prompt=512, with_vq=False, options=opts)
del sd
cond, neg = synthetic.layoutsam_prompts(dims, B, seed=1234)
ids, mask = synthetic.collate_cfg_batch(cond, neg, dims.pad_id, dims.n_img_tokens)
emb = eng.language_model.get_input_embeddings()(ids.to(dev))
step = int(os.environ.get("PG_STEP", "300")); NS = 256
prof = torch.zeros(4 * NS, dtype=torch.int64, device=dev)
nd = torch.zeros(NS * 64 * 8, dtype=torch.int64, device=dev)
eng.set_option("prof_ptr", prof.data_ptr()); eng.set_option("prof_step", step)
eng.set_option("norm_dbg_ptr", nd.data_ptr()); eng.set_option("norm_dbg_step", step)
eng.sample_image(emb, B, step + 3, mask.to(dev), 5.0, 1.0, generator=0)
torch.cuda.synchronize()
eng.set_option("norm_dbg_ptr", 0)
snap = prof[2 * NS:].cpu().numpy().astype(np.float64)
beg, end = snap[:NS], snap[NS:]
ndc = nd.cpu().numpy().astype(np.float64).reshape(NS, 64, 8)
per_layer = ["qkv", "attn", "o", "norm1", "gu", "down", "norm2"]
names = ["entry", "past wait", "loads done (w0)", "block sum done", "stores issued"]
for L in (10,):
    base = 2 + L * 7
    t0 = beg[base]
    print(f"--- layer {L} (us relative to the layer's QKV begin)")
    for k, nm in enumerate(per_layer):
        print(f"  {nm:6s} begin {(beg[base + k] - t0) / 1e3:8.2f}  end {(end[base + k] - t0) / 1e3:8.2f}")
    for which, off in (("norm1", 3), ("norm2", 6)):
        t = ndc[base + off, :2 * B, :5]
        if not (t[:, 0] > 0).all():
            print("  no stamps for", which); continue
        print(f"  {which}: predecessor end {(end[base + off - 1] - t0) / 1e3:8.2f}")
        for k in range(5):
            v = (t[:, k] - t0) / 1e3
            print(f"    {names[k]:18s} min {v.min():8.2f} p50 {np.median(v):8.2f} max {v.max():8.2f}")
