"""In-kernel timeline of one tcgen05 contraction inside the decode graph (per-CTA %globaltimer stamps).
PG_N selects the launch by its weight-row count: 11264 gate|up, 6144 qkv, 2048 o / down (last one wins)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from plangen_b200 import JANUS_1P3B, synthetic
from plangen_b200.engine import FastJanus
B = 16; dims = JANUS_1P3B; dev = torch.device("cuda", 0)
sd = synthetic.random_state_dict(dims, dev, seed=0, with_vq=False)
eng = FastJanus(sd, dims, mode="bf16", max_batch=B, max_prompt=512, with_vq=False)
del sd
cond, neg = synthetic.layoutsam_prompts(dims, B, seed=1234)
ids, mask = synthetic.collate_cfg_batch(cond, neg, dims.pad_id, dims.n_img_tokens)
emb = eng.language_model.get_input_embeddings()(ids.to(dev))
names = ["entry", "X producer past wait", "first stage full", "last stage full", "accumulator done", "epilogue done", "exit", "first TMEM chunk read (SwiGLU)"]
for n in [int(x) for x in os.environ.get("PG_N", "11264,6144,2048").split(",")]:
    dbg = torch.zeros(2048 * 8 + 512, dtype=torch.int64, device=dev)
    eng.set_option("gemm_dbg_n", n); eng.set_option("gemm_dbg_ptr", dbg.data_ptr())
    eng.sample_image(emb, B, 40, mask.to(dev), 5.0, 1.0, generator=0)
    torch.cuda.synchronize()
    eng.set_option("gemm_dbg_ptr", 0)
    raw = dbg.cpu().numpy().astype(np.float64)
    t = raw[:2048 * 8].reshape(-1, 8)
    fine = raw[2048 * 8:]
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    print(f"--- weight rows {n}: {t.shape[0]} CTAs")
    for k in range(8):
        if not (t[:, k] > 0).any(): continue
        v = (t[:, k] - t0) / 1e3
        print(f"  {names[k]:22s} min {v.min():6.2f}  p50 {np.median(v):6.2f}  p90 {np.percentile(v, 90):6.2f}  max {v.max():6.2f} us")
    for nm, o in (("MMA saw stage full", 0), ("W producer issued", 64), ("X producer issued", 128)):
        v = fine[o:o + 64]; v = v[v > 0]
        if v.size: print(f"  CTA 5 {nm:20s}", " ".join(f"{(x - t0) / 1e3:.2f}" for x in v[:40]))
