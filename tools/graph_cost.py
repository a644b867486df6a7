"""Cost of a decode-step graph capture + instantiation: sample_image (8 tokens) on a shape seen before vs a new padded prompt length."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plangen_b200 import JANUS_1P3B, synthetic
from plangen_b200.engine import FastJanus
B = 16; dims = JANUS_1P3B; dev = torch.device("cuda", 0)
sd = synthetic.random_state_dict(dims, dev, seed=0, with_vq=False)
eng = FastJanus(sd, dims, mode="bf16", max_batch=B, max_prompt=512, with_vq=False)
del sd
def batch(seed):
    cond, neg = synthetic.layoutsam_prompts(dims, B, seed=seed)
    ids, mask = synthetic.collate_cfg_batch(cond, neg, dims.pad_id, dims.n_img_tokens)
    return eng.language_model.get_input_embeddings()(ids.to(dev)), mask.to(dev), ids.shape[1]
def run(emb, mask):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.sample_image(emb, B, 8, mask, 5.0, 1.0, generator=0)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
e1, m1, P1 = batch(1); e2, m2, P2 = batch(2); e3, m3, P3 = batch(3)
print("P", P1, P2, P3)
print("first (capture)", run(e1, m1)); print("same shape again", run(e1, m1)); print("same shape again", run(e1, m1))
print("new P (capture)", run(e2, m2)); print("again", run(e2, m2)); print("new P (capture)", run(e3, m3)); print("back to first (cached)", run(e1, m1))
