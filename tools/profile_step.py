"""Short run of the hot path for ncu: prefill + a few decode steps at BASELINE configs[1] shapes
(graph capture off so every kernel is a plain launch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plangen_b200 import JANUS_1P3B, synthetic
from plangen_b200.engine import FastJanus

B = int(os.environ.get("PG_B", "16"))
STEPS = int(os.environ.get("PG_STEPS", "3"))
VQ = int(os.environ.get("PG_VQ", "0"))
dims = JANUS_1P3B
dev = torch.device("cuda", 0)
sd = synthetic.random_state_dict(dims, dev, seed=0, with_vq=True)
eng = FastJanus(sd, dims, mode="bf16", max_batch=B, max_prompt=512, options={"use_graph": int(os.environ.get("PG_GRAPH", "0"))})
del sd
cond, neg = synthetic.layoutsam_prompts(dims, B, seed=1234)
ids, mask = synthetic.collate_cfg_batch(cond, neg, dims.pad_id, dims.n_img_tokens)
emb = eng.language_model.get_input_embeddings()(ids.to(dev))
torch.cuda.synchronize()
toks = eng.sample_image(emb, B, STEPS, mask.to(dev), 5.0, 1.0, generator=0)
if VQ:
    full = torch.randint(0, dims.img_vocab, (B, dims.n_img_tokens), dtype=torch.int32, device=dev)
    eng.gen_vision_model.decode_code(full, shape=[B, dims.code_dim, dims.grid, dims.grid])
torch.cuda.synchronize()
print("done", toks[:, :STEPS].tolist()[0])
