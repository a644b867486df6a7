import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import janus_oracle as O
sys.path.insert(0, "tests")
from tests.test_gpu_fullsize import _engine, _batch
eng, sd = _engine(); d = O.JANUS_1P3B
B, steps = 4, 5
ids, mask = _batch(B)
tr16, tr32 = {}, {}
ref_tok, _ = O.t2i(sd, d, ids, mask, sampler=O.greedy_sampler, mode="autocast", image_token_num_per_image=steps, decode=False, trace=tr16)
forced = torch.zeros(B, steps, dtype=torch.long)
O.t2i(sd, d, ids, mask, sampler=O.greedy_sampler, mode="fp32", image_token_num_per_image=steps, decode=False, trace=tr32, edit_region=forced, gt_labels=ref_tok)
r16 = torch.stack(tr16["raw_logits"]).numpy(); r32 = torch.stack(tr32["raw_logits"]).numpy()
emb = eng.language_model.get_input_embeddings()(ids)
outputs, got = None, []
for i in range(steps):
    outputs = eng.language_model.model(inputs_embeds=emb, attention_mask=mask, use_cache=True, past_key_values=outputs.past_key_values if i != 0 else None)
    got.append(eng.gen_head(outputs.last_hidden_state[:, -1, :]).float().cpu())
    tok = ref_tok[:, i].long()
    emb = eng.prepare_gen_img_embeds(torch.stack([tok, tok], 1).view(-1)).unsqueeze(1)
g = torch.stack(got).numpy()
def st(a, b, name):
    e = np.abs(a - b)
    print(f"{name:22s} max {e.max():.4f} mean {e.mean():.5f} p99.9 {np.quantile(e, 0.999):.4f}  per-step max {[round(float(e[s].max()),4) for s in range(steps)]}")
print("max|ref32|", np.abs(r32).max())
st(g, r16, "engine vs ref bf16"); st(r16, r32, "ref bf16 vs ref fp32"); st(g, r32, "engine vs ref fp32")
