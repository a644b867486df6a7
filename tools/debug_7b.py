"""Per-step error of the 7B-architecture engine vs the autocast / fp32 oracle on the same GPU (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import janus_oracle as O
from plangen_b200 import synthetic
from plangen_b200.config import Dims
from plangen_b200.engine import FastJanus
name = os.environ.get("PG_MODEL", "janus-pro-7b")
od = O.PRESETS[name]
d = Dims.from_any(od)
B, steps = int(os.environ.get("PG_B", "2")), 5
sd = synthetic.random_state_dict(d, torch.device("cuda", 0), seed=0, with_vq=False)
opts = {k[4:].lower(): int(v) for k, v in os.environ.items() if k.startswith("OPT_")}
eng = FastJanus(sd, d, mode="bf16", max_batch=B, max_prompt=512, with_vq=False, options=opts)
cond, neg = synthetic.layoutsam_prompts(d, B, seed=1234, lo=150, hi=480)
ids, mask = synthetic.collate_cfg_batch(cond, neg, d.pad_id, d.n_img_tokens)
ids, mask = ids.cuda(), mask.cuda()
tr16, tr32 = {}, {}
ref_tok, _ = O.t2i(sd, od, ids, mask, sampler=O.greedy_sampler, mode="autocast", image_token_num_per_image=steps, decode=False, trace=tr16)
forced = torch.zeros(B, steps, dtype=torch.long)
O.t2i(sd, od, ids, mask, sampler=O.greedy_sampler, mode="fp32", image_token_num_per_image=steps, decode=False, trace=tr32, edit_region=forced, gt_labels=ref_tok)
r16 = torch.stack(tr16["raw_logits"]).numpy(); r32 = torch.stack(tr32["raw_logits"]).numpy()
h16 = torch.stack(tr16["hidden"]).numpy(); h32 = torch.stack(tr32["hidden"]).numpy()
emb = eng.language_model.get_input_embeddings()(ids)
outputs = None
for i in range(steps):
    outputs = eng.language_model.model(inputs_embeds=emb, attention_mask=mask, use_cache=True, past_key_values=outputs.past_key_values if i != 0 else None)
    h = outputs.last_hidden_state[:, -1, :]
    lg = eng.gen_head(h).float().cpu().numpy()
    hh = h.float().cpu().numpy()
    tol = 2e-2 * np.abs(r16[i]) + 2e-2 * np.abs(r16[i]).max()
    print(f"step {i}: logits max|ref| {np.abs(r16[i]).max():.3f} frac_ok {(np.abs(lg - r16[i]) <= tol).mean():.4f} "
          f"| err vs fp32: mine mean {np.abs(lg - r32[i]).mean():.4e} max {np.abs(lg - r32[i]).max():.4e}; ref16 mean {np.abs(r16[i] - r32[i]).mean():.4e} max {np.abs(r16[i] - r32[i]).max():.4e}"
          f" | hidden: max|ref| {np.abs(h32[i]).max():.3f} mine-fp32 mean {np.abs(hh - h32[i]).mean():.4e}; ref16-fp32 mean {np.abs(h16[i] - h32[i]).mean():.4e}")
    for r in range(2 * B):
        print(f"    row {r}: hidden err mean {np.abs(hh[r] - h32[i][r]).mean():.4e} (ref16 {np.abs(h16[i][r] - h32[i][r]).mean():.4e})")
    tok = ref_tok[:, i].long()
    emb = eng.prepare_gen_img_embeds(torch.stack([tok, tok], 1).view(-1)).unsqueeze(1)
