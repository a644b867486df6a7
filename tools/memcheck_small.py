"""Small-shape run of every entry point for `compute-sanitizer --tool memcheck` (bf16 / tcgen05 regime):
prefill (tcgen05 attention) + CFG image decode loop + VQ decode, stage-1 text decode, VQ encode + editing t2i."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import janus_oracle as O
from plangen_b200.config import Dims
from plangen_b200.engine import FastJanus
d = O.SMALL
sd = O.init_state_dict(d, seed=0, with_vq=True, with_lm_head=True, with_vq_encoder=True)
eng = FastJanus(sd, Dims.from_any(d), mode="bf16", max_batch=4, max_prompt=160, max_steps=48)
lens = [3, 150, 129]
g = torch.Generator().manual_seed(1)
prompts = [torch.randint(0, d.pad_id, (n,), generator=g).tolist() for n in lens]
neg = [[1, 2, 3, 4, 5]] * 3
ids, mask = O.t2i_infer_collate_batch(prompts, neg, d.pad_id, d.n_img_tokens)
dec, _ = eng.t2i(tokens=ids.cuda(), mask=mask.cuda())
tids, tmask = O.pad_input_ids(prompts, d.pad_id)
emb = eng.language_model.get_input_embeddings()(tids.cuda())
txt = eng.language_model.generate(inputs_embeds=emb, attention_mask=tmask.cuda(), pad_token_id=d.vocab - 1, eos_token_id=d.vocab - 1, max_new_tokens=20)
side = d.grid * 2 ** (len(d.vq_ch_mult) - 1)
img = torch.rand(3, 3, side, side, generator=g).cuda() * 2 - 1
eng.t2i(tokens=ids.cuda(), mask=mask.cuda(), gt_image=img, batch={"edit_region": torch.zeros(3, d.n_img_tokens, dtype=torch.int64)}, use_teacher_forcing=True)
torch.cuda.synchronize()
print("ok", tuple(dec.shape), tuple(txt.shape), eng.last_tokens[0, :4].tolist())
