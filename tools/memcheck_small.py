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
# round 2: top-k sampler, stream-K gate|up (forced), padded prefill (packing off), drop-in per-call API, mmu front-end
eng.t2i(tokens=ids.cuda(), mask=mask.cuda(), top_k=7)
eng.set_option("gu_streamk", 2)
eng.t2i(tokens=ids.cuda(), mask=mask.cuda())
eng.set_option("gu_streamk", 1)
eng.set_option("prefill_pack", 0)
eng.t2i(tokens=ids.cuda(), mask=mask.cuda())
eng.set_option("prefill_pack", 1)
out = eng.language_model.model(inputs_embeds=eng.language_model.get_input_embeddings()(ids.cuda()), attention_mask=mask.cuda(), use_cache=True)
lg = eng.gen_head(out.last_hidden_state[:, -1, :])
nx = eng.prepare_gen_img_embeds(torch.zeros(ids.shape[0], dtype=torch.int64, device="cuda")).unsqueeze(1)
out2 = eng.language_model.model(inputs_embeds=nx, attention_mask=mask.cuda(), use_cache=True, past_key_values=out.past_key_values)
u8 = eng.images_to_uint8(dec)
del eng
v = O.SigLIPDims(name="siglip-small-hd64", width=128, layers=2, heads=2, patch=16, image=96)
sdv = {**O.init_state_dict(O.TINY, seed=0, with_vq=False, with_lm_head=True), **O.init_siglip_state_dict(v, O.TINY, seed=0)}
ev = FastJanus(sdv, Dims.from_any(O.TINY, vision=v), mode="bf16", max_batch=4, max_prompt=96, with_vq=False, max_images=3)
n = v.n_patches
pix = torch.rand(3, 1, 3, v.image, v.image, generator=g).cuda() * 2 - 1
T = n + 9
mids = torch.randint(1, 900, (3, T), generator=g)
seq = torch.zeros(3, T, dtype=torch.bool); seq[:, 4:4 + n] = True
mids[seq] = -1
x = ev.prepare_inputs_embeds(mids.cuda(), pix, seq.cuda(), torch.ones(3, 1, n, dtype=torch.bool).cuda())
t2 = ev.language_model.generate(inputs_embeds=x, attention_mask=torch.ones(3, T, dtype=torch.int32).cuda(), pad_token_id=7, eos_token_id=7, max_new_tokens=6)
torch.cuda.synchronize()
print("ok round-2 paths", tuple(lg.shape), tuple(out2.last_hidden_state.shape), tuple(u8.shape), tuple(x.shape), tuple(t2.shape))
