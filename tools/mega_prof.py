"""Per-phase timing inside the persistent step kernel (globaltimer stamps of every CTA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plangen_b200 import JANUS_1P3B, synthetic
from plangen_b200.engine import FastJanus
B = 16; dims = JANUS_1P3B; dev = torch.device("cuda", 0)
sd = synthetic.random_state_dict(dims, dev, seed=0, with_vq=False)
eng = FastJanus(sd, dims, mode="bf16", max_batch=B, max_prompt=512, with_vq=False, options={"use_mega": 1})
del sd
cond, neg = synthetic.layoutsam_prompts(dims, B, seed=1234)
ids, mask = synthetic.collate_cfg_batch(cond, neg, dims.pad_id, dims.n_img_tokens)
emb = eng.language_model.get_input_embeddings()(ids.to(dev))
n = int(os.environ.get("PG_STEPS", "300"))
G = eng.counter("num_sms"); NP = dims.L * 8 + 1
prof = torch.zeros(G * NP + 16, dtype=torch.int64, device=dev)
eng.set_option("sk_prof_ptr", prof.data_ptr())
eng.sample_image(emb, B, n, mask.to(dev), 5.0, 1.0, generator=0)
torch.cuda.synchronize()
fine = prof[G * NP:].cpu().double()
t = prof[:G * NP].view(G, NP).cpu().double()          # stamps of the LAST launch
t0 = t[:, 0].min()
end = t[:, 1:]                   # [G, L*8] phase end per CTA
# phase duration measured globally: max over CTAs of phase end - max over CTAs of previous phase end
gend = end.max(0).values
gprev = torch.cat([t[:, 0].max().view(1), gend[:-1]])
dur = (gend - gprev).view(dims.L, 8) / 1e3
names = ["qkv", "attn", "o", "norm1", "gu", "swiglu", "down", "norm2"]
print("total us", float(gend[-1] - t0) / 1e3)
print("mean per phase (us):", {n_: round(float(dur[:, i].mean()), 2) for i, n_ in enumerate(names)})
print("layer sum us", round(float(dur.sum(1).mean()), 2))
# spread of arrival within a phase (last - first CTA)
spread = (end.max(0).values - end.min(0).values).view(dims.L, 8) / 1e3
print("mean arrival spread (us):", {n_: round(float(spread[:, i].mean()), 2) for i, n_ in enumerate(names)})
lay = 12
a_end = end[:, lay * 8 + 1]; a_start = end[:, lay * 8 + 0].max()
d = ((a_end - a_start) / 1e3)
print("attn per-CTA duration us (layer 12), every 4th CTA:", [round(float(x), 1) for x in d[::4]])
q_end = end[:, lay * 8 + 0]; q_start = end[:, lay * 8 - 1].max()
print("qkv per-CTA us:", [round(float(x), 1) for x in ((q_end - q_start) / 1e3)[::8]])
g_end = end[:, lay * 8 + 4]; g_start = end[:, lay * 8 + 3].max()
print("gu per-CTA us:", [round(float(x), 1) for x in ((g_end - g_start) / 1e3)[::8]])
n_end = end[:, lay * 8 + 3]; n_start = end[:, lay * 8 + 2].max()
print("norm1 per-CTA us:", [round(float(x), 1) for x in ((n_end - n_start) / 1e3)[::8]])

print("norm1 fine (CTA 0, layer 12) us: wait %.2f loads %.2f reduce %.2f stores %.2f done(fence+bar+atomic) %.2f" % tuple(float(fine[i + 1] - fine[i]) / 1e3 for i in range(5)))
