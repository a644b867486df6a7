"""Stage-1 layout-text decode (language_model.generate -> pg_generate_greedy) at full size: ms per greedy step.
PG_B rows (BASELINE configs[2] stage 1: 64 per GPU), PG_P prompt length, PG_N new tokens (eos never hit: vocab-1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plangen_b200 import JANUS_1P3B, synthetic
from plangen_b200.engine import FastJanus

B = int(os.environ.get("PG_B", "64")); P = int(os.environ.get("PG_P", "160")); N = int(os.environ.get("PG_N", "200"))
dims = JANUS_1P3B
dev = torch.device("cuda", 0)
sd = synthetic.random_state_dict(dims, dev, seed=0, with_vq=False, with_lm_head=True)
eng = FastJanus(sd, dims, mode="bf16", max_batch=(B + 1) // 2, max_prompt=512, with_vq=False)
del sd
g = torch.Generator().manual_seed(5)
lens = torch.randint(P // 2, P + 1, (B,), generator=g).tolist()
lens[0] = P
ids = torch.full((B, P), dims.pad_id, dtype=torch.int32)
mask = torch.zeros(B, P, dtype=torch.int32)
for r, n in enumerate(lens):
    ids[r, P - n:] = torch.randint(0, dims.pad_id, (n,), generator=g, dtype=torch.int32)
    mask[r, P - n:] = 1
ids, mask = ids.to(dev), mask.to(dev)
emb = eng.language_model.get_input_embeddings()(ids)
eos = dims.vocab - 1
st = torch.cuda.current_stream(dev)


def run(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    out = eng.language_model.generate(inputs_embeds=emb, attention_mask=mask, pad_token_id=eos, eos_token_id=eos, max_new_tokens=n)
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), out


run(N)
t1, _ = run(1)
tn, out = run(N)
step_ms = (tn - t1) / (N - 1)
kvb = 2 * dims.L * dims.D * 2
mean_T = sum(lens) / B + N / 2
bytes_step = eng.weight_bytes_per_step - (dims.img_embed * dims.D + dims.img_vocab * dims.img_embed) * 2 + dims.vocab * dims.D * 2 + B * mean_T * kvb
print(f"x2t rows={B} P={P} new={N}: prefill+1 {t1:.2f} ms, {step_ms:.4f} ms/step, {B / step_ms * 1e3:.0f} tokens/s, "
      f"algorithmic {bytes_step / 1e9:.3f} GB/step = {bytes_step / step_ms / 1e6:.0f} GB/s; tokens {out.shape}")
