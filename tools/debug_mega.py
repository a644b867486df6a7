"""Compare intermediates of the persistent step kernel with the per-op path on one decode step."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plangen_b200 import synthetic, _lib
from plangen_b200.config import Dims
from plangen_b200.engine import FastJanus

L = int(os.environ.get("PG_L", "1")); H = int(os.environ.get("PG_H", "2")); Dm = int(os.environ.get("PG_D", "256")); F = int(os.environ.get("PG_F", "512"))
dims = Dims(name="dbg", D=Dm, L=L, H=H, F=F, vocab=1000, img_vocab=2048, img_embed=256, grid=4, vq_ch=32, vq_ch_mult=(1, 2), vq_z=32, pad_id=999)
B = int(os.environ.get("PG_B", "3")); P = int(os.environ.get("PG_P", "40"))
dev = torch.device("cuda", 0)
sd = synthetic.random_state_dict(dims, dev, seed=0, with_vq=False)
engs = {m: FastJanus(sd, dims, mode="bf16", max_batch=B, max_prompt=64, max_steps=64, with_vq=False, options={"use_mega": m}) for m in (0, 1)}
g = torch.Generator().manual_seed(1)
ids = torch.randint(0, 900, (2 * B, P), generator=g, dtype=torch.int32).to(dev)
mask = torch.ones(2 * B, P + 64, dtype=torch.int32, device=dev)
mask[1, :7] = 0; mask[3, :20] = 0
xin = torch.randn(2 * B, 1, dims.D, device=dev).to(torch.bfloat16).float()

def grab(e, name, shape, dtype):
    t = torch.empty(shape, dtype=dtype, device=dev)
    _lib.check(e._lib.pg_debug_copy(e._h, name.encode(), C.c_void_p(t.data_ptr()), t.numel() * t.element_size(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return t.float()

out = {}
for m, e in engs.items():
    emb = e.language_model.get_input_embeddings()(ids)
    o = e.language_model.model(inputs_embeds=emb, attention_mask=mask, use_cache=True)
    cur = xin
    for step in range(int(os.environ.get("PG_STEPS", "1"))):
        o2 = e.language_model.model(inputs_embeds=cur, attention_mask=mask, use_cache=True, past_key_values=o.past_key_values)
    torch.cuda.synchronize()
    R = 2 * B
    out[m] = {"hidden": o2.last_hidden_state.float().view(R, -1),
              "attn_out": grab(e, "attn_out", (R, dims.H * 128), torch.bfloat16),
              "xn": grab(e, "xn", (R, dims.D), torch.bfloat16),
              "hbuf": grab(e, "hbuf", (R, dims.F), torch.bfloat16),
              "x_dec": grab(e, "x_dec", (R, dims.D), torch.float32)}
for k in out[0]:
    a, b = out[0][k], out[1][k]
    d = (a - b).abs()
    rows = d.max(dim=1).values
    print(f"{k:9s} max|ref| {a.abs().max():.4f}  max err {d.max():.5f}  per-row max err {[round(float(x),4) for x in rows]}")
    if d.max() > 0.05 * a.abs().max():
        r = int(rows.argmax())
        cols = (d[r] > 0.05 * a.abs().max()).nonzero().flatten()
        print("   bad row", r, "bad cols count", len(cols), "first", cols[:16].tolist(), "last", cols[-4:].tolist())
