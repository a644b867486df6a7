"""Standalone decode-attention time vs KV length: separates fixed overhead from per-byte cost."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plangen_b200 import JANUS_1P3B, synthetic, _lib
from plangen_b200.engine import FastJanus
B = 16; dims = JANUS_1P3B; dev = torch.device("cuda", 0)
sd = synthetic.random_state_dict(dims, dev, seed=0, with_vq=False)
eng = FastJanus(sd, dims, mode="bf16", max_batch=B, max_prompt=512, with_vq=False)
del sd
cond, neg = synthetic.layoutsam_prompts(dims, B, seed=1234)
ids, mask = synthetic.collate_cfg_batch(cond, neg, dims.pad_id, dims.n_img_tokens)
P = ids.shape[1]
kvs = (mask[:, :P] == 0).sum(1).to(torch.int32).to(dev).contiguous()
lens = (mask[:, :P] != 0).sum(1).tolist()
st = torch.cuda.current_stream(dev); sp = C.c_void_p(st.cuda_stream)
_lib.check(eng._lib.pg_debug_zero_part(eng._h, 2 * B * 3 * dims.H * dims.head_dim * 4, sp))
kv_tok = 2 * dims.H * dims.head_dim * 2
for pos in (P + 8, P + 150, P + 288, P + 450, P + 700):
    def one_pass():
        for l in range(dims.L):
            _lib.check(eng._lib.pg_test_attn_decode(eng._h, C.c_void_p(kvs.data_ptr()), 2 * B, pos, l, sp))
    for _ in range(2): one_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(4): one_pass()
    e1.record(st); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (4 * dims.L)
    mb = sum((ln + pos - P) * kv_tok for ln in lens) / 1e6
    print(f"pos {pos:5d}  {mb:7.1f} MB  {us:6.2f} us/launch  {mb/us*1e-3*1e3:7.1f} GB/s", flush=True)
