import os, sys, torch
sys.path.insert(0, "/root/repo")
from plangen_b200 import JANUS_1P3B, synthetic
from plangen_b200.engine import FastJanus
import dataclasses
d = JANUS_1P3B; dev = torch.device("cuda", 0)
# tiny LM, real VQ: only the VQ decoder is timed
dd = dataclasses.replace(d, D=256, L=2, H=2, F=512, vocab=1000, img_embed=256, pad_id=999, name="tinylm-vq16")
sd = synthetic.random_state_dict(dd, dev, seed=0, with_vq=True)
eng = FastJanus(sd, dd, mode="bf16", max_batch=16, max_prompt=64)
codes = torch.randint(0, dd.img_vocab, (16, 576), dtype=torch.int32, device=dev)
def t(n=3):
    eng.gen_vision_model.decode_code(codes, shape=[16, 8, 24, 24]); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): eng.gen_vision_model.decode_code(codes, shape=[16, 8, 24, 24])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("default", round(t(), 2))
for ch in (8, 16):
    eng2 = FastJanus(sd, dd, mode="bf16", max_batch=16, max_prompt=64, options={"vq_chunk": ch})
    eng, keep = eng2, eng
    print("vq_chunk (at construction)", ch, round(t(), 2))
    eng = keep
    del eng2
for k, v in [("fuse_conv_epilogue", 1), ("vq_chunk", 2)]:
    eng.set_option(k, v)
    try:
        print(k, v, round(t(), 2))
    except Exception as ex:
        print(k, v, "failed", str(ex)[:100])
    eng.set_option(k, 0)
