"""In-kernel timeline of the decode-attention kernel (per CTA / group %globaltimer stamps)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
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
st = torch.cuda.current_stream(dev); sp = C.c_void_p(st.cuda_stream)
_lib.check(eng._lib.pg_debug_zero_part(eng._h, 2 * B * 3 * dims.H * dims.head_dim * 4, sp))
nsm = torch.cuda.get_device_properties(0).multi_processor_count
names = ["entry", "prologue done", "first q ready", "first tile full", "stream end", "consumer end", "helper end", "producer end", "helper q0 out", "helper waited"]
pos = P + int(os.environ.get("PG_STEP", "288"))
for impl in (3,):
    for flags in (0,):
        eng.set_option("attn_impl", impl)
        eng.set_option("attn_test_flags", flags)
        dbg = torch.zeros(nsm * 4 * 16, dtype=torch.int64, device=dev)
        for l in range(6):
            _lib.check(eng._lib.pg_test_attn_decode(eng._h, C.c_void_p(kvs.data_ptr()), 2 * B, pos, l, sp))
        torch.cuda.synchronize()
        eng.set_option("attn_dbg_ptr", dbg.data_ptr())
        _lib.check(eng._lib.pg_test_attn_decode(eng._h, C.c_void_p(kvs.data_ptr()), 2 * B, pos, 7, sp))
        torch.cuda.synchronize()
        eng.set_option("attn_dbg_ptr", 0)
        t = dbg.cpu().numpy().reshape(nsm, 4, 16).astype(np.float64)
        t0 = t[:, :, 0][t[:, :, 0] > 0].min()
        os.makedirs("gpurun_out", exist_ok=True)
        np.save(f"gpurun_out/attn_tl_{impl}_{flags}.npy", t - t0)
        print(f"--- impl {impl} flags {flags} pos {pos}")
        for k in range(10):
            v = t[:, :, k]; v = v[v > 0]
            if v.size == 0: continue
            v = (v - t0) / 1e3
            print(f"  {names[k]:16s} min {v.min():6.2f}  p50 {np.median(v):6.2f}  p90 {np.percentile(v, 90):6.2f}  max {v.max():6.2f} us")
eng.set_option("attn_impl", 3); eng.set_option("attn_test_flags", 0)
