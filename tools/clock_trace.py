"""Decode loop repeated for ~12 s with nvidia-smi sampled every 100 ms: ms/step per loop next to SM clock, power and
throttle reasons - is the sustained step time a clock effect?"""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plangen_b200 import JANUS_1P3B, synthetic
from plangen_b200.engine import FastJanus
B = 16; dims = JANUS_1P3B; dev = torch.device("cuda", 0)
sd = synthetic.random_state_dict(dims, dev, seed=0, with_vq=False)
eng = FastJanus(sd, dims, mode="bf16", max_batch=B, max_prompt=512, with_vq=False)
del sd
cond, neg = synthetic.layoutsam_prompts(dims, B, seed=1234)
ids, mask = synthetic.collate_cfg_batch(cond, neg, dims.pad_id, dims.n_img_tokens)
ids, mask = ids.to(dev), mask.to(dev)
emb = eng.language_model.get_input_embeddings()(ids)
Q = "clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown"
rows = []
p = subprocess.Popen(["nvidia-smi", "-i", "0", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
def rd():
    for line in p.stdout:
        rows.append((time.perf_counter(), line.strip()))
threading.Thread(target=rd, daemon=True).start()
st = torch.cuda.current_stream(dev)
if os.environ.get("PG_IDLE"):
    time.sleep(float(os.environ["PG_IDLE"]))
t_begin = time.perf_counter()
for it in range(int(os.environ.get("PG_LOOPS", "12"))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(st)
    eng.sample_image(emb, B, 576, mask, 5.0, 1.0, generator=0)
    e1.record(st)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    mine = [r for (t, r) in rows if t0 <= t <= t1]
    print(f"loop {it:2d} t={t0 - t_begin:6.2f}s  {e0.elapsed_time(e1) / 576:.4f} ms/step incl. prefill | " + " | ".join(mine[len(mine) // 2:len(mine) // 2 + 1]), flush=True)
p.terminate()
