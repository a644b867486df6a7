"""-m gpu: parity at BASELINE.json's full sizes (Janus-1.3B architecture, random-init weights).

The oracle (reference PyTorch path) runs on the same GPU under autocast(bf16) for a few decode steps of
a real-shaped batch; longer runs are covered through size-independent properties: determinism of the
fused loop, graph vs plain launches, per-op kernels vs the persistent step kernel (bit-identical), and
teacher forcing (edit_region all zero reproduces gt_labels exactly)."""
import numpy as np
import pytest
import torch

from oracle import janus_oracle as O
from tests.gpu_util import assert_close, product_dims

pytestmark = pytest.mark.gpu

_STATE = {}


def _engine():
    if "eng" not in _STATE:
        from plangen_b200 import synthetic
        from plangen_b200.engine import FastJanus
        d = product_dims(O.JANUS_1P3B)
        sd = synthetic.random_state_dict(d, torch.device("cuda", 0), seed=0, with_vq=True)
        _STATE["sd"] = sd
        _STATE["eng"] = FastJanus(sd, d, mode="bf16", max_batch=16, max_prompt=512)
    return _STATE["eng"], _STATE["sd"]


def _batch(B, seed=1234, lo=150, hi=480):
    from plangen_b200 import synthetic
    d = product_dims(O.JANUS_1P3B)
    cond, neg = synthetic.layoutsam_prompts(d, B, seed=seed, lo=lo, hi=hi)
    ids, mask = synthetic.collate_cfg_batch(cond, neg, d.pad_id, d.n_img_tokens)
    return ids.cuda(), mask.cuda()


def test_fullsize_bf16_logits_vs_autocast_reference():
    """configs[1] shapes (B=4 of the 16 to keep the oracle cheap): gen_head logits of prefill + 4 decode steps,
    driven through the drop-in API exactly as System.sample_image drives vl_gpt and teacher-forced on the
    reference's tokens, vs the reference PyTorch path (fp32 master weights, autocast bf16) on the same GPU.
    Tolerance.  At 24 layers no two bf16 evaluations agree element-wise to rtol 2e-2: the reference's OWN
    autocast path deviates from its fp32 path by max 3.3e-2 / mean 4.9e-3 on these logits (max|logit| 0.91;
    measured on the B200, tools/fullsize_noise.py), the engine by max 3.2e-2 / mean 4.6e-3.  So the test asks:
    (1) >= 99.9 % of the logits within rtol 2e-2 + 2e-2 * max|ref| of the reference bf16 path (north_star's
    2e-2), (2) the engine's error against the fp32 reference not larger than 1.25x the reference bf16 path's
    own error against fp32 (mean and max) - i.e. the engine is as good a bf16 evaluation as the reference."""
    eng, sd = _engine()
    d = O.JANUS_1P3B
    B, steps = 4, 5
    ids, mask = _batch(B)
    tr16, tr32 = {}, {}
    ref_tok, _ = O.t2i(sd, d, ids, mask, sampler=O.greedy_sampler, mode="autocast", image_token_num_per_image=steps,
                       decode=False, trace=tr16)
    forced = torch.zeros(B, steps, dtype=torch.long)
    O.t2i(sd, d, ids, mask, sampler=O.greedy_sampler, mode="fp32", image_token_num_per_image=steps, decode=False,
          trace=tr32, edit_region=forced, gt_labels=ref_tok)
    ref_raw = torch.stack(tr16["raw_logits"]).numpy()
    ref_raw32 = torch.stack(tr32["raw_logits"]).numpy()
    ref_cfg = torch.stack(tr16["logits"]).numpy()
    # drop-in API, teacher-forced
    emb = eng.language_model.get_input_embeddings()(ids)
    outputs, got_raw = None, []
    for i in range(steps):
        outputs = eng.language_model.model(inputs_embeds=emb, attention_mask=mask, use_cache=True,
                                           past_key_values=outputs.past_key_values if i != 0 else None)
        logits = eng.gen_head(outputs.last_hidden_state[:, -1, :])
        got_raw.append(logits.float().cpu())
        tok = ref_tok[:, i].long()
        emb = eng.prepare_gen_img_embeds(torch.stack([tok, tok], 1).view(-1)).unsqueeze(1)
    got_raw = torch.stack(got_raw).numpy()
    tol = 2e-2 * np.abs(ref_raw) + 2e-2 * np.abs(ref_raw).max()
    frac_ok = float((np.abs(got_raw - ref_raw) <= tol).mean())
    assert frac_ok >= 0.999, f"only {frac_ok:.5f} of the logits within rtol 2e-2"
    e_mine, e_ref = np.abs(got_raw - ref_raw32), np.abs(ref_raw - ref_raw32)
    assert e_mine.mean() <= 1.25 * e_ref.mean(), (e_mine.mean(), e_ref.mean())
    assert e_mine.max() <= 1.25 * e_ref.max(), (e_mine.max(), e_ref.max())
    # fused loop: CFG logits
    dbg = torch.zeros(steps, B, d.img_vocab, device="cuda")
    eng.set_option("dbg_logits_ptr", dbg.data_ptr())
    try:
        emb = eng.language_model.get_input_embeddings()(ids)
        fb = {"edit_region": torch.zeros(B, steps, dtype=torch.int32)}
        got = eng.sample_image(emb, B, steps, mask, 5.0, 1.0, generator=0, batch=fb, gt_labels=ref_tok, greedy=True)
        torch.cuda.synchronize()
    finally:
        eng.set_option("dbg_logits_ptr", 0)
    assert got.cpu().tolist() == ref_tok.cpu().tolist()
    # CFG logits u + 5 (c - u) = 5c - 4u amplify per-row deviations up to 9x
    tol = 2e-2 * np.abs(ref_cfg) + 9 * 2e-2 * np.abs(ref_raw).max()
    frac_ok = float((np.abs(dbg.cpu().numpy() - ref_cfg) <= tol).mean())
    assert frac_ok >= 0.999, f"only {frac_ok:.5f} of the CFG logits within tolerance"


def test_fullsize_loop_properties():
    """Full 576-token loop at B=16: deterministic across runs, identical with and without CUDA-graph replay,
    the persistent step kernel agrees with the per-op kernels to rounding level, token ids in range."""
    eng, _ = _engine()
    d = O.JANUS_1P3B
    B = 16
    ids, mask = _batch(B)
    emb = eng.language_model.get_input_embeddings()(ids)
    a = eng.sample_image(emb, B, 576, mask, 5.0, 1.0, generator=0).cpu()
    b = eng.sample_image(emb, B, 576, mask, 5.0, 1.0, generator=0).cpu()
    assert torch.equal(a, b), "fused loop is not deterministic"
    assert int(a.min()) >= 0 and int(a.max()) < d.img_vocab and a.float().std() > 1000
    eng.set_option("use_graph", 0)
    try:
        c = eng.sample_image(emb, B, 576, mask, 5.0, 1.0, generator=0).cpu()
    finally:
        eng.set_option("use_graph", 1)
    assert torch.equal(a, c), "graph replay and plain launches disagree"
    # persistent step kernel vs per-op kernels: same arithmetic, different split-K / merge orders, so compare
    # teacher-forced CFG logits (not sampled ids, which diverge after the first rounding-level flip)
    n = 24
    fb = {"edit_region": torch.zeros(B, n, dtype=torch.int32)}
    logs = []
    for mega in (0, 1):
        dbg = torch.zeros(n, B, d.img_vocab, device="cuda")
        eng.set_option("use_mega", mega)
        eng.set_option("dbg_logits_ptr", dbg.data_ptr())
        try:
            eng.sample_image(emb, B, n, mask, 5.0, 1.0, generator=0, batch=fb, gt_labels=a[:, :n].contiguous())
            torch.cuda.synchronize()
        finally:
            eng.set_option("dbg_logits_ptr", 0)
            eng.set_option("use_mega", 0)
        logs.append(dbg.cpu().numpy())
    err = np.abs(logs[0] - logs[1])
    assert err.mean() < 5e-3 * np.abs(logs[0]).max() and np.quantile(err, 0.999) < 5e-2 * np.abs(logs[0]).max(), \
        (err.mean(), err.max(), np.abs(logs[0]).max())
    # a different seed gives a different sample; teacher forcing reproduces the labels exactly
    s2 = eng.sample_image(emb, B, 32, mask, 5.0, 1.0, generator=1).cpu()
    assert not torch.equal(a[:, :32], s2)
    forced = {"edit_region": torch.zeros(B, 32, dtype=torch.int32)}
    gt = torch.randint(0, d.img_vocab, (B, 32), dtype=torch.int32)
    f = eng.sample_image(emb, B, 32, mask, 5.0, 1.0, generator=0, batch=forced, gt_labels=gt).cpu()
    assert torch.equal(f, gt)


def test_fullsize_vq_decode_vs_autocast_reference():
    """VQ-16 at the real 24x24 grid (2 images): engine bf16 vs the reference decoder under autocast."""
    eng, sd = _engine()
    d = O.JANUS_1P3B
    g = torch.Generator().manual_seed(9)
    codes = torch.randint(0, d.img_vocab, (2, 576), generator=g, dtype=torch.int32).cuda()
    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
        want = O.decode_code(sd, d, codes, [2, 8, 24, 24]).float()
    with torch.inference_mode():
        want32 = O.decode_code(sd, d, codes, [2, 8, 24, 24]).float()
    out = eng.gen_vision_model.decode_code(codes, shape=[2, 8, 24, 24]).float()
    assert out.shape == (2, 3, 384, 384)
    a, b, b32 = out.cpu().numpy(), want.cpu().numpy(), want32.cpu().numpy()
    # Same reasoning as the LM test above: ~60 bf16 layers with GroupNorm do not agree element-wise in the
    # tail between any two bf16 evaluations (a single pixel of 884 736 moved across the fixed tolerance when
    # only the GroupNorm summation order changed).  So: (1) >= 99.99 % of the pixels within rtol 2e-2 +
    # 6e-2 * max|ref| of the reference bf16 path, (2) mean error small, (3) the engine's error against the
    # fp32 reference decoder no larger than 1.25x the reference bf16 path's own error against fp32.
    tol = 2e-2 * np.abs(b) + 6e-2 * np.abs(b).max()
    frac_ok = float((np.abs(a - b) <= tol).mean())
    assert frac_ok >= 0.9999, f"VQ-16 bf16 384x384: only {frac_ok:.6f} of the pixels within tolerance"
    assert np.abs(a - b).mean() < 1e-2 * np.abs(b).max()
    e_mine, e_ref = np.abs(a - b32), np.abs(b - b32)
    assert e_mine.mean() <= 1.25 * e_ref.mean(), (e_mine.mean(), e_ref.mean())
    assert np.quantile(e_mine, 0.9999) <= 1.25 * np.quantile(e_ref, 0.9999), (np.quantile(e_mine, 0.9999), np.quantile(e_ref, 0.9999))
    assert e_mine.max() <= 1.5 * e_ref.max(), (e_mine.max(), e_ref.max())
    # idempotence / batch independence: decoding image 0 alone gives the same pixels
    solo = eng.gen_vision_model.decode_code(codes[:1], shape=[1, 8, 24, 24]).float()
    assert torch.equal(solo[0], out[0])
