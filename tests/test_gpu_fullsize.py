"""-m gpu: parity at BASELINE.json's full sizes (Janus-1.3B architecture, random-init weights).

The oracle (reference PyTorch path) runs on the same GPU under autocast(bf16) for a few decode steps of
a real-shaped batch; longer runs are covered through size-independent properties: determinism of the
fused loop, graph vs plain launches, and teacher forcing (edit_region all zero reproduces gt_labels exactly)."""
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


def _logits_parity(eng, sd, d, B, steps, batch_seed=1234, cfg_quantile=0.999):
    """gen_head logits of prefill + (steps - 1) decode steps, driven through the drop-in API exactly as
    System.sample_image drives vl_gpt and teacher-forced on the reference's tokens, vs the reference PyTorch path
    (fp32 master weights, autocast bf16) on the same GPU; then the fused loop's CFG logits the same way.
    Tolerance actually enforced (DESIGN.md section 2 states it as a deviation from a bare rtol 2e-2).  At 24-30 layers
    no two bf16 evaluations agree element-wise to rtol 2e-2: the reference's OWN autocast path deviates from its fp32
    path by max 3.3e-2 / mean 4.9e-3 on these logits (max|logit| 0.91; measured on the B200, tools/fullsize_noise.py),
    the engine by max 3.2e-2 / mean 4.6e-3.  So: (1) >= 99.9 % of the raw logits within rtol 2e-2 + max(2e-2 * max|ref|,
    2 x q99.9 of the reference's own bf16-vs-fp32 error) of the reference bf16 path (north_star's 2e-2; the second
    term only matters at the 7B architecture), (2) the engine's error against the fp32 reference not larger than
    1.25x the reference bf16 path's own error against fp32 (mean and max) - i.e. the engine is as good a bf16
    evaluation as the reference, (3) CFG logits u + 5 (c - u) = 5c - 4u: the same two criteria on the combined values
    (the bound scales with the measured error of the reference's own bf16 CFG logits against fp32, not with the 9x
    worst case)."""
    from plangen_b200 import synthetic
    pd = product_dims(d)
    cond, neg = synthetic.layoutsam_prompts(pd, B, seed=batch_seed, lo=150, hi=480)
    ids, mask = synthetic.collate_cfg_batch(cond, neg, pd.pad_id, pd.n_img_tokens)
    ids, mask = ids.cuda(), mask.cuda()
    tr16, tr32 = {}, {}
    ref_tok, _ = O.t2i(sd, d, ids, mask, sampler=O.greedy_sampler, mode="autocast", image_token_num_per_image=steps,
                       decode=False, trace=tr16)
    forced = torch.zeros(B, steps, dtype=torch.long)
    O.t2i(sd, d, ids, mask, sampler=O.greedy_sampler, mode="fp32", image_token_num_per_image=steps, decode=False,
          trace=tr32, edit_region=forced, gt_labels=ref_tok)
    ref_raw = torch.stack(tr16["raw_logits"]).numpy()
    ref_raw32 = torch.stack(tr32["raw_logits"]).numpy()
    ref_cfg = torch.stack(tr16["logits"]).numpy()
    ref_cfg32 = torch.stack(tr32["logits"]).numpy()
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
    e_mine, e_ref = np.abs(got_raw - ref_raw32), np.abs(ref_raw - ref_raw32)
    # additive term: north_star's 2e-2 of the logit range, or - where the reference's own bf16 noise is larger than that
    # (7B architecture: 30 layers x 4096 wide, its autocast path sits mean 1.3e-2 / max 1.1e-1 from its fp32 path on
    # logits of range 0.98, tools/debug_7b.py) - twice the 99.9 % quantile of that noise: two independent bf16
    # evaluations differ by sqrt(2) x the single-evaluation error
    add = max(2e-2 * np.abs(ref_raw).max(), 2.0 * np.quantile(e_ref, 0.999))
    tol = 2e-2 * np.abs(ref_raw) + add
    frac_ok = float((np.abs(got_raw - ref_raw) <= tol).mean())
    assert frac_ok >= 0.999, f"only {frac_ok:.5f} of the logits within rtol 2e-2 + {add:.3e}"
    assert e_mine.mean() <= 1.25 * e_ref.mean(), (e_mine.mean(), e_ref.mean())
    assert e_mine.max() <= 1.25 * e_ref.max(), (e_mine.max(), e_ref.max())
    # fused loop: CFG logits
    dbg = torch.zeros(steps, B, d.img_vocab, device="cuda")
    eng.set_option("dbg_logits_ptr", dbg.data_ptr())
    try:
        emb = eng.language_model.get_input_embeddings()(ids)
        fb = {"edit_region": torch.zeros(B, steps, dtype=torch.int32)}
        got = eng.sample_image(emb, B, steps, mask, 5.0, 1.0, generator=0, batch=fb, gt_labels=ref_tok, greedy=True,
                               use_teacher_forcing=True)
        torch.cuda.synchronize()
    finally:
        eng.set_option("dbg_logits_ptr", 0)
    assert got.cpu().tolist() == ref_tok.cpu().tolist()
    got_cfg = dbg.cpu().numpy()
    c_mine, c_ref = np.abs(got_cfg - ref_cfg32), np.abs(ref_cfg - ref_cfg32)
    assert c_mine.mean() <= 1.25 * c_ref.mean(), (c_mine.mean(), c_ref.mean())
    assert np.quantile(c_mine, cfg_quantile) <= 1.25 * np.quantile(c_ref, cfg_quantile), \
        (np.quantile(c_mine, cfg_quantile), np.quantile(c_ref, cfg_quantile))
    assert c_mine.max() <= 1.5 * c_ref.max(), (c_mine.max(), c_ref.max())
    # and directly against the reference's bf16 CFG logits: twice its own measured deviation from fp32 (two bf16
    # evaluations each that far from the truth), on top of north_star's relative 2e-2
    tol = 2e-2 * np.abs(ref_cfg) + 2.0 * np.quantile(c_ref, cfg_quantile)
    frac_ok = float((np.abs(got_cfg - ref_cfg) <= tol).mean())
    assert frac_ok >= 0.999, f"only {frac_ok:.5f} of the CFG logits within tolerance"


def test_fullsize_bf16_logits_vs_autocast_reference():
    """configs[1] architecture, B=4 of the 16 (keeps the oracle cheap), prefill + 4 decode steps."""
    eng, sd = _engine()
    _logits_parity(eng, sd, O.JANUS_1P3B, B=4, steps=5)


def test_fullsize_bf16_logits_bench_shape_b16():
    """The bench shape itself: B=16 (R=32 rows, the split-K / attention-cut schedules of the headline run), prefill +
    1 decode step against the autocast reference on the same GPU."""
    eng, sd = _engine()
    _logits_parity(eng, sd, O.JANUS_1P3B, B=16, steps=2, batch_seed=1239)


def test_fullsize_7b_bf16_logits_vs_autocast_reference():
    """BASELINE configs[4] architecture (Janus-Pro-7B: D=4096, L=30, H=32, F=11008 - other tile counts, split-K
    schedules and attention cuts than the 1.3B model), B=2, prefill + 4 decode steps.  VQ is shared with 1.3B: skipped."""
    from plangen_b200 import synthetic
    from plangen_b200.engine import FastJanus
    d = product_dims(O.JANUS_7B)
    sd = synthetic.random_state_dict(d, torch.device("cuda", 0), seed=0, with_vq=False)
    eng = FastJanus(sd, d, mode="bf16", max_batch=2, max_prompt=512, with_vq=False)
    try:
        _logits_parity(eng, sd, O.JANUS_7B, B=2, steps=5)
    finally:
        del eng, sd
        torch.cuda.empty_cache()


def test_fullsize_fp32_check_mode_greedy_tokens_identical():
    """BASELINE configs[0] at its full size: fp32 check mode, Janus-1.3B architecture, B=1 (R=2 rows), P=256, 40 greedy
    steps: token ids IDENTICAL to the oracle's fp32 path (run on this GPU, TF32 off) and CFG logits within 1e-4 -
    where the CUDA-core split-K summation order meets 24 layers."""
    from plangen_b200 import synthetic
    from plangen_b200.engine import FastJanus
    d = O.JANUS_1P3B
    pd = product_dims(d)
    _, sd = _engine()
    steps = 40
    cond, neg = synthetic.layoutsam_prompts(pd, 1, seed=4321, lo=256, hi=256)
    ids, mask = synthetic.collate_cfg_batch(cond, neg, pd.pad_id, pd.n_img_tokens)
    ids, mask = ids.cuda(), mask.cuda()
    assert ids.shape == (2, 256)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        trace = {}
        want, _ = O.t2i(sd, d, ids, mask, sampler=O.greedy_sampler, mode="fp32", image_token_num_per_image=steps,
                        decode=False, trace=trace)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    want_logits = torch.stack(trace["logits"]).float().cpu().numpy()
    eng = FastJanus(sd, pd, mode="fp32", max_batch=1, max_prompt=256, max_steps=64, with_vq=False)
    try:
        dbg = torch.zeros(steps, 1, d.img_vocab, device="cuda")
        eng.set_option("dbg_logits_ptr", dbg.data_ptr())
        emb = eng.language_model.get_input_embeddings()(ids)
        got = eng.sample_image(emb, 1, steps, mask, 5.0, 1.0, generator=0, greedy=True)
        torch.cuda.synchronize()
        eng.set_option("dbg_logits_ptr", 0)
        assert got.cpu().tolist() == want.cpu().tolist(), "fp32 greedy token ids differ from the oracle at 1.3B dims"
        assert_close(dbg.cpu().numpy(), want_logits, 1e-4, 1e-4, "fp32 CFG logits at 1.3B dims")
    finally:
        del eng
        torch.cuda.empty_cache()


def test_fullsize_loop_properties():
    """Full 576-token loop at B=16: deterministic across runs, identical with and without CUDA-graph replay,
    token ids in range."""
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
    # scheduling switches that move no arithmetic: the shared-prompt shortcuts (prefill once, K / V read from one copy) and
    # the down projection's ring depth - the same 96 tokens of all 16 images
    for opts in ({"prefill_dedup": 0}, {"attn_alias": 0}, {"down_stages": 0}):
        for k, v in opts.items():
            eng.set_option(k, v)
        try:
            e2 = eng.sample_image(emb, B, 96, mask, 5.0, 1.0, generator=0).cpu()
        finally:
            for k in opts:
                eng.set_option(k, {"prefill_dedup": 1, "attn_alias": 1, "down_stages": 4}[k])
        assert torch.equal(a[:, :96], e2), opts
    # a different seed gives a different sample; teacher forcing reproduces the labels exactly
    s2 = eng.sample_image(emb, B, 32, mask, 5.0, 1.0, generator=1).cpu()
    assert not torch.equal(a[:, :32], s2)
    forced = {"edit_region": torch.zeros(B, 32, dtype=torch.int32)}
    gt = torch.randint(0, d.img_vocab, (B, 32), dtype=torch.int32)
    f = eng.sample_image(emb, B, 32, mask, 5.0, 1.0, generator=0, batch=forced, gt_labels=gt, use_teacher_forcing=True).cpu()
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
