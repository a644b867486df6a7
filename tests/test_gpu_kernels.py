"""-m gpu: unit parity of the individual CUDA kernels through the C-ABI."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import janus_oracle as O
from oracle import philox as PX
from tests.gpu_util import get_engine, assert_close, rel_err

pytestmark = pytest.mark.gpu


def _gemm(eng, impl, X, W, splits):
    from plangen_b200 import _lib
    M, K = X.shape
    N = W.shape[0]
    out = torch.empty(splits, M, N, device=X.device, dtype=torch.float32)
    _lib.check(eng._lib.pg_test_gemm(eng._h, impl, int(X.dtype == torch.bfloat16), C.c_void_p(X.data_ptr()),
                                     C.c_void_p(W.data_ptr()), M, N, K, splits, C.c_void_p(out.data_ptr()),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return out.sum(0)


@pytest.mark.parametrize("M,N,K,splits", [
    (2, 128, 64, 1), (2, 256, 256, 1), (16, 256, 512, 2), (32, 384, 2048, 3), (5, 130, 72, 1),
    (64, 1024, 1408, 4), (128, 512, 512, 1), (200, 256, 320, 1), (576, 576, 512, 1), (1000, 3, 1152, 1),
    (32, 6144, 2048, 3), (32, 2048, 5632, 9),
    # ring reuse with every token-tile width (the ring depth must stay even, see gemm_tc_kernel's invariant)
    (8, 256, 4096, 1), (16, 384, 5632, 1), (48, 256, 4096, 1), (100, 256, 2048, 1), (192, 256, 2304, 1), (300, 128, 4096, 1),
    # CTA-pair kernel (gemm_tc2.cuh: M > 128, N > 128, no split): odd weight-tile counts, ragged M / N / K, ring reuse
    (200, 256, 320, 1), (576, 576, 512, 1), (257, 384, 1024, 1), (1000, 1024, 4096, 1), (513, 200, 72, 1), (2048, 3072, 1024, 1),
])
def test_gemm_tcgen05_matches_torch(M, N, K, splits):
    eng = get_engine(O.TINY, "bf16")
    g = torch.Generator(device="cuda").manual_seed(M * 7919 + N * 31 + K)
    X = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    W = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    got = _gemm(eng, 1, X, W, splits)
    want = X.double() @ W.double().T            # bf16 products are exact; only the fp32 summation order differs
    assert_close(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol_frac=2e-6, what=f"tc gemm {M}x{N}x{K}/{splits}")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("M,N,K,splits", [(2, 64, 16, 1), (3, 70, 8, 1), (33, 200, 1024, 4), (130, 96, 72, 2)])
def test_gemm_simt_matches_torch(mode, M, N, K, splits):
    eng = get_engine(O.TINY, mode)
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    X = torch.randn(M, K, device="cuda", generator=g).to(dt)
    W = torch.randn(N, K, device="cuda", generator=g).to(dt)
    got = _gemm(eng, 0, X, W, splits)
    want = X.double() @ W.double().T
    assert_close(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-5, atol_frac=2e-6, what=f"simt gemm {mode}")


@pytest.mark.parametrize("B,V", [(1, 2048), (3, 16384), (16, 16384), (20, 16384)])
def test_sampler_is_bit_exact_with_torch_multinomial(B, V):
    """Index work: the fused sampler must pick exactly the tokens torch.multinomial picks on this GPU
    for the same (seed, offset) and the same probabilities, over several consecutive calls."""
    d = O.JanusDims(**{**O.TINY.__dict__, "name": f"tiny-v{V}", "img_vocab": V})
    eng = get_engine(d, "fp32", max_batch=max(B, 4), with_vq=False)
    props = torch.cuda.get_device_properties(0)
    gen = torch.Generator(device="cuda").manual_seed(1234)
    tg = torch.Generator(device="cuda").manual_seed(99)
    n_steps = 5
    toks = torch.zeros(B, n_steps, dtype=torch.int32, device="cuda")
    off = 0
    for step in range(n_steps):
        logits = torch.randn(2 * B, V, device="cuda", generator=tg) * 2.0
        cfg = logits[1::2] + 5.0 * (logits[0::2] - logits[1::2])
        probs = torch.softmax(cfg / 1.0, dim=-1)
        assert gen.get_offset() == off
        want = torch.multinomial(probs, 1, generator=gen).squeeze(-1)
        # oracle's CPU port of the CUDA algorithm
        port, off2 = PX.cuda_multinomial1(probs.cpu().numpy(), 1234, off, props.multi_processor_count,
                                          props.max_threads_per_multi_processor)
        eng.cfg_sample_embed(logits, 5.0, 1.0, 1234, off, step, n_steps, toks)
        torch.cuda.synchronize()
        assert toks[:, step].tolist() == want.tolist(), f"step {step}"
        assert port.tolist() == want.tolist(), f"oracle port, step {step}"
        off = gen.get_offset()
        assert off2 == off and eng.philox_offset_per_step(B) == off2 // (step + 1)


@pytest.mark.parametrize("B,V,k", [(3, 16384, 50), (2, 16384, 1), (4, 2048, 2047), (16, 16384, 1000)])
def test_sampler_top_k_matches_torch_topk_masked_multinomial(B, V, k):
    """north_star (4) top-k: `logits[logits < topk(logits, k).values[..., -1:]] = -inf` before the softmax, then the same
    Philox draw: tokens equal torch.multinomial on the masked distribution for the same (seed, offset); every sampled
    token lies inside the top-k set; top_k = 0 and top_k >= V reproduce the unmasked sampler bit for bit."""
    d = O.JanusDims(**{**O.TINY.__dict__, "name": f"tiny-v{V}", "img_vocab": V})
    eng = get_engine(d, "fp32", max_batch=max(B, 4), with_vq=False)
    gen = torch.Generator(device="cuda").manual_seed(77)
    tg = torch.Generator(device="cuda").manual_seed(5 + k)
    n_steps = 4
    toks = torch.zeros(B, n_steps, dtype=torch.int32, device="cuda")
    off = 0
    for step in range(n_steps):
        logits = torch.randn(2 * B, V, device="cuda", generator=tg) * 2.0
        if step == 1:                                       # ties at the threshold are kept (like the torch idiom)
            logits[:, : V // 2] = logits[:, V // 2: 2 * (V // 2)]
        cfg = logits[1::2] + 5.0 * (logits[0::2] - logits[1::2])
        kth = torch.topk(cfg, k, dim=-1).values[:, -1:]
        masked = cfg.masked_fill(cfg < kth, float("-inf"))
        probs = torch.softmax(masked / 1.0, dim=-1)
        want = torch.multinomial(probs, 1, generator=gen).squeeze(-1)
        eng.cfg_sample_embed(logits, 5.0, 1.0, 77, off, step, n_steps, toks, top_k=k)
        torch.cuda.synchronize()
        assert toks[:, step].tolist() == want.tolist(), f"step {step}"
        assert bool((cfg.gather(1, toks[:, step:step + 1].long()) >= kth).all())
        off = gen.get_offset()
    # off / degenerate settings are the plain sampler
    logits = torch.randn(2 * B, V, device="cuda", generator=tg)
    outs = []
    for kk in (0, V, V + 5):
        t = torch.zeros(B, 1, dtype=torch.int32, device="cuda")
        eng.cfg_sample_embed(logits, 5.0, 1.0, 3, 0, 0, 1, t, top_k=kk)
        outs.append(t.cpu())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_sampler_teacher_forcing_and_greedy():
    eng = get_engine(O.TINY, "fp32")
    B, V, n = 2, O.TINY.img_vocab, 3
    logits = torch.randn(2 * B, V, device="cuda")
    toks = torch.zeros(B, n, dtype=torch.int32, device="cuda")
    er = torch.tensor([[1, 0, 1], [0, 0, 1]], dtype=torch.int32, device="cuda")
    gt = torch.tensor([[7, 8, 9], [17, 18, 19]], dtype=torch.int32, device="cuda")
    x = eng.cfg_sample_embed(logits, 5.0, 1.0, 0, 0, 1, n, toks, greedy=True, edit_region=er, gt_labels=gt)
    assert toks[:, 1].tolist() == [8, 18]
    x2 = eng.cfg_sample_embed(logits, 5.0, 1.0, 0, 0, 2, n, toks, greedy=True, edit_region=er, gt_labels=gt)
    cfg = logits[1::2] + 5.0 * (logits[0::2] - logits[1::2])
    assert toks[:, 2].tolist() == cfg.argmax(-1).tolist()
    # next inputs are gen_aligner(gen_embed(tok)) duplicated to the cond/uncond rows
    sd = {k: v.cuda() for k, v in O.init_state_dict(O.TINY, seed=0, with_vq=False).items() if k.startswith("gen_")}
    want = O.prepare_gen_img_embeds(sd, toks[:, 2].long())
    assert_close(x2[0::2].cpu().numpy(), want.cpu().numpy(), 1e-4, 1e-5, "embed cond rows")
    assert torch.equal(x2[0::2], x2[1::2])


@pytest.mark.parametrize("lens", [[3, 150, 290, 399, 129, 128, 257, 1], [128, 128], [256, 1, 255], [1], [384, 383, 385, 127]])
def test_prefill_attention_tensor_core_path_long_ragged_prompts(lens):
    """Prompt prefill through the tcgen05 attention kernel (attn_prefill_tc.cuh) on prompts that span several
    128-query tiles and 128-key blocks: rows whose left padding ends inside the first, a middle and the last key
    block, a row with no padding, and tiles that are all padding.  Final hidden states of the valid positions vs
    (a) the reference PyTorch path (fp32 master weights under autocast bf16) on the same GPU within the bf16
    tolerance and (b) the CUDA-core attention kernel of the same engine (two bf16 evaluations of the same
    arithmetic: closer to each other than either is required to be to the reference)."""
    import numpy as np
    from oracle import janus_oracle as O
    from tests.gpu_util import get_engine, assert_close
    d = O.SMALL
    g = torch.Generator().manual_seed(21)
    prompts = [torch.randint(0, d.pad_id, (n,), generator=g).tolist() for n in lens]
    ids, mask = O.pad_input_ids(prompts, d.pad_id)
    P = ids.shape[1]
    assert P == max(lens)
    sd = O.init_state_dict(d, seed=0, with_vq=False)
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
        ref = O.llama_model_forward(sdc, d, O.embed_tokens(sdc, ids.cuda()), attention_mask=mask.cuda()).last_hidden_state.float()
    eng = get_engine(d, "bf16", max_batch=4, max_prompt=400, with_vq=False)
    emb = eng.language_model.get_input_embeddings()(ids.cuda())
    outs = {}
    for tc in (1, 0):
        eng.set_option("prefill_attn_tc", tc)
        try:
            outs[tc] = eng.language_model.model(inputs_embeds=emb, attention_mask=mask.cuda(), use_cache=True).last_hidden_state.float()
            torch.cuda.synchronize()
        finally:
            eng.set_option("prefill_attn_tc", 1)
    valid = (mask.cuda() != 0)
    r, a, b = ref[valid].cpu().numpy(), outs[1][valid].cpu().numpy(), outs[0][valid].cpu().numpy()
    assert np.isfinite(outs[1].cpu().numpy()).all()
    assert_close(a, r, 2e-2, 3e-2, "prefill hidden states, tcgen05 attention vs autocast reference")
    assert_close(b, r, 2e-2, 3e-2, "prefill hidden states, CUDA-core attention vs autocast reference")
    assert np.abs(a - b).mean() <= 1.5 * max(np.abs(a - r).mean(), np.abs(b - r).mean())
