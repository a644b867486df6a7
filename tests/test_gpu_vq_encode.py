"""-m gpu: VQ encode side of the editing path (`vl_gpt.gen_vision_model.encode(img)[-1][-1]`, plangen_base.py:532;
SURVEY §8f rank 3) through the C-ABI (pg_vq_encode) against the goldens made with the reference's own classes and the
oracle restatement."""
import os

import numpy as np
import pytest
import torch

from oracle import janus_oracle as O
from tests.gpu_util import product_dims

pytestmark = pytest.mark.gpu

_ENG = {}


def _engine(d, mode):
    from plangen_b200.engine import FastJanus
    key = (d.name, mode)
    if key not in _ENG:
        sd = O.init_state_dict(d, seed=0, with_vq=True, with_vq_encoder=True)
        _ENG[key] = (FastJanus(sd, product_dims(d), mode=mode, max_batch=4, max_prompt=64, with_vq=True), sd)
    return _ENG[key]


def _excess_distance(sd, z_ref, idx):
    """fp64 distance of the chosen code to the reference's normalised z, minus the distance of the best code."""
    zf = torch.einsum("b c h w -> b h w c", z_ref.double()).reshape(-1, z_ref.shape[1])
    zf = zf / zf.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    emb = sd["gen_vision_model.quantize.embedding.weight"].double().to(zf.device)
    emb = emb / emb.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    dist = 2.0 - 2.0 * zf @ emb.T
    chosen = dist.gather(1, idx.reshape(-1, 1).long().to(zf.device)).squeeze(1)
    return (chosen - dist.min(1).values).cpu().numpy()


@pytest.mark.parametrize("name,dims", [("vqenc_tiny.npz", O.TINY), ("vqenc_small.npz", O.SMALL)])
def test_fp32_encode_matches_reference_classes_golden(golden_dir, name, dims):
    """fp32 check mode: code indices identical to the reference's Encoder + quant_conv + VectorQuantizer; where a
    position differs (fp32 summation order of a near-tie) the chosen code must be as near as the reference's to 1e-6."""
    g = np.load(os.path.join(golden_dir, name))
    eng, sd = _engine(dims, "fp32")
    img = torch.from_numpy(g["img"]).cuda()
    out = eng.gen_vision_model.encode(img)
    idx = out[-1][-1]
    assert idx.dtype == torch.int64 and idx.shape == (g["indices"].size,)
    same = (idx.cpu().numpy() == g["indices"])
    assert same.mean() >= 0.99, f"only {same.mean():.4f} of the indices agree"
    exc = _excess_distance(sd, torch.from_numpy(g["z"]), idx.cpu())
    assert exc.max() <= 1e-6, exc.max()


@pytest.mark.parametrize("dims", [O.TINY, O.SMALL])
def test_bf16_encode_vs_autocast_reference(dims):
    """Reference regime (autocast bf16) on the same GPU.  The nearest-code search runs on bf16-rounded inner products,
    so near-ties are common and an index can legitimately differ; what must hold is that every chosen code is
    (almost) as near to the fp32 reference's z as the best code, and that most indices agree with the reference's
    own bf16 run."""
    eng, sd = _engine(dims, "bf16")
    sdc = {k: v.cuda() for k, v in sd.items() if k.startswith("gen_vision_model.")}
    g = torch.Generator().manual_seed(9)
    side = dims.grid * 2 ** (len(dims.vq_ch_mult) - 1)
    img = (torch.rand(3, 3, side, side, generator=g) * 2 - 1).cuda()
    with torch.inference_mode():
        z32 = O.vq_encoder_forward(sdc, dims, img)
    ref16 = O.vq_encode(sdc, dims, img, mode="autocast")
    idx = eng.gen_vision_model.encode(img)[-1][-1]
    exc_mine = _excess_distance(sd, z32.cpu(), idx.cpu())
    exc_ref = _excess_distance(sd, z32.cpu(), ref16.cpu())
    agree = float((idx == ref16).float().mean())
    print(f"agree {agree:.3f} excess mine mean {exc_mine.mean():.4g} max {exc_mine.max():.4g} | ref mean {exc_ref.mean():.4g} max {exc_ref.max():.4g}")
    assert exc_mine.mean() <= 1.5 * exc_ref.mean() + 1e-3 and exc_mine.max() <= 1.5 * exc_ref.max() + 2e-2
    assert agree >= 0.5


def test_encode_properties_and_errors():
    from plangen_b200.engine import FastJanus
    from plangen_b200 import _lib
    d = O.SMALL
    eng, sd = _engine(d, "bf16")
    g = torch.Generator().manual_seed(3)
    img = (torch.rand(6, 3, 16, 24, generator=g) * 2 - 1).cuda()        # 6 images > one 4-image chunk, non-square
    a = eng.gen_vision_model.encode(img)[-1][-1]
    b = eng.gen_vision_model.encode(img)[-1][-1]
    down = 2 ** (len(d.vq_ch_mult) - 1)
    assert a.shape == (6 * (16 // down) * (24 // down),) and torch.equal(a, b)
    assert int(a.min()) >= 0 and int(a.max()) < d.img_vocab
    one = eng.gen_vision_model.encode(img[4:5])[-1][-1]
    n = (16 // down) * (24 // down)
    assert torch.equal(one, a[4 * n:5 * n]), "result depends on the batch composition / chunking"
    with pytest.raises(_lib.PgError, match="multiples"):
        eng.gen_vision_model.encode(torch.zeros(1, 3, 18, 24, device="cuda"))
    no_enc = FastJanus(O.init_state_dict(d, seed=0, with_vq=True), product_dims(d), mode="bf16", max_batch=2, max_prompt=64)
    with pytest.raises(_lib.PgError, match="encoder"):
        no_enc.gen_vision_model.encode(torch.zeros(1, 3, 24, 24, device="cuda"))


def test_fullsize_vq16_encode_vs_autocast_reference():
    """The real VQ-16 encoder shapes (ch 128, mult (1,1,2,2,4), z 256, codebook 16384 x 8) on 384 x 384 images - the
    editing path's `encode(gt_image)` - vs the reference PyTorch path under autocast(bf16) on the same GPU; the LM
    part of the engine is tiny (it plays no role here)."""
    import dataclasses
    import time
    from plangen_b200.engine import FastJanus
    d = dataclasses.replace(O.TINY, name="tiny-lm-vq16", img_vocab=16384, img_embed=256, grid=24, vq_ch=128,
                            vq_ch_mult=(1, 1, 2, 2, 4), vq_z=256)
    sd = O.init_state_dict(d, seed=0, with_vq=True, with_vq_encoder=True)
    eng = FastJanus(sd, product_dims(d), mode="bf16", max_batch=2, max_prompt=64, max_steps=16, with_vq=True)
    sdc = {k: v.cuda() for k, v in sd.items() if k.startswith("gen_vision_model.")}
    g = torch.Generator().manual_seed(17)
    # smooth images (random low-resolution pattern upsampled) in [-1, 1]
    img = torch.nn.functional.interpolate(torch.rand(2, 3, 24, 24, generator=g) * 2 - 1, size=(384, 384), mode="bilinear").cuda()
    with torch.inference_mode():
        z32 = O.vq_encoder_forward(sdc, d, img)
    ref16 = O.vq_encode(sdc, d, img, mode="autocast")
    idx = eng.gen_vision_model.encode(img)[-1][-1]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.gen_vision_model.encode(img)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert idx.shape == (2 * 576,)
    exc_mine = _excess_distance(sd, z32.cpu(), idx.cpu())
    exc_ref = _excess_distance(sd, z32.cpu(), ref16.cpu())
    agree = float((idx == ref16).float().mean())
    print(f"VQ-16 encode 2 x 384^2: {dt * 1e3:.1f} ms; agree {agree:.3f}; excess mine mean {exc_mine.mean():.4g} max {exc_mine.max():.4g} | "
          f"ref mean {exc_ref.mean():.4g} max {exc_ref.max():.4g}")
    assert exc_mine.mean() <= 1.5 * exc_ref.mean() + 1e-3 and exc_mine.max() <= 1.5 * exc_ref.max() + 2e-2
    assert agree >= 0.5
    # round trip through the decoder: codes -> image -> codes is stable in shape / range (no identity for random weights)
    dec = eng.gen_vision_model.decode_code(idx.reshape(2, 576).int(), shape=[2, d.code_dim, 24, 24])
    again = eng.gen_vision_model.encode(dec.float())[-1][-1]
    assert again.shape == idx.shape and int(again.min()) >= 0 and int(again.max()) < d.img_vocab


def test_t2i_editing_branch_teacher_forcing_from_gt_image():
    """System.t2i with use_teacher_forcing (plangen_base.py:528-532, :593-598, :557-562): gt_image is VQ-encoded on
    the device, positions whose edit_region is 0 take the ground-truth code, the others are sampled; edit_region all 0
    reproduces encode(gt_image), all 1 reproduces the free run; the returned mask image is the upscaled edit region."""
    d = O.SMALL
    eng, sd = _engine(d, "bf16")
    B, n, side = 3, d.n_img_tokens, d.grid * 2 ** (len(d.vq_ch_mult) - 1)
    cond, neg = O.synthetic_prompts(d, B, seed=5, lo=9, hi=40, neg_len=13)
    ids, mask = O.t2i_infer_collate_batch(cond, neg, d.pad_id, n)
    g = torch.Generator().manual_seed(2)
    gt_image = (torch.rand(B, 3, side, side, generator=g) * 2 - 1).cuda()
    labels = eng.gen_vision_model.encode(gt_image.to(torch.bfloat16))[-1][-1].reshape(B, -1)
    free, _ = eng.t2i(tokens=ids.cuda(), mask=mask.cuda()), None
    free_tok = eng.last_tokens.clone()
    dec0, m0 = eng.t2i(tokens=ids.cuda(), mask=mask.cuda(), gt_image=gt_image, batch={"edit_region": torch.zeros(B, n, dtype=torch.int64)},
                      use_teacher_forcing=True)
    assert torch.equal(eng.last_tokens.long(), labels) and m0.shape == dec0.shape and float(m0.abs().max()) == 0.0
    dec1, m1 = eng.t2i(tokens=ids.cuda(), mask=mask.cuda(), gt_image=gt_image, batch={"edit_region": torch.ones(B, n, dtype=torch.int64)},
                      use_teacher_forcing=True)
    assert torch.equal(eng.last_tokens, free_tok) and float(m1.min()) == 1.0
    er = (torch.rand(B, n, generator=g) > 0.5).long()
    eng.t2i(tokens=ids.cuda(), mask=mask.cuda(), gt_image=gt_image, batch={"edit_region": er}, use_teacher_forcing=True)
    got = eng.last_tokens.long().cpu()
    keep = er == 0
    assert torch.equal(got[keep], labels.cpu()[keep])


def test_t2i_ignores_edit_region_when_teacher_forcing_is_off():
    """the reference gates teacher forcing on args.use_teacher_forcing (plangen_base.py:528,556,593),
    and its datasets emit an all-zero edit_region for plain layout-to-image samples (data_hico.py:355,367) while
    uni_generate always passes gt_image.  With the flag off the call must be a free generation: same tokens as without
    gt_image / batch, no mask image.  With the flag on and parallel_size = 2, only the first bs rows are overridden."""
    d = O.SMALL
    eng, sd = _engine(d, "bf16")
    B, n, side = 2, d.n_img_tokens, d.grid * 2 ** (len(d.vq_ch_mult) - 1)
    cond, neg = O.synthetic_prompts(d, B, seed=6, lo=9, hi=40, neg_len=13)
    ids, mask = O.t2i_infer_collate_batch(cond, neg, d.pad_id, n)
    g = torch.Generator().manual_seed(3)
    gt_image = (torch.rand(B, 3, side, side, generator=g) * 2 - 1).cuda()
    batch = {"edit_region": torch.zeros(B, n, dtype=torch.int64)}
    eng.t2i(tokens=ids.cuda(), mask=mask.cuda())
    free_tok = eng.last_tokens.clone()
    dec, mask_image = eng.t2i(tokens=ids.cuda(), mask=mask.cuda(), gt_image=gt_image, batch=batch)
    assert mask_image is None and torch.equal(eng.last_tokens, free_tok)
    labels = eng.gen_vision_model.encode(gt_image.to(torch.bfloat16))[-1][-1].reshape(B, -1)
    assert not torch.equal(free_tok.long(), labels)
    # flag on, two parallel copies: rows 0..B-1 reproduce the labels, rows B..2B-1 are sampled freely
    eng.t2i(tokens=ids.cuda(), mask=mask.cuda(), parallel_size=2, gt_image=gt_image, batch=batch, use_teacher_forcing=True)
    tok = eng.last_tokens.long()
    assert tok.shape == (2 * B, n) and torch.equal(tok[:B], labels) and not torch.equal(tok[B:], labels)
    # mask image is binary (resize_pt rounds back to integers, plangen_base.py:560)
    er = (torch.rand(B, n, generator=g) > 0.5).long()
    _, m = eng.t2i(tokens=ids.cuda(), mask=mask.cuda(), gt_image=gt_image, batch={"edit_region": er}, use_teacher_forcing=True)
    assert set(m.float().unique().tolist()) <= {0.0, 1.0}
