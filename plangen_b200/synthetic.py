"""Synthetic weights and prompts for benchmarks and smoke tests (no checkpoints / tokenizer / dataset
are available offline).  Weights: random init of the named architecture, generated directly on the
device (HF init N(0, 0.02) for the LM, torch default U(+-1/sqrt(fan_in)) for gen_head / gen_aligner /
conv, N(0,1) gen_embed, L2-normalised codebook).  Prompts: LayoutSAM-shaped token-id lists
(SURVEY.md §8d): caption 20-60 tokens + 4-8 boxes x (8-30 description + ~20 markup) + ~12 template
tokens; one shared ~110-token negative prompt (cfg/base.py:129); LEFT padded, rows interleaved
[cond0, neg0, cond1, neg1, ...] and 576 ones appended to the mask (plangen_base.py:636-725)."""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

from .config import Dims


def state_dict_names(d: Dims, with_vq: bool = True, with_vq_encoder: bool = False) -> List[Tuple[str, Tuple[int, ...], str]]:
    s: List[Tuple[str, Tuple[int, ...], str]] = []
    lm = "language_model.model."
    HD = d.H * d.head_dim
    s.append((lm + "embed_tokens.weight", (d.vocab, d.D), "lm"))
    for i in range(d.L):
        l = lm + f"layers.{i}."
        s += [(l + "input_layernorm.weight", (d.D,), "norm_w"),
              (l + "self_attn.q_proj.weight", (HD, d.D), "lm"), (l + "self_attn.k_proj.weight", (HD, d.D), "lm"),
              (l + "self_attn.v_proj.weight", (HD, d.D), "lm"), (l + "self_attn.o_proj.weight", (d.D, HD), "lm"),
              (l + "post_attention_layernorm.weight", (d.D,), "norm_w"),
              (l + "mlp.gate_proj.weight", (d.F, d.D), "lm"), (l + "mlp.up_proj.weight", (d.F, d.D), "lm"),
              (l + "mlp.down_proj.weight", (d.D, d.F), "lm")]
    s.append((lm + "norm.weight", (d.D,), "norm_w"))
    s += [("gen_head.output_mlp_projector.weight", (d.img_embed, d.D), "fan"),
          ("gen_head.output_mlp_projector.bias", (d.img_embed,), "bias:%d" % d.D),
          ("gen_head.vision_head.weight", (d.img_vocab, d.img_embed), "fan"),
          ("gen_head.vision_head.bias", (d.img_vocab,), "bias:%d" % d.img_embed),
          ("gen_embed.weight", (d.img_vocab, d.code_dim), "normal1"),
          ("gen_aligner.layers.0.weight", (d.D, d.code_dim), "fan"),
          ("gen_aligner.layers.0.bias", (d.D,), "bias:%d" % d.code_dim),
          ("gen_aligner.layers.2.weight", (d.D, d.D), "fan"),
          ("gen_aligner.layers.2.bias", (d.D,), "bias:%d" % d.D)]
    if not with_vq:
        return s
    p = "gen_vision_model."
    s.append((p + "quantize.embedding.weight", (d.img_vocab, d.code_dim), "codebook"))
    s.append((p + "post_quant_conv.weight", (d.vq_z, d.code_dim, 1, 1), "fan"))
    s.append((p + "post_quant_conv.bias", (d.vq_z,), "bias:%d" % d.code_dim))
    nres = len(d.vq_ch_mult)
    block_in = d.vq_ch * d.vq_ch_mult[nres - 1]

    def conv(n, cin, cout, k):
        s.append((n + ".weight", (cout, cin, k, k), "fan"))
        s.append((n + ".bias", (cout,), "bias:%d" % (cin * k * k)))

    def norm(n, c):
        s.append((n + ".weight", (c,), "norm_w"))
        s.append((n + ".bias", (c,), "norm_b"))

    def res(n, cin, cout):
        norm(n + ".norm1", cin); conv(n + ".conv1", cin, cout, 3)
        norm(n + ".norm2", cout); conv(n + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(n + ".nin_shortcut", cin, cout, 1)

    def attn(n, c):
        norm(n + ".norm", c)
        for t in ("q", "k", "v", "proj_out"):
            conv(n + "." + t, c, c, 1)

    dp = p + "decoder."
    conv(dp + "conv_in", d.vq_z, block_in, 3)
    res(dp + "mid.0", block_in, block_in); attn(dp + "mid.1", block_in); res(dp + "mid.2", block_in, block_in)
    for idx, i_level in enumerate(reversed(range(nres))):
        block_out = d.vq_ch * d.vq_ch_mult[i_level]
        for j in range(d.vq_res_blocks + 1):
            res(dp + f"conv_blocks.{idx}.res.{j}", block_in, block_out)
            block_in = block_out
            if i_level == nres - 1:
                attn(dp + f"conv_blocks.{idx}.attn.{j}", block_in)
        if i_level != 0:
            conv(dp + f"conv_blocks.{idx}.upsample.conv", block_in, block_in, 3)
    norm(dp + "norm_out", block_in)
    conv(dp + "conv_out", block_in, 3, 3)
    if with_vq_encoder:      # encode side of the editing path (gen_vision_model.encode): Encoder + quant_conv
        ep = p + "encoder."
        conv(ep + "conv_in", 3, d.vq_ch, 3)
        in_mult = (1,) + tuple(d.vq_ch_mult)
        cin = d.vq_ch
        for lvl in range(nres):
            cin, cout = d.vq_ch * in_mult[lvl], d.vq_ch * d.vq_ch_mult[lvl]
            for j in range(d.vq_res_blocks):
                res(ep + f"conv_blocks.{lvl}.res.{j}", cin, cout)
                cin = cout
                if lvl == nres - 1:
                    attn(ep + f"conv_blocks.{lvl}.attn.{j}", cin)
            if lvl != nres - 1:
                conv(ep + f"conv_blocks.{lvl}.downsample.conv", cin, cin, 3)
        res(ep + "mid.0", cin, cin); attn(ep + "mid.1", cin); res(ep + "mid.2", cin, cin)
        norm(ep + "norm_out", cin)
        conv(ep + "conv_out", cin, d.vq_z, 3)
        conv(p + "quant_conv", d.vq_z, d.code_dim, 1)
    return s


def vision_state_dict_names(d: Dims) -> List[Tuple[str, Tuple[int, ...], str]]:
    """vision_model.vision_tower.* (VisionTransformer, siglip_vit.py:262-440: no class token, qkv bias, LayerNorm affine)
    and the understanding `aligner` (projector.py:39-45)."""
    p = "vision_model.vision_tower."
    W, hid, pp = d.sig_width, d.sig_mlp, d.sig_patch
    s: List[Tuple[str, Tuple[int, ...], str]] = [
        (p + "pos_embed", (1, d.sig_patches, W), "lm"),
        (p + "patch_embed.proj.weight", (W, 3, pp, pp), "fan"),
        (p + "patch_embed.proj.bias", (W,), "bias:%d" % (3 * pp * pp))]
    for i in range(d.sig_layers):
        b = p + f"blocks.{i}."
        s += [(b + "norm1.weight", (W,), "norm_w"), (b + "norm1.bias", (W,), "norm_b"),
              (b + "attn.qkv.weight", (3 * W, W), "lm"), (b + "attn.qkv.bias", (3 * W,), "norm_b"),
              (b + "attn.proj.weight", (W, W), "lm"), (b + "attn.proj.bias", (W,), "norm_b"),
              (b + "norm2.weight", (W,), "norm_w"), (b + "norm2.bias", (W,), "norm_b"),
              (b + "mlp.fc1.weight", (hid, W), "lm"), (b + "mlp.fc1.bias", (hid,), "norm_b"),
              (b + "mlp.fc2.weight", (W, hid), "lm"), (b + "mlp.fc2.bias", (W,), "norm_b")]
    s += [(p + "norm.weight", (W,), "norm_w"), (p + "norm.bias", (W,), "norm_b"),
          ("aligner.layers.0.weight", (d.D, W), "fan"), ("aligner.layers.0.bias", (d.D,), "bias:%d" % W),
          ("aligner.layers.2.weight", (d.D, d.D), "fan"), ("aligner.layers.2.bias", (d.D,), "bias:%d" % d.D)]
    return s


def random_state_dict(d: Dims, device, seed: int = 0, with_vq: bool = True, with_lm_head: bool = False,
                      with_vq_encoder: bool = False, with_vision: bool = False) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device).manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    names = state_dict_names(d, with_vq, with_vq and with_vq_encoder)
    if with_vision:
        names = names + vision_state_dict_names(d)
    if with_lm_head:       # untied text head, only needed by language_model.generate (stage-1 layout-text decode)
        names = names + [("language_model.lm_head.weight", (d.vocab, d.D), "lm")]
    for name, shape, kind in names:
        t = torch.empty(shape, device=device, dtype=torch.float32)
        if kind == "lm":
            t.normal_(0.0, 0.02, generator=g)
        elif kind == "norm_w":
            t.normal_(0.0, 0.1, generator=g).add_(1.0)
        elif kind == "norm_b":
            t.normal_(0.0, 0.1, generator=g)
        elif kind == "normal1":
            t.normal_(0.0, 1.0, generator=g)
        elif kind == "codebook":
            t.uniform_(-1.0 / shape[0], 1.0 / shape[0], generator=g)
            t = torch.nn.functional.normalize(t, p=2, dim=-1)
        elif kind == "fan":
            b = 1.0 / math.sqrt(int(math.prod(shape[1:])))
            t.uniform_(-b, b, generator=g)
        else:
            b = 1.0 / math.sqrt(int(kind.split(":")[1]))
            t.uniform_(-b, b, generator=g)
        sd[name] = t
    return sd


def layoutsam_prompts(d: Dims, batch: int, seed: int = 1234, lo: int = 150, hi: int = 480, neg_len: int = 110):
    g = torch.Generator(device="cpu").manual_seed(seed)

    def ri(a, b):
        return int(torch.randint(a, b + 1, (1,), generator=g).item())

    def ids(n):
        t = torch.randint(0, d.vocab - 1, (n,), generator=g)
        return torch.where(t >= d.pad_id, t + 1, t).clamp_(max=d.vocab - 1).tolist()

    neg = ids(neg_len)
    cond = []
    for _ in range(batch):
        n = ri(20, 60) + 12
        for _ in range(ri(4, 8)):
            n += ri(8, 30) + 20
        cond.append(ids(max(lo, min(hi, n))))
    return cond, [list(neg) for _ in range(batch)]


def collate_cfg_batch(cond: List[List[int]], neg: List[List[int]], pad_id: int, n_img_tokens: int):
    """Host mirror of t2i_infer_collate_batch + pad_input_ids (plangen_base.py:636-725) on token-id
    lists: returns ids (2B, P) int32 and mask (2B, P + n_img_tokens) int32."""
    bs = len(cond)
    P = max(max(map(len, cond)), max(map(len, neg)))
    ids = torch.full((bs, 2, P), pad_id, dtype=torch.int32)
    mask = torch.zeros((bs, 2, P + n_img_tokens), dtype=torch.int32)
    mask[:, :, P:] = 1
    for i in range(bs):
        for j, seq in enumerate((cond[i], neg[i])):
            n = len(seq)
            if n:
                ids[i, j, P - n:] = torch.tensor(seq, dtype=torch.int32)
                mask[i, j, P - n:P] = 1
    return ids.view(bs * 2, P), mask.view(bs * 2, -1)
