"""Packing of the reference state_dict (names in SURVEY.md §8b) into the device tensors the engine
registers through `pg_engine_set_tensor`.

  language_model.model.layers.{i}.self_attn.{q,k,v}_proj.weight -> l{i}.wqkv  [3*H*128, D]  (rows q | k | v)
  ...mlp.{gate,up}_proj.weight                                   -> l{i}.wgu   [2*F, D]      (fp32: rows gate | up;
                                                                    bf16: interleaved in blocks of 64 rows)
  conv weights [Cout, Cin, kh, kw]                               -> [Cout, kh*kw*Cin]        (channels-last taps)
  vision_model.vision_tower.* / aligner.* (mmu front-end)        -> sig.* / ualign.*          (nn.Linear layout kept)

Weights are stored in the engine's operand type: bf16 (round-to-nearest-even of the fp32 master
weights - exactly the cast torch.autocast performs at every Linear/conv call, plangen_base.py:360)
or fp32 (check mode).  Norm scales, biases, the text embedding table, gen_embed and the VQ codebook
stay fp32 as they do in the reference."""
from __future__ import annotations

from typing import Dict

import torch

from .config import Dims


VQ_IN_CPAD = 8      # csrc/engine.cu VQ_IN_CPAD


def rope_tables(dims: Dims, tmax: int):
    """cos/sin of LlamaRotaryEmbedding (HF modeling_llama.py:124-136) for positions 0..tmax-1, fp32
    [tmax, head_dim/2] (the two halves of HF's `emb = cat(freqs, freqs)` are identical)."""
    inv_freq = 1.0 / (dims.rope_theta ** (torch.arange(0, dims.head_dim, 2, dtype=torch.int64).to(torch.float) / dims.head_dim))
    pos = torch.arange(tmax, dtype=torch.float)
    freqs = pos[:, None] * inv_freq[None, :]
    return freqs.cos().contiguous(), freqs.sin().contiguous()


def pack_state_dict(sd: Dict[str, torch.Tensor], dims: Dims, mode: str, device, tmax: int,
                    with_vq: bool = True, with_vision: bool = False) -> Dict[str, torch.Tensor]:
    wt = torch.bfloat16 if mode == "bf16" else torch.float32
    out: Dict[str, torch.Tensor] = {}

    def w(t):   # operand-typed weight
        return t.detach().to(device=device, dtype=torch.float32).to(wt).contiguous()

    def f(t):   # stays fp32
        return t.detach().to(device=device, dtype=torch.float32).contiguous()

    lm = "language_model.model."
    out["embed_tokens"] = f(sd[lm + "embed_tokens.weight"])
    for i in range(dims.L):
        p = lm + f"layers.{i}."
        out[f"l{i}.ln1"] = f(sd[p + "input_layernorm.weight"])
        out[f"l{i}.ln2"] = f(sd[p + "post_attention_layernorm.weight"])
        out[f"l{i}.wqkv"] = w(torch.cat([sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.k_proj.weight"],
                                         sd[p + "self_attn.v_proj.weight"]], dim=0))
        out[f"l{i}.wo"] = w(sd[p + "self_attn.o_proj.weight"])
        g_, u_ = sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"]
        if mode == "bf16" and dims.F % 64 == 0:
            # rows interleaved in blocks of 64 (g[0:64], u[0:64], g[64:128], ...): one 128-row MMA tile then holds
            # gate AND up of the same 64 features, which the tcgen05 epilogue combines (fused SwiGLU)
            gu = torch.stack([g_.reshape(dims.F // 64, 64, dims.D), u_.reshape(dims.F // 64, 64, dims.D)], dim=1)
            out[f"l{i}.wgu"] = w(gu.reshape(2 * dims.F, dims.D))
        else:
            out[f"l{i}.wgu"] = w(torch.cat([g_, u_], dim=0))
        out[f"l{i}.wd"] = w(sd[p + "mlp.down_proj.weight"])
    out["norm"] = f(sd[lm + "norm.weight"])
    if "language_model.lm_head.weight" in sd:      # optional: stage-1 layout-text decode (language_model.generate)
        out["lm_head"] = w(sd["language_model.lm_head.weight"])
    out["head.w0"] = w(sd["gen_head.output_mlp_projector.weight"])
    out["head.b0"] = f(sd["gen_head.output_mlp_projector.bias"])
    out["head.w1"] = w(sd["gen_head.vision_head.weight"])
    out["head.b1"] = f(sd["gen_head.vision_head.bias"])
    out["gen_embed"] = f(sd["gen_embed.weight"])
    out["align.w0"] = w(sd["gen_aligner.layers.0.weight"])
    out["align.b0"] = f(sd["gen_aligner.layers.0.bias"])
    out["align.w1"] = w(sd["gen_aligner.layers.2.weight"])
    out["align.b1"] = f(sd["gen_aligner.layers.2.bias"])
    cos, sin = rope_tables(dims, tmax)
    out["rope_cos"] = cos.to(device)
    out["rope_sin"] = sin.to(device)
    if with_vision:
        # understanding side: SigLIP tower (vision_model.vision_tower.*, siglip_vit.py) + `aligner` (projector.py:39-45)
        p = "vision_model.vision_tower."
        out["sig.pos"] = f(sd[p + "pos_embed"]).reshape(dims.sig_patches, dims.sig_width).contiguous()
        out["sig.patch.w"] = w(sd[p + "patch_embed.proj.weight"].reshape(dims.sig_width, -1))      # k = c*p*p + ky*p + kx
        out["sig.patch.b"] = f(sd[p + "patch_embed.proj.bias"])
        for i in range(dims.sig_layers):
            b = p + f"blocks.{i}."
            out[f"sig.{i}.ln1.w"], out[f"sig.{i}.ln1.b"] = f(sd[b + "norm1.weight"]), f(sd[b + "norm1.bias"])
            out[f"sig.{i}.ln2.w"], out[f"sig.{i}.ln2.b"] = f(sd[b + "norm2.weight"]), f(sd[b + "norm2.bias"])
            out[f"sig.{i}.qkv.w"], out[f"sig.{i}.qkv.b"] = w(sd[b + "attn.qkv.weight"]), f(sd[b + "attn.qkv.bias"])
            out[f"sig.{i}.proj.w"], out[f"sig.{i}.proj.b"] = w(sd[b + "attn.proj.weight"]), f(sd[b + "attn.proj.bias"])
            out[f"sig.{i}.fc1.w"], out[f"sig.{i}.fc1.b"] = w(sd[b + "mlp.fc1.weight"]), f(sd[b + "mlp.fc1.bias"])
            out[f"sig.{i}.fc2.w"], out[f"sig.{i}.fc2.b"] = w(sd[b + "mlp.fc2.weight"]), f(sd[b + "mlp.fc2.bias"])
        out["sig.norm.w"], out["sig.norm.b"] = f(sd[p + "norm.weight"]), f(sd[p + "norm.bias"])
        out["ualign.w0"], out["ualign.b0"] = w(sd["aligner.layers.0.weight"]), f(sd["aligner.layers.0.bias"])
        out["ualign.w1"], out["ualign.b1"] = w(sd["aligner.layers.2.weight"]), f(sd["aligner.layers.2.bias"])
    if with_vq:
        g = "gen_vision_model."
        out["vq.codebook"] = f(sd[g + "quantize.embedding.weight"])
        out["vq.pqc.w"] = w(sd[g + "post_quant_conv.weight"].reshape(dims.vq_z, dims.code_dim))
        out["vq.pqc.b"] = f(sd[g + "post_quant_conv.bias"])
        for k, v in sd.items():
            # decoder always; encoder / quant_conv when present (editing path: gen_vision_model.encode)
            if not (k.startswith(g + "decoder.") or k.startswith(g + "encoder.") or k.startswith(g + "quant_conv.")):
                continue
            name = "vq." + k[len(g):]
            if v.dim() == 4:       # conv weight -> [Cout, kh*kw*Cin]
                t = v.permute(0, 2, 3, 1)
                if k == g + "encoder.conv_in.weight":
                    # image channels padded 3 -> 8 with zero taps (pg_vq_encode pads the NHWC image the same way)
                    t = torch.nn.functional.pad(t, (0, VQ_IN_CPAD - t.shape[-1]))
                out[name] = w(t.reshape(v.shape[0], -1))
            else:                  # conv bias / GroupNorm scale+shift
                out[name] = f(v)
    return out
