"""ORACLE (test infrastructure, NOT product code).

CPU/torch restatement of the PlanGen CFG image-token decode path.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` leg may import this module; `plangen_b200/` never does.

What it restates (reference paths relative to /root/reference):

  * `System.t2i`           project/plangen/plangen_base.py:525-565
  * `System.sample_image`  project/plangen/plangen_base.py:567-607
  * `t2i_infer_collate_batch` / `pad_input_ids`   plangen_base.py:636-725
  * `vision_head` (gen_head)   three_party/Janus/janus/models/modeling_vlm.py:36-51
  * `MlpProjector` mlp_gelu    three_party/Janus/janus/models/projector.py:39-45,86
  * `prepare_gen_img_embeds`   modeling_vlm.py:270-271 (gen_embed :214-216)
  * `VQModel.decode_code`      three_party/Janus/janus/models/vq_model.py:505-508
      -> get_codebook_entry :284-299, decode :500-503, Decoder.forward :193-214,
         ResnetBlock :337-352, AttnBlock :366-390, Upsample :417-427,
         Normalize :398-405, nonlinearity :393-395
  * the LM arithmetic, which lives in a THIRD-PARTY dependency that is not under
    /root/reference: HF `transformers` (pinned ==4.48.3 in requirements.txt:296;
    5.5.0 installed here) `LlamaModel.forward` and friends
    (transformers/models/llama/modeling_llama.py: LlamaRMSNorm :53-67,
    LlamaRotaryEmbedding :124-136, rotate_half/apply_rotary_pos_emb :138-168,
    LlamaMLP :182-184, eager_attention_forward :199-221, LlamaAttention :251-290,
    LlamaDecoderLayer :303-331, LlamaModel.forward :375-427;
    cache_utils.DynamicLayer.update :102-121).

Pinning: the reference ships no tests / golden vectors for this path
(SURVEY.md §4), so the oracle is pinned against outputs of the real pieces run
in the authoring container: the installed HF `LlamaModel` and the reference's
`vq_model.py` loaded by file path (`oracle/make_golden.py` -> `tests/golden/`),
and `tests/test_oracle.py` re-checks it against both whenever they are present.

Precision regimes:
  mode="fp32"      no autocast, everything fp32 (BASELINE config 1, check mode)
  mode="autocast"  the reference's regime (plangen_base.py:360): fp32 master
                   weights under torch.autocast(bf16)
"""
from __future__ import annotations

import contextlib
import math
from dataclasses import dataclass, field, asdict
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- dims
@dataclass(frozen=True)
class JanusDims:
    """Architecture numbers (SURVEY.md §8 preamble)."""
    name: str = "janus-1.3b"
    D: int = 2048            # hidden size
    L: int = 24              # decoder layers
    H: int = 16              # attention heads (= kv heads, MHA)
    head_dim: int = 128
    F: int = 5632            # FFN inner size
    vocab: int = 102400      # text vocabulary
    img_vocab: int = 16384   # image_token_size / codebook size
    code_dim: int = 8        # codebook_embed_dim / gen_embed width
    img_embed: int = 2048    # gen_head hidden width (image_token_embed)
    rms_eps: float = 1e-6
    rope_theta: float = 10000.0
    grid: int = 24           # image tokens per side (384/16)
    vq_ch: int = 128
    vq_ch_mult: Tuple[int, ...] = (1, 1, 2, 2, 4)
    vq_z: int = 256
    vq_res_blocks: int = 2
    pad_id: int = 100002     # <｜▁pad▁｜> (processing_vlm.py:91,206-213); synthetic value

    @property
    def n_img_tokens(self) -> int:
        return self.grid * self.grid

    @property
    def img_size(self) -> int:
        return self.grid * 2 ** (len(self.vq_ch_mult) - 1)


JANUS_1P3B = JanusDims()
JANUS_7B = JanusDims(name="janus-pro-7b", D=4096, L=30, H=32, F=11008, img_embed=4096)
# small shapes for tests; same structure (head_dim stays 128, GroupNorm(32) needs C%32==0)
TINY = JanusDims(name="tiny", D=256, L=2, H=2, F=512, vocab=1000, img_vocab=2048,
                 img_embed=256, grid=4, vq_ch=32, vq_ch_mult=(1, 2), vq_z=32, pad_id=999)
SMALL = JanusDims(name="small", D=512, L=4, H=4, F=1408, vocab=4096, img_vocab=16384,
                  img_embed=512, grid=6, vq_ch=32, vq_ch_mult=(1, 1, 2), vq_z=64, pad_id=4095)

PRESETS = {d.name: d for d in (JANUS_1P3B, JANUS_7B, TINY, SMALL)}


# --------------------------------------------------------------------- weight init
def _vq_decoder_shapes(d: JanusDims) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) for every tensor of VQModel's decode side
    (vq_model.py:127-187 Decoder.__init__, :481-492 VQModel.__init__)."""
    out: List[Tuple[str, Tuple[int, ...], str]] = []
    p = "gen_vision_model."
    out.append((p + "quantize.embedding.weight", (d.img_vocab, d.code_dim), "codebook"))
    out.append((p + "post_quant_conv.weight", (d.vq_z, d.code_dim, 1, 1), "conv"))
    out.append((p + "post_quant_conv.bias", (d.vq_z,), "bias:%d" % d.code_dim))
    nres = len(d.vq_ch_mult)
    block_in = d.vq_ch * d.vq_ch_mult[nres - 1]

    def conv(name, cin, cout, k):
        out.append((name + ".weight", (cout, cin, k, k), "conv"))
        out.append((name + ".bias", (cout,), "bias:%d" % (cin * k * k)))

    def norm(name, c):
        out.append((name + ".weight", (c,), "norm_w"))
        out.append((name + ".bias", (c,), "norm_b"))

    def res(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cin, cout, 3)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".nin_shortcut", cin, cout, 1)

    def attn(name, c):
        norm(name + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(name + "." + n, c, c, 1)

    dp = p + "decoder."
    conv(dp + "conv_in", d.vq_z, block_in, 3)
    res(dp + "mid.0", block_in, block_in)
    attn(dp + "mid.1", block_in)
    res(dp + "mid.2", block_in, block_in)
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
    return out


def _vq_encoder_shapes(d: JanusDims) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) for VQModel's encode side (vq_model.py:46-106 Encoder.__init__, :493 quant_conv)."""
    out: List[Tuple[str, Tuple[int, ...], str]] = []
    p = "gen_vision_model."

    def conv(name, cin, cout, k):
        out.append((name + ".weight", (cout, cin, k, k), "conv"))
        out.append((name + ".bias", (cout,), "bias:%d" % (cin * k * k)))

    def norm(name, c):
        out.append((name + ".weight", (c,), "norm_w"))
        out.append((name + ".bias", (c,), "norm_b"))

    def res(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cin, cout, 3)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".nin_shortcut", cin, cout, 1)

    def attn(name, c):
        norm(name + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(name + "." + n, c, c, 1)

    ep = p + "encoder."
    nres = len(d.vq_ch_mult)
    conv(ep + "conv_in", 3, d.vq_ch, 3)
    in_mult = (1,) + tuple(d.vq_ch_mult)
    block_in = d.vq_ch
    for i_level in range(nres):
        block_in = d.vq_ch * in_mult[i_level]
        block_out = d.vq_ch * d.vq_ch_mult[i_level]
        for j in range(d.vq_res_blocks):
            res(ep + f"conv_blocks.{i_level}.res.{j}", block_in, block_out)
            block_in = block_out
            if i_level == nres - 1:
                attn(ep + f"conv_blocks.{i_level}.attn.{j}", block_in)
        if i_level != nres - 1:
            conv(ep + f"conv_blocks.{i_level}.downsample.conv", block_in, block_in, 3)
    res(ep + "mid.0", block_in, block_in)
    attn(ep + "mid.1", block_in)
    res(ep + "mid.2", block_in, block_in)
    norm(ep + "norm_out", block_in)
    conv(ep + "conv_out", block_in, d.vq_z, 3)
    conv(p + "quant_conv", d.vq_z, d.code_dim, 1)
    return out


def tensor_specs(d: JanusDims, with_vq: bool = True, with_lm_head: bool = False,
                 with_vq_encoder: bool = False) -> List[Tuple[str, Tuple[int, ...], str]]:
    """All state-dict tensors the decode path touches, reference naming.  `with_lm_head` adds the
    untied text head used by the stage-1 layout-text decode (x2t, §8f rank 1)."""
    s: List[Tuple[str, Tuple[int, ...], str]] = []
    lm = "language_model.model."
    s.append((lm + "embed_tokens.weight", (d.vocab, d.D), "lm"))
    if with_lm_head:
        s.append(("language_model.lm_head.weight", (d.vocab, d.D), "lm"))
    HD = d.H * d.head_dim
    for i in range(d.L):
        l = lm + f"layers.{i}."
        s.append((l + "input_layernorm.weight", (d.D,), "norm_w"))
        s.append((l + "self_attn.q_proj.weight", (HD, d.D), "lm"))
        s.append((l + "self_attn.k_proj.weight", (HD, d.D), "lm"))
        s.append((l + "self_attn.v_proj.weight", (HD, d.D), "lm"))
        s.append((l + "self_attn.o_proj.weight", (d.D, HD), "lm"))
        s.append((l + "post_attention_layernorm.weight", (d.D,), "norm_w"))
        s.append((l + "mlp.gate_proj.weight", (d.F, d.D), "lm"))
        s.append((l + "mlp.up_proj.weight", (d.F, d.D), "lm"))
        s.append((l + "mlp.down_proj.weight", (d.D, d.F), "lm"))
    s.append((lm + "norm.weight", (d.D,), "norm_w"))
    s.append(("gen_head.output_mlp_projector.weight", (d.img_embed, d.D), "linear"))
    s.append(("gen_head.output_mlp_projector.bias", (d.img_embed,), "bias:%d" % d.D))
    s.append(("gen_head.vision_head.weight", (d.img_vocab, d.img_embed), "linear"))
    s.append(("gen_head.vision_head.bias", (d.img_vocab,), "bias:%d" % d.img_embed))
    s.append(("gen_embed.weight", (d.img_vocab, d.code_dim), "normal1"))
    s.append(("gen_aligner.layers.0.weight", (d.D, d.code_dim), "linear"))
    s.append(("gen_aligner.layers.0.bias", (d.D,), "bias:%d" % d.code_dim))
    s.append(("gen_aligner.layers.2.weight", (d.D, d.D), "linear"))
    s.append(("gen_aligner.layers.2.bias", (d.D,), "bias:%d" % d.D))
    if with_vq:
        s.extend(_vq_decoder_shapes(d))
    if with_vq_encoder:
        s.extend(_vq_encoder_shapes(d))
    return s


def init_state_dict(d: JanusDims, seed: int = 0, with_vq: bool = True,
                    lm_std: float = 0.02, only: Optional[str] = None, with_lm_head: bool = False,
                    with_vq_encoder: bool = False) -> Dict[str, torch.Tensor]:
    """Deterministic random-init fp32 weights (CPU generator; identical on every
    box with the same torch build).  Every tensor has its own generator seeded
    from (seed, crc32(name)), so any subset (`only` = name prefix) reproduces the
    same values.  LM: N(0, lm_std) (HF init); gen_head / gen_aligner / conv: torch
    default U(+-1/sqrt(fan_in)); gen_embed N(0,1); codebook U(+-1/n) then
    L2-normalised (vq_model.py:228-232).  Norm scales are 1 + 0.1 N(0,1) and
    GroupNorm biases 0.1 N(0,1) rather than exactly 1 / 0 so a kernel that drops
    them is caught."""
    import zlib
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in tensor_specs(d, with_vq, with_lm_head, with_vq_encoder):
        if only is not None and not name.startswith(only):
            continue
        g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
        if kind == "lm":
            t = torch.empty(shape).normal_(0.0, lm_std, generator=g)
        elif kind == "norm_w":
            t = 1.0 + 0.1 * torch.empty(shape).normal_(0.0, 1.0, generator=g)
        elif kind == "norm_b":
            t = 0.1 * torch.empty(shape).normal_(0.0, 1.0, generator=g)
        elif kind == "normal1":
            t = torch.empty(shape).normal_(0.0, 1.0, generator=g)
        elif kind == "codebook":
            t = torch.empty(shape).uniform_(-1.0 / shape[0], 1.0 / shape[0], generator=g)
            t = F.normalize(t, p=2, dim=-1)
        elif kind in ("linear", "conv"):
            fan_in = int(math.prod(shape[1:]))
            b = 1.0 / math.sqrt(fan_in)
            t = torch.empty(shape).uniform_(-b, b, generator=g)
        elif kind.startswith("bias:"):
            b = 1.0 / math.sqrt(int(kind.split(":")[1]))
            t = torch.empty(shape).uniform_(-b, b, generator=g)
        else:  # pragma: no cover
            raise ValueError(kind)
        sd[name] = t
    return sd


# ------------------------------------------------------------------ LM restatement
def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """LlamaRMSNorm.forward (HF modeling_llama.py:60-65): fp32 statistics, cast
    back to the input dtype BEFORE the multiply by `weight`."""
    input_dtype = x.dtype
    h = x.to(torch.float32)
    variance = h.pow(2).mean(-1, keepdim=True)
    h = h * torch.rsqrt(variance + eps)
    return w * h.to(input_dtype)


def rope_cos_sin(position_ids: torch.Tensor, d: JanusDims, dtype: torch.dtype):
    """LlamaRotaryEmbedding.forward (HF :124-136): fp32 outer product, cat(freqs,
    freqs), cos/sin, cast to the activation dtype.  position_ids: (1, q)."""
    dev = position_ids.device
    inv_freq = 1.0 / (d.rope_theta ** (torch.arange(0, d.head_dim, 2, dtype=torch.int64)
                                       .to(device=dev, dtype=torch.float) / d.head_dim))
    inv_freq_expanded = inv_freq[None, :, None].float().expand(position_ids.shape[0], -1, 1)
    position_ids_expanded = position_ids[:, None, :].float()
    with torch.autocast(device_type=dev.type, enabled=False):
        freqs = (inv_freq_expanded.float() @ position_ids_expanded.float()).transpose(1, 2)
        emb = torch.cat((freqs, freqs), dim=-1)
        cos, sin = emb.cos(), emb.sin()
    return cos.to(dtype), sin.to(dtype)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(q, k, cos, sin):
    cos = cos.unsqueeze(1)
    sin = sin.unsqueeze(1)
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


def build_additive_mask(attention_mask: Optional[torch.Tensor], q_pos: torch.Tensor,
                        kv_len: int, dtype: torch.dtype, rows: int) -> torch.Tensor:
    """What HF `create_causal_mask` hands eager attention: key j is visible to the
    query at absolute position p iff j <= p and attention_mask[r, j] == 1.  The
    reference passes the FULL (R, P+576) mask on every call (plangen_base.py:573);
    only its first kv_len columns matter."""
    dev = q_pos.device
    j = torch.arange(kv_len, device=dev)
    allowed = (j[None, :] <= q_pos[:, None])[None].expand(rows, -1, -1)  # (R, q, kv)
    if attention_mask is not None:
        allowed = allowed & (attention_mask[:, None, :kv_len] != 0)
    m = torch.zeros(allowed.shape, dtype=dtype, device=dev)
    m.masked_fill_(~allowed, torch.finfo(dtype).min)
    return m[:, None]  # (R, 1, q, kv)


@dataclass
class LMOutput:
    """BaseModelOutputWithPast stand-in."""
    last_hidden_state: torch.Tensor
    past_key_values: List[Tuple[torch.Tensor, torch.Tensor]]


def llama_model_forward(sd: Dict[str, torch.Tensor], d: JanusDims, inputs_embeds: torch.Tensor,
                        attention_mask: Optional[torch.Tensor] = None,
                        past_key_values: Optional[List[Tuple[torch.Tensor, torch.Tensor]]] = None,
                        use_cache: bool = True, position_ids: Optional[torch.Tensor] = None) -> LMOutput:
    """LlamaModel.forward (HF :375-427) as the reference calls it
    (plangen_base.py:571-576): inputs_embeds + full-length 0/1 mask, no
    position_ids => positions = past_len + arange(q).  `position_ids` (R, q) is what
    generate() passes (x2t): they only feed the rotary embedding; the causal mask is
    built from the cache positions either way (HF :394-400)."""
    lm = "language_model.model."
    R, q_len, _ = inputs_embeds.shape
    past_len = 0 if not past_key_values else past_key_values[0][0].shape[2]
    cache_position = (torch.arange(q_len, device=inputs_embeds.device) + past_len).unsqueeze(0)
    rope_position = cache_position if position_ids is None else position_ids
    position_ids = cache_position
    kv_len = past_len + q_len
    h = inputs_embeds
    cos, sin = rope_cos_sin(rope_position, d, h.dtype)
    new_cache: List[Tuple[torch.Tensor, torch.Tensor]] = []
    scaling = d.head_dim ** -0.5
    mask4 = None
    for i in range(d.L):
        l = lm + f"layers.{i}."
        residual = h
        x = rms_norm(h, sd[l + "input_layernorm.weight"], d.rms_eps)
        shape = (R, q_len, -1, d.head_dim)
        qs = F.linear(x, sd[l + "self_attn.q_proj.weight"]).view(shape).transpose(1, 2)
        ks = F.linear(x, sd[l + "self_attn.k_proj.weight"]).view(shape).transpose(1, 2)
        vs = F.linear(x, sd[l + "self_attn.v_proj.weight"]).view(shape).transpose(1, 2)
        qs, ks = apply_rope(qs, ks, cos, sin)
        if past_key_values:                       # DynamicLayer.update: cat on dim -2
            ks = torch.cat([past_key_values[i][0], ks], dim=-2)
            vs = torch.cat([past_key_values[i][1], vs], dim=-2)
        new_cache.append((ks, vs))
        # eager_attention_forward (HF :199-221)
        aw = torch.matmul(qs, ks.transpose(2, 3)) * scaling
        if mask4 is None or mask4.dtype != aw.dtype:
            mask4 = build_additive_mask(attention_mask, position_ids[0], kv_len, aw.dtype, R)
        aw = aw + mask4
        aw = F.softmax(aw, dim=-1, dtype=torch.float32).to(qs.dtype)
        ao = torch.matmul(aw, vs).transpose(1, 2).contiguous().reshape(R, q_len, -1)
        ao = F.linear(ao, sd[l + "self_attn.o_proj.weight"])
        h = residual + ao
        residual = h
        x = rms_norm(h, sd[l + "post_attention_layernorm.weight"], d.rms_eps)
        x = F.linear(F.silu(F.linear(x, sd[l + "mlp.gate_proj.weight"]))
                     * F.linear(x, sd[l + "mlp.up_proj.weight"]), sd[l + "mlp.down_proj.weight"])
        h = residual + x
    h = rms_norm(h, sd[lm + "norm.weight"], d.rms_eps)
    return LMOutput(h, new_cache if use_cache else [])


# ------------------------------------------------------------ heads and projectors
def gen_head(sd, x: torch.Tensor) -> torch.Tensor:
    """vision_head.forward (modeling_vlm.py:47-51): Linear -> GELU(erf) -> Linear."""
    x = F.linear(x, sd["gen_head.output_mlp_projector.weight"], sd["gen_head.output_mlp_projector.bias"])
    x = F.gelu(x)
    return F.linear(x, sd["gen_head.vision_head.weight"], sd["gen_head.vision_head.bias"])


def prepare_gen_img_embeds(sd, image_ids: torch.Tensor) -> torch.Tensor:
    """modeling_vlm.py:270-271: gen_aligner(gen_embed(ids)); MlpProjector mlp_gelu
    depth 2 (projector.py:39-45): Linear(8,D) -> GELU -> Linear(D,D)."""
    x = F.embedding(image_ids.long(), sd["gen_embed.weight"])
    x = F.linear(x, sd["gen_aligner.layers.0.weight"], sd["gen_aligner.layers.0.bias"])
    x = F.gelu(x)
    return F.linear(x, sd["gen_aligner.layers.2.weight"], sd["gen_aligner.layers.2.bias"])


def embed_tokens(sd, ids: torch.Tensor) -> torch.Tensor:
    return F.embedding(ids.long(), sd["language_model.model.embed_tokens.weight"])


# ------------------------------------------------------- mmu front-end - ORACLE ONLY
# The CUDA path for this row is NOT built yet (plangen_b200.engine.FastJanus.prepare_inputs_embeds raises); the
# restatement and its pin are here so the row can start from a checked oracle.
@dataclass(frozen=True)
class SigLIPDims:
    """SigLIP_MODEL_CONFIG["siglip_large_patch16_384"] (siglip_vit.py:628-637): the vision tower of Janus-1.3B."""
    name: str = "siglip_large_patch16_384"
    width: int = 1024
    layers: int = 24
    heads: int = 16
    patch: int = 16
    image: int = 384
    mlp_ratio: float = 4.0

    @property
    def n_patches(self) -> int:
        return (self.image // self.patch) ** 2


SIGLIP_L16_384 = SigLIPDims()
SIGLIP_TINY = SigLIPDims(name="siglip-tiny", width=64, layers=2, heads=2, patch=16, image=32)


def siglip_tensor_specs(v: SigLIPDims, d: JanusDims) -> List[Tuple[str, Tuple[int, ...], str]]:
    """State-dict names of `vision_model.vision_tower` (VisionTransformer, siglip_vit.py:262-440, class_token=False,
    qkv_bias=True, no pre-norm, LayerNorm eps 1e-6) and of the understanding `aligner` (MlpProjector mlp_gelu depth 2,
    projector.py:39-45)."""
    p = "vision_model.vision_tower."
    hid = int(v.width * v.mlp_ratio)
    s: List[Tuple[str, Tuple[int, ...], str]] = [
        (p + "pos_embed", (1, v.n_patches, v.width), "lm"),
        (p + "patch_embed.proj.weight", (v.width, 3, v.patch, v.patch), "conv"),
        (p + "patch_embed.proj.bias", (v.width,), "bias:%d" % (3 * v.patch * v.patch)),
    ]
    for i in range(v.layers):
        b = p + f"blocks.{i}."
        s += [(b + "norm1.weight", (v.width,), "norm_w"), (b + "norm1.bias", (v.width,), "norm_b"),
              (b + "attn.qkv.weight", (3 * v.width, v.width), "lm"), (b + "attn.qkv.bias", (3 * v.width,), "norm_b"),
              (b + "attn.proj.weight", (v.width, v.width), "lm"), (b + "attn.proj.bias", (v.width,), "norm_b"),
              (b + "norm2.weight", (v.width,), "norm_w"), (b + "norm2.bias", (v.width,), "norm_b"),
              (b + "mlp.fc1.weight", (hid, v.width), "lm"), (b + "mlp.fc1.bias", (hid,), "norm_b"),
              (b + "mlp.fc2.weight", (v.width, hid), "lm"), (b + "mlp.fc2.bias", (v.width,), "norm_b")]
    s += [(p + "norm.weight", (v.width,), "norm_w"), (p + "norm.bias", (v.width,), "norm_b")]
    s += [("aligner.layers.0.weight", (d.D, v.width), "linear"), ("aligner.layers.0.bias", (d.D,), "bias:%d" % v.width),
          ("aligner.layers.2.weight", (d.D, d.D), "linear"), ("aligner.layers.2.bias", (d.D,), "bias:%d" % d.D)]
    return s


def init_siglip_state_dict(v: SigLIPDims, d: JanusDims, seed: int = 0) -> Dict[str, torch.Tensor]:
    import zlib
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in siglip_tensor_specs(v, d):
        g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
        if kind == "lm":
            t = torch.empty(shape).normal_(0.0, 0.02, generator=g)
        elif kind == "norm_w":
            t = 1.0 + 0.1 * torch.empty(shape).normal_(0.0, 1.0, generator=g)
        elif kind == "norm_b":
            t = 0.1 * torch.empty(shape).normal_(0.0, 1.0, generator=g)
        else:
            fan = int(math.prod(shape[1:])) if not kind.startswith("bias:") else int(kind.split(":")[1])
            b = 1.0 / math.sqrt(fan)
            t = torch.empty(shape).uniform_(-b, b, generator=g)
        sd[name] = t
    return sd


def siglip_forward(sd, v: SigLIPDims, images: torch.Tensor) -> torch.Tensor:
    """CLIPVisionTower.forward (clip_encoder.py:107-122; no image_norm in the Janus config, select_feature "same") ->
    VisionTransformer.forward with ignore_head (siglip_vit.py:584-606): patch embedding (conv16/16), + learned position
    embedding, `layers` x [x += proj(SDPA(qkv(LN1 x))); x += fc2(GELU(fc1(LN2 x)))], final LayerNorm.
    (B,3,H,W) -> (B, n_patches, width)."""
    p = "vision_model.vision_tower."
    x = F.conv2d(images, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=v.patch)
    x = x.flatten(2).transpose(1, 2)
    x = x + sd[p + "pos_embed"]
    B, N, C = x.shape
    hd = C // v.heads
    for i in range(v.layers):
        b = p + f"blocks.{i}."
        h = F.layer_norm(x, (C,), sd[b + "norm1.weight"], sd[b + "norm1.bias"], eps=1e-6)
        qkv = F.linear(h, sd[b + "attn.qkv.weight"], sd[b + "attn.qkv.bias"]).reshape(B, N, 3, v.heads, hd).permute(2, 0, 3, 1, 4)
        q, k, vv = qkv.unbind(0)
        a = F.scaled_dot_product_attention(q, k, vv).transpose(1, 2).reshape(B, N, C)
        x = x + F.linear(a, sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"])
        h = F.layer_norm(x, (C,), sd[b + "norm2.weight"], sd[b + "norm2.bias"], eps=1e-6)
        h = F.linear(F.gelu(F.linear(h, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"])), sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"])
        x = x + h
    return F.layer_norm(x, (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], eps=1e-6)


def understanding_aligner(sd, x: torch.Tensor) -> torch.Tensor:
    """`aligner` = MlpProjector mlp_gelu depth 2 (projector.py:39-45): Linear(width, D) -> GELU -> Linear(D, D)."""
    x = F.linear(x, sd["aligner.layers.0.weight"], sd["aligner.layers.0.bias"])
    return F.linear(F.gelu(x), sd["aligner.layers.2.weight"], sd["aligner.layers.2.bias"])


def prepare_inputs_embeds(sd, v: SigLIPDims, input_ids: torch.Tensor, pixel_values: torch.Tensor,
                          images_seq_mask: torch.Tensor, images_emb_mask: torch.Tensor, mode: str = "fp32") -> torch.Tensor:
    """MultiModalityCausalLM.prepare_inputs_embeds (modeling_vlm.py:221-268): image features of every image
    (b, n, 3, h, w) through the vision tower and the aligner, scattered into the text embeddings at the
    `<image_placeholder>` positions; negative ids (image slots) are embedded as id 0 first."""
    bs, n = pixel_values.shape[:2]
    images = pixel_values.reshape(bs * n, *pixel_values.shape[2:])
    if mode == "autocast":
        images = images.bfloat16()                                     # `images.bfloat16()` (:247)
    with _autocast_ctx(mode, images.device):
        emb = understanding_aligner(sd, siglip_forward(sd, v, images.float() if mode == "fp32" else images))
    emb = emb.reshape(bs, n * emb.shape[1], emb.shape[2])
    emb_mask = images_emb_mask.reshape(bs, -1).bool()
    ids = input_ids.clone()
    ids[ids < 0] = 0
    out = embed_tokens(sd, ids).clone()
    out[images_seq_mask.bool()] = emb[emb_mask].to(out.dtype)
    return out


# --------------------------------------------------- stage-1 layout-text decode (x2t)
def generate_greedy(sd, d: JanusDims, inputs_embeds: torch.Tensor, attention_mask: torch.Tensor,
                    max_new_tokens: int, eos_token_id: int, pad_token_id: int, mode: str = "fp32",
                    return_logits: bool = False):
    """System.x2t (plangen_base.py:513-523): `language_model.generate(inputs_embeds=, attention_mask=,
    pad_token_id=eos, bos_token_id=, eos_token_id=eos, max_new_tokens=512, do_sample=False, use_cache=True)`
    = HF GenerationMixin greedy search (third-party `transformers`, pinned ==4.48.3, generation/utils.py
    `_sample` with do_sample=False; `prepare_inputs_for_generation`):
      * position_ids = attention_mask.cumsum(-1) - 1, pads filled with 1; one more per generated token;
      * next_token_logits = logits[:, -1, :].float(); next_tokens = argmax;
      * finished rows keep emitting pad_token_id; a row finishes when it emits eos_token_id;
      * the loop stops when every row has finished or after max_new_tokens;
      * with inputs_embeds and no input_ids the returned sequences hold the NEW tokens only.
    Returns int64 (R, n_generated) [, list of per-step fp32 logits]."""
    dev = inputs_embeds.device
    R = inputs_embeds.shape[0]
    mask = attention_mask.to(dev).long()
    head = sd["language_model.lm_head.weight"]
    logits_all = []
    with torch.inference_mode(), _autocast_ctx(mode, dev):
        pos = (mask.cumsum(-1) - 1).masked_fill(mask == 0, 1)
        out = llama_model_forward(sd, d, inputs_embeds, attention_mask=mask, position_ids=pos)
        pkv = out.past_key_values
        hidden = out.last_hidden_state[:, -1, :]
        unfinished = torch.ones(R, dtype=torch.long, device=dev)
        seq = torch.zeros(R, 0, dtype=torch.long, device=dev)
        for i in range(max_new_tokens):
            logits = F.linear(hidden, head).float()
            if return_logits:
                logits_all.append(logits.clone())
            nxt = torch.argmax(logits, dim=-1)
            nxt = nxt * unfinished + pad_token_id * (1 - unfinished)
            seq = torch.cat([seq, nxt[:, None]], dim=-1)
            unfinished = unfinished & (nxt != eos_token_id).long()
            if int(unfinished.max()) == 0 or i == max_new_tokens - 1:
                break
            mask = torch.cat([mask, torch.ones(R, 1, dtype=torch.long, device=dev)], dim=-1)
            pos = mask.sum(-1, keepdim=True) - 1
            out = llama_model_forward(sd, d, embed_tokens(sd, nxt)[:, None, :], attention_mask=mask,
                                      past_key_values=pkv, position_ids=pos)
            pkv = out.past_key_values
            hidden = out.last_hidden_state[:, -1, :]
    return (seq, logits_all) if return_logits else seq


# ------------------------------------------------------------------- VQ decode side
def _gn(x, sd, name):
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps=1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def _conv(x, sd, name, pad):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=1, padding=pad)


def _resblock(x, sd, name):
    h = _conv(_swish(_gn(x, sd, name + ".norm1")), sd, name + ".conv1", 1)
    h = _conv(_swish(_gn(h, sd, name + ".norm2")), sd, name + ".conv2", 1)   # dropout p=0
    if (name + ".nin_shortcut.weight") in sd:
        x = _conv(x, sd, name + ".nin_shortcut", 0)
    return x + h


def _attnblock(x, sd, name):
    h_ = _gn(x, sd, name + ".norm")
    q = _conv(h_, sd, name + ".q", 0)
    k = _conv(h_, sd, name + ".k", 0)
    v = _conv(h_, sd, name + ".v", 0)
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(h_, sd, name + ".proj_out", 0)


def _upsample(x, sd, name):
    if x.dtype != torch.float32:                 # vq_model.py:418-421
        x = F.interpolate(x.to(torch.float), scale_factor=2.0, mode="nearest").to(torch.bfloat16)
    else:
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
    return _conv(x, sd, name + ".conv", 1)


def get_codebook_entry(sd, indices: torch.Tensor, shape: Sequence[int]) -> torch.Tensor:
    """VectorQuantizer.get_codebook_entry (vq_model.py:284-299), channel_first."""
    emb = F.normalize(sd["gen_vision_model.quantize.embedding.weight"], p=2, dim=-1)
    z_q = emb[indices.long().reshape(-1)]
    z_q = z_q.reshape(shape[0], shape[2], shape[3], shape[1])
    return z_q.permute(0, 3, 1, 2).contiguous()


def vq_decode(sd, d: JanusDims, quant: torch.Tensor) -> torch.Tensor:
    """VQModel.decode (vq_model.py:500-503) + Decoder.forward (:193-214)."""
    p = "gen_vision_model."
    h = _conv(quant, sd, p + "post_quant_conv", 0)
    dp = p + "decoder."
    h = _conv(h, sd, dp + "conv_in", 1)
    h = _resblock(h, sd, dp + "mid.0")
    h = _attnblock(h, sd, dp + "mid.1")
    h = _resblock(h, sd, dp + "mid.2")
    nres = len(d.vq_ch_mult)
    for idx in range(nres):
        for j in range(d.vq_res_blocks + 1):
            h = _resblock(h, sd, dp + f"conv_blocks.{idx}.res.{j}")
            if idx == 0:
                h = _attnblock(h, sd, dp + f"conv_blocks.{idx}.attn.{j}")
        if idx != nres - 1:
            h = _upsample(h, sd, dp + f"conv_blocks.{idx}.upsample")
    h = _swish(_gn(h, sd, dp + "norm_out"))
    return _conv(h, sd, dp + "conv_out", 1)


def _downsample(x, sd, name):
    """Downsample.forward (vq_model.py:440-445): asymmetric zero pad (right, bottom), 3x3 stride-2 conv."""
    x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
    return F.conv2d(x, sd[name + ".conv.weight"], sd[name + ".conv.bias"], stride=2, padding=0)


def vq_encoder_forward(sd, d: JanusDims, x: torch.Tensor) -> torch.Tensor:
    """Encoder.forward (vq_model.py:108-124) + quant_conv (:495-496): (B,3,H,W) -> (B,code_dim,H/16,W/16)."""
    p = "gen_vision_model."
    ep = p + "encoder."
    nres = len(d.vq_ch_mult)
    h = _conv(x, sd, ep + "conv_in", 1)
    for i_level in range(nres):
        for j in range(d.vq_res_blocks):
            h = _resblock(h, sd, ep + f"conv_blocks.{i_level}.res.{j}")
            if i_level == nres - 1:
                h = _attnblock(h, sd, ep + f"conv_blocks.{i_level}.attn.{j}")
        if i_level != nres - 1:
            h = _downsample(h, sd, ep + f"conv_blocks.{i_level}.downsample")
    h = _resblock(h, sd, ep + "mid.0")
    h = _attnblock(h, sd, ep + "mid.1")
    h = _resblock(h, sd, ep + "mid.2")
    h = _swish(_gn(h, sd, ep + "norm_out"))
    h = _conv(h, sd, ep + "conv_out", 1)
    return _conv(h, sd, p + "quant_conv", 0)


def vq_quantize_indices(sd, z: torch.Tensor, return_distances: bool = False):
    """VectorQuantizer.forward (vq_model.py:236-262), inference branch, l2_norm=True: nearest code of every
    position.  z (B,C,H,W) -> flat int64 indices (B*H*W,) in (b, h, w) order."""
    z = torch.einsum("b c h w -> b h w c", z).contiguous()
    z_flattened = z.view(-1, z.shape[-1])
    z_flattened = F.normalize(z_flattened, p=2, dim=-1)
    embedding = F.normalize(sd["gen_vision_model.quantize.embedding.weight"], p=2, dim=-1)
    dist = (torch.sum(z_flattened ** 2, dim=1, keepdim=True) + torch.sum(embedding ** 2, dim=1)
            - 2 * torch.einsum("bd,dn->bn", z_flattened, torch.einsum("n d -> d n", embedding)))
    idx = torch.argmin(dist, dim=1)
    return (idx, dist) if return_distances else idx


def vq_encode(sd, d: JanusDims, img: torch.Tensor, mode: str = "fp32") -> torch.Tensor:
    """`vl_gpt.gen_vision_model.encode(img)[-1][-1]` (plangen_base.py:532; VQModel.encode vq_model.py:494-498):
    (B,3,H,W) image in [-1,1] -> flat int64 code indices (B * H/16 * W/16,)."""
    with torch.inference_mode(), _autocast_ctx(mode, img.device):
        return vq_quantize_indices(sd, vq_encoder_forward(sd, d, img))


def decode_code(sd, d: JanusDims, code_b: torch.Tensor, shape: Sequence[int]) -> torch.Tensor:
    """VQModel.decode_code (vq_model.py:505-508)."""
    return vq_decode(sd, d, get_codebook_entry(sd, code_b, shape))


# ------------------------------------------------------------ host prompt layout
def pad_input_ids(all_inputs_ids: Sequence[Sequence[int]], pad_id: int,
                  max_length: Optional[int] = None):
    """plangen_base.py:699-725 (test branch: no truncation): LEFT padding."""
    bs = len(all_inputs_ids)
    if max_length is None:
        max_length = max(map(len, all_inputs_ids))
    ids = torch.ones((bs, max_length)) * pad_id
    mask = torch.zeros((bs, max_length))
    for i, inputs_ids in enumerate(all_inputs_ids):
        n = len(inputs_ids)
        ids[i, max_length - n:] = torch.as_tensor(list(inputs_ids), dtype=ids.dtype)
        mask[i, max_length - n:] = 1
    return ids.int(), mask.int()


def t2i_infer_collate_batch(cond_ids: Sequence[Sequence[int]], neg_ids: Sequence[Sequence[int]],
                            pad_id: int, n_img_tokens: int):
    """plangen_base.py:636-697 on already-tokenised prompts: cond and negative rows
    LEFT-padded to a common P (:661-667,678-684 + pad_input_ids), 576 ones appended
    to the mask (:668,686; the cond mask already carries them from mmu_collate),
    rows interleaved [cond0, neg0, cond1, neg1, ...] (:690-691).
    Returns ids (2B, P) int32 and mask (2B, P+n_img_tokens) int32."""
    bs = len(cond_ids)
    assert len(neg_ids) == bs
    P = max(max(map(len, cond_ids)), max(map(len, neg_ids)))
    c_ids, c_mask = pad_input_ids(cond_ids, pad_id, P)
    n_ids, n_mask = pad_input_ids(neg_ids, pad_id, P)
    ones = torch.ones((bs, n_img_tokens), dtype=c_mask.dtype)
    c_mask = torch.cat([c_mask, ones], dim=-1)
    n_mask = torch.cat([n_mask, ones], dim=-1)
    ids = torch.stack([c_ids, n_ids], dim=1).view(bs * 2, -1)
    mask = torch.stack([c_mask, n_mask], dim=1).view(bs * 2, -1)
    return ids.int(), mask.int()


# ---------------------------------------------------------------- the decode loop
Sampler = Callable[[torch.Tensor, int], torch.Tensor]


def make_torch_sampler(generator: torch.Generator) -> Sampler:
    """plangen_base.py:591 verbatim: device-native torch.multinomial."""
    def _s(probs: torch.Tensor, step: int) -> torch.Tensor:
        return torch.multinomial(probs, num_samples=1, generator=generator)
    return _s


def greedy_sampler(probs: torch.Tensor, step: int) -> torch.Tensor:
    """north_star's fp32 'greedy' check mode: argmax instead of multinomial."""
    return torch.argmax(probs, dim=-1, keepdim=True)


def _autocast_ctx(mode: str, device: torch.device):
    if mode == "fp32":
        return contextlib.nullcontext()
    if mode == "autocast":
        return torch.autocast(device_type=device.type, dtype=torch.bfloat16)
    raise ValueError(mode)


@torch.inference_mode()
def sample_image(sd, d: JanusDims, inputs_embeds: torch.Tensor, num_gen: int,
                 image_token_num_per_image: int, mask: Optional[torch.Tensor],
                 cfg_weight: float, temperature: float, sampler: Sampler,
                 edit_region: Optional[torch.Tensor] = None,
                 gt_labels: Optional[torch.Tensor] = None,
                 mode: str = "fp32", trace: Optional[dict] = None,
                 lm_forward: Optional[Callable] = None) -> torch.Tensor:
    """System.sample_image (plangen_base.py:567-607).  `trace`, if given, collects
    per-step CFG logits / probs / tokens for golden vectors.  `lm_forward` lets the
    golden generator drive the real HF LlamaModel through the very same loop."""
    dev = inputs_embeds.device
    generated_tokens = torch.zeros((num_gen, image_token_num_per_image), dtype=torch.int, device=dev)
    fwd = lm_forward or (lambda **kw: llama_model_forward(sd, d, **kw))
    outputs = None
    with _autocast_ctx(mode, dev):
        for i in range(image_token_num_per_image):
            outputs = fwd(inputs_embeds=inputs_embeds, attention_mask=mask, use_cache=True,
                          past_key_values=outputs.past_key_values if i != 0 else None)
            hidden_states = outputs.last_hidden_state
            logits = gen_head(sd, hidden_states[:, -1, :])
            if trace is not None:
                trace.setdefault("raw_logits", []).append(logits.float().cpu())
            logit_cond = logits[0::2, :]
            logit_uncond = logits[1::2, :]
            logits = logit_uncond + cfg_weight * (logit_cond - logit_uncond)
            probs = torch.softmax(logits / temperature, dim=-1)
            next_token = sampler(probs, i)
            if edit_region is not None:                       # teacher forcing :593-598
                for bid in range(len(edit_region)):
                    if edit_region[bid, i].item() == 0:
                        next_token[bid, 0] = gt_labels[bid, i]
            generated_tokens[:, i] = next_token.squeeze(dim=-1)
            if trace is not None:
                trace.setdefault("hidden", []).append(hidden_states[:, -1, :].float().cpu())
                trace.setdefault("logits", []).append(logits.float().cpu())
                trace.setdefault("probs", []).append(probs.float().cpu())
                trace.setdefault("tokens", []).append(next_token.squeeze(-1).cpu())
            next_token = torch.cat([next_token.unsqueeze(dim=1), next_token.unsqueeze(dim=1)], dim=1).view(-1)
            img_embeds = prepare_gen_img_embeds(sd, next_token)
            inputs_embeds = img_embeds.unsqueeze(dim=1)
    return generated_tokens


@torch.inference_mode()
def t2i(sd, d: JanusDims, tokens: torch.Tensor, mask: torch.Tensor, parallel_size: int = 1,
        cfg_weight: float = 5.0, temperature: float = 1.0, seed: int = 0,
        sampler: Optional[Sampler] = None, mode: str = "fp32",
        edit_region: Optional[torch.Tensor] = None, gt_labels: Optional[torch.Tensor] = None,
        image_token_num_per_image: Optional[int] = None, decode: bool = True,
        trace: Optional[dict] = None):
    """System.t2i (plangen_base.py:525-565), `tokens is not None` branch (the one
    uni_generate uses, :401-406): reseed, tile by parallel_size, embed, loop, VQ."""
    dev = tokens.device
    if sampler is None:
        generator = torch.Generator(device=dev).manual_seed(seed)      # :526
        sampler = make_torch_sampler(generator)
    n_tok = image_token_num_per_image or d.n_img_tokens
    tokens = torch.cat([tokens] * parallel_size)
    mask = torch.cat([mask] * parallel_size)
    inputs_embeds = embed_tokens(sd, tokens)
    num_gen = inputs_embeds.shape[0] // 2
    generated = sample_image(sd, d, inputs_embeds, num_gen, n_tok, mask, cfg_weight, temperature,
                             sampler, edit_region, gt_labels, mode, trace)
    if not decode:
        return generated, None
    with _autocast_ctx(mode, dev):
        dec = decode_code(sd, d, generated.to(dtype=torch.int),
                          shape=[num_gen, d.code_dim, d.grid, d.grid])
    return generated, dec


# ------------------------------------------------------------- synthetic prompts
def synthetic_prompts(d: JanusDims, batch: int, seed: int = 1234, lo: int = 150, hi: int = 480,
                      neg_len: int = 110) -> Tuple[List[List[int]], List[List[int]]]:
    """LayoutSAM-shaped synthetic prompts: caption 20-60 tokens + 4-8
    boxes x (8-30 desc + ~20 markup) + ~12 template tokens, clipped to [lo, hi];
    one fixed negative prompt shared by all samples (cfg/base.py:129)."""
    g = torch.Generator(device="cpu").manual_seed(seed)

    def ri(a, b):
        return int(torch.randint(a, b + 1, (1,), generator=g).item())

    def ids(n):
        t = torch.randint(0, d.vocab - 1, (n,), generator=g)
        t = torch.where(t >= d.pad_id, t + 1, t).clamp_(max=d.vocab - 1)
        return t.tolist()

    neg = ids(neg_len)
    cond = []
    for _ in range(batch):
        n = ri(20, 60) + 12
        for _ in range(ri(4, 8)):
            n += ri(8, 30) + 20
        n = max(lo, min(hi, n))
        cond.append(ids(n))
    return cond, [list(neg) for _ in range(batch)]
