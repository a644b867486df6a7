"""Architecture numbers of the Janus models PlanGen decodes with (SURVEY.md §8 preamble;
three_party/Janus/janus_pro_tech_report.pdf Table 1; vq_model.py:31-43)."""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Tuple


@dataclass(frozen=True)
class Dims:
    name: str = "janus-1.3b"
    D: int = 2048
    L: int = 24
    H: int = 16
    head_dim: int = 128
    F: int = 5632
    vocab: int = 102400
    img_vocab: int = 16384
    code_dim: int = 8
    img_embed: int = 2048
    rms_eps: float = 1e-6
    rope_theta: float = 10000.0
    grid: int = 24
    vq_ch: int = 128
    vq_ch_mult: Tuple[int, ...] = (1, 1, 2, 2, 4)
    vq_z: int = 256
    vq_res_blocks: int = 2
    pad_id: int = 100002
    # understanding-side vision tower: SigLIP_MODEL_CONFIG["siglip_large_patch16_384"] (siglip_vit.py:628-637), the tower
    # of both Janus-1.3B and Janus-Pro-7B; `sig_mlp` = width * mlp_ratio
    sig_width: int = 1024
    sig_layers: int = 24
    sig_heads: int = 16
    sig_patch: int = 16
    sig_image: int = 384
    sig_mlp: int = 4096

    @property
    def n_img_tokens(self) -> int:
        return self.grid * self.grid

    @property
    def img_size(self) -> int:
        return self.grid * 2 ** (len(self.vq_ch_mult) - 1)

    @property
    def sig_patches(self) -> int:
        return (self.sig_image // self.sig_patch) ** 2

    @classmethod
    def from_any(cls, other, vision=None) -> "Dims":
        """Build from any object carrying the same attribute names (fields it lacks keep their defaults); `vision`:
        an object with width / layers / heads / patch / image / mlp_ratio (the oracle's SigLIPDims)."""
        kw = {f.name: (tuple(getattr(other, f.name)) if f.name == "vq_ch_mult" else getattr(other, f.name))
              for f in fields(cls) if hasattr(other, f.name)}
        if vision is not None:
            kw.update(sig_width=vision.width, sig_layers=vision.layers, sig_heads=vision.heads, sig_patch=vision.patch,
                      sig_image=vision.image, sig_mlp=int(vision.width * vision.mlp_ratio))
        return cls(**kw)


JANUS_1P3B = Dims()
JANUS_7B = Dims(name="janus-pro-7b", D=4096, L=30, H=32, F=11008, img_embed=4096)
