"""ORACLE tooling (test infrastructure, NOT product code): the handful of `timm` names the reference's
`three_party/Janus/janus/models/siglip_vit.py` imports, so that file - which defines its own Attention / Block /
VisionTransformer - can be executed in a container without timm.  Third-party dependency: `timm` (requirements of
Janus: timm>=0.9.16; absent offline).  Restated from timm's published definitions:
  PatchEmbed  Conv2d(in_chans, embed_dim, kernel = stride = patch) -> flatten(2).transpose(1, 2)   (norm: Identity)
  Mlp         fc1 -> act -> drop1 -> norm(Identity) -> fc2 -> drop2
  DropPath / PatchDropout   identity at inference; AttentionPoolLatent: only constructed, never called with ignore_head
`install()` injects the stub into sys.modules only when the real timm is missing."""
import sys
import types

import torch.nn as nn


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, flatten=True, bias=True,
                 dynamic_img_pad=False, **kw):
        super().__init__()
        img = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        pat = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.img_size, self.patch_size = img, pat
        self.grid_size = (img[0] // pat[0], img[1] // pat[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=pat, stride=pat, bias=bias)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        return self.norm(self.proj(x).flatten(2).transpose(1, 2))


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, norm_layer=None, bias=True,
                 drop=0.0, **kw):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.norm = norm_layer(hidden_features) if norm_layer else nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class _Identity(nn.Module):
    def __init__(self, *a, **kw):
        super().__init__()

    def forward(self, x):
        return x


def install():
    try:
        import timm  # noqa: F401
        return False
    except ImportError:
        pass
    import importlib.machinery

    def mk(name):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)     # importlib.util.find_spec() rejects spec-less modules
        return m

    timm, layers, models, manip = mk("timm"), mk("timm.layers"), mk("timm.models"), mk("timm.models._manipulate")
    timm.__version__ = "0.0.0-oracle-stub"
    layers.PatchEmbed, layers.Mlp = PatchEmbed, Mlp
    layers.DropPath = layers.PatchDropout = layers.AttentionPoolLatent = _Identity
    layers.LayerType = object
    layers.resample_abs_pos_embed = lambda pos, *a, **kw: pos
    manip.checkpoint_seq = lambda fn, x, **kw: fn(x)
    manip.named_apply = lambda fn, module, **kw: module
    timm.layers, timm.models, models._manipulate = layers, models, manip
    sys.modules.update({"timm": timm, "timm.layers": layers, "timm.models": models, "timm.models._manipulate": manip})
    return True
