"""Import shims that let the UNMODIFIED reference (/root/reference) run on this image -- TEST INFRASTRUCTURE.

Used only by oracle/gen_golden.py (fixture generation, in the build container where /root/reference exists). Nothing
here travels into the product; nothing here edits a reference file. Every shim is a stand-in for a THIRD-PARTY
package the reference pins but this image lacks:

  timm==0.4.12        PatchEmbed (Conv2d(3,D,P,P) -> flatten(2).transpose(1,2)), DropPath, trunc_normal_, _cfg, ...
  fairscale           checkpoint_wrapper (identity)
  tkinter.messagebox  stray import in models/blip_retrieval.py:1
  transformers 4.15   apply_chunking_to_forward / find_pruneable_heads_and_indices / prune_linear_layer moved out of
                      transformers.modeling_utils; PreTrainedModel.get_head_mask removed
  BertTokenizer       ./pretrained/bert-base-uncased/ is absent -> a fake tokenizer with the ids blip.py:222-224 adds
"""
from __future__ import annotations

import os
import sys
import types

import torch
from torch import nn

def _resolve_root() -> str:
    """$MADTP_REFERENCE_ROOT, else /root/reference (build container), else oracle/_ref (the byte-for-byte staging of the
    hot-path files made by oracle/make_ref.py, which is what exists on the GPU box)."""
    env = os.environ.get("MADTP_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/models"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _resolve_root()
_installed = False


def _module(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class PatchEmbed(nn.Module):
    """timm 0.4.12 timm/models/layers/patch_embed.py semantics."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, flatten=True):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        x = self.proj(x)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


class FakeTokenizer:
    """Deterministic stand-in for BertTokenizer + the two tokens models/blip.py:222-224 adds (vocab 30522 + 2)."""
    bos_token_id = 30522
    enc_token_id = 30523
    pad_token_id = 0
    sep_token_id = 102
    cls_token_id = 101
    additional_special_tokens_ids = [30523]

    def __init__(self):
        self.next_ids = None       # set by the fixture generator: (input_ids [B,L] long, attention_mask [B,L] long)

    def __call__(self, text, padding=None, truncation=None, max_length=None, return_tensors="pt", **kw):
        assert self.next_ids is not None, "FakeTokenizer: set .next_ids = (input_ids, attention_mask) first"
        ids, mask = self.next_ids

        class _Enc:
            def __init__(s, i, m):
                s.input_ids, s.attention_mask = i.clone(), m.clone()

            def to(s, device):
                s.input_ids, s.attention_mask = s.input_ids.to(device), s.attention_mask.to(device)
                return s
        return _Enc(ids, mask)


def install():
    """Idempotent. Must run before any `import models.*` of the reference."""
    global _installed
    if _installed:
        return
    import transformers  # noqa: F401  (must be imported BEFORE timm is stubbed: it probes timm.__spec__)
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu

    for name in ("apply_chunking_to_forward", "find_pruneable_heads_and_indices", "prune_linear_layer"):
        if not hasattr(mu, name) and hasattr(pu, name):
            setattr(mu, name, getattr(pu, name))
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        mu.find_pruneable_heads_and_indices = lambda *a, **k: (set(), None)
    if not hasattr(mu.PreTrainedModel, "get_head_mask"):
        def get_head_mask(self, head_mask, num_hidden_layers, is_attention_chunked=False):
            assert head_mask is None
            return [None] * num_hidden_layers
        mu.PreTrainedModel.get_head_mask = get_head_mask

    def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)

    _module("timm")
    _module("timm.models")
    _module("timm.models.vision_transformer", _cfg=lambda **kw: dict(kw), PatchEmbed=PatchEmbed)
    _module("timm.models.registry", register_model=lambda f: f)
    _module("timm.models.layers", trunc_normal_=trunc_normal_, DropPath=DropPath)
    _module("timm.models.helpers", named_apply=lambda *a, **k: None, adapt_input_conv=lambda *a, **k: None)
    _module("timm.models.hub", download_cached_file=lambda *a, **k: None)
    _module("fairscale")
    _module("fairscale.nn")
    _module("fairscale.nn.checkpoint")
    _module("fairscale.nn.checkpoint.checkpoint_activations", checkpoint_wrapper=lambda m, *a, **k: m)
    if "tkinter" not in sys.modules:
        try:
            import tkinter.messagebox  # noqa: F401
        except Exception:
            _module("tkinter")
            _module("tkinter.messagebox", NO="no")
    try:
        import ftfy  # noqa: F401
    except Exception:
        _module("ftfy", fix_text=lambda s: s)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def build_blip_nlvr(image_size: int = 384):
    """Unmodified reference BLIP_NLVR(evaluate=True) with the fake tokenizer. Returns (model, tokenizer)."""
    install()
    import models.blip as blip
    tok = FakeTokenizer()
    blip.init_tokenizer = lambda: tok
    import models.blip_nlvr as bn
    bn.init_tokenizer = lambda: tok
    model = bn.BLIP_NLVR(med_config=os.path.join(REFERENCE_ROOT, "configs/med_config.json"), image_size=image_size,
                         vit="base", evaluate=True)
    model.eval()
    return model, tok


def install_clip():
    """Extra stand-ins for clip/mock.py, which forks torch 1.11's nn.MultiheadAttention and imports 1.11 privates
    (clip/mock.py:1-4). Only `_scaled_dot_product_attention(q, k, v, attn_mask, dropout_p) -> (out, attn)` is gone in
    torch 2.x; it is restated here from torch 1.11 torch/nn/functional.py: q / sqrt(E), baddbmm with the additive mask,
    softmax, dropout, bmm."""
    install()
    import math
    import typing
    import warnings

    import torch.nn.functional as F
    import torch.nn.modules.activation as act

    for name, val in (("torch", torch), ("Optional", typing.Optional), ("Tuple", typing.Tuple)):
        if not hasattr(act, name):
            setattr(act, name, val)
    if not hasattr(F, "_scaled_dot_product_attention"):
        def _scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0):
            B, Nt, E = q.shape
            q = q / math.sqrt(E)
            if attn_mask is not None:
                attn = torch.baddbmm(attn_mask, q, k.transpose(-2, -1))
            else:
                attn = torch.bmm(q, k.transpose(-2, -1))
            attn = F.softmax(attn, dim=-1)
            if dropout_p > 0.0:
                attn = F.dropout(attn, p=dropout_p)
            return torch.bmm(attn, v), attn
        F._scaled_dot_product_attention = _scaled_dot_product_attention
    for name, val in (("math", math), ("warnings", warnings)):
        if not hasattr(F, name):
            setattr(F, name, val)
