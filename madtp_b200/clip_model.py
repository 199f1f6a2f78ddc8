"""Mirror of the pruned encoder path of the reference's clip/model.py (+ the patched nn.MultiheadAttention of
clip/mock.py): LayerNorm (:160-166), QuickGELU (:169-171), ResidualAttentionBlock (:174-261), Transformer (:264-272),
VisionTransformer (:275-313) and CLIP.encode_image / encode_text (:482-503), with the reference's constructor
arguments, tuple-in / tuple-out block protocol, [N, B, C] tensor layout at the module boundary and state-dict keys
(`attn.in_proj_weight`, `attn.out_proj.*`, `ln_1`, `mlp.c_fc`, `mlp.c_proj`, `ln_2`, `query_model.q_map.0.*`).
The ResNet towers, momentum twins, queues and losses of that file are out of scope.

Per block (clip/model.py:236-261): the block's own Query_model with the learned q_map scores the tokens BEFORE
attention; pre-LN multi-head attention (q pre-scaled by 1/sqrt(head_dim), additive causal mask cropped to the current
length for text, clip/mock.py:309-310); Reduce_token with the `topk_num <= max_keep` guard (:220); QuickGELU MLP.
The vision tower runs the tensor-core attention path; the 77-token causal text tower the fp32 CUDA-core one.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
from torch import nn

from . import _lib as L
from . import functional as Fn
from .graphs import GraphedForward
from .utils import Query_model, vector_gather  # noqa: F401
from .vit import _eval_only


class LayerNorm(nn.LayerNorm):
    """clip/model.py:160-166 (fp32 LayerNorm); evaluated by the row kernel."""

    def forward(self, x: torch.Tensor):
        Fn.require_cuda(x, "x")
        shape = x.shape
        return Fn.layernorm_rows(x.reshape(-1, shape[-1]).contiguous(), self.weight, self.bias, self.eps,
                                 f32=True)["y"].view(shape)


class QuickGELU(nn.Module):
    def forward(self, x: torch.Tensor):
        return x * torch.sigmoid(1.702 * x)


class MultiheadAttention(nn.Module):
    """Parameter layout and accessors of the patched nn.MultiheadAttention (clip/mock.py:252-358)."""

    def __init__(self, embed_dim, num_heads):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.head_dim = embed_dim // num_heads
        if self.head_dim != 64:
            raise RuntimeError("madtp_b200: the attention kernels are built for head_dim 64")
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        self.attention_map = None
        self.cls_attn = None
        self._cache = Fn.WeightCache()

    def save_attention_map(self, attention_map):
        self.attention_map = attention_map

    def get_attention_map(self):
        return self.attention_map

    def save_cls_attn(self, cls_attn):
        self.cls_attn = cls_attn

    def get_cls_attn(self):
        return self.cls_attn

    def _prepared(self):
        qkv = self._cache.get("in", [self.in_proj_weight, self.in_proj_bias],
                              lambda: Fn.PreparedLinear(self.in_proj_weight, self.in_proj_bias, split=True))
        out = self._cache.get("out", [self.out_proj.weight, self.out_proj.bias],
                              lambda: Fn.PreparedLinear(self.out_proj.weight, self.out_proj.bias, f16=True))
        return qkv, out

    def rows(self, y_hi, y_lo, B, N, residual, want_stats, causal, n_dev=None):
        qkv_w, out_w = self._prepared()
        C = self.embed_dim
        scale = self.head_dim ** -0.5        # q / sqrt(E) then q.k: exact for head_dim 64 (a power of two)
        # image tower and the causal text tower (clip/mock.py:302-340) both run the tensor-core scoring-lane attention
        ctx16, stats = Fn.self_attention_tc(y_hi, y_lo, qkv_w, B, N, self.num_heads, scale, want_stats, causal=causal,
                                            n_dev=n_dev)
        self.save_attention_map(stats)
        self.save_cls_attn(None if stats is None else stats.cls_attn[:, 1:])
        return Fn.linear_f16(ctx16.view(B * N, C), out_w, residual=residual, m_dev=n_dev, m_mult=B)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None, sd_dim=768):
        super().__init__()
        self.attn = MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask           # only its presence matters: the text tower's mask is causal (:452-457)
        self.query_model = Query_model(ft_dim=d_model, sd_dim=sd_dim, temperature=1, att_func_type='sparsemax',
                                       pool_type='max', map_func=True)
        self.last_prune = None
        self._cache = Fn.WeightCache()

    def _mlp(self):
        fc, pj = self.mlp.c_fc, self.mlp.c_proj
        a = self._cache.get("fc", [fc.weight, fc.bias], lambda: Fn.PreparedLinear(fc.weight, fc.bias, f16=True))
        b = self._cache.get("pj", [pj.weight, pj.bias], lambda: Fn.PreparedLinear(pj.weight, pj.bias, f16=True))
        return a, b

    def forward_bnc(self, x, space_dict, temperature, sd_ft_all, max_keep, n_dev=None, n_out=None, k_out=None):
        """x [B, N, C] contiguous fp32 -> (x', sd_ft_all). n_dev / n_out / k_out: device-resident lengths (x is a
        capacity-sized buffer of packed sequences of *n_dev tokens; the block writes the next length to n_out and its
        topk_num to k_out -- functional.dtp_finish); needs the pruned path (a codebook and temperature > 0)."""
        B, N, C = x.shape
        with_dict = space_dict is not None
        prune = with_dict and temperature > 0
        if n_dev is not None and not prune:
            raise RuntimeError("madtp_b200: device-resident lengths are the PRUNED path")
        ln1 = Fn.layernorm_rows(x.view(B * N, C), self.ln_1.weight, self.ln_1.bias, self.ln_1.eps, split=True,
                                split_x=with_dict, n_dev=n_dev, n_mult=B)
        token_attn = None
        if with_dict:                                                                    # :239-245
            token_attn, sd_ft_all = self.query_model.forward_rows(x, ln1["x_hi"], ln1["x_lo"], space_dict, sd_ft_all,
                                                                  n_dev=n_dev)
        x1 = self.attn.rows(ln1["y_hi"], ln1["y_lo"], B, N, x.view(B * N, C), prune, self.attn_mask is not None,
                            n_dev=n_dev)
        x1 = x1.view(B, N, C)
        self.last_prune = None
        nd2 = n_dev
        if prune:                                                                        # :254-258
            pend = Fn.dtp_score_async(self.attn.get_attention_map(), token_attn, float(temperature), N - 1, n_dev=n_dev)
            res = Fn.dtp_finish(x1, pend, max_keep=int(max_keep), n_dev=n_dev, n_out=n_out, k_out=k_out,
                                ln=(self.ln_2.weight, self.ln_2.bias, self.ln_2.eps))
            self.last_prune = res
            x1 = res.x
            if n_dev is not None:
                nd2 = n_out
        N2 = x1.shape[1]
        x2d = x1.view(B * N2, C)
        fc, pj = self._mlp()
        if prune and res.ln16 is not None:      # ln_2 came out of the fused select + gather kernel
            y16 = res.ln16.view(B * N2, C)
        else:
            y16 = Fn.layernorm_rows(x2d, self.ln_2.weight, self.ln_2.bias, self.ln_2.eps, f16=True, n_dev=nd2,
                                    n_mult=B)["y16"]
        h = Fn.linear_f16(y16, fc, out_dtype=torch.float16, act=L.ACT_QUICKGELU, m_dev=nd2, m_mult=B)
        return Fn.linear_f16(h, pj, residual=x2d, m_dev=nd2, m_mult=B).view(B, N2, C), sd_ft_all

    @torch.no_grad()
    def forward(self, inputs):
        # x: (N, B, C)
        x, space_dict, temperature, sd_ft_all, max_keep = inputs
        Fn.require_cuda(x, "x")
        _eval_only(self)
        y, sd_ft_all = self.forward_bnc(x.permute(1, 0, 2).contiguous(), space_dict, temperature, sd_ft_all, max_keep)
        return y.permute(1, 0, 2), space_dict, temperature, sd_ft_all, max_keep


def _blocks_device(blocks, y, space_dict, temperature, max_keep):
    """The residual blocks with device-resident lengths: y [B, N, C] stays a capacity-sized buffer of packed sequences,
    nothing is read back. Returns (y, sd_ft_all, Trajectory, device scalar with the final tokens per sequence)."""
    B, N, _ = y.shape
    depth = len(blocks)
    lens = L.empty((2 * depth + 1,), torch.int32, y.device)
    dims, ks = lens[:depth + 1], lens[depth + 1:]
    dims[:1].fill_(N)
    ks.fill_(-1)
    traj = Fn.Trajectory(dims, ks)
    sd_ft_all = None
    for i, blk in enumerate(blocks):
        y, sd_ft_all = blk.forward_bnc(y, space_dict, temperature, sd_ft_all, max_keep, n_dev=dims[i:i + 1],
                                       n_out=dims[i + 1:i + 2], k_out=ks[i:i + 1])
        blk.last_prune = Fn.LazyPrune(traj, i, blk.last_prune, B)
    return y, sd_ft_all, traj, dims[depth:depth + 1]


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None, sd_dim=768):
        super().__init__()
        self.width = width
        self.layers = layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask, sd_dim=sd_dim)
                                         for _ in range(layers)])

    @torch.no_grad()
    def forward(self, x: torch.Tensor, space_dict=None, temperature=0, sd_ft_all=None, max_keep=1):
        """x [N, B, C] -> the reference's 5-tuple (:271-272); one layout change in, one out."""
        Fn.require_cuda(x, "x")
        _eval_only(self)
        y = x.permute(1, 0, 2).contiguous()
        for blk in self.resblocks:
            y, sd_ft_all = blk.forward_bnc(y, space_dict, temperature, sd_ft_all, max_keep)
        return y.permute(1, 0, 2), space_dict, temperature, sd_ft_all, max_keep


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int,
                 sd_dim=768):
        super().__init__()
        self.input_resolution = input_resolution
        self.output_dim = output_dim
        self.patch_size = patch_size
        self.conv1 = nn.Conv2d(in_channels=3, out_channels=width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads, sd_dim=sd_dim)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self._cache = Fn.WeightCache()

    @torch.no_grad()
    def forward(self, x: torch.Tensor, space_dict=None, temperature=0, max_keep=1, _device=False):
        Fn.require_cuda(x, "image")
        _eval_only(self)
        B, _, Hh, Ww = x.shape
        P = self.patch_size
        conv = self._cache.get("conv1", [self.conv1.weight], lambda: Fn.PreparedLinear(
            self.conv1.weight.reshape(self.conv1.weight.shape[0], -1), None, split=True))
        hi, lo = L.patchify(x.contiguous(), P)
        n = (Hh // P) * (Ww // P)
        patches = Fn.linear_split(hi, lo, conv)
        C = patches.shape[1]
        tok = L.assemble_tokens(patches, self.class_embedding.detach(), self.positional_embedding.detach(), B, n, C)
        y = self.ln_pre(tok)
        if _device:     # device-resident lengths: the CLS rows are taken from the packed stream, nothing is read back
            y, sd_img_ft_all, traj, n_dev = _blocks_device(self.transformer.resblocks, y, space_dict, temperature, max_keep)
            cls = self.ln_post(L.take_token(y.view(-1, C), B, y.shape[1], 0, n_dev=n_dev))
            proj = self._cache.get("proj", [self.proj], lambda: Fn.PreparedLinear(self.proj.t(), None, f32=True))
            return (Fn.linear_f32(cls, proj), sd_img_ft_all), [traj]
        sd_img_ft_all = None
        for blk in self.transformer.resblocks:
            y, sd_img_ft_all = blk.forward_bnc(y, space_dict, temperature if space_dict is not None else 0,
                                               sd_img_ft_all, max_keep)
        cls = self.ln_post(y[:, 0, :].contiguous())
        proj = self._cache.get("proj", [self.proj], lambda: Fn.PreparedLinear(self.proj.t(), None, f32=True))
        return Fn.linear_f32(cls, proj), sd_img_ft_all


class CLIP(GraphedForward, nn.Module):
    """The two encoders of clip/model.py:CLIP (ViT towers only) with `encode_image` / `encode_text` (:482-503).
    Constructor arguments and their order are the reference's (:317-332), including the positional `evaluate`; the
    momentum twins, queues and the ResNet towers (:383-437, training only) are out of scope, so their state-dict keys
    (`*_m.*`, `*_queue`) are reported as unexpected by `load_state_dict(strict=False)` exactly like any extra key."""

    def __init__(self, embed_dim: int, image_resolution: int, vision_layers: int, vision_width: int,
                 vision_patch_size: int, context_length: int, vocab_size: int, transformer_width: int,
                 transformer_heads: int, transformer_layers: int, evaluate: bool = False, config=None):
        super().__init__()
        if isinstance(evaluate, dict):
            raise TypeError("madtp_b200.CLIP: the 11th positional argument is `evaluate` (clip/model.py:330); "
                            "pass the config dict as `config=`")
        if isinstance(vision_layers, (tuple, list)):
            raise NotImplementedError("madtp_b200: the ModifiedResNet towers of clip/model.py are out of scope")
        self.sd_num, self.sd_dim = (100, 768) if config is None else (config['sd_num'], config['sd_dim'])
        self.space_dict = nn.Parameter(torch.randn(self.sd_num, self.sd_dim))
        self.context_length = context_length
        self.visual = VisionTransformer(input_resolution=image_resolution, patch_size=vision_patch_size,
                                        width=vision_width, layers=vision_layers, heads=vision_width // 64,
                                        output_dim=embed_dim, sd_dim=self.sd_dim)
        self.transformer = Transformer(width=transformer_width, layers=transformer_layers, heads=transformer_heads,
                                       attn_mask=self.build_attention_mask(), sd_dim=self.sd_dim)
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592)            # log(1 / 0.07), :381
        if not evaluate:
            self.initialize_parameters()
        self.tokenize = None
        self.vision_layers = vision_layers
        self.transformer_layers = transformer_layers
        self.embed_dim = embed_dim
        self._cache = Fn.WeightCache()

    def initialize_parameters(self):
        """clip/model.py:439-470 (ViT towers)."""
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (self.transformer.width ** -0.5) * ((2 * self.transformer.layers) ** -0.5)
        attn_std = self.transformer.width ** -0.5
        fc_std = (2 * self.transformer.width) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=self.transformer.width ** -0.5)

    def build_attention_mask(self):
        mask = torch.empty(self.context_length, self.context_length)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    @torch.no_grad()
    def encode_image(self, image, space_dict=None, temperature=0):
        image = image.type(self.dtype)
        if space_dict is not None and self._device_path(temperature, image):
            # device-resident lengths end to end (the reference syncs once per block, clip/model.py:219); a CUDA-graph
            # replay when enable_cuda_graphs(True)
            t = float(temperature)
            emb, sd_ft = self._run_device("image", lambda im: self.visual(im, space_dict=space_dict, temperature=t,
                                                                          _device=True), [image], t,
                                          extra=(space_dict.data_ptr(),))
            return emb, sd_ft
        return self.visual(image, space_dict=space_dict, temperature=temperature)

    @torch.no_grad()
    def encode_text(self, text, space_dict=None, temperature=0):
        if not text.is_cuda:
            raise RuntimeError("madtp_b200: text ids must be a CUDA tensor -- this package has no CPU fallback")
        max_keep = int(text.argmax(dim=-1).max()) + 2                                      # :495
        if space_dict is not None and self._device_path(temperature, text):
            # max_keep is a property of the input ids (one read before anything is launched) and part of the graph key
            t = float(temperature)
            emb, sd_ft = self._run_device("text", lambda ids: self._encode_text_device(ids, space_dict, t, max_keep),
                                          [text.to(torch.int64)], t, extra=(max_keep, space_dict.data_ptr()))
            return emb, sd_ft
        x = L.bert_embed(text.to(torch.int64), self.token_embedding.weight.detach(),
                         self.positional_embedding.detach())                               # :490-492
        y = x
        sd_txt_ft_all = None
        for blk in self.transformer.resblocks:
            y, sd_txt_ft_all = blk.forward_bnc(y, space_dict, temperature if space_dict is not None else 0,
                                               sd_txt_ft_all, max_keep)
        y = self.ln_final(y)
        eot = y[torch.arange(y.shape[0], device=y.device), text.argmax(dim=-1)].contiguous()   # :501
        proj = self._cache.get("tp", [self.text_projection],
                               lambda: Fn.PreparedLinear(self.text_projection.t(), None, f32=True))
        return Fn.linear_f32(eot, proj), sd_txt_ft_all

    def _encode_text_device(self, text, space_dict, temperature, max_keep):
        """CLIP.encode_text with device-resident lengths: the blocks run on capacity-sized packed buffers, ln_final on the
        dynamic row count, and the EOT row of every sequence (clip/model.py:501: index argmax(text) into the PRUNED
        sequence) is gathered from the packed stream with indices formed on the device."""
        B, N = text.shape
        y = L.bert_embed(text, self.token_embedding.weight.detach(), self.positional_embedding.detach())
        y, sd_ft, traj, n_dev = _blocks_device(self.transformer.resblocks, y, space_dict, temperature, max_keep)
        C = y.shape[-1]
        yn = Fn.layernorm_rows(y.view(B * N, C), self.ln_final.weight, self.ln_final.bias, self.ln_final.eps, f32=True,
                               n_dev=n_dev, n_mult=B)["y"]
        rows = torch.arange(B, device=y.device, dtype=torch.int64) * n_dev.to(torch.int64) + text.argmax(dim=-1)
        eot = yn.index_select(0, rows)
        proj = self._cache.get("tp", [self.text_projection],
                               lambda: Fn.PreparedLinear(self.text_projection.t(), None, f32=True))
        return (Fn.linear_f32(eot, proj), sd_ft), [traj]


def convert_weights(model: nn.Module):
    """clip/model.py:655-676: Conv / Linear / attention in-projection weights and the two projection matrices to fp16.
    `clip.load` (clip/clip.py:144-148) calls `.float()` on the result, so the net effect on a checkpoint is ONE fp16
    rounding of those weights -- reproduced literally, because it changes the scores."""
    def _convert(l):
        if isinstance(l, (nn.Conv1d, nn.Conv2d, nn.Linear)):
            l.weight.data = l.weight.data.half()
            if l.bias is not None:
                l.bias.data = l.bias.data.half()
        if isinstance(l, (nn.MultiheadAttention, MultiheadAttention)):
            for attr in ["in_proj_weight", "q_proj_weight", "k_proj_weight", "v_proj_weight", "in_proj_bias", "bias_k",
                         "bias_v"]:
                t = getattr(l, attr, None)
                if t is not None:
                    t.data = t.data.half()
        for name in ["text_projection", "proj"]:
            if hasattr(l, name):
                attr = getattr(l, name)
                if attr is not None and torch.is_tensor(attr):
                    attr.data = attr.data.half()
    model.apply(_convert)


def build_model(state_dict: dict, evaluate: bool = False, config=None):
    """clip/model.py:678-716: derive the architecture from a checkpoint's shapes, convert to fp16, load (strict=False).
    The caller (clip/clip.py:144-148) moves the model to the device and calls `.float()`."""
    if "visual.proj" not in state_dict:
        raise NotImplementedError("madtp_b200: the ModifiedResNet towers of clip/model.py are out of scope")
    vision_width = state_dict["visual.conv1.weight"].shape[0]
    vision_layers = len([k for k in state_dict.keys() if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
    vision_patch_size = state_dict["visual.conv1.weight"].shape[-1]
    grid_size = round((state_dict["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    image_resolution = vision_patch_size * grid_size
    embed_dim = state_dict["text_projection"].shape[1]
    context_length = state_dict["positional_embedding"].shape[0]
    vocab_size = state_dict["token_embedding.weight"].shape[0]
    transformer_width = state_dict["ln_final.weight"].shape[0]
    transformer_heads = transformer_width // 64
    transformer_layers = len(set(k.split(".")[2] for k in state_dict if k.startswith("transformer.resblocks")))
    model = CLIP(embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size, context_length, vocab_size,
                 transformer_width, transformer_heads, transformer_layers, evaluate, config)
    for key in ["input_resolution", "context_length", "vocab_size"]:
        if key in state_dict:
            del state_dict[key]
    convert_weights(model)
    model.load_state_dict(state_dict, strict=False)
    return model
