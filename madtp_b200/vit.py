"""Mirror of the reference's models/vit.py hot path: Mlp (:15-36), Attention (:39-103), Block (:106-207) and
VisionTransformer (:210-315) with the same constructor arguments, forward signatures, return arity and state-dict
keys, executing on the sm_100a kernels of libmadtp_b200.so.

Differences that are part of the contract:
  * evaluation forward only (`module.eval()`, no autograd through the kernels); dropout / DropPath are identities;
  * the [B,H,N,N] attention map is never materialised: `Attention.get_attention_map()` returns an `AttnStats` handle
    (per-tile column sums of max_h P and cls_attn), which is what `Block.Reduce_token` consumes;
  * survivors are kept in ascending token order (the reference's `topk(sorted=False)` order is implementation-defined);
  * `register_hook=True` (Grad-CAM hooks on P) is not supported on this path.
"""
from __future__ import annotations

import os
from functools import partial

import torch
from torch import nn

from . import _lib as L
from . import functional as Fn
from .utils import Query_model, vector_gather  # noqa: F401  (models/vit.py does `from models.utils import *`)


def _act_code(act: nn.Module) -> int:
    if isinstance(act, nn.GELU):
        return L.FFN_GELU
    if isinstance(act, nn.ReLU):
        return L.ACT_RELU
    if isinstance(act, nn.Identity):
        return L.ACT_NONE
    if type(act).__name__ == "QuickGELU":
        return L.ACT_QUICKGELU
    raise RuntimeError(f"madtp_b200: unsupported activation {type(act).__name__}")


def _eval_only(m: nn.Module):
    if m.training:
        raise RuntimeError("madtp_b200 implements the evaluation forward only: call .eval() first")


class Mlp(nn.Module):
    """fc2(act(fc1(x)))  (reference models/vit.py:15-36); fp16 tensor-core GEMMs with fp32 accumulation."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.hidden_features = hidden_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        self._cache = Fn.WeightCache()

    def _prepared(self):
        fc1 = self._cache.get("fc1", [self.fc1.weight, self.fc1.bias],
                              lambda: Fn.PreparedLinear(self.fc1.weight, self.fc1.bias, f16=True))
        fc2 = self._cache.get("fc2", [self.fc2.weight, self.fc2.bias],
                              lambda: Fn.PreparedLinear(self.fc2.weight, self.fc2.bias, f16=True))
        return fc1, fc2

    def forward_rows(self, y16, residual=None, m_dev=None, m_mult=1):
        """y16 [rows, C] fp16 (already normalised) -> fp32 [rows, C] (+ residual). m_dev / m_mult: device-resident
        row count rows = *m_dev * m_mult."""
        fc1, fc2 = self._prepared()
        h = Fn.linear_f16(y16, fc1, out_dtype=torch.float16, act=_act_code(self.act), m_dev=m_dev, m_mult=m_mult)
        return Fn.linear_f16(h, fc2, residual=residual, m_dev=m_dev, m_mult=m_mult)

    def forward_rows_split(self, y_hi, y_lo, residual=None, m_dev=None, m_mult=1):
        """The same FFN at the scoring lane's precision (functional.value_lane_split): hi/lo planes in, exact erf GELU,
        fp32 hidden activations re-split for fc2."""
        fc1 = self._cache.get("fc1s", [self.fc1.weight, self.fc1.bias],
                              lambda: Fn.PreparedLinear(self.fc1.weight, self.fc1.bias, split=True))
        fc2 = self._cache.get("fc2s", [self.fc2.weight, self.fc2.bias],
                              lambda: Fn.PreparedLinear(self.fc2.weight, self.fc2.bias, split=True))
        act = _act_code(self.act)
        h = Fn.linear_split(y_hi, y_lo, fc1, act=L.ACT_GELU if act == L.FFN_GELU else act, m_dev=m_dev, m_mult=m_mult)
        # the row kernel takes rows of <= 1024 elements: split the dense [rows, 4 C] activations as [4 rows, C]
        rows, dh = h.shape
        parts = dh // y_hi.shape[1]
        h_hi, h_lo = Fn.split_rows(h.view(rows * parts, dh // parts), n_dev=m_dev, n_mult=m_mult * parts)
        h_hi, h_lo = h_hi.view(rows, dh), h_lo.view(rows, dh)
        return Fn.linear_split(h_hi, h_lo, fc2, residual=residual, m_dev=m_dev, m_mult=m_mult)

    def forward(self, x):
        Fn.require_cuda(x, "x")
        _eval_only(self)
        shape = x.shape
        y16 = L.cast_f16(x.reshape(-1, shape[-1]))
        return self.forward_rows(y16).view(*shape[:-1], -1)


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.dim = dim
        self.num_heads = num_heads
        head_dim = dim // num_heads
        if head_dim != 64:
            raise RuntimeError("madtp_b200: the attention kernels are built for head_dim 64")
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj_drop = nn.Dropout(proj_drop)
        self.attn_gradients = None
        self.attention_map = None
        self.cls_attn = None
        self._cache = Fn.WeightCache()

    # reference accessors (models/vit.py:57-73)
    def save_attn_gradients(self, attn_gradients):
        self.attn_gradients = attn_gradients

    def get_attn_gradients(self):
        return self.attn_gradients

    def save_attention_map(self, attention_map):
        self.attention_map = attention_map

    def get_attention_map(self):
        return self.attention_map

    def save_cls_attn(self, cls_attn):
        self.cls_attn = cls_attn

    def get_cls_attn(self):
        return self.cls_attn

    def _prepared(self):
        qkv = self._cache.get("qkv", [self.qkv.weight, self.qkv.bias],
                              lambda: Fn.PreparedLinear(self.qkv.weight, self.qkv.bias, split=True))
        proj = self._cache.get("proj", [self.proj.weight, self.proj.bias],
                               lambda: Fn.PreparedLinear(self.proj.weight, self.proj.bias, f16=True))
        return qkv, proj

    def forward_rows(self, y_hi, y_lo, B, N, residual=None, want_stats=True):
        """y_hi/y_lo: fp16 hi/lo split of the normalised input rows [B*N, C]. Returns fp32 [B*N, C] = proj(ctx) (+residual)
        and stores the pruning statistics (models/vit.py:83,96-101)."""
        ctx16 = self.attend_rows(y_hi, y_lo, B, N, want_stats)
        return self.project_rows(ctx16, B, N, residual)

    def attend_rows(self, y_hi, y_lo, B, N, want_stats=True, n_dev=None):
        """q|k|v projection, attention and (optionally) the pruning statistics; returns the fp16 context [B,N,C]."""
        qkv_w, _ = self._prepared()
        # opt-in value lane at scoring precision: the kernel also writes the context in fp32
        self._ctx32 = L.empty((B, N, self.dim), torch.float32, y_hi.device) if Fn.value_lane_split() else None
        ctx16, stats = Fn.self_attention_tc(y_hi, y_lo, qkv_w, B, N, self.num_heads, self.scale, want_stats,
                                            n_dev=n_dev, ctx32=self._ctx32)
        self.save_attention_map(stats)
        self.save_cls_attn(None if stats is None else stats.cls_attn[:, 1:])
        return ctx16

    def project_rows(self, ctx16, B, N, residual=None, n_dev=None):
        _, proj_w = self._prepared()
        ctx32 = getattr(self, "_ctx32", None)
        if ctx32 is not None and Fn.value_lane_split():
            self._ctx32 = None
            w = self._cache.get("projs", [self.proj.weight, self.proj.bias],
                                lambda: Fn.PreparedLinear(self.proj.weight, self.proj.bias, split=True))
            c_hi, c_lo = Fn.split_rows(ctx32.view(B * N, self.dim), n_dev=n_dev, n_mult=B)
            return Fn.linear_split(c_hi, c_lo, w, residual=residual, m_dev=n_dev, m_mult=B)
        return Fn.linear_f16(ctx16.view(B * N, self.dim), proj_w, residual=residual, m_dev=n_dev, m_mult=B)

    def forward(self, x, register_hook=False):
        Fn.require_cuda(x, "x")
        _eval_only(self)
        if register_hook:
            raise NotImplementedError("madtp_b200: register_hook needs the materialised attention map")
        B, N, C = x.shape
        y_hi, y_lo = Fn.split_rows(x.reshape(B * N, C))
        return self.forward_rows(y_hi, y_lo, B, N).view(B, N, C)


class Block(nn.Module):

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, use_grad_checkpointing=False):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)
        self.drop_path = nn.Identity()      # DropPath is the identity in evaluation
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop)
        self.last_prune = None              # Fn.PruneResult of the most recent forward (diagnostics / parity tests)

    def Reduce_token(self, x, reduce_num=0, temperature=0, self_attn=None, cls_attn=None, token_attn=None):
        """x [B,n,d] prunable tokens -> [B,k+1,d] (survivors in ascending order + merged token), or x unchanged.
        `self_attn` is the AttnStats handle from Attention.get_attention_map() (or a materialised [B,H,N,N] map)."""
        Fn.require_cuda(x, "x")
        B, n, d = x.shape
        if isinstance(self_attn, Fn.AttnStats):
            stats = self_attn
        else:   # a materialised map: reduce it to the same statistics (off the fast path)
            a = self_attn[:, :, 1:, 1:].max(1)[0].sum(dim=1)
            col = torch.zeros(B, 1, n + 1, device=x.device, dtype=torch.float32)
            col[:, 0, 1:] = a
            ca = torch.zeros(B, n + 1, device=x.device, dtype=torch.float32)
            ca[:, 1:] = cls_attn
            stats = Fn.AttnStats(col, ca)
        xin = torch.cat([x[:, :1, :], x], dim=1).contiguous()      # slot 0 is a placeholder for the CLS row
        res = Fn.dtp_prune(xin, stats, token_attn, float(temperature))
        self.last_prune = res
        return res.x[:, 1:, :] if res.pruned else x

    def forward_rows(self, x, ln1, temperature=0, token_attn=None, n_dev=None, n_out=None, k_out=None):
        """x [B,N,C] fp32 contiguous; ln1 = functional.layernorm_rows(..., split=True) of norm1(x).
        n_dev / n_out / k_out: device-resident lengths (x is a capacity-sized buffer of packed sequences, *n_dev tokens
        each; the layer writes the next length to n_out and its topk_num to k_out -- see functional.dtp_finish)."""
        B, N, C = x.shape
        prune = temperature > 0
        ctx16 = self.attn.attend_rows(ln1["y_hi"], ln1["y_lo"], B, N, want_stats=prune, n_dev=n_dev)
        # the score kernel and the read-back of topk_num go first; the output projection does not depend on them and
        # runs while the host waits for the four bytes
        pend = Fn.dtp_score_async(self.attn.get_attention_map(), token_attn, float(temperature), N - 1,
                                  n_dev=n_dev) if prune else None
        x1 = self.attn.project_rows(ctx16, B, N, residual=x.view(B * N, C), n_dev=n_dev).view(B, N, C)
        self.last_prune = None
        nd2 = n_dev
        y16 = None
        if prune:   # select + gather + merged token + norm2 in one kernel (dtp_apply.cu)
            res = Fn.dtp_finish(x1, pend, n_dev=n_dev, n_out=n_out, k_out=k_out,
                                ln=(self.norm2.weight, self.norm2.bias, self.norm2.eps))
            self.last_prune = res
            x1 = res.x
            if res.ln16 is not None:
                y16 = res.ln16.view(-1, C)
            if n_dev is not None:
                nd2 = n_out
        N2 = x1.shape[1]
        x2d = x1.view(B * N2, C)
        if Fn.value_lane_split():   # diagnostic: norm2 as hi/lo planes, FFN on the error-compensated lane
            ln2 = Fn.layernorm_rows(x2d, self.norm2.weight, self.norm2.bias, self.norm2.eps, split=True, n_dev=nd2,
                                    n_mult=B)
            return self.mlp.forward_rows_split(ln2["y_hi"], ln2["y_lo"], residual=x2d, m_dev=nd2,
                                               m_mult=B).view(B, N2, C)
        if y16 is None:     # nothing was pruned (host-side early-out) or pruning is off
            y16 = Fn.layernorm_rows(x2d, self.norm2.weight, self.norm2.bias, self.norm2.eps, f16=True, n_dev=nd2,
                                    n_mult=B)["y16"]
        return self.mlp.forward_rows(y16, residual=x2d, m_dev=nd2, m_mult=B).view(B, N2, C)

    def forward(self, x, register_hook=False, reduce_num=0, temperature=0, token_attn=None):
        Fn.require_cuda(x, "x")
        _eval_only(self)
        if register_hook:
            raise NotImplementedError("madtp_b200: register_hook needs the materialised attention map")
        if temperature > 0 and token_attn is None:
            raise RuntimeError("madtp_b200: temperature > 0 needs token_attn (models/vit.py:131)")
        x = x.contiguous()
        B, N, C = x.shape
        ln1 = Fn.layernorm_rows(x.view(B * N, C), self.norm1.weight, self.norm1.bias, self.norm1.eps, split=True)
        return self.forward_rows(x, ln1, temperature, token_attn)


class PatchEmbed(nn.Module):
    """timm 0.4.12 PatchEmbed (Conv2d(in_chans, D, P, stride P) -> flatten -> transpose); same parameter names."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.Identity()
        self._cache = Fn.WeightCache()

    def forward(self, x):
        Fn.require_cuda(x, "image")
        B, Cin, H, W = x.shape
        P = self.patch_size[0]
        w = self._cache.get("proj", [self.proj.weight, self.proj.bias],
                            lambda: Fn.PreparedLinear(self.proj.weight.reshape(self.proj.weight.shape[0], -1),
                                                      self.proj.bias, split=True))
        hi, lo = L.patchify(x, P)
        return Fn.linear_split(hi, lo, w).view(B, (H // P) * (W // P), -1)


class VisionTransformer(nn.Module):
    """reference models/vit.py:210-315 (ViT encoder with per-layer Dynamic Token Pruning)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=True, qk_scale=None, representation_size=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., norm_layer=None,
                 use_grad_checkpointing=False, ckpt_layer=0, evaluate=False, sd_dim=768, map_func=False):
        super().__init__()
        self.num_features = self.embed_dim = embed_dim
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=0.0, norm_layer=norm_layer)
            for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.depth = depth
        if not evaluate:
            nn.init.trunc_normal_(self.pos_embed, std=.02)
            nn.init.trunc_normal_(self.cls_token, std=.02)
            self.apply(self._init_weights)
        self.img_query_model = Query_model(ft_dim=embed_dim, sd_dim=sd_dim, temperature=1, att_func_type='sparsemax',
                                           pool_type='max', map_func=map_func)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    @torch.no_grad()
    def forward_device(self, x, space_dict, temperature, pack_groups=0, keep_f32=False):
        """The pruned forward with DEVICE-RESIDENT token counts: no host read-back anywhere (the reference syncs once
        per layer, models/vit.py:145), so the whole pass can be enqueued ahead or captured in a CUDA graph. Every
        buffer is sized for the unpruned token count and holds B packed sequences of the current, device-side length.
        Returns a DeviceEncoded record; pack_groups > 0 writes the final LayerNorm straight into the fp16 operand
        layout of the cross-attention K / V^T projections (madtp_layernorm_pack), split into that many equal groups of
        sequences (BLIP-NLVR: 2 = image0 / image1)."""
        Fn.require_cuda(x, "image")
        _eval_only(self)
        if not (temperature > 0) or space_dict is None:
            raise RuntimeError("madtp_b200: forward_device is the PRUNED path (temperature > 0 with a codebook)")
        B = x.shape[0]
        patches = self.patch_embed(x.contiguous())
        n, C = patches.shape[1], patches.shape[2]
        if n + 1 > self.pos_embed.shape[1]:
            raise RuntimeError("madtp_b200: image has more patches than pos_embed rows")
        x = L.assemble_tokens(patches, self.cls_token.detach().reshape(-1), self.pos_embed.detach().reshape(-1, C),
                              B, n, C)
        N = n + 1
        depth = len(self.blocks)
        lens = L.empty((2 * depth + 2,), torch.int32, x.device)     # dims [depth + 1] | ks [depth] | P
        dims, ks = lens[:depth + 1], lens[depth + 1:2 * depth + 1]
        dims[:1].fill_(N)
        ks.fill_(-1)
        traj = Fn.Trajectory(dims, ks)
        sd_img_ft_all = None
        for i, blk in enumerate(self.blocks):
            n_dev, n_out, k_out = dims[i:i + 1], dims[i + 1:i + 2], ks[i:i + 1]
            ln1 = Fn.layernorm_rows(x.view(B * N, C), blk.norm1.weight, blk.norm1.bias, blk.norm1.eps, split=True,
                                    split_x=True, n_dev=n_dev, n_mult=B)
            token_attn, sd_img_ft_all = self.img_query_model.forward_rows(x, ln1["x_hi"], ln1["x_lo"], space_dict,
                                                                          sd_img_ft_all, n_dev=n_dev)
            x = blk.forward_rows(x, ln1, temperature, token_attn, n_dev=n_dev, n_out=n_out, k_out=k_out)
            blk.last_prune = Fn.LazyPrune(traj, i, blk.last_prune, B)
        n_dev = dims[depth:depth + 1]
        enc = DeviceEncoded(B, N, C, n_dev, traj, sd_img_ft_all)
        if pack_groups:
            if B % pack_groups:
                raise RuntimeError("madtp_b200: the batch does not split into equal groups")
            P = (N + 7) // 8 * 8
            per = B // pack_groups
            enc.y16 = L.empty((pack_groups, per * P, C), torch.float16, x.device)
            enc.p_dev = lens[2 * depth + 1:2 * depth + 2]
            enc.per_group, enc.P = per, P
            if keep_f32:
                enc.y = L.empty((B * N, C), torch.float32, x.device)
            L.layernorm_pack(x.view(B * N, C), B, N, self.norm.weight, self.norm.bias, self.norm.eps, enc.y16, per,
                             per * P * C, y_f32=enc.y, p_out=enc.p_dev, n_dev=n_dev)
        else:
            enc.y = Fn.layernorm_rows(x.view(B * N, C), self.norm.weight, self.norm.bias, self.norm.eps, f32=True,
                                      n_dev=n_dev, n_mult=B)["y"]
        return enc

    @torch.no_grad()
    def forward(self, x, register_blk=-1, space_dict=None, temperature=0):
        Fn.require_cuda(x, "image")
        _eval_only(self)
        if register_blk >= 0:
            raise NotImplementedError("madtp_b200: register_blk needs the materialised attention map")
        if temperature > 0 and space_dict is not None and device_lengths_enabled():
            # device-resident lengths: ONE host read-back (the final token count, to shape the returned tensor)
            # instead of one per layer
            with L.arena_for(self, tuple(x.shape)):
                enc = self.forward_device(x, space_dict, temperature)
                return enc.narrowed().clone(), enc.sd_ft.clone()      # the arena's buffers are reused by the next call
        B = x.shape[0]
        patches = self.patch_embed(x.contiguous())
        n, C = patches.shape[1], patches.shape[2]
        if n + 1 > self.pos_embed.shape[1]:
            raise RuntimeError("madtp_b200: image has more patches than pos_embed rows")
        x = L.assemble_tokens(patches, self.cls_token.detach().reshape(-1), self.pos_embed.detach().reshape(-1, C),
                              B, n, C)
        sd_img_ft_all = None
        for blk in self.blocks:
            N = x.shape[1]
            with_dict = space_dict is not None
            ln1 = Fn.layernorm_rows(x.view(B * N, C), blk.norm1.weight, blk.norm1.bias, blk.norm1.eps, split=True,
                                    split_x=with_dict)
            if with_dict:
                token_attn, sd_img_ft_all = self.img_query_model.forward_rows(x, ln1["x_hi"], ln1["x_lo"], space_dict,
                                                                              sd_img_ft_all)
                x = blk.forward_rows(x, ln1, temperature, token_attn)
            else:
                x = blk.forward_rows(x, ln1)
        N = x.shape[1]
        x = Fn.layernorm_rows(x.view(B * N, C), self.norm.weight, self.norm.bias, self.norm.eps, f32=True)["y"]
        return x.view(B, N, C), sd_img_ft_all


class DeviceEncoded:
    """Result of an encoder pass with device-resident lengths: capacity-sized buffers of B packed sequences, the device
    scalar holding the final tokens-per-sequence, and the trajectory."""

    def __init__(self, B, cap, C, n_dev, traj, sd_ft):
        self.B, self.cap, self.C, self.n_dev, self.traj, self.sd_ft = B, cap, C, n_dev, traj, sd_ft
        self.y = None          # packed fp32 final states [B * cap, C]
        self.y16 = None        # or: fp16 cross-attention operand layout [groups, per_group * P, C] (layernorm_pack)
        self.p_dev = None      # device scalar P = final length rounded up to 8
        self.per_group = self.P = 0

    def narrowed(self):
        """[B, N_final, C] view of the packed final states (one device -> host read of the trajectory)."""
        n = self.traj.host()[0][-1]
        return self.y.reshape(-1)[:self.B * n * self.C].view(self.B, n, self.C)


_DEVICE_LENGTHS = [os.environ.get("MADTP_HOST_LENGTHS", "0") != "1"]


def device_lengths_enabled(enable=None) -> bool:
    """Module-level switch: encoders keep the pruned token counts on the device (default) or read them back once per
    layer like the reference does (MADTP_HOST_LENGTHS=1 / device_lengths_enabled(False): the round-1 path, kept as a
    cross-check -- both produce bit-identical results)."""
    if enable is not None:
        _DEVICE_LENGTHS[0] = bool(enable)
    return _DEVICE_LENGTHS[0]


def interpolate_pos_embed(pos_embed_checkpoint, visual_encoder):
    """reference models/vit.py:398-422 (checkpoint I/O helper; bicubic resize of the patch position grid)."""
    embedding_size = pos_embed_checkpoint.shape[-1]
    num_patches = visual_encoder.patch_embed.num_patches
    num_extra_tokens = visual_encoder.pos_embed.shape[-2] - num_patches
    orig_size = int((pos_embed_checkpoint.shape[-2] - num_extra_tokens) ** 0.5)
    new_size = int(num_patches ** 0.5)
    if orig_size == new_size:
        return pos_embed_checkpoint
    extra_tokens = pos_embed_checkpoint[:, :num_extra_tokens]
    pos_tokens = pos_embed_checkpoint[:, num_extra_tokens:]
    pos_tokens = pos_tokens.reshape(-1, orig_size, orig_size, embedding_size).permute(0, 3, 1, 2)
    pos_tokens = torch.nn.functional.interpolate(pos_tokens, size=(new_size, new_size), mode='bicubic',
                                                 align_corners=False)
    pos_tokens = pos_tokens.permute(0, 2, 3, 1).flatten(1, 2)
    return torch.cat((extra_tokens, pos_tokens), dim=1)
