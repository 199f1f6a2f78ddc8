"""Compositions of the C-ABI kernels (madtp_b200/_lib.py) that the module mirrors in vit.py / nlvr_encoder.py /
med.py / utils.py share. Everything here runs on the GPU through libmadtp_b200.so; there is no CPU path.

Precision lanes (hard part 1):
  scoring lane  LayerNorm -> q/k/v projection -> attention statistics, and token . codebook^T: fp32-accurate
                (error-compensated fp16 hi/lo tensor-core GEMMs and attention with chunk-drained fp32 accumulation;
                fp32 CUDA-core attention for short text sequences)
  value lane    attention output projection, FFN, cross-attention: fp16 operands, fp32 accumulation
"""
from __future__ import annotations

import math
import os
import weakref
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib as L
from . import dist as _dist

Tensor = torch.Tensor
# MADTP_SDFT_PLANES=0 keeps the round-1 aggregation kernel (fp32 x re-laid out K-major by builder warps)
SDFT_PLANES = os.environ.get("MADTP_SDFT_PLANES", "1") != "0"
# Opt-in diagnostic: run the ViT's VALUE lane (attention output projection, FFN) at the scoring lane's precision too --
# fp32 attention context, error-compensated fp16 hi/lo planes for every operand, exact erf GELU -- to show what the
# free-running keep-masks do without the fp16 rounding of the value lane (MADTP_VALUE_LANE=split, bench.py --value-lane).
def value_lane_split(enable=None) -> bool:
    if enable is not None:
        L._value_split[0] = bool(enable)
    return L._value_split[0]


TA_LD = 128  # row pitch of token_att buffers (T = 100 codebook entries padded to a 16-byte multiple of columns)


def require_cuda(t: Tensor, name: str):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"madtp_b200: {name} must be a CUDA tensor -- this package has no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"madtp_b200: {name} must be float32 (the reference's evaluation dtype), got {t.dtype}")


# ------------------------------------------------------------------------------------------------------------------
# prepared weights
# ------------------------------------------------------------------------------------------------------------------
def pow2_scale(w: Tensor, target: float = 16384.0) -> float:
    """Power of two s with max|s * w| <= target: puts a weight's magnitude where both fp16 planes of its hi/lo split
    are normal numbers. One host sync, at weight-preparation time only."""
    m = float(w.detach().abs().max()) if w.numel() else 0.0
    if not (m > 0.0) or not math.isfinite(m):
        return 1.0
    return 2.0 ** max(-24, min(24, math.floor(math.log2(target / m))))


class PreparedLinear:
    """GEMM-ready copies of an nn.Linear: fp16 hi/lo split of scale * W (scoring lane; `split=True` for historical
    reasons selects this error-compensated lane) and/or plain fp16 (value lane)."""
    __slots__ = ("hi", "lo", "scale", "w16", "w32", "bias", "out_features", "in_features")

    def __init__(self, weight: Tensor, bias: Optional[Tensor], split: bool = False, f16: bool = False,
                 f32: bool = False, bias_scale: float = 1.0):
        w = weight.detach().to(torch.float32).contiguous()
        self.out_features, self.in_features = w.shape
        self.hi = self.lo = self.w16 = self.w32 = None
        self.scale = 1.0
        if split:
            self.scale = pow2_scale(w)
            self.hi, self.lo = L.split_f16(w, self.scale)
        if f16:
            self.w16 = L.cast_f16(w)
        if f32:
            self.w32 = w
        self.bias = None if bias is None else (bias.detach().to(torch.float32) * bias_scale).contiguous()


class WeightCache:
    """Per-module cache of PreparedLinear objects, invalidated when any source tensor is modified in place (its
    autograd version counter moves), replaced by another tensor object (checked through weak references, so a
    temporary whose allocation is recycled by a later tensor of the same shape can never hit a stale entry) or
    re-allocated.

    One limitation is inherent to torch: writes through `param.data` (which carries its OWN version counter) are
    invisible here. Code that updates weights that way after a first forward -- EMA, manual reloads -- must call
    `clear_caches(module)` afterwards; `load_state_dict` and `madtp_b200.dist.broadcast_parameters` copy through
    `param.detach()` / `param.copy_`, which share the counter, and need nothing."""

    def __init__(self):
        self._store = {}

    @staticmethod
    def _stamp(params):
        return tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in params if p is not None)

    def get(self, key, params, build):
        stamp = self._stamp(params)
        hit = self._store.get(key)
        if hit is not None and hit[0] == stamp:
            live = [p for p in params if p is not None]
            if len(live) == len(hit[2]) and all(r() is p for r, p in zip(hit[2], live)):
                return hit[1]
        val = build()
        self._store[key] = (stamp, val, [weakref.ref(p) for p in params if p is not None])
        return val

    def clear(self):
        self._store.clear()


def clear_caches(module) -> int:
    """Drops every prepared-weight cache below `module` (call after changing weights through `.data`). Returns the
    number of caches cleared."""
    n = 0
    for m in module.modules():
        for v in vars(m).values():
            if isinstance(v, WeightCache):
                v.clear()
                n += 1
    return n


# ------------------------------------------------------------------------------------------------------------------
# row operations
# ------------------------------------------------------------------------------------------------------------------
def layernorm_rows(x2d: Tensor, gamma: Optional[Tensor], beta: Optional[Tensor], eps: float, *, f32=False, split=False,
                   f16=False, split_x=False, n_dev: Optional[Tensor] = None, n_mult: int = 1):
    """LayerNorm over the rows of x2d [rows, d] with the requested operand copies.
    Returns dict with any of y, y_hi, y_lo, y16, x_hi, x_lo. n_dev / n_mult: device-resident row count
    rows = *n_dev * n_mult (x2d.shape[0] is then the capacity)."""
    rows, d = x2d.shape
    dev = x2d.device
    out = {}

    def new(dt=torch.float32):
        return L.empty((rows, d), dt, dev)
    if f32:
        out["y"] = new()
    if split:
        out["y_hi"], out["y_lo"] = new(torch.float16), new(torch.float16)
    if f16:
        out["y16"] = new(torch.float16)
    if split_x:
        out["x_hi"], out["x_lo"] = new(torch.float16), new(torch.float16)
    L.layernorm(x2d, gamma, beta, eps, y_f32=out.get("y"), y_hi=out.get("y_hi"), y_lo=out.get("y_lo"),
                y_f16=out.get("y16"), x_hi=out.get("x_hi"), x_lo=out.get("x_lo"), n_dev=n_dev, n_mult=n_mult)
    return out


def split_rows(x2d: Tensor, n_dev: Optional[Tensor] = None, n_mult: int = 1) -> Tuple[Tensor, Tensor]:
    """fp16 hi/lo split of fp32 rows (the operand format of the F16x3 GEMM)."""
    o = layernorm_rows(x2d, None, None, 0.0, split_x=True, n_dev=n_dev, n_mult=n_mult)
    return o["x_hi"], o["x_lo"]


def linear_split(a_hi: Tensor, a_lo: Tensor, lin: PreparedLinear, out: Optional[Tensor] = None, *, residual=None,
                act=L.ACT_NONE, alpha=1.0, m_dev: Optional[Tensor] = None, m_mult: int = 1) -> Tensor:
    if out is None:
        out = L.empty((a_hi.shape[0], lin.out_features), torch.float32, a_hi.device)
    return L.gemm(L.GEMM_F16X3, a_hi, lin.hi, out, a_lo=a_lo, b_lo=lin.lo, bias=lin.bias, residual=residual, act=act,
                  alpha=alpha / lin.scale, m_dev=m_dev, m_mult=m_mult)


def linear_f16(a16: Tensor, lin: PreparedLinear, out: Optional[Tensor] = None, *, out_dtype=torch.float32,
               residual=None, act=L.ACT_NONE, alpha=1.0, m_dev: Optional[Tensor] = None, m_mult: int = 1) -> Tensor:
    if out is None:
        out = L.empty((a16.shape[0], lin.out_features), out_dtype, a16.device)
    return L.gemm(L.GEMM_F16, a16, lin.w16, out, bias=lin.bias, residual=residual, act=act, alpha=alpha, m_dev=m_dev,
                  m_mult=m_mult)


def linear_f32(a: Tensor, lin: PreparedLinear, *, act=L.ACT_NONE) -> Tensor:
    """fp32 CUDA-core GEMM for tiny heads (cls_head)."""
    out = L.empty((a.shape[0], lin.out_features), torch.float32, a.device)
    return L.gemm(L.GEMM_SIMT, a, lin.w32, out, bias=lin.bias, act=act)


# ------------------------------------------------------------------------------------------------------------------
# Query_model  (reference models/utils.py:147-183)
# ------------------------------------------------------------------------------------------------------------------
def prepare_codebook(space_dict: Tensor):
    """space_dict [T, sd_dim] -> (fp16 hi, lo planes of scale * book, T, scale), rows padded to TA_LD so token_att rows
    are 16-byte aligned."""
    T, d = space_dict.shape
    if T > TA_LD:
        raise RuntimeError(f"madtp_b200: codebook size {T} exceeds the supported {TA_LD}")
    pad = torch.zeros(TA_LD, d, dtype=torch.float32, device=space_dict.device)
    pad[:T] = space_dict.detach()
    scale = pow2_scale(pad)
    hi, lo = L.split_f16(pad, scale)
    return hi, lo, T, scale


def query_model_rows(x_hi: Tensor, x_lo: Tensor, x3d: Tensor, book, sd_dim: int, sd_ft: Optional[Tensor],
                     first_token: int = 1, n_dev: Optional[Tensor] = None):
    """token_att for EVERY row of x3d [B,N,d] (one GEMM over the flat rows), then the over-token softmax
    aggregation over tokens first_token..N-1.  Returns (token_att view [B, N-first_token, T], sd_ft [B,T,d]).
    n_dev: device-resident N (x3d is then a capacity-sized buffer holding B packed sequences)."""
    B, N, d = x3d.shape
    hi, lo, T, scale = book
    ta = L.empty((B * N, TA_LD), torch.float32, x3d.device)
    L.gemm(L.GEMM_F16X3, x_hi, hi, ta, a_lo=x_lo, b_lo=lo, alpha=1.0 / scale, m_dev=n_dev, m_mult=B)
    return query_model_from_token_att(ta.view(B, N, TA_LD), x3d, T, sd_dim, sd_ft, first_token, n_dev=n_dev,
                                      planes=(x_hi, x_lo))


def query_model_from_token_att(ta_full: Tensor, x3d: Tensor, T: int, sd_dim: int, sd_ft: Optional[Tensor],
                               first_token: int = 1, n_dev: Optional[Tensor] = None, planes=None):
    """Second half of Query_model given token_att for every row (ta_full [B, N, >=T] view, unit inner stride):
    over-token softmax statistics and the aggregated feature. Returns (token_att view [B, n, T], sd_ft)."""
    B, N, d = x3d.shape
    ta3 = ta_full[:, first_token:, :]
    n = N - first_token
    div = math.sqrt(sd_dim)
    cm, cs = L.token_colstats(ta3, n, T, div, n_dev=n_dev, n_sub=first_token)
    accumulate = sd_ft is not None
    if sd_ft is None:
        sd_ft = L.empty((B, T, d), torch.float32, x3d.device)
    if n >= 64 and planes is not None and d % 64 == 0 and SDFT_PLANES:
        # tensor-core path over the fp16 hi/lo planes the LayerNorm kernel wrote (MN-major operands, no transposition)
        L.query_sdft_planes(ta3, cm, cs, planes[0], planes[1], N, first_token, n, T, div, sd_ft, accumulate, n_dev=n_dev)
    elif n >= 64 and x3d.is_contiguous() and d % 32 == 0:     # tensor-core path over the dense fp32 rows of x3d
        L.query_sdft_tc(ta3, cm, cs, x3d.view(B * N, d), N, first_token, n, T, div, sd_ft, accumulate, n_dev=n_dev)
    else:
        L.query_sdft(ta3, cm, cs, x3d[:, first_token:, :], n, T, div, sd_ft, accumulate, n_dev=n_dev,
                     n_sub=first_token)
    return ta3[:, :, :T], sd_ft


# ------------------------------------------------------------------------------------------------------------------
# attention + pruning statistics
# ------------------------------------------------------------------------------------------------------------------
@dataclass
class AttnStats:
    """What Reduce_token needs from the self-attention instead of the materialised [B,H,N,N] map
    (reference models/vit.py:83,101): per-query-tile partial column sums of max_h P, and cls_attn."""
    col_part: Tensor   # [B, n_parts, N]
    cls_attn: Tensor   # [B, N]  (entry 0 unused)
    parts_tile: int = 0   # query-tile height of the producer (n_parts = ceil(N / parts_tile)); 0: one part


def self_attention(q: Tensor, k: Tensor, v: Tensor, H: int, scale: float, key_mask: Optional[Tensor],
                   want_stats: bool, ctx16: Optional[Tensor] = None, causal: bool = False,
                   l_dev: Optional[Tensor] = None):
    """q,k,v: [B,N,H*64] fp32 views. Returns (ctx16 [B,N,H*64] fp16, AttnStats or None).
    l_dev: device-resident N (short-sequence path only: packed q / k / v / mask / outputs)."""
    B, N, C = q.shape
    dev = q.device
    if ctx16 is None:
        ctx16 = L.empty((B, N, C), torch.float16, dev)
    if N <= 64:     # short text sequences: per-(head, sequence) CTAs + one statistics pass
        if not want_stats:
            L.attn_small_self(q, k, v, H, scale, ctx16, key_mask=key_mask, causal=causal, l_dev=l_dev)
            return ctx16, None
        col_part = L.empty((B, 1, N), torch.float32, dev)
        cls_attn = L.empty((B, N), torch.float32, dev)
        L.attn_small_self(q, k, v, H, scale, ctx16, key_mask=key_mask, col_sum=col_part, cls_attn=cls_attn,
                          causal=causal, l_dev=l_dev)
        return ctx16, AttnStats(col_part, cls_attn, 0)
    if l_dev is not None:
        raise RuntimeError("madtp_b200: device-resident lengths need the short-sequence (N <= 64) or the tensor-core "
                           "self-attention path")
    if not want_stats:
        L.attn_fwd(q, k, v, H, scale, ctx16, key_mask=key_mask, causal=causal)
        return ctx16, None
    rows = torch.empty(3, B, H, N, dtype=torch.float32, device=dev)
    stats = (rows[0], rows[1], rows[2])
    L.attn_fwd(q, k, v, H, scale, ctx16, key_mask=key_mask, stats=stats, causal=causal)
    n_parts = (N + 63) // 64
    col_part = torch.empty(B, n_parts, N, dtype=torch.float32, device=dev)
    cls_attn = torch.empty(B, N, dtype=torch.float32, device=dev)
    L.attn_stats(q, k, H, scale, stats, col_part, cls_attn, key_mask=key_mask, causal=causal)
    return ctx16, AttnStats(col_part, cls_attn, 64)


def self_attention_tc(y_hi: Tensor, y_lo: Tensor, qkv: PreparedLinear, B: int, N: int, H: int, scale: float,
                      want_stats: bool, n_dev: Optional[Tensor] = None, causal: bool = False,
                      ctx32: Optional[Tensor] = None):
    """Tensor-core scoring-lane self-attention from the fp16 hi/lo planes of the normalised rows [B*N, C]:
    fused q|k|v projection (split / transposed epilogue) -> attention -> (optionally) pruning statistics.
    Returns (ctx16 [B,N,H*64] fp16, AttnStats or None). n_dev: device-resident N (N is then the capacity)."""
    dev = y_hi.device
    qk_hi, qk_lo, vt_hi, vt_lo = L.gemm_qkv(y_hi, y_lo, qkv.hi, qkv.lo, qkv.bias, N, H, alpha=1.0 / qkv.scale,
                                            n_dev=n_dev)
    ctx16 = L.empty((B, N, H * 64), torch.float16, dev)
    rows = L.empty((3 if want_stats else 2, B, H, N), torch.float32, dev)
    if not want_stats:
        L.attn_tc_fwd(qk_hi, qk_lo, vt_hi, vt_lo, B, H, N, scale, ctx16, rows[0], rows[1], n_dev=n_dev, causal=causal,
                      out_f32=ctx32)
        return ctx16, None
    cls_tile_max = L.empty((B, H, (N + 63) // 64), torch.float32, dev)
    L.attn_tc_fwd(qk_hi, qk_lo, vt_hi, vt_lo, B, H, N, scale, ctx16, rows[0], rows[1], cls_p=rows[2],
                  cls_tile_max=cls_tile_max, n_dev=n_dev, causal=causal, out_f32=ctx32)
    n_parts = (N + 127) // 128
    col_part = L.empty((B, n_parts, N), torch.float32, dev)
    cls_attn = L.empty((B, N), torch.float32, dev)
    L.attn_tc_stats(qk_hi, qk_lo, B, H, N, scale, rows[0], rows[1], col_part, cls_attn, rows[2], cls_tile_max,
                    n_dev=n_dev, causal=causal)
    return ctx16, AttnStats(col_part, cls_attn, 128)


@dataclass
class PruneResult:
    x: Tensor                       # [B, k+2, d] (or the input when nothing was pruned)
    pruned: bool
    k: int
    score: Tensor                   # [B, n]
    threshold: Tensor               # [B]
    count: Tensor                   # [B] int32
    keep: Optional[Tensor] = None   # [B, n] uint8
    mask: Optional[Tensor] = None   # [B, k+2] additive mask (text)
    x16: Optional[Tensor] = None    # fp16 copy of x when the caller asked for it (dtp_finish(want_f16=True))
    ln16: Optional[Tensor] = None   # fp16 LayerNorm(x) when the caller passed ln=(gamma, beta, eps)


class PendingPrune:
    """Score kernel launched, topk_num on its way to the host (dtp_score_async); dtp_finish completes the pruning."""
    __slots__ = ("score", "thr", "cnt", "topk", "handle")

    def __init__(self, score, thr, cnt, topk, handle):
        self.score, self.thr, self.cnt, self.topk, self.handle = score, thr, cnt, topk, handle


def dtp_score_async(stats: AttnStats, token_att: Tensor, temperature: float, n: int,
                    n_dev: Optional[Tensor] = None) -> PendingPrune:
    """First half of Reduce_token: importance score, threshold, survivor counts and their batch maximum (reference
    models/vit.py:126-145), plus an asynchronous read-back of that maximum. Launch work that does not depend on the
    pruning decision between this and dtp_finish: it runs while the host waits for topk_num.
    n_dev: device-resident N = n + 1 -- nothing is read back, dtp_finish consumes topk_num on the device."""
    T = token_att.shape[2]
    score, thr, cnt, topk = L.dtp_score(stats.col_part, stats.cls_attn, token_att, n, T, temperature, n_dev=n_dev,
                                        parts_tile=stats.parts_tile)
    _dist.allreduce_topk_(topk)         # no-op unless madtp_b200.dist.global_topk(True) (strict multi-GPU mode)
    return PendingPrune(score, thr, cnt, topk, None if n_dev is not None else L.readback_begin(topk))


def dtp_finish(x: Tensor, pend: PendingPrune, *, mask_mode: int = 0, mask_in: Optional[Tensor] = None,
               max_keep: int = 0, want_f16: bool = False, n_dev: Optional[Tensor] = None,
               n_out: Optional[Tensor] = None, k_out: Optional[Tensor] = None, ln=None) -> PruneResult:
    """Second half of Reduce_token on x [B, n+1, d] (position 0 always survives): select, gather, merge.
    n_dev / n_out / k_out (device-resident lengths): x is a capacity-sized buffer of packed sequences; the kernels read
    topk_num and N on the device, write the next layer's N to n_out and the trajectory entry to k_out, and the result
    is again capacity-sized (`pruned` and `k` of the returned record are unknown to the host: None / -1)."""
    B, N, d = x.shape
    n = N - 1
    score, thr, cnt, topk = pend.score, pend.thr, pend.cnt, pend.topk
    # select + gather + merged token (+ the LayerNorm that follows, ln = (gamma, beta, eps)) are ONE kernel: dtp_apply.cu
    if n_dev is not None:
        out, out16, ln16, keep, mask_out = L.dtp_apply(x, score, topk, n - 1, mask_mode=mask_mode, mask_in=mask_in,
                                                       max_keep=max_keep, want_f16=want_f16, ln=ln, n_dev=n_dev,
                                                       n_out=n_out, k_out=k_out)
        return PruneResult(out, None, -1, score, thr, cnt, keep, mask_out, out16, ln16)
    k = L.readback_wait(pend.handle)        # the reference's one host sync per pruned layer (models/vit.py:145)
    if k <= max_keep or n - k <= 1:         # models/vit.py:148-149 (max_keep = 0); clip/model.py:220
        return PruneResult(x, False, k, score, thr, cnt, None, mask_in)
    out, out16, ln16, keep, mask_out = L.dtp_apply(x, score, topk, k, mask_mode=mask_mode, mask_in=mask_in,
                                                   max_keep=max_keep, want_f16=want_f16, ln=ln)
    return PruneResult(out, True, k, score, thr, cnt, keep, None if mask_out is None else mask_out[:, :k + 2], out16,
                       ln16)


class Trajectory:
    """Device-resident lengths of one encoder pass: dims[i] = tokens per sequence (incl. CLS / [ENC]) entering layer i
    (dims[depth] = leaving the last layer), ks[i] = topk_num of layer i or -1 when it did not prune. `host()` copies
    both to the host ONCE (the only device -> host transfer of such a pass) and caches them."""

    def __init__(self, dims: Tensor, ks: Tensor):
        self.dims, self.ks = dims, ks
        self._host = None

    def invalidate(self):
        self._host = None

    def host(self):
        if self._host is None:
            v = torch.cat([self.dims, self.ks]).cpu().tolist()
            nd = self.dims.numel()
            self._host = (v[:nd], v[nd:])
        return self._host


class LazyPrune:
    """PruneResult of one layer of a device-resident-length pass. The buffers are capacity-sized and packed with the
    dynamic lengths; attribute access narrows them with the trajectory read from the device on first use."""

    def __init__(self, traj: Trajectory, layer: int, res: PruneResult, B: int):
        self._traj, self._layer, self._res, self._B = traj, layer, res, B

    def _packed(self, t, per_seq, tail=()):
        return t.reshape(-1)[:self._B * per_seq * int(math.prod(tail))].view(self._B, per_seq, *tail)

    @property
    def k(self):
        return self._traj.host()[1][self._layer]

    @property
    def pruned(self):
        return self.k >= 0

    @property
    def n_in(self):
        return self._traj.host()[0][self._layer] - 1

    @property
    def n_out_tokens(self):
        return self._traj.host()[0][self._layer + 1]

    @property
    def score(self):
        return self._packed(self._res.score, self.n_in)

    @property
    def keep(self):
        return self._packed(self._res.keep, self.n_in) if self.pruned else None

    @property
    def threshold(self):
        return self._res.threshold

    @property
    def count(self):
        return self._res.count

    @property
    def x(self):
        return self._packed(self._res.x, self.n_out_tokens, (self._res.x.shape[-1],))

    @property
    def mask(self):
        return None if self._res.mask is None else self._packed(self._res.mask, self.n_out_tokens)


def dtp_prune(x: Tensor, stats: AttnStats, token_att: Tensor, temperature: float, *, mask_mode: int = 0,
              mask_in: Optional[Tensor] = None, max_keep: int = 0, ln=None) -> PruneResult:
    """Reduce_token on x [B, n+1, d] (position 0 always survives). reference models/vit.py:123-163,
    models/nlvr_encoder.py:400-454 (mask_mode 1), models/med.py:345-391 (mask_mode 2)."""
    pend = dtp_score_async(stats, token_att, temperature, x.shape[1] - 1)
    return dtp_finish(x, pend, mask_mode=mask_mode, mask_in=mask_in, max_keep=max_keep, ln=ln)
