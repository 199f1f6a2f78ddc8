"""Mirror of the hot-path part of the reference's models/utils.py: `vector_gather` (:13-33) and `Query_model`
(:109-183) -- same constructor arguments, forward signature, return arity and (absence of) parameters, backed by the
sm_100a kernels. Sparsemax and the contrastive losses of that file are dead code on the pruned forward path
 and are not mirrored.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import _lib as L
from . import functional as Fn


def vector_gather(vectors: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """out[b,i,:] = vectors[b, indices[b,i], :]   (reference models/utils.py:13-33; [B,L,D],[B,K] -> [B,K,D]).

    Runs the DTP gather kernel with an identity selection: slot i of the output copies source row indices[b,i]."""
    Fn.require_cuda(vectors, "vectors")
    if vectors.dim() != 3 or indices.dim() != 2 or indices.shape[0] != vectors.shape[0]:
        raise RuntimeError("vector_gather: expected vectors [B,L,D] and indices [B,K]")
    B, Ltok, D = vectors.shape
    K = indices.shape[1]
    if K == 0:
        return vectors.new_empty(B, 0, D)
    if int(indices.min()) < 0 or int(indices.max()) >= Ltok:
        raise IndexError("vector_gather: index out of range")
    return L.gather_rows(vectors.contiguous(), indices.to(torch.int32).contiguous())


class Query_model(nn.Module):
    """The DTP alignment head: token_att = ft . sd^T (returned raw), att_ft = softmax_tokens(token_att/sqrt(sd_dim)) ft.

    `temperature`, `att_func_type`, `pool_type` and the forward's `mask`/`temperature` are accepted and ignored exactly
    as in the reference (models/utils.py:147-183 hard-codes softmax and never reads them)."""

    def __init__(self, ft_dim, sd_dim, temperature=1, att_func_type='softmax', pool_type='sum', map_func=False):
        super().__init__()
        assert att_func_type in ['softmax', 'sigmoid', 'sparsemax']
        assert pool_type in ['mean', 'max', 'sum']
        self.att_func_type = att_func_type
        self.pool_type = pool_type
        self.att_dim = sd_dim
        self.temperature = temperature
        self.map_func = map_func
        if self.map_func:
            self.q_map = nn.Sequential(nn.Linear(ft_dim, sd_dim))
        self._cache = Fn.WeightCache()

    # -- prepared operands ------------------------------------------------------------------------------------
    def _book(self, sd: torch.Tensor):
        return self._cache.get("book", [sd], lambda: Fn.prepare_codebook(sd))

    def _qmap(self):
        lin = self.q_map[0]
        return self._cache.get("q_map", [lin.weight, lin.bias],
                               lambda: Fn.PreparedLinear(lin.weight, lin.bias, split=True))

    def forward_rows(self, x3d, x_hi, x_lo, sd, sd_ft_acc=None, first_token=1, n_dev=None):
        """Fast path used by the encoders: x3d [B,N,d] with its fp16 hi/lo split already produced by the LayerNorm kernel;
        tokens first_token.. are the prunable ones. Returns (token_att view [B,n,T], sd_ft (accumulated))."""
        if self.map_func:
            B, N, d = x3d.shape
            q = Fn.linear_split(x_hi, x_lo, self._qmap(), m_dev=n_dev, m_mult=B)
            q_hi, q_lo = Fn.split_rows(q, n_dev=n_dev, n_mult=B)
            return Fn.query_model_rows(q_hi, q_lo, q.view(B, N, -1), self._book(sd), self.att_dim, sd_ft_acc,
                                       first_token, n_dev=n_dev)
        return Fn.query_model_rows(x_hi, x_lo, x3d, self._book(sd), self.att_dim, sd_ft_acc, first_token, n_dev=n_dev)

    def forward(self, ft, sd, mask=None, return_token_att=False, temperature=1):
        """ft [B, n, ft_dim], sd [T, sd_dim] -> (token_att [B,n,T] | att_weight [B,T,n], att_ft [B,T,sd_dim], sd)."""
        Fn.require_cuda(ft, "ft")
        Fn.require_cuda(sd, "sd")
        B, n, d = ft.shape
        x = ft.contiguous()
        x_hi, x_lo = Fn.split_rows(x.view(B * n, d))
        token_att, att_ft = self.forward_rows(x, x_hi, x_lo, sd, None, first_token=0)
        if return_token_att:
            return token_att, att_ft, sd
        # Not on the pruned forward path (every encoder passes return_token_att=True): plain softmax over tokens.
        att_weight = torch.softmax((token_att / math.sqrt(self.att_dim)).permute(0, 2, 1), dim=-1)
        return att_weight, att_ft, sd
