"""Mirror of the reference's models/nlvr_encoder.py hot path -- the BERT text encoder with twin cross-attention used
by BLIP-NLVR: BertEmbeddings (:43-85), BertSelfAttention (:88-237), BertSelfOutput (:240-271), BertAttention
(:274-349), BertIntermediate/BertOutput (:352-385), BertLayer (:388-559), BertEncoder (:562-687), BertModel
(:760-1015). Same constructor arguments, forward signatures (including nlvr_encoder's positional order, which differs
from med.py), return arity and state-dict keys; executed by the sm_100a kernels of libmadtp_b200.so.

Contract differences: evaluation forward only; `past_key_value`, `head_mask`, `output_attentions`,
relative position embeddings are not on the pruned encoder path and raise; the (unpruned) decoder mode is `is_decoder=True`;
attention maps are AttnStats handles; survivors keep ascending token order; the key/value cache slot of the returned
tuples is None.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
from torch import nn

from . import _lib as L
from . import functional as Fn
from .utils import Query_model, vector_gather  # noqa: F401
from .vit import _eval_only


class EncoderOutput:
    """Stand-in for transformers' BaseModelOutputWithPoolingAndCrossAttentions: attribute and [0] access."""

    def __init__(self, last_hidden_state, pooler_output=None):
        self.last_hidden_state = last_hidden_state
        self.pooler_output = pooler_output
        self.past_key_values = None
        self.hidden_states = None
        self.attentions = None
        self.cross_attentions = None

    def __getitem__(self, i):
        return (self.last_hidden_state, self.pooler_output)[i]


def _unsupported(**kw):
    for name, val in kw.items():
        if val is not None and val is not False:
            raise NotImplementedError(f"madtp_b200: `{name}` is not part of the pruned encoder forward path")


def _key_mask(ext_mask: Optional[torch.Tensor], B: int, N: int) -> Optional[torch.Tensor]:
    """additive extended mask [B,1,1,N] (or [B,N]) -> contiguous [B,N] fp32."""
    if ext_mask is None:
        return None
    m = ext_mask.reshape(B, -1)
    if m.shape[1] != N:
        raise RuntimeError("madtp_b200: attention mask must broadcast as [B,1,1,N] over keys")
    return m.to(torch.float32).contiguous()


class CrossKV:
    """Operands of the tensor-core cross-attention for one layer and branch: k16 [B,Nk,C] (or [Nk,C] when broadcast)
    fp16 view, vt16 [C, >= B*P] fp16 (V^T, keys contiguous, P = keys_per_batch columns per sequence, 0 = broadcast),
    v_bias [C] fp32. `ragged` = (k_start, k_len, key0_bias, max_len): sequences of DIFFERENT key counts packed back to
    back in k16 [rows, C] / vt16 [C, rows] (see RaggedImageFeatures)."""
    __slots__ = ("k16", "vt16", "v_bias", "keys_per_batch", "ragged")

    def __init__(self, k16, vt16, v_bias, keys_per_batch, ragged=None):
        self.k16, self.vt16, self.v_bias, self.keys_per_batch, self.ragged = k16, vt16, v_bias, keys_per_batch, ragged


class RaggedImageFeatures:
    """Pruned image features of DIFFERENT lengths as ONE packed tensor -- the encoder states of the t2i ITM rerank (one
    caption against k_test images from different evaluation batches, compress_retrieval_dtp.py:186-200). The reference
    pads every image to the longest one of the whole evaluation set with copies of its CLS token (:142-154) and runs a
    dense cross-attention; here the images stay packed (`cu_seqlens` layout, rows aligned to 8) and the padding is
    evaluated in closed form: `pad` copies of key 0 are the same as key 0 with ln(1 + pad) added to its logit.

    feats: list of [N_i, d] fp32 CUDA tensors (token 0 = CLS); pad_to: the length the reference would pad to (default:
    the longest of the list)."""

    def __init__(self, feats, pad_to=None):
        if not feats:
            raise RuntimeError("madtp_b200: RaggedImageFeatures needs at least one image")
        lens = [int(f.shape[0]) for f in feats]
        self.pad_to = max(lens) if pad_to is None else int(pad_to)
        if self.pad_to < max(lens):
            raise RuntimeError("madtp_b200: pad_to is shorter than the longest image")
        dev, d = feats[0].device, feats[0].shape[1]
        starts, total = [], 0
        for n in lens:
            starts.append(total)
            total += (n + 7) // 8 * 8
        self.rows = total
        self.max_len = max(lens)
        packed = torch.zeros(total, d, dtype=torch.float32, device=dev)
        for s, n, f in zip(starts, lens, feats):
            Fn.require_cuda(f, "image feature")
            packed[s:s + n] = f
        self.e16 = L.cast_f16(packed)
        self.k_start = torch.tensor(starts, dtype=torch.int32, device=dev)
        self.k_len = torch.tensor(lens, dtype=torch.int32, device=dev)
        self.key0_bias = torch.tensor([math.log(1 + self.pad_to - n) for n in lens], dtype=torch.float32, device=dev)
        self.B = len(feats)


class LayerLengths:
    """Device-resident lengths of one text layer (see functional.dtp_finish): tokens per sequence entering the layer,
    where the layer writes the count it leaves and its topk_num, and the (device-resident) number of image tokens the
    cross-attention reads, if that is dynamic too."""
    __slots__ = ("l_in", "l_out", "k_out", "nk_dev")

    def __init__(self, l_in, l_out, k_out, nk_dev=None):
        self.l_in, self.l_out, self.k_out, self.nk_dev = l_in, l_out, k_out, nk_dev


class BertEmbeddings(nn.Module):
    """word + position embeddings -> LayerNorm (models/nlvr_encoder.py:43-85)."""

    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))
        self.position_embedding_type = getattr(config, "position_embedding_type", "absolute")
        self.config = config

    def forward(self, input_ids=None, position_ids=None, inputs_embeds=None, past_key_values_length=0):
        _unsupported(position_ids=position_ids, inputs_embeds=inputs_embeds)
        if past_key_values_length != 0 or self.position_embedding_type != "absolute":
            raise NotImplementedError("madtp_b200: only absolute positions without a key/value cache")
        if not input_ids.is_cuda:
            raise RuntimeError("madtp_b200: input_ids must be a CUDA tensor -- this package has no CPU fallback")
        B, Ltok = input_ids.shape
        if Ltok > self.position_embeddings.weight.shape[0]:
            raise RuntimeError(f"madtp_b200: sequence length {Ltok} exceeds max_position_embeddings "
                               f"{self.position_embeddings.weight.shape[0]}")
        e = L.bert_embed(input_ids.to(torch.int64), self.word_embeddings.weight.detach(),
                         self.position_embeddings.weight.detach())
        d = e.shape[-1]
        y = Fn.layernorm_rows(e.view(B * Ltok, d), self.LayerNorm.weight, self.LayerNorm.bias, self.LayerNorm.eps,
                              f32=True)["y"]
        return y.view(B, Ltok, d)


class BertSelfAttention(nn.Module):
    CROSS_ATTENTION_MASK = True     # nlvr_encoder adds the encoder mask in cross-attention (:196); med.py does not (:197)

    def __init__(self, config, is_cross_attention):
        super().__init__()
        self.config = config
        if config.hidden_size % config.num_attention_heads != 0 and not hasattr(config, "embedding_size"):
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)" %
                             (config.hidden_size, config.num_attention_heads))
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = int(config.hidden_size / config.num_attention_heads)
        if self.attention_head_size != 64:
            raise RuntimeError("madtp_b200: the attention kernels are built for head_dim 64")
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        kv_in = config.encoder_width if is_cross_attention else config.hidden_size
        self.key = nn.Linear(kv_in, self.all_head_size)
        self.value = nn.Linear(kv_in, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)
        self.position_embedding_type = getattr(config, "position_embedding_type", "absolute")
        if self.position_embedding_type != "absolute":
            raise NotImplementedError("madtp_b200: relative position embeddings are not on the pruned path")
        self.save_attention = False
        self.attention_map = None
        self.cls_attn = None
        self.is_cross_attention = bool(is_cross_attention)
        self._cache = Fn.WeightCache()

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

    # -- prepared operands ------------------------------------------------------------------------------------
    def _qkv_split(self):
        ps = [self.query.weight, self.query.bias, self.key.weight, self.key.bias, self.value.weight, self.value.bias]
        return self._cache.get("qkv", ps, lambda: Fn.PreparedLinear(
            torch.cat([self.query.weight, self.key.weight, self.value.weight], 0),
            torch.cat([self.query.bias, self.key.bias, self.value.bias], 0), split=True))

    def _q_f16(self):
        return self._cache.get("q16", [self.query.weight, self.query.bias],
                               lambda: Fn.PreparedLinear(self.query.weight, self.query.bias, f16=True))

    def _kv_f16(self):
        ps = [self.key.weight, self.key.bias, self.value.weight, self.value.bias]
        return self._cache.get("kv16", ps, lambda: Fn.PreparedLinear(
            torch.cat([self.key.weight, self.value.weight], 0), torch.cat([self.key.bias, self.value.bias], 0),
            f16=True))

    # -- kernels ----------------------------------------------------------------------------------------------
    def _qkv_book_split(self, space_dict):
        """[Wq; Wk; Wv; codebook (zero-padded to 128 rows)]: BERT's q/k/v and the Query_model dots share the operand
        h, so one GEMM produces both (the ViT cannot do this: its q/k/v read LayerNorm(x), its codebook dots read x)."""
        ps = [self.query.weight, self.query.bias, self.key.weight, self.key.bias, self.value.weight, self.value.bias,
              space_dict]

        def build():
            pad = torch.zeros(Fn.TA_LD, space_dict.shape[1], dtype=torch.float32, device=space_dict.device)
            pad[:space_dict.shape[0]] = space_dict.detach()
            w = torch.cat([self.query.weight, self.key.weight, self.value.weight, pad], 0)
            b = torch.cat([self.query.bias, self.key.bias, self.value.bias,
                           torch.zeros(Fn.TA_LD, dtype=torch.float32, device=space_dict.device)], 0)
            return Fn.PreparedLinear(w, b, split=True)
        return self._cache.get("qkv_book", ps, build)

    def project_qkv_and_token_att(self, h_hi, h_lo, B, Ltok, space_dict, l_dev=None):
        """One split-operand (fp16 hi/lo) GEMM -> (qkv view [B, L, 3C], token_att view [B, L, 128]) over the same rows."""
        C = self.all_head_size
        out = Fn.linear_split(h_hi, h_lo, self._qkv_book_split(space_dict), m_dev=l_dev,
                              m_mult=B).view(B, Ltok, 3 * C + Fn.TA_LD)
        return out[..., :3 * C], out[..., 3 * C:]

    def self_rows(self, h_hi, h_lo, B, Ltok, key_mask, want_stats=True, qkv=None, causal=False, l_dev=None):
        """Self-attention on the scoring lane. Returns ctx16 [B,L,C]; stores AttnStats + cls_attn (:213-235).
        causal=True: decoder self-attention (key j visible to query i only if j <= i, models/med.py:749-771).
        l_dev: device-resident L (packed sequences in capacity-sized buffers)."""
        C = self.all_head_size
        if qkv is None:
            qkv = Fn.linear_split(h_hi, h_lo, self._qkv_split(), m_dev=l_dev, m_mult=B).view(B, Ltok, 3 * C)
        ctx16, stats = Fn.self_attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], self.num_attention_heads,
                                         1.0 / math.sqrt(self.attention_head_size), key_mask, want_stats, causal=causal,
                                         l_dev=l_dev)
        self.save_attention_map(stats)
        self.save_cls_attn(None if stats is None else stats.cls_attn[:, 1:])
        return ctx16

    def project_kv(self, enc16_rows):
        """enc16_rows [B*Nk, width] fp16 -> fp32 [B*Nk, 2C] = [key | value] projections of the encoder states."""
        return Fn.linear_f16(enc16_rows, self._kv_f16())

    def cross_rows(self, q, k, v, key_mask, out16):
        """q [B,Lq,C], k/v [B,Nk,C] fp32 views -> context written into out16 [B,Lq,C] (fp16 view)."""
        L.attn_fwd(q, k, v, self.num_attention_heads, 1.0 / math.sqrt(self.attention_head_size), out16,
                   key_mask=key_mask)
        return out16

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_value=None, output_attentions=False):
        Fn.require_cuda(hidden_states, "hidden_states")
        _eval_only(self)
        _unsupported(head_mask=head_mask, past_key_value=past_key_value, output_attentions=output_attentions)
        B, Ltok, d = hidden_states.shape
        C = self.all_head_size
        if encoder_hidden_states is None:
            h = hidden_states.contiguous()
            h_hi, h_lo = Fn.split_rows(h.view(B * Ltok, d))
            ctx16 = self.self_rows(h_hi, h_lo, B, Ltok, _key_mask(attention_mask, B, Ltok))
        else:
            enc = encoder_hidden_states.contiguous()
            Nk = enc.shape[1]
            kv = self.project_kv(L.cast_f16(enc.view(B * Nk, -1))).view(B, Nk, 2 * C)
            q = Fn.linear_f16(L.cast_f16(hidden_states.reshape(B * Ltok, d)), self._q_f16()).view(B, Ltok, C)
            ctx16 = torch.empty(B, Ltok, C, dtype=torch.float16, device=q.device)
            em = encoder_attention_mask if self.CROSS_ATTENTION_MASK else None
            self.cross_rows(q, kv[..., :C], kv[..., C:], _key_mask(em, B, Nk), ctx16)
        return (ctx16.float(), None)


class BertSelfOutput(nn.Module):
    def __init__(self, config, twin=False, merge=False):
        super().__init__()
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        if twin:
            self.dense0 = nn.Linear(config.hidden_size, config.hidden_size)
            self.dense1 = nn.Linear(config.hidden_size, config.hidden_size)
        else:
            self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.twin = twin
        if merge:
            self.act = nn.GELU()
            self.merge_layer = nn.Linear(config.hidden_size * 2, config.hidden_size)
            self.merge = True
        else:
            self.merge = False
        self._cache = Fn.WeightCache()

    def _dense(self):
        return self._cache.get("dense", [self.dense.weight, self.dense.bias],
                               lambda: Fn.PreparedLinear(self.dense.weight, self.dense.bias, f16=True))

    def _twin_avg(self):
        """(dense0(c0) + dense1(c1)) / 2 as ONE GEMM over K = 2C: 0.5 * [c0|c1] . [W0|W1]^T + 0.5 (b0 + b1)."""
        ps = [self.dense0.weight, self.dense0.bias, self.dense1.weight, self.dense1.bias]
        return self._cache.get("avg", ps, lambda: Fn.PreparedLinear(
            torch.cat([self.dense0.weight, self.dense1.weight], 1), self.dense0.bias + self.dense1.bias, f16=True,
            bias_scale=0.5))

    def _twin_sep(self):
        d0 = self._cache.get("d0", [self.dense0.weight, self.dense0.bias],
                             lambda: Fn.PreparedLinear(self.dense0.weight, self.dense0.bias, f16=True))
        d1 = self._cache.get("d1", [self.dense1.weight, self.dense1.bias],
                             lambda: Fn.PreparedLinear(self.dense1.weight, self.dense1.bias, f16=True))
        mg = self._cache.get("mg", [self.merge_layer.weight, self.merge_layer.bias],
                             lambda: Fn.PreparedLinear(self.merge_layer.weight, self.merge_layer.bias, f16=True))
        return d0, d1, mg

    def rows(self, ctx16, residual, *, f16=False, split=False, m_dev=None, m_mult=1):
        """ctx16 [rows, C] (or [rows, 2C] = [ctx0|ctx1] for the twin); residual fp32 [rows, C].
        Returns the layernorm_rows dict of LayerNorm(dense(ctx) + residual) (always with 'y').
        m_dev / m_mult: device-resident row count rows = *m_dev * m_mult."""
        dyn = dict(m_dev=m_dev, m_mult=m_mult)
        if not self.twin:
            pre = Fn.linear_f16(ctx16, self._dense(), residual=residual, **dyn)
        elif not self.merge:
            pre = Fn.linear_f16(ctx16, self._twin_avg(), residual=residual, alpha=0.5, **dyn)
        else:
            d0, d1, mg = self._twin_sep()
            C = d0.out_features
            d01 = L.empty((ctx16.shape[0], 2 * C), torch.float16, ctx16.device)
            Fn.linear_f16(ctx16[:, :C], d0, out=d01[:, :C], **dyn)
            Fn.linear_f16(ctx16[:, C:], d1, out=d01[:, C:], **dyn)
            pre = Fn.linear_f16(d01, mg, residual=residual, **dyn)
        return Fn.layernorm_rows(pre, self.LayerNorm.weight, self.LayerNorm.bias, self.LayerNorm.eps, f32=True,
                                 f16=f16, split=split, n_dev=m_dev, n_mult=m_mult)

    def forward(self, hidden_states, input_tensor):
        _eval_only(self)
        shape = input_tensor.shape
        C = shape[-1]
        res = input_tensor.reshape(-1, C).contiguous()
        if type(hidden_states) == list:
            ctx16 = torch.cat([L.cast_f16(hidden_states[0].reshape(-1, C)), L.cast_f16(hidden_states[1].reshape(-1, C))],
                              dim=1)
        else:
            ctx16 = L.cast_f16(hidden_states.reshape(-1, C))
        return self.rows(ctx16, res)["y"].view(shape)


class BertAttention(nn.Module):
    def __init__(self, config, is_cross_attention=False, layer_num=-1):
        super().__init__()
        if is_cross_attention:
            self.self0 = BertSelfAttention(config, is_cross_attention)
            self.self1 = BertSelfAttention(config, is_cross_attention)
        else:
            self.self = BertSelfAttention(config, is_cross_attention)
        self.output = BertSelfOutput(config, twin=is_cross_attention, merge=(is_cross_attention and layer_num >= 6))
        self.is_cross_attention = bool(is_cross_attention)
        self.pruned_heads = set()
        self._qcache = Fn.WeightCache()

    def prune_heads(self, heads):
        if len(heads):
            raise NotImplementedError("madtp_b200: head pruning is not supported")

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_value=None, output_attentions=False, space_dict=None):
        if type(encoder_hidden_states) == list:
            em = encoder_attention_mask if type(encoder_attention_mask) == list else [encoder_attention_mask] * 2
            o0 = self.self0(hidden_states, attention_mask, head_mask, encoder_hidden_states[0], em[0],
                            past_key_value, output_attentions)
            o1 = self.self1(hidden_states, attention_mask, head_mask, encoder_hidden_states[1], em[1],
                            past_key_value, output_attentions)
            attention_output = self.output([o0[0], o1[0]], hidden_states)
            return (attention_output,) + o0[1:]
        o = self.self(hidden_states, attention_mask, head_mask, encoder_hidden_states, encoder_attention_mask,
                      past_key_value, output_attentions)
        return (self.output(o[0], hidden_states),) + o[1:]


class BertIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)
        act = config.hidden_act
        if act not in ("gelu", "relu"):
            raise NotImplementedError(f"madtp_b200: hidden_act {act!r}")
        self.act_code = L.FFN_GELU if act == "gelu" else L.ACT_RELU
        self.intermediate_act_fn = nn.GELU() if act == "gelu" else nn.ReLU()
        self._cache = Fn.WeightCache()

    def rows(self, y16, m_dev=None, m_mult=1):
        w = self._cache.get("dense", [self.dense.weight, self.dense.bias],
                            lambda: Fn.PreparedLinear(self.dense.weight, self.dense.bias, f16=True))
        return Fn.linear_f16(y16, w, out_dtype=torch.float16, act=self.act_code, m_dev=m_dev, m_mult=m_mult)

    def forward(self, hidden_states):
        _eval_only(self)
        shape = hidden_states.shape
        return self.rows(L.cast_f16(hidden_states.reshape(-1, shape[-1]))).float().view(*shape[:-1], -1)


class BertOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self._cache = Fn.WeightCache()

    def rows(self, inter16, residual, m_dev=None, m_mult=1, planes=False):
        """planes=True: the LayerNorm launch also writes the fp16 hi/lo planes of its output -- the operand of the NEXT
        layer's fused q|k|v + codebook GEMM -- and the dict of outputs is returned instead of the fp32 rows."""
        w = self._cache.get("dense", [self.dense.weight, self.dense.bias],
                            lambda: Fn.PreparedLinear(self.dense.weight, self.dense.bias, f16=True))
        pre = Fn.linear_f16(inter16, w, residual=residual, m_dev=m_dev, m_mult=m_mult)
        o = Fn.layernorm_rows(pre, self.LayerNorm.weight, self.LayerNorm.bias, self.LayerNorm.eps, f32=True,
                              split=planes, n_dev=m_dev, n_mult=m_mult)
        return o if planes else o["y"]

    def forward(self, hidden_states, input_tensor):
        _eval_only(self)
        shape = input_tensor.shape
        return self.rows(L.cast_f16(hidden_states.reshape(-1, hidden_states.shape[-1])),
                         input_tensor.reshape(-1, shape[-1]).contiguous()).view(shape)


class BertLayer(nn.Module):
    MASK_MODE = 1   # nlvr_encoder.py:451-452: slot r of the pruned mask <- mask of the r-th ranked token

    def __init__(self, config, layer_num):
        super().__init__()
        self.config = config
        self.chunk_size_feed_forward = getattr(config, "chunk_size_feed_forward", 0)
        self.seq_len_dim = 1
        self.attention = BertAttention(config)
        self.layer_num = layer_num
        if self.config.add_cross_attention:
            self.crossattention = BertAttention(config, is_cross_attention=self.config.add_cross_attention,
                                                layer_num=layer_num)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)
        self.last_prune = None

    def Reduce_token(self, x, reduce_num, temperature=0, self_attn=None, cls_attn=None, token_attn=None, mask=None):
        """x [B,n,d] prunable tokens, mask [B,n] additive -> (x' [B,k+1,d], mask' [B,k+1]) or the inputs unchanged."""
        Fn.require_cuda(x, "x")
        if not isinstance(self_attn, Fn.AttnStats):
            raise RuntimeError("madtp_b200: Reduce_token expects the AttnStats handle of get_attention_map()")
        B, n, d = x.shape
        xin = torch.cat([x[:, :1, :], x], dim=1).contiguous()
        min_ = None
        if mask is not None:
            min_ = torch.cat([mask[:, :1], mask], dim=1).to(torch.float32).contiguous()
        res = Fn.dtp_prune(xin, self_attn, token_attn, float(temperature),
                           mask_mode=self.MASK_MODE if mask is not None else 0, mask_in=min_)
        self.last_prune = res
        if not res.pruned:
            return x, mask
        return res.x[:, 1:, :], (None if mask is None else res.mask[:, 1:])

    def forward(self, hidden_states, attention_mask=None, space_dict=None, head_mask=None,
                encoder_hidden_states=None, encoder_attention_mask=None, past_key_value=None,
                output_attentions=False, mode=None, token_attn=None, reduce_num=0, temperature=0, _kv=None):
        """Returns (layer_output, None, attention_mask') -- the last element is the pruned extended mask
        (models/nlvr_encoder.py:551-553), consumed by BertEncoder."""
        return self._forward_impl(hidden_states, attention_mask, head_mask, encoder_hidden_states,
                                  encoder_attention_mask, past_key_value, output_attentions, mode, token_attn,
                                  temperature, _kv)

    def _forward_impl(self, hidden_states, attention_mask, head_mask, encoder_hidden_states, encoder_attention_mask,
                      past_key_value, output_attentions, mode, token_attn, temperature, _kv, _qkv=None,
                      _causal=False, _dyn: Optional[LayerLengths] = None, _want_planes=False):
        Fn.require_cuda(hidden_states, "hidden_states")
        _eval_only(self)
        _unsupported(head_mask=head_mask, past_key_value=past_key_value, output_attentions=output_attentions)
        h = hidden_states.contiguous()
        B, Ltok, d = h.shape
        key_mask = _key_mask(attention_mask, B, Ltok)
        prune = temperature > 0
        if prune and token_attn is None:
            raise RuntimeError("madtp_b200: temperature > 0 needs token_attn")

        # device-resident lengths (functional.dtp_finish): h / key_mask are capacity-sized buffers of packed sequences
        l_dev = None if _dyn is None else _dyn.l_in
        # self-attention + output LayerNorm (:501-509)
        sa = self.attention.self
        if _qkv is not None:      # BertEncoder already projected q|k|v together with the codebook dots
            ctx16 = sa.self_rows(None, None, B, Ltok, key_mask, want_stats=prune, qkv=_qkv, causal=_causal, l_dev=l_dev)
        else:
            h_hi, h_lo = Fn.split_rows(h.view(B * Ltok, d), n_dev=l_dev, n_mult=B)
            ctx16 = sa.self_rows(h_hi, h_lo, B, Ltok, key_mask, want_stats=prune, causal=_causal, l_dev=l_dev)
        # score kernel + read-back of topk_num first: the output dense + LayerNorm below do not depend on them
        pend = Fn.dtp_score_async(sa.get_attention_map(), token_attn, float(temperature), Ltok - 1,
                                  n_dev=l_dev) if prune else None
        att = self.attention.output.rows(ctx16.view(B * Ltok, d), h.view(B * Ltok, d), f16=True, m_dev=l_dev, m_mult=B)
        att_f32, att16 = att["y"].view(B, Ltok, d), att.get("y16")

        # dynamic token pruning between self- and cross-attention (:519-533)
        self.last_prune = None
        if prune:
            if key_mask is None:
                key_mask = L.zeros((B, Ltok), torch.float32, h.device)
            res = Fn.dtp_finish(att_f32, pend, mask_mode=self.MASK_MODE, mask_in=key_mask, want_f16=True,
                                n_dev=l_dev, n_out=None if _dyn is None else _dyn.l_out,
                                k_out=None if _dyn is None else _dyn.k_out)
            self.last_prune = res
            att_f32 = res.x
            if _dyn is not None:    # capacity-sized results, packed with the length the select kernel wrote to l_out
                key_mask = res.mask
                att16 = res.x16.view(B * Ltok, d)
                l_dev = _dyn.l_out
            elif res.pruned:        # the gather kernel also wrote the fp16 operand of the next GEMM
                key_mask = res.mask.contiguous()
                Ltok = att_f32.shape[1]
                att16 = res.x16.view(B * Ltok, d)
            attention_mask = key_mask.view(B, 1, 1, Ltok)
        att_rows = att_f32.reshape(B * Ltok, d)

        if mode == 'multimodal':
            assert encoder_hidden_states is not None, "encoder_hidden_states must be given for cross-attention layers"
            att_rows, att16 = self._cross(att_rows, att16, B, Ltok, encoder_hidden_states, encoder_attention_mask, _kv,
                                          l_dev=l_dev, nk_dev=None if _dyn is None else _dyn.nk_dev)

        inter16 = self.intermediate.rows(att16, m_dev=l_dev, m_mult=B)
        self.out_planes = None
        if _want_planes:
            o = self.output.rows(inter16, att_rows, m_dev=l_dev, m_mult=B, planes=True)
            out = o["y"].view(B, Ltok, d)
            self.out_planes = (o["y_hi"], o["y_lo"], out)
        else:
            out = self.output.rows(inter16, att_rows, m_dev=l_dev, m_mult=B).view(B, Ltok, d)
        return (out, None, attention_mask)

    def _cross(self, att_rows, att16, B, Ltok, enc, enc_mask, kv, l_dev=None, nk_dev=None):
        """Twin (list-valued encoder states, :312-335) or single cross-attention + output LayerNorm."""
        ca = self.crossattention
        d = att_rows.shape[1]
        if isinstance(enc, RaggedImageFeatures):
            selfs, enc, masks = [ca.self], [None], [None]
        elif hasattr(enc, "y16"):   # vit.DeviceEncoded: one group of packed image tokens per branch, no encoder mask
            selfs = [ca.self0, ca.self1] if enc.y16.shape[0] == 2 else [ca.self]
            enc, masks = [None] * len(selfs), [None] * len(selfs)
        elif type(enc) == list:
            selfs = [ca.self0, ca.self1]
            masks = enc_mask if type(enc_mask) == list else [enc_mask, enc_mask]
        else:
            selfs, enc, masks = [ca.self], [enc], [enc_mask]
        C = selfs[0].all_head_size
        nb = len(selfs)
        ctx = L.empty((B, Ltok, nb * C), torch.float16, att_rows.device)
        use_tc = kv is not None and all(isinstance(kv[i], CrossKV) for i in range(nb))   # tensor-core kernel operands
        if (l_dev is not None or nk_dev is not None) and not use_tc:
            raise RuntimeError("madtp_b200: device-resident lengths need the tensor-core cross-attention operands")
        dyn = dict(m_dev=l_dev, m_mult=B)
        qdt = torch.float16 if use_tc else torch.float32
        if nb == 2:   # both query projections read the same rows: one GEMM over [Wq0; Wq1]
            ps = [p for s in selfs for p in (s.query.weight, s.query.bias)]
            wq = ca._qcache.get("q01", ps, lambda: Fn.PreparedLinear(
                torch.cat([s.query.weight for s in selfs], 0), torch.cat([s.query.bias for s in selfs], 0), f16=True))
            q_all = Fn.linear_f16(att16, wq, out_dtype=qdt, **dyn).view(B, Ltok, nb * C)
        for i, s in enumerate(selfs):
            q = q_all[..., i * C:(i + 1) * C] if nb == 2 else Fn.linear_f16(att16, s._q_f16(), out_dtype=qdt,
                                                                             **dyn).view(B, Ltok, C)
            em = masks[i] if s.CROSS_ATTENTION_MASK else None
            out = ctx[..., i * C:(i + 1) * C]
            if use_tc and kv[i].ragged is not None:
                ckv = kv[i]
                k_start, k_len, key0_bias, max_len = ckv.ragged
                L.attn_cross_tc_ragged(q, ckv.k16, ckv.vt16, s.num_attention_heads,
                                       1.0 / math.sqrt(s.attention_head_size), out, k_start, k_len, max_len,
                                       v_bias=ckv.v_bias, key0_bias=key0_bias, lq_dev=l_dev)
                continue
            if use_tc:
                ckv = kv[i]
                Nk = ckv.k16.shape[-2]
                L.attn_cross_tc(q, ckv.k16, ckv.vt16, s.num_attention_heads, 1.0 / math.sqrt(s.attention_head_size), out,
                                keys_per_batch=ckv.keys_per_batch, v_bias=ckv.v_bias, key_mask=_key_mask(em, B, Nk),
                                lq_dev=l_dev, nk_dev=nk_dev)
                continue
            Nk = enc[i].shape[1]
            if kv is not None:
                k, v = kv[i]
            else:
                e16 = L.cast_f16(enc[i].contiguous().view(B * Nk, -1))
                kvp = s.project_kv(e16).view(B, Nk, 2 * C)
                k, v = kvp[..., :C], kvp[..., C:]
            s.cross_rows(q, k, v, _key_mask(em, B, Nk), out)
        o = ca.output.rows(ctx.view(B * Ltok, nb * C), att_rows, f16=True, **dyn)
        return o["y"], o["y16"]

    def feed_forward_chunk(self, attention_output):
        return self.output(self.intermediate(attention_output), attention_output)


class BertEncoder(nn.Module):
    REQUIRE_SPACE_DICT = True
    LAYER_CLS = None    # set below (BertLayer); med.py overrides

    def __init__(self, config, sd_dim=768, map_func=False):
        super().__init__()
        self.config = config
        layer_cls = self.LAYER_CLS or BertLayer
        self.layer = nn.ModuleList([layer_cls(config, i) for i in range(config.num_hidden_layers)])
        self.gradient_checkpointing = False
        self.txt_query_model = Query_model(ft_dim=config.hidden_size, sd_dim=sd_dim, temperature=1,
                                           att_func_type='sparsemax', pool_type='max', map_func=map_func)
        self._cache = Fn.WeightCache()

    def _all_kv_weights(self, which: str):
        """[key; value] weights of crossattention.<which> of every layer stacked along N: one GEMM projects the image
        tokens for all layers at once (they do not depend on the text stream)."""
        selfs = [getattr(l.crossattention, which) for l in self.layer]
        ps = [p for s in selfs for p in (s.key.weight, s.key.bias, s.value.weight, s.value.bias)]
        return self._cache.get("kv_" + which, ps, lambda: Fn.PreparedLinear(
            torch.cat([torch.cat([s.key.weight, s.value.weight], 0) for s in selfs], 0),
            torch.cat([torch.cat([s.key.bias, s.value.bias], 0) for s in selfs], 0), f16=True))

    def _all_k_vt_weights(self, which: str):
        """Tensor-core cross-attention operands: key weights of every layer stacked along N (K = X . Wk^T, fp16 out),
        value weights stacked along M (V^T = Wv . X^T, keys contiguous) and the value biases [layers, C], which the
        attention kernel adds to its normalised output (rows of P sum to one)."""
        selfs = [getattr(l.crossattention, which) for l in self.layer]
        ps = [p for s in selfs for p in (s.key.weight, s.key.bias, s.value.weight, s.value.bias)]

        def build():
            wk = Fn.PreparedLinear(torch.cat([s.key.weight for s in selfs], 0),
                                   torch.cat([s.key.bias for s in selfs], 0), f16=True)
            wv16 = L.cast_f16(torch.cat([s.value.weight for s in selfs], 0).detach().float())
            vb = torch.stack([s.value.bias.detach().float() for s in selfs], 0).contiguous()
            return wk, wv16, vb
        return self._cache.get("kvt_" + which, ps, build)

    def _project_encoder_states(self, enc, Lq: int = 1 << 30):
        """Returns for every layer a list with one entry per cross-attention branch: CrossKV (fp16 K, V^T and value
        bias for the tensor-core kernel) when the problem fits it, else (k, v) fp32 views [B, Nk, C]."""
        names = ["self0", "self1"] if type(enc) == list else ["self"]
        encs = enc if type(enc) == list else [enc]
        C = self.config.hidden_size
        per_layer = [[] for _ in self.layer]
        use_tc = C % 64 == 0 and all(L.cross_tc_supported(Lq, e.shape[1]) for e in encs)
        for name, e in zip(names, encs):
            Fn.require_cuda(e, "encoder_hidden_states")
            B, Nk, w = e.shape
            # the same encoder states for every text of the batch (ITM rerank: one image against k_test captions,
            # compress_retrieval_dtp.py:166-176): project the image ONCE and let the attention kernels broadcast it
            # -- the reference recomputes the K/V projections k_test times
            broadcast = B > 1 and e.stride(0) == 0
            if use_tc:
                # every sequence owns P = Nk rounded up to 8 key rows / V^T columns: TMA box origins must be 16-byte
                # aligned; the padding rows are zero and masked out by the kernel (keys >= Nk)
                wk, wv16, vb = self._all_k_vt_weights(name)
                Bk = 1 if broadcast else B
                P = (Nk + 7) // 8 * 8
                src = (e[:1] if broadcast else e)
                if P == Nk:
                    e16 = L.cast_f16(src.contiguous().view(Bk * Nk, w))
                else:
                    e16 = torch.zeros(Bk, P, w, dtype=torch.float16, device=e.device)
                    e16[:, :Nk] = L.cast_f16(src.contiguous().view(Bk * Nk, w)).view(Bk, Nk, w)
                    e16 = e16.view(Bk * P, w)
                allk = Fn.linear_f16(e16, wk, out_dtype=torch.float16).view(Bk, P, -1)        # [Bk, P, layers*C]
                allvt = torch.empty(wv16.shape[0], Bk * P, dtype=torch.float16, device=e.device)  # [layers*C, Bk*P]
                L.gemm(L.GEMM_F16, wv16, e16, allvt)
                for i in range(len(self.layer)):
                    k16 = allk[:, :Nk, i * C:(i + 1) * C]
                    per_layer[i].append(CrossKV(k16[0] if broadcast else k16, allvt[i * C:(i + 1) * C], vb[i],
                                                0 if broadcast else P))
                continue
            e16 = L.cast_f16((e[0] if broadcast else e).contiguous().view(-1, w))
            allkv = Fn.linear_f16(e16, self._all_kv_weights(name))
            allkv = allkv.view(1, Nk, -1).expand(B, Nk, -1) if broadcast else allkv.view(B, Nk, -1)
            for i in range(len(self.layer)):
                o = i * 2 * C
                per_layer[i].append((allkv[..., o:o + C], allkv[..., o + C:o + 2 * C]))
        return per_layer

    def _project_encoder_states_ragged(self, enc: "RaggedImageFeatures"):
        """K / V^T of every layer over the packed rows of a RaggedImageFeatures (one GEMM each)."""
        C = self.config.hidden_size
        wk, wv16, vb = self._all_k_vt_weights("self")
        allk = Fn.linear_f16(enc.e16, wk, out_dtype=torch.float16)                         # [rows, layers * C]
        allvt = L.empty((wv16.shape[0], enc.rows), torch.float16, enc.e16.device)
        L.gemm(L.GEMM_F16, wv16, enc.e16, allvt)
        rag = (enc.k_start, enc.k_len, enc.key0_bias, enc.max_len)
        return [[CrossKV(allk[:, i * C:(i + 1) * C], allvt[i * C:(i + 1) * C], vb[i], 0, rag)]
                for i in range(len(self.layer))]

    def _project_encoder_states_device(self, enc):
        """The same operands from a vit.DeviceEncoded record whose final LayerNorm was written by madtp_layernorm_pack:
        fp16 image tokens [groups, per_group * P, w] with P = the DEVICE-RESIDENT token count rounded up to 8. One
        group per cross-attention branch (BLIP-NLVR: image0 -> self0, image1 -> self1); the K and V^T projections of
        all layers run as one GEMM each over the dynamic row count."""
        names = ["self0", "self1"] if enc.y16.shape[0] == 2 else ["self"]
        if enc.y16.shape[0] != len(names):
            raise RuntimeError("madtp_b200: one group of image tokens per cross-attention branch")
        C = self.config.hidden_size
        per, Pcap = enc.per_group, enc.P
        per_layer = [[] for _ in self.layer]
        for gi, name in enumerate(names):
            wk, wv16, vb = self._all_k_vt_weights(name)
            e16 = enc.y16[gi]                                                                  # [per * Pcap, w]
            allk = Fn.linear_f16(e16, wk, out_dtype=torch.float16, m_dev=enc.p_dev, m_mult=per).view(per, Pcap, -1)
            allvt = L.empty((wv16.shape[0], per * Pcap), torch.float16, e16.device)
            L.gemm(L.GEMM_F16, wv16, e16, allvt, n_dev=enc.p_dev, n_mult=per)
            for i in range(len(self.layer)):
                per_layer[i].append(CrossKV(allk[:, :enc.cap, i * C:(i + 1) * C], allvt[i * C:(i + 1) * C], vb[i], Pcap))
        return per_layer

    def forward(self, hidden_states, attention_mask=None, space_dict=None, temperature=0, head_mask=None,
                encoder_hidden_states=None, encoder_attention_mask=None, past_key_values=None, use_cache=None,
                output_attentions=False, output_hidden_states=False, return_dict=True, mode='multimodal',
                _causal=False, _device=False):
        """_device=True (internal): device-resident lengths -- `hidden_states` stays a capacity-sized buffer of packed
        sequences, nothing is read back, and the return value is (final states [B, L_cap, d], sd_txt_ft, Trajectory,
        device scalar with the final length). `encoder_hidden_states` may then be a vit.DeviceEncoded record."""
        _unsupported(past_key_values=past_key_values, use_cache=use_cache, output_attentions=output_attentions,
                     output_hidden_states=output_hidden_states)
        Fn.require_cuda(hidden_states, "hidden_states")
        if space_dict is None and self.REQUIRE_SPACE_DICT:
            raise RuntimeError("madtp_b200: nlvr_encoder.BertEncoder always queries the codebook (:605-608)")
        B = hidden_states.shape[0]
        token_num = hidden_states.shape[-2]
        reduce_num = int((token_num - 1) // self.config.num_hidden_layers)
        kv = None
        nk_dev = None
        enc_dev = encoder_hidden_states if hasattr(encoder_hidden_states, "y16") else None
        if mode == 'multimodal' and isinstance(encoder_hidden_states, RaggedImageFeatures):
            kv = self._project_encoder_states_ragged(encoder_hidden_states)
        elif mode == 'multimodal' and enc_dev is not None:
            kv = self._project_encoder_states_device(enc_dev)
            nk_dev = enc_dev.n_dev
        elif mode == 'multimodal' and encoder_hidden_states is not None:
            kv = self._project_encoder_states(encoder_hidden_states, hidden_states.shape[1])
        dims = ks = traj = None
        depth = len(self.layer)
        if _device:
            lens = L.empty((2 * depth + 1,), torch.int32, hidden_states.device)
            dims, ks = lens[:depth + 1], lens[depth + 1:]
            dims[:1].fill_(token_num)
            ks.fill_(-1)
            traj = Fn.Trajectory(dims, ks)
        sd_txt_ft_all = None
        planes = None
        fuse_planes = space_dict is not None and not self.txt_query_model.map_func
        for i, layer_module in enumerate(self.layer):
            h = hidden_states.contiguous()
            Ltok, d = h.shape[1], h.shape[2]
            token_attn = qkv = None
            dyn = LayerLengths(dims[i:i + 1], dims[i + 1:i + 2], ks[i:i + 1], nk_dev) if _device else None
            l_dev = None if dyn is None else dyn.l_in
            if space_dict is not None and not self.txt_query_model.map_func:
                # q|k|v and the codebook dots share the operand h: one split-operand GEMM (see _qkv_book_split). The
                # planes of h come out of the previous layer's output LayerNorm launch when there is one.
                if planes is not None and planes[2].data_ptr() == h.data_ptr() and planes[2].shape == h.shape:
                    h_hi, h_lo = planes[0], planes[1]
                else:
                    h_hi, h_lo = Fn.split_rows(h.view(B * Ltok, d), n_dev=l_dev, n_mult=B)
                qkv, ta_full = layer_module.attention.self.project_qkv_and_token_att(h_hi, h_lo, B, Ltok, space_dict,
                                                                                     l_dev=l_dev)
                token_attn, sd_txt_ft_all = Fn.query_model_from_token_att(ta_full, h, space_dict.shape[0],
                                                                          self.txt_query_model.att_dim, sd_txt_ft_all,
                                                                          n_dev=l_dev)
            elif space_dict is not None:
                h_hi, h_lo = Fn.split_rows(h.view(B * Ltok, d), n_dev=l_dev, n_mult=B)
                token_attn, sd_txt_ft_all = self.txt_query_model.forward_rows(h, h_hi, h_lo, space_dict,
                                                                              sd_txt_ft_all, n_dev=l_dev)
            layer_outputs = layer_module._forward_impl(h, attention_mask, None, encoder_hidden_states,
                                                       encoder_attention_mask, None, False, mode, token_attn,
                                                       temperature, None if kv is None else kv[i], _qkv=qkv,
                                                       _causal=_causal, _dyn=dyn,
                                                       _want_planes=(fuse_planes and i + 1 < depth))
            planes = getattr(layer_module, "out_planes", None)
            if _device:
                layer_module.last_prune = Fn.LazyPrune(traj, i, layer_module.last_prune, B)
            hidden_states = layer_outputs[0]
            attention_mask = layer_outputs[-1]
        if _device:
            return hidden_states, sd_txt_ft_all, traj, dims[depth:depth + 1]
        if not return_dict:
            return (hidden_states,), sd_txt_ft_all
        return EncoderOutput(hidden_states), sd_txt_ft_all


class BertModel(nn.Module):
    """models/nlvr_encoder.py:760-1015 without the HuggingFace PreTrainedModel machinery (no pooler on this path)."""

    def __init__(self, config, add_pooling_layer=True, sd_dim=768):
        super().__init__()
        if add_pooling_layer:
            raise NotImplementedError("madtp_b200: the pooler is not used by BLIP (add_pooling_layer=False)")
        self.config = config
        self.embeddings = BertEmbeddings(config)
        self.encoder = BertEncoder(config, sd_dim)
        self.pooler = None

    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    @staticmethod
    def get_extended_attention_mask(attention_mask, input_shape=None, device=None, is_decoder=False):
        """(1 - mask) * -10000 broadcastable over heads and queries (:826-871). For a decoder (`is_decoder=True`) this
        is only the key-padding factor of the reference's mask; the causal factor (models/med.py:749-771) is applied
        inside the attention kernels (`causal` flag), never materialised."""
        if attention_mask.dim() != 2:
            raise NotImplementedError("madtp_b200: only [batch, seq] attention masks")
        return (1.0 - attention_mask[:, None, None, :].to(torch.float32)) * -10000.0

    invert_attention_mask = get_extended_attention_mask

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, space_dict=None, temperature=0, position_ids=None,
                head_mask=None, inputs_embeds=None, encoder_embeds=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_values=None, use_cache=None, output_attentions=None,
                output_hidden_states=None, return_dict=None, is_decoder=False, mode='multimodal'):
        _eval_only(self)
        _unsupported(position_ids=position_ids, head_mask=head_mask, inputs_embeds=inputs_embeds,
                     past_key_values=past_key_values, use_cache=use_cache, output_attentions=output_attentions,
                     output_hidden_states=output_hidden_states)
        if is_decoder and temperature > 0:
            raise NotImplementedError("madtp_b200: the decoder is unpruned in the reference (blip_vqa.py:161-190)")
        if input_ids is not None:
            B, Ltok = input_ids.shape
            device = input_ids.device
        elif encoder_embeds is not None:
            B, Ltok = encoder_embeds.shape[:2]
            device = encoder_embeds.device
        else:
            raise ValueError("You have to specify either input_ids or encoder_embeds")
        if attention_mask is None:
            attention_mask = torch.ones((B, Ltok), device=device)
        ext = self.get_extended_attention_mask(attention_mask)
        enc_ext = None
        if encoder_hidden_states is not None:
            if type(encoder_attention_mask) == list:
                enc_ext = [self.invert_attention_mask(m) for m in encoder_attention_mask]
            elif encoder_attention_mask is not None:
                enc_ext = self.invert_attention_mask(encoder_attention_mask)
            if enc_ext is not None and type(encoder_hidden_states) == list and type(enc_ext) != list:
                enc_ext = [enc_ext] * len(encoder_hidden_states)
        emb = self.embeddings(input_ids=input_ids) if encoder_embeds is None else encoder_embeds
        from .vit import device_lengths_enabled
        prunes = temperature > 0 and (space_dict is not None or self.encoder.REQUIRE_SPACE_DICT)
        if prunes and not is_decoder and device_lengths_enabled() and Ltok <= 64:
            # device-resident lengths: ONE host read-back (the final length, to shape the returned tensor) instead of
            # one per layer (models/nlvr_encoder.py:432 / models/med.py:369)
            enc_shape = None
            if torch.is_tensor(encoder_hidden_states):      # a zero batch stride (ITM rerank broadcast) changes the launches
                enc_shape = (tuple(encoder_hidden_states.shape), encoder_hidden_states.stride(0) == 0)
            elif type(encoder_hidden_states) == list:
                enc_shape = tuple((tuple(e.shape), e.stride(0) == 0) for e in encoder_hidden_states)
            elif isinstance(encoder_hidden_states, RaggedImageFeatures):
                enc_shape = ("ragged", encoder_hidden_states.rows, encoder_hidden_states.B)
            with L.arena_for(self, (B, Ltok, mode, enc_shape, enc_ext is None)):
                h, sd_txt_ft, traj, l_dev = self.encoder(emb, attention_mask=ext, space_dict=space_dict,
                                                         temperature=temperature,
                                                         encoder_hidden_states=encoder_hidden_states,
                                                         encoder_attention_mask=enc_ext, mode=mode, _device=True)
                n = traj.host()[0][-1]
                d = h.shape[-1]
                # clones: the arena's buffers are reused by the next call with the same shapes
                return (EncoderOutput(h.reshape(-1)[:B * n * d].view(B, n, d).clone()),
                        None if sd_txt_ft is None else sd_txt_ft.clone())
        out, sd_txt_ft = self.encoder(emb, attention_mask=ext, space_dict=space_dict, temperature=temperature,
                                      encoder_hidden_states=encoder_hidden_states, encoder_attention_mask=enc_ext,
                                      mode=mode, _causal=bool(is_decoder))
        return out, sd_txt_ft

    @torch.no_grad()
    def forward_device(self, input_ids, attention_mask, enc, space_dict, temperature, mode='multimodal'):
        """Pruned text encoder with device-resident lengths end to end: `enc` is the vit.DeviceEncoded record of the
        image encoder (packed fp16 image tokens + their device-resident count). Nothing is read back. Returns
        (final states [B, L_cap, d] packed, sd_txt_ft, Trajectory, device scalar with the final length)."""
        _eval_only(self)
        if attention_mask is None:
            attention_mask = torch.ones(input_ids.shape, device=input_ids.device)
        ext = self.get_extended_attention_mask(attention_mask)
        emb = self.embeddings(input_ids=input_ids)
        return self.encoder(emb, attention_mask=ext, space_dict=space_dict, temperature=temperature,
                            encoder_hidden_states=enc, encoder_attention_mask=None, mode=mode, _device=True)
