"""Mirror of the encoder half of the reference's models/med.py (BLIP retrieval / VQA / caption text encoder):
BertEmbeddings (:43-86), BertSelfAttention (:89-236), BertSelfOutput (:239-250), BertAttention (:253-299),
BertIntermediate / BertOutput (:302-329), BertLayer (:332-467), BertEncoder (:470-598), BertModel (:686-929), with
med.py's own signatures (its BertLayer / BertEncoder argument order differs from nlvr_encoder.py's) on top of the
same sm_100a kernels, plus the evaluation forward of BertLMHeadModel (:933-1094) that VQA answer ranking needs
(models/blip_vqa.py:156-203); `generate` (beam search over past_key_values) is out of scope.

Differences from nlvr_encoder.py that this file carries (hard part 5):
  * single cross-attention (`crossattention.self`, `crossattention.output.dense`);
  * cross-attention adds NO mask (med.py:197) -- which is what lets VQA pass a stale full-length mask with pruned states;
  * Reduce_token takes `topk(k+1)` and keeps its first k (:377-378): the kept-token masks travel with their tokens and
    the merged slot inherits the mask of the (k+1)-th ranked token (:388-390)  -> dtp_select mask_mode 2;
  * `mode='text'` skips cross-attention; the codebook query is optional (`space_dict=None`, :513-524).
"""
from __future__ import annotations

import torch
from torch import nn

from . import functional as Fn
from . import nlvr_encoder as _ne
from .configuration import BertConfig  # noqa: F401
from .nlvr_encoder import BertEmbeddings, BertIntermediate, BertOutput, EncoderOutput  # noqa: F401
from .utils import Query_model, vector_gather  # noqa: F401


class BertSelfAttention(_ne.BertSelfAttention):
    CROSS_ATTENTION_MASK = False


class BertSelfOutput(_ne.BertSelfOutput):
    def __init__(self, config):
        super().__init__(config, twin=False, merge=False)


class BertAttention(nn.Module):
    def __init__(self, config, is_cross_attention=False):
        super().__init__()
        self.self = BertSelfAttention(config, is_cross_attention)
        self.output = BertSelfOutput(config)
        self.is_cross_attention = bool(is_cross_attention)
        self.pruned_heads = set()

    def prune_heads(self, heads):
        if len(heads):
            raise NotImplementedError("madtp_b200: head pruning is not supported")

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_value=None, output_attentions=False):
        o = self.self(hidden_states, attention_mask, head_mask, encoder_hidden_states, encoder_attention_mask,
                      past_key_value, output_attentions)
        return (self.output(o[0], hidden_states),) + o[1:]


class BertLayer(_ne.BertLayer):
    MASK_MODE = 2

    def __init__(self, config, layer_num):
        nn.Module.__init__(self)
        self.config = config
        self.chunk_size_feed_forward = getattr(config, "chunk_size_feed_forward", 0)
        self.seq_len_dim = 1
        self.attention = BertAttention(config)
        self.layer_num = layer_num
        if self.config.add_cross_attention:
            self.crossattention = BertAttention(config, is_cross_attention=self.config.add_cross_attention)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)
        self.last_prune = None

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_value=None, output_attentions=False, mode=None,
                space_dict=None, token_attn=None, reduce_num=0, temperature=0, _kv=None):
        """(layer_output, None, attention_mask') -- models/med.py:393-462; the last element is the pruned mask."""
        return self._forward_impl(hidden_states, attention_mask, head_mask, encoder_hidden_states,
                                  encoder_attention_mask, past_key_value, output_attentions, mode, token_attn,
                                  temperature, _kv)


class BertEncoder(_ne.BertEncoder):
    REQUIRE_SPACE_DICT = False
    LAYER_CLS = BertLayer

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_values=None, use_cache=None, output_attentions=False,
                output_hidden_states=False, return_dict=True, mode='multimodal', space_dict=None, temperature=0,
                _causal=False, _device=False):
        return _ne.BertEncoder.forward(self, hidden_states, attention_mask=attention_mask, space_dict=space_dict,
                                       temperature=temperature, head_mask=head_mask,
                                       encoder_hidden_states=encoder_hidden_states,
                                       encoder_attention_mask=encoder_attention_mask, past_key_values=past_key_values,
                                       use_cache=use_cache, output_attentions=output_attentions,
                                       output_hidden_states=output_hidden_states, return_dict=return_dict, mode=mode,
                                       _causal=_causal, _device=_device)


class BertModel(_ne.BertModel):
    def __init__(self, config, add_pooling_layer=True, sd_dim=768, map_func=False):
        nn.Module.__init__(self)
        if add_pooling_layer:
            raise NotImplementedError("madtp_b200: the pooler is not used by BLIP (add_pooling_layer=False)")
        self.config = config
        self.embeddings = BertEmbeddings(config)
        self.encoder = BertEncoder(config, sd_dim, map_func)
        self.pooler = None

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, head_mask=None, inputs_embeds=None,
                encoder_embeds=None, encoder_hidden_states=None, encoder_attention_mask=None, past_key_values=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None, is_decoder=False,
                mode='multimodal', space_dict=None, temperature=0):
        return _ne.BertModel.forward(self, input_ids=input_ids, attention_mask=attention_mask, space_dict=space_dict,
                                     temperature=temperature, position_ids=position_ids, head_mask=head_mask,
                                     inputs_embeds=inputs_embeds, encoder_embeds=encoder_embeds,
                                     encoder_hidden_states=encoder_hidden_states,
                                     encoder_attention_mask=encoder_attention_mask, past_key_values=past_key_values,
                                     use_cache=use_cache, output_attentions=output_attentions,
                                     output_hidden_states=output_hidden_states, return_dict=return_dict,
                                     is_decoder=is_decoder, mode=mode)


# ---------------------------------------------------------------------------------------------------------------
# Answer decoder for VQA ranking  (models/med.py:615-657, 933-1094)
# ---------------------------------------------------------------------------------------------------------------
class BertPredictionHeadTransform(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class BertLMPredictionHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.decoder = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))
        self.decoder.bias = self.bias          # models/med.py:644
        self._cache = Fn.WeightCache()

    def rows(self, h2d):
        """h2d [R, d] fp32 -> logits [R, vocab] fp32: dense -> GELU(erf) -> LayerNorm -> tied decoder + bias, all on the
        scoring lane (the ranking compares summed log-probabilities of candidates)."""
        t = self.transform
        dense = self._cache.get("dense", [t.dense.weight, t.dense.bias],
                                lambda: Fn.PreparedLinear(t.dense.weight, t.dense.bias, split=True))
        dec = self._cache.get("dec", [self.decoder.weight, self.bias],
                              lambda: Fn.PreparedLinear(self.decoder.weight, self.bias, split=True))
        hi, lo = Fn.split_rows(h2d)
        y = Fn.linear_split(hi, lo, dense, act=Fn.L.ACT_GELU)
        ln = Fn.layernorm_rows(y, t.LayerNorm.weight, t.LayerNorm.bias, t.LayerNorm.eps, split=True)
        return Fn.linear_split(ln["y_hi"], ln["y_lo"], dec)

    def forward(self, hidden_states):
        Fn.require_cuda(hidden_states, "hidden_states")
        shape = hidden_states.shape
        return self.rows(hidden_states.reshape(-1, shape[-1]).contiguous()).view(*shape[:-1], -1)


class BertOnlyMLMHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.predictions = BertLMPredictionHead(config)

    def forward(self, sequence_output):
        return self.predictions(sequence_output)


class LMOutput:
    """The fields of CausalLMOutputWithCrossAttentions that the evaluation path fills."""

    def __init__(self, loss, logits):
        self.loss, self.logits = loss, logits

    def __getitem__(self, i):
        return (self.loss, self.logits)[i] if self.loss is not None else (self.logits,)[i]


class BertLMHeadModel(nn.Module):
    """models/med.py:933-1094, evaluation forward only (`labels` + reduction 'none' or plain logits)."""

    def __init__(self, config, sd_dim=768):
        super().__init__()
        self.config = config
        self.bert = BertModel(config, add_pooling_layer=False, sd_dim=sd_dim)
        self.cls = BertOnlyMLMHead(config)
        self.tie_weights()

    def tie_weights(self):
        self.cls.predictions.decoder.weight = self.bert.embeddings.word_embeddings.weight     # models/med.py:947-950

    def get_output_embeddings(self):
        return self.cls.predictions.decoder

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, head_mask=None, inputs_embeds=None,
                encoder_hidden_states=None, encoder_attention_mask=None, labels=None, past_key_values=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None,
                return_logits=False, is_decoder=True, reduction='mean', mode='multimodal', space_dict=None,
                temperature=0, train=False):
        if train:
            raise NotImplementedError("madtp_b200: evaluation only")
        out, _ = self.bert(input_ids, attention_mask=attention_mask, position_ids=position_ids, head_mask=head_mask,
                           inputs_embeds=inputs_embeds, encoder_hidden_states=encoder_hidden_states,
                           encoder_attention_mask=encoder_attention_mask, past_key_values=past_key_values,
                           use_cache=use_cache, output_attentions=output_attentions,
                           output_hidden_states=output_hidden_states, return_dict=True, is_decoder=is_decoder,
                           mode=mode, space_dict=space_dict, temperature=temperature)
        seq = out.last_hidden_state
        B, Lt, d = seq.shape
        logits = self.cls(seq)                                       # [B, L, vocab]
        if return_logits:
            return logits[:, :-1, :].contiguous()
        loss = None
        if labels is not None:                                       # :1040-1047
            if reduction != 'none':
                raise NotImplementedError("madtp_b200: only reduction='none' (per-sequence sums) is on the eval path")
            # position t predicts token t+1; the last position predicts nothing (label -100), so the logits are used in
            # place instead of being copied into a shifted tensor
            lab = torch.full((B, Lt), -100, dtype=torch.int64, device=logits.device)
            lab[:, :-1] = labels[:, 1:]
            tok_loss, _ = Fn.L.lm_nll(logits.view(B * Lt, -1), lab.view(-1), 0.1)
            loss = tok_loss.view(B, Lt).sum(1)
        return LMOutput(loss, logits)
