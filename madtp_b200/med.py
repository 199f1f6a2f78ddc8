"""Mirror of the encoder half of the reference's models/med.py (BLIP retrieval / VQA / caption text encoder):
BertEmbeddings (:43-86), BertSelfAttention (:89-236), BertSelfOutput (:239-250), BertAttention (:253-299),
BertIntermediate / BertOutput (:302-329), BertLayer (:332-467), BertEncoder (:470-598), BertModel (:686-929), with
med.py's own signatures (its BertLayer / BertEncoder argument order differs from nlvr_encoder.py's) on top of the
same sm_100a kernels. BertLMHeadModel (:933-1094, the unpruned decoder used by generation) is out of scope.

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
                output_hidden_states=False, return_dict=True, mode='multimodal', space_dict=None, temperature=0):
        return _ne.BertEncoder.forward(self, hidden_states, attention_mask=attention_mask, space_dict=space_dict,
                                       temperature=temperature, head_mask=head_mask,
                                       encoder_hidden_states=encoder_hidden_states,
                                       encoder_attention_mask=encoder_attention_mask, past_key_values=past_key_values,
                                       use_cache=use_cache, output_attentions=output_attentions,
                                       output_hidden_states=output_hidden_states, return_dict=return_dict, mode=mode)


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
