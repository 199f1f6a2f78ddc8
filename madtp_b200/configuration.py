"""Minimal stand-in for transformers' BertConfig as the reference uses it (models/med.py, models/nlvr_encoder.py):
attribute access over the keys of configs/med_config.json, `from_json_file`, and the few defaults the encoder reads."""
from __future__ import annotations

import json

# configs/med_config.json of the reference (BERT-base with cross-attention, vocab 30522 + [DEC] + [ENC])
MED_CONFIG_DEFAULTS = dict(
    architectures=["BertModel"], attention_probs_dropout_prob=0.1, hidden_act="gelu", hidden_dropout_prob=0.1,
    hidden_size=768, initializer_range=0.02, intermediate_size=3072, layer_norm_eps=1e-12,
    max_position_embeddings=512, model_type="bert", num_attention_heads=12, num_hidden_layers=12, pad_token_id=0,
    type_vocab_size=2, vocab_size=30524, encoder_width=768, add_cross_attention=True,
)


class BertConfig:
    def __init__(self, **kw):
        vals = dict(MED_CONFIG_DEFAULTS)
        vals.update(kw)
        vals.setdefault("chunk_size_feed_forward", 0)
        vals.setdefault("position_embedding_type", "absolute")
        vals.setdefault("output_attentions", False)
        vals.setdefault("output_hidden_states", False)
        vals.setdefault("use_return_dict", True)
        vals.setdefault("evaluate", False)
        self.__dict__.update(vals)

    @classmethod
    def from_json_file(cls, path):
        with open(path) as f:
            return cls(**json.load(f))

    def to_dict(self):
        return dict(self.__dict__)
