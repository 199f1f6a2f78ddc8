"""Deterministic synthetic weights and inputs (plain seeded tensors; no algorithm of the reference lives here).

There is no network, so neither released checkpoints nor datasets exist here or on the GPU box. The oracle, the
parity tests, bench.py and the CUDA path are all driven from a state dict generated on the CPU from a fixed seed, with the reference's key names
and shapes (SURVEY.md section 8b "State-dict keys") and the distributions the reference modules get by default
construction with evaluate=True (nn.Linear: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias; LayerNorm 1/0;
nn.Embedding N(0,1); space_dict = randn, models/blip_nlvr.py:46). cls_token / pos_embed are zeros in the reference
under evaluate=True (models/vit.py:244-245,257-260); a small N(0, 0.02) is used here so positions are not degenerate.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

Tensor = torch.Tensor


def _linear(sd: Dict[str, Tensor], g: torch.Generator, name: str, out_f: int, in_f: int, bias: bool = True):
    bound = 1.0 / math.sqrt(in_f)
    sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
    if bias:
        sd[name + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * bound


def _ln(sd: Dict[str, Tensor], name: str, d: int):
    sd[name + ".weight"] = torch.ones(d)
    sd[name + ".bias"] = torch.zeros(d)


def vit_state_dict(g: torch.Generator, prefix: str = "", img_size: int = 384, patch: int = 16, d: int = 768,
                   depth: int = 12, mlp_ratio: int = 4) -> Dict[str, Tensor]:
    """Keys of models/vit.py:VisionTransformer (timm naming)."""
    sd: Dict[str, Tensor] = {}
    n = (img_size // patch) ** 2
    sd[prefix + "cls_token"] = torch.randn(1, 1, d, generator=g) * 0.02
    sd[prefix + "pos_embed"] = torch.randn(1, n + 1, d, generator=g) * 0.02
    bound = 1.0 / math.sqrt(3 * patch * patch)
    sd[prefix + "patch_embed.proj.weight"] = (torch.rand(d, 3, patch, patch, generator=g) * 2 - 1) * bound
    sd[prefix + "patch_embed.proj.bias"] = (torch.rand(d, generator=g) * 2 - 1) * bound
    for i in range(depth):
        p = f"{prefix}blocks.{i}"
        _ln(sd, p + ".norm1", d)
        _linear(sd, g, p + ".attn.qkv", 3 * d, d)
        _linear(sd, g, p + ".attn.proj", d, d)
        _ln(sd, p + ".norm2", d)
        _linear(sd, g, p + ".mlp.fc1", mlp_ratio * d, d)
        _linear(sd, g, p + ".mlp.fc2", d, mlp_ratio * d)
    _ln(sd, prefix + "norm", d)
    return sd


def block_state_dict(seed: int = 1234, d: int = 768, mlp_ratio: int = 4) -> Dict[str, Tensor]:
    """One models/vit.py:Block (BASELINE config 1)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    _ln(sd, "norm1", d)
    _linear(sd, g, "attn.qkv", 3 * d, d)
    _linear(sd, g, "attn.proj", d, d)
    _ln(sd, "norm2", d)
    _linear(sd, g, "mlp.fc1", mlp_ratio * d, d)
    _linear(sd, g, "mlp.fc2", d, mlp_ratio * d)
    return sd


def nlvr_text_state_dict(g: torch.Generator, prefix: str = "", d: int = 768, depth: int = 12, dff: int = 3072,
                         vocab: int = 30524, max_pos: int = 512) -> Dict[str, Tensor]:
    """Keys of models/nlvr_encoder.py:BertModel(add_pooling_layer=False)."""
    sd: Dict[str, Tensor] = {}
    e = prefix + "embeddings."
    sd[e + "word_embeddings.weight"] = torch.randn(vocab, d, generator=g)
    sd[e + "word_embeddings.weight"][0].zero_()        # padding_idx = 0
    sd[e + "position_embeddings.weight"] = torch.randn(max_pos, d, generator=g)
    sd[e + "position_ids"] = torch.arange(max_pos).unsqueeze(0)
    _ln(sd, e + "LayerNorm", d)
    for i in range(depth):
        p = f"{prefix}encoder.layer.{i}"
        for nm in ("query", "key", "value"):
            _linear(sd, g, f"{p}.attention.self.{nm}", d, d)
        _linear(sd, g, p + ".attention.output.dense", d, d)
        _ln(sd, p + ".attention.output.LayerNorm", d)
        for s in ("self0", "self1"):
            for nm in ("query", "key", "value"):
                _linear(sd, g, f"{p}.crossattention.{s}.{nm}", d, d)
        _linear(sd, g, p + ".crossattention.output.dense0", d, d)
        _linear(sd, g, p + ".crossattention.output.dense1", d, d)
        if i >= 6:
            _linear(sd, g, p + ".crossattention.output.merge_layer", d, 2 * d)
        _ln(sd, p + ".crossattention.output.LayerNorm", d)
        _linear(sd, g, p + ".intermediate.dense", dff, d)
        _linear(sd, g, p + ".output.dense", d, dff)
        _ln(sd, p + ".output.LayerNorm", d)
    return sd


def med_text_state_dict(g: torch.Generator, prefix: str = "", d: int = 768, depth: int = 12, dff: int = 3072,
                        vocab: int = 30524, max_pos: int = 512) -> Dict[str, Tensor]:
    """Keys of models/med.py:BertModel(add_pooling_layer=False) with add_cross_attention=True."""
    sd: Dict[str, Tensor] = {}
    e = prefix + "embeddings."
    sd[e + "word_embeddings.weight"] = torch.randn(vocab, d, generator=g)
    sd[e + "word_embeddings.weight"][0].zero_()
    sd[e + "position_embeddings.weight"] = torch.randn(max_pos, d, generator=g)
    sd[e + "position_ids"] = torch.arange(max_pos).unsqueeze(0)
    _ln(sd, e + "LayerNorm", d)
    for i in range(depth):
        p = f"{prefix}encoder.layer.{i}"
        for blk in ("attention", "crossattention"):
            for nm in ("query", "key", "value"):
                _linear(sd, g, f"{p}.{blk}.self.{nm}", d, d)
            _linear(sd, g, f"{p}.{blk}.output.dense", d, d)
            _ln(sd, f"{p}.{blk}.output.LayerNorm", d)
        _linear(sd, g, p + ".intermediate.dense", dff, d)
        _linear(sd, g, p + ".output.dense", d, dff)
        _ln(sd, p + ".output.LayerNorm", d)
    return sd


def retrieval_state_dict(seed: int = 4321, img_size: int = 384, sd_num: int = 100, sd_dim: int = 768,
                         depth: int = 12) -> Dict[str, Tensor]:
    """The encoder part of models/blip_retrieval.py:BLIP_Retrieval: space_dict, visual_encoder.*, text_encoder.*,
    itm_head.* (the projections / momentum twins / queues are training-only and not generated)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {"space_dict": torch.randn(sd_num, sd_dim, generator=g)}
    sd.update(vit_state_dict(g, "visual_encoder.", img_size=img_size, depth=depth))
    sd.update(med_text_state_dict(g, "text_encoder.", depth=depth))
    _linear(sd, g, "itm_head", 2, 768)
    return sd


def vqa_state_dict(seed: int = 99, img_size: int = 480, sd_num: int = 100, sd_dim: int = 768,
                   depth: int = 12) -> Dict[str, Tensor]:
    """models/blip_vqa.py:BLIP_VQA: space_dict, visual_encoder.*, text_encoder.* and the answer decoder
    text_decoder.{bert.*, cls.predictions.*} (models/med.py:BertLMHeadModel; the output embedding is tied to the input
    embedding, models/med.py:947-950)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {"space_dict": torch.randn(sd_num, sd_dim, generator=g)}
    sd.update(vit_state_dict(g, "visual_encoder.", img_size=img_size, depth=depth))
    sd.update(med_text_state_dict(g, "text_encoder.", depth=depth))
    sd.update(med_text_state_dict(g, "text_decoder.bert.", depth=depth))
    # small-norm embeddings keep the tied LM head's logits in a softmax regime that still separates candidates
    sd["text_decoder.bert.embeddings.word_embeddings.weight"] *= 0.05
    p = "text_decoder.cls.predictions"
    _linear(sd, g, p + ".transform.dense", 768, 768)
    _ln(sd, p + ".transform.LayerNorm", 768)
    sd[p + ".decoder.weight"] = sd["text_decoder.bert.embeddings.word_embeddings.weight"]
    sd[p + ".bias"] = (torch.rand(sd[p + ".decoder.weight"].shape[0], generator=g) * 2 - 1) * 0.02
    sd[p + ".decoder.bias"] = sd[p + ".bias"]
    return sd


def vqa_answer_candidates(n_answers: int = 6, max_len: int = 5, seed: int = 0, bos_id: int = 30522):
    """Tokenised answer list as the VQA driver builds it (`padding='longest'`, first token = [DEC] bos): returns
    (input_ids [n, L] long, attention_mask [n, L] long); every answer ends with [SEP] = 102, pads are 0."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros(n_answers, max_len, dtype=torch.long)
    mask = torch.zeros(n_answers, max_len, dtype=torch.long)
    for a in range(n_answers):
        n_tok = int(torch.randint(1, max_len - 1, (1,), generator=g))       # answer words
        ids[a, 0] = bos_id
        ids[a, 1:1 + n_tok] = torch.randint(1000, 30000, (n_tok,), generator=g)
        ids[a, 1 + n_tok] = 102
        mask[a, :2 + n_tok] = 1
    return ids, mask


def retrieval_inputs(batch: int, img_size: int = 384, max_len: int = 35, seed: int = 0):
    """BASELINE config 3 inputs: images ~ N(0,1); text padded to max_len (`padding='max_length'`,
    models/blip_retrieval.py:107) with true lengths ~ U[8, 20)."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, img_size, img_size, generator=g)
    lens = torch.randint(8, 20, (batch,), generator=g)
    ids = torch.zeros(batch, max_len, dtype=torch.long)
    mask = torch.zeros(batch, max_len, dtype=torch.long)
    for b in range(batch):
        ids[b, :lens[b]] = torch.randint(1000, 30000, (int(lens[b]),), generator=g)
        ids[b, 0] = 101
        mask[b, :lens[b]] = 1
    return images, ids, mask


def clip_block_state_dict(sd: Dict[str, Tensor], g: torch.Generator, p: str, d: int, sd_dim: int):
    """Keys of clip/model.py:ResidualAttentionBlock (patched nn.MultiheadAttention + per-block Query_model.q_map)."""
    bound = 1.0 / math.sqrt(d)
    sd[p + ".attn.in_proj_weight"] = (torch.rand(3 * d, d, generator=g) * 2 - 1) * bound
    sd[p + ".attn.in_proj_bias"] = (torch.rand(3 * d, generator=g) * 2 - 1) * bound
    _linear(sd, g, p + ".attn.out_proj", d, d)
    _ln(sd, p + ".ln_1", d)
    _linear(sd, g, p + ".mlp.c_fc", 4 * d, d)
    _linear(sd, g, p + ".mlp.c_proj", d, 4 * d)
    _ln(sd, p + ".ln_2", d)
    _linear(sd, g, p + ".query_model.q_map.0", sd_dim, d)


def clip_state_dict(seed: int = 777, img_size: int = 224, patch: int = 16, vision_width: int = 768,
                    vision_layers: int = 12, text_width: int = 512, text_layers: int = 12, embed_dim: int = 512,
                    context: int = 77, vocab: int = 49408, sd_num: int = 100, sd_dim: int = 768) -> Dict[str, Tensor]:
    """The encoder part of clip/model.py:CLIP with ViT-B/16 shapes: space_dict, visual.*, transformer.*,
    token_embedding, positional_embedding, ln_final, text_projection."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {"space_dict": torch.randn(sd_num, sd_dim, generator=g)}
    scale = vision_width ** -0.5
    n = (img_size // patch) ** 2
    sd["visual.conv1.weight"] = (torch.rand(vision_width, 3, patch, patch, generator=g) * 2 - 1) / math.sqrt(3 * patch * patch)
    sd["visual.class_embedding"] = scale * torch.randn(vision_width, generator=g)
    sd["visual.positional_embedding"] = scale * torch.randn(n + 1, vision_width, generator=g)
    _ln(sd, "visual.ln_pre", vision_width)
    _ln(sd, "visual.ln_post", vision_width)
    sd["visual.proj"] = scale * torch.randn(vision_width, embed_dim, generator=g)
    for i in range(vision_layers):
        clip_block_state_dict(sd, g, f"visual.transformer.resblocks.{i}", vision_width, sd_dim)
    sd["token_embedding.weight"] = torch.randn(vocab, text_width, generator=g) * 0.02
    sd["positional_embedding"] = torch.randn(context, text_width, generator=g) * 0.01
    _ln(sd, "ln_final", text_width)
    sd["text_projection"] = torch.randn(text_width, embed_dim, generator=g) * text_width ** -0.5
    for i in range(text_layers):
        clip_block_state_dict(sd, g, f"transformer.resblocks.{i}", text_width, sd_dim)
    return sd


def clip_inputs(batch: int, img_size: int = 224, context: int = 77, vocab: int = 49408, seed: int = 0):
    """images ~ N(0,1); token ids: SOT (49406), random words, EOT (49407 = the arg-max id, clip/clip.py:235-239), zero pad."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, img_size, img_size, generator=g)
    text = torch.zeros(batch, context, dtype=torch.long)
    lens = torch.randint(6, 30, (batch,), generator=g)
    for b in range(batch):
        text[b, 0] = vocab - 2
        text[b, 1:lens[b]] = torch.randint(1000, 40000, (int(lens[b]) - 1,), generator=g)
        text[b, lens[b]] = vocab - 1
    return images, text


def blip_nlvr_state_dict(seed: int = 1234, img_size: int = 384, sd_num: int = 100, sd_dim: int = 768,
                         depth: int = 12) -> Dict[str, Tensor]:
    """Keys of models/blip_nlvr.py:BLIP_NLVR."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {"space_dict": torch.randn(sd_num, sd_dim, generator=g)}
    sd.update(vit_state_dict(g, "visual_encoder.", img_size=img_size, depth=depth))
    sd.update(nlvr_text_state_dict(g, "text_encoder.", depth=depth))
    _linear(sd, g, "cls_head.0", 768, 768)
    _linear(sd, g, "cls_head.2", 2, 768)
    return sd


def nlvr_inputs(pairs: int, img_size: int = 384, text_len: int = 20, seed: int = 0, pad_to: int = 0):
    """Synthetic batch: images = cat(image0, image1) ~ N(0,1); ids ~ U[1000, 30000), ids[:,0]=101.

    pad_to > text_len appends pad tokens (id 0, mask 0) with per-row true lengths in [text_len//2, text_len]."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(2 * pairs, 3, img_size, img_size, generator=g)
    ids = torch.randint(1000, 30000, (pairs, text_len), generator=g)
    ids[:, 0] = 101
    mask = torch.ones(pairs, text_len, dtype=torch.long)
    if pad_to > text_len:
        lens = torch.randint(max(2, text_len // 2), text_len + 1, (pairs,), generator=g)
        lens[0] = text_len
        full_ids = torch.zeros(pairs, pad_to, dtype=torch.long)
        full_mask = torch.zeros(pairs, pad_to, dtype=torch.long)
        for b in range(pairs):
            full_ids[b, :lens[b]] = ids[b, :lens[b]]
            full_mask[b, :lens[b]] = 1
        ids, mask = full_ids, full_mask
    return images, ids, mask


def block_inputs(seed: int = 0, B: int = 2, N: int = 197, d: int = 768, T: int = 100):
    """BASELINE config 1 inputs: layer input x [B,N,d] and codebook [T,d], both ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, N, d, generator=g), torch.randn(T, d, generator=g)


def tensor_digest(*tensors) -> str:
    """sha256 prefix of the raw bytes -- fixtures store it so a drifting RNG is detected instead of mis-compared."""
    import hashlib
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()[:16]
