"""CPU oracle for MADTP's pruned forward path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain PyTorch-CPU fp32 restatement of the reference algorithm (double125/MADTP), written functionally over a
state dict with the reference's key names. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; nothing under madtp_b200/ does.

Parity status: PINNED against the reference itself. The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so oracle/gen_golden.py imports the unmodified reference modules from /root/reference under
import shims (oracle/ref_shims.py), feeds both sides the same seeded weights and inputs, asserts agreement, and
writes the fixtures in tests/golden/. tests/test_oracle.py re-checks this module against those fixtures on every run.

Each function cites the reference lines it follows (paths relative to the reference tree).

One deliberate canonicalisation: the reference keeps survivors in the order `topk(sorted=False)` returns, which is
implementation-defined (differs between ATen CPU and CUDA). Every later layer is permutation-equivariant over the
non-CLS tokens, so the oracle emits survivors in ascending token order (what stream compaction produces) and tests
compare keep-masks as boolean vectors. The one place where order is observable in the reference -- the
nlvr_encoder mask gather by sorted index (models/nlvr_encoder.py:451-452) -- is restated literally.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ---------------------------------------------------------------------------------------------------------------
# small helpers
# ---------------------------------------------------------------------------------------------------------------
def linear(x: Tensor, sd: SD, prefix: str) -> Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def layer_norm(x: Tensor, sd: SD, prefix: str, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def split_heads(x: Tensor, H: int) -> Tensor:
    B, N, C = x.shape
    return x.view(B, N, H, C // H).permute(0, 2, 1, 3)


def merge_heads(x: Tensor) -> Tensor:
    B, H, N, D = x.shape
    return x.permute(0, 2, 1, 3).reshape(B, N, H * D)


@dataclass
class PruneTrace:
    """Per-layer diagnostics used by the parity tests (not part of the reference's outputs)."""
    n_in: int = 0                       # prunable tokens entering the layer
    k: int = 0                          # topk_num
    pruned: bool = False
    score: Optional[Tensor] = None      # [B, n] Importance_score
    threshold: Optional[Tensor] = None  # [B]
    count: Optional[Tensor] = None      # [B]
    keep: Optional[Tensor] = None       # [B, n] bool
    layer_input: Optional[Tensor] = None
    layer_output: Optional[Tensor] = None
    mask_in: Optional[Tensor] = None
    mask_out: Optional[Tensor] = None


# ---------------------------------------------------------------------------------------------------------------
# Query_model  (models/utils.py:147-183)
# ---------------------------------------------------------------------------------------------------------------
def query_model(ft: Tensor, space_dict: Tensor, sd_dim: int, q_map: Optional[SD] = None):
    """Returns (token_att [B,n,T] raw dot products, att_ft [B,T,d])."""
    q = ft if q_map is None else linear(ft, q_map, "q_map.0")          # utils.py:159-162 (map_func, CLIP only)
    dots = torch.matmul(q, space_dict.t().unsqueeze(0))                 # utils.py:164-170
    token_att = dots                                                    # returned un-scaled, utils.py:172-173
    weights = torch.softmax((dots / math.sqrt(sd_dim)).permute(0, 2, 1), dim=-1)  # over tokens, utils.py:174-177
    att_ft = torch.bmm(weights, q)                                      # utils.py:178
    return token_att, att_ft


# ---------------------------------------------------------------------------------------------------------------
# Attention statistics shared by every encoder
# ---------------------------------------------------------------------------------------------------------------
def cls_attention(probs: Tensor, ctx_heads: Tensor) -> Tensor:
    """vit.py:96-100 / nlvr_encoder.py:229-235 / med.py:229-235.  probs [B,H,N,N], ctx_heads [B,H,N,dh]."""
    cls_row = probs[:, :, 0, 1:]
    head_imp = ctx_heads[..., 1:, :].norm(dim=-1)
    head_imp = head_imp / (head_imp.sum(dim=1, keepdim=True) + 1e-8)
    return (cls_row * head_imp).sum(dim=1)


def importance_score(probs: Tensor, cls_attn: Tensor, token_attn: Tensor) -> Tensor:
    """vit.py:125-134 (identical in nlvr_encoder.py:404-414, med.py:348-358)."""
    a = probs[:, :, 1:, 1:].max(1)[0].sum(dim=1)
    a = a / (a.sum(dim=1, keepdim=True) + 1e-8)
    b = token_attn.max(2)[0]
    b = b / (b.sum(dim=1, keepdim=True) + 1e-8)
    return (a + b + cls_attn) / 3.0


def prune_decision(score: Tensor, token_attn: Tensor, temperature: float):
    """vit.py:137-146: threshold, per-row counts and the batch-max k."""
    w = torch.softmax(token_attn / temperature, dim=1).permute(0, 2, 1)      # [B,T,n]
    score_weight = torch.bmm(w, score.unsqueeze(-1))                         # [B,T,1]
    threshold = torch.min(score_weight, dim=1)[0]                            # [B,1]
    count = (score > threshold).sum(dim=1)
    return threshold.squeeze(-1), count, int(count.max().item())


def select_and_merge(x: Tensor, score: Tensor, k: int):
    """vit.py:153-161 with survivors in ascending token order. Returns (x_out [B,k+1,d], keep [B,n], order [B,n])."""
    B, n, d = x.shape
    order = torch.sort(score, dim=1, descending=True, stable=True)[1]        # full descending ranking (vit.py:155)
    keep = torch.zeros(B, n, dtype=torch.bool)
    keep.scatter_(1, order[:, :k], True)                                     # the top-k *set* (vit.py:153)
    idx = keep.nonzero()[:, 1].view(B, k)                                    # ascending token order
    x_topk = torch.gather(x, 1, idx.unsqueeze(-1).expand(B, k, d))           # utils.py:13-33
    tail_idx = order[:, k:]
    w = torch.gather(score, 1, tail_idx)
    w = w / (w.sum(dim=1, keepdim=True) + 1e-8)                              # vit.py:157-159
    x_tail = torch.gather(x, 1, tail_idx.unsqueeze(-1).expand(B, n - k, d))
    merged = torch.bmm(w.unsqueeze(1), x_tail)                               # vit.py:160
    return torch.cat([x_topk, merged], dim=1), keep, order


def reduce_token(x: Tensor, probs: Tensor, cls_attn: Tensor, token_attn: Tensor, temperature: float,
                 mask: Optional[Tensor] = None, variant: str = "vit", trace: Optional[PruneTrace] = None,
                 max_keep: int = 0):
    """Reduce_token of vit.py:123-163 (variant 'vit'), nlvr_encoder.py:400-454 ('nlvr'), med.py:345-391 ('med').

    x [B,n,d] prunable tokens; mask [B,n] additive (text only). Returns (x', mask')."""
    n = x.shape[1]
    score = importance_score(probs, cls_attn, token_attn)
    threshold, count, k = prune_decision(score, token_attn, temperature)
    if trace is not None:
        trace.n_in, trace.k, trace.score, trace.threshold, trace.count = n, k, score, threshold, count
    if k <= max_keep or n - k <= 1:                                          # vit.py:148-149; clip/model.py:220
        if trace is not None:
            trace.keep = torch.ones(x.shape[0], n, dtype=torch.bool)
        return x, mask
    x_out, keep, order = select_and_merge(x, score, k)
    if trace is not None:
        trace.keep, trace.pruned = keep, True
    if mask is not None:
        if variant == "nlvr":      # mask of the r-th ranked token lands in slot r, r = 0..k   (nlvr_encoder.py:451-452)
            mask = torch.gather(mask, 1, order[:, :k + 1])
        elif variant == "med":     # masks travel with their tokens; merged slot takes rank k's   (med.py:377-390)
            idx = keep.nonzero()[:, 1].view(x.shape[0], k)
            mask = torch.cat([torch.gather(mask, 1, idx), torch.gather(mask, 1, order[:, k:k + 1])], dim=1)
        else:
            raise ValueError(variant)
    return x_out, mask


# ---------------------------------------------------------------------------------------------------------------
# ViT  (models/vit.py)
# ---------------------------------------------------------------------------------------------------------------
def vit_attention(x: Tensor, sd: SD, prefix: str, H: int):
    """vit.py:75-103. Returns (projected output, probs [B,H,N,N], cls_attn [B,N-1])."""
    B, N, C = x.shape
    qkv = linear(x, sd, prefix + ".qkv").reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    probs = ((q @ k.transpose(-2, -1)) * (C // H) ** -0.5).softmax(dim=-1)
    ctx = probs @ v
    out = linear(merge_heads(ctx), sd, prefix + ".proj")
    return out, probs, cls_attention(probs, ctx)


def vit_block(x: Tensor, sd: SD, prefix: str, H: int, temperature: float = 0.0, token_attn: Optional[Tensor] = None,
              trace: Optional[PruneTrace] = None, eps: float = 1e-6) -> Tensor:
    """vit.py:183-207."""
    if trace is not None:
        trace.layer_input = x
    attn_out, probs, cls_attn = vit_attention(layer_norm(x, sd, prefix + ".norm1", eps), sd, prefix + ".attn", H)
    x = x + attn_out
    if temperature > 0:
        patches, _ = reduce_token(x[:, 1:, :], probs, cls_attn, token_attn, temperature, trace=trace)
        x = torch.cat([x[:, :1, :], patches], dim=1)
    h = F.gelu(linear(layer_norm(x, sd, prefix + ".norm2", eps), sd, prefix + ".mlp.fc1"))
    x = x + linear(h, sd, prefix + ".mlp.fc2")
    if trace is not None:
        trace.layer_output = x
    return x


def vit_embed(img: Tensor, sd: SD, prefix: str, patch: int = 16) -> Tensor:
    """timm 0.4.12 PatchEmbed + cls + pos (vit.py:283-289)."""
    x = F.conv2d(img, sd[prefix + "patch_embed.proj.weight"], sd[prefix + "patch_embed.proj.bias"], stride=patch)
    x = x.flatten(2).transpose(1, 2)
    cls = sd[prefix + "cls_token"].expand(img.shape[0], -1, -1)
    x = torch.cat([cls, x], dim=1)
    return x + sd[prefix + "pos_embed"][:, :x.size(1), :]


def vit_forward(img: Tensor, sd: SD, prefix: str, space_dict: Optional[Tensor], temperature: float, depth: int = 12,
                H: int = 12, traces: Optional[List[PruneTrace]] = None):
    """vit.py:281-310. Returns (tokens after the final norm, accumulated sd_img_ft)."""
    x = vit_embed(img, sd, prefix)
    sd_ft_all = None
    for i in range(depth):
        token_attn = None
        if space_dict is not None:
            token_attn, sd_ft = query_model(x[:, 1:, :], space_dict, space_dict.shape[1])
            sd_ft_all = sd_ft if sd_ft_all is None else sd_ft_all + sd_ft
        tr = None
        if traces is not None:
            tr = PruneTrace()
            traces.append(tr)
        x = vit_block(x, sd, f"{prefix}blocks.{i}", H, temperature if space_dict is not None else 0.0, token_attn, tr)
    return layer_norm(x, sd, prefix + "norm", 1e-6), sd_ft_all


# ---------------------------------------------------------------------------------------------------------------
# BERT text encoder with twin cross-attention  (models/nlvr_encoder.py) and the med.py single cross-attention
# ---------------------------------------------------------------------------------------------------------------
def bert_self_attention(h: Tensor, ext_mask: Optional[Tensor], sd: SD, prefix: str, H: int,
                        enc: Optional[Tensor] = None, enc_mask: Optional[Tensor] = None):
    """nlvr_encoder.py:142-237 / med.py:143-236. ext_mask, enc_mask are additive [B,1,1,L]. Returns
    (context [B,L,d], probs, cls_attn or None)."""
    q = split_heads(linear(h, sd, prefix + ".query"), H)
    src = h if enc is None else enc
    k = split_heads(linear(src, sd, prefix + ".key"), H)
    v = split_heads(linear(src, sd, prefix + ".value"), H)
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    m = ext_mask if enc is None else enc_mask
    if m is not None:
        scores = scores + m
    probs = torch.softmax(scores, dim=-1)
    ctx = torch.matmul(probs, v)
    cls_attn = cls_attention(probs, ctx) if enc is None else None
    return merge_heads(ctx), probs, cls_attn


def nlvr_layer(h: Tensor, ext_mask: Tensor, sd: SD, prefix: str, layer_num: int, enc: List[Tensor],
               temperature: float, token_attn: Tensor, H: int = 12, eps: float = 1e-12,
               trace: Optional[PruneTrace] = None):
    """nlvr_encoder.py:484-559. Returns (layer_output, pruned ext_mask)."""
    if trace is not None:
        trace.layer_input, trace.mask_in = h, ext_mask
    ctx, probs, cls_attn = bert_self_attention(h, ext_mask, sd, prefix + ".attention.self", H)
    att = layer_norm(linear(ctx, sd, prefix + ".attention.output.dense") + h, sd,
                     prefix + ".attention.output.LayerNorm", eps)                        # :240-271 (non-twin path)
    if temperature > 0:                                                                  # :519-533
        tokens, pm = reduce_token(att[:, 1:, :], probs, cls_attn, token_attn, temperature,
                                  mask=ext_mask[:, 0, 0, 1:], variant="nlvr", trace=trace)
        att = torch.cat([att[:, :1, :], tokens], dim=1)
        ext_mask = torch.cat([ext_mask[:, :, :, :1], pm[:, None, None, :]], dim=-1)
    # twin cross-attention (:535-545, :277-282, :258-270); encoder masks are all-ones -> additive zeros
    c0, _, _ = bert_self_attention(att, None, sd, prefix + ".crossattention.self0", H, enc=enc[0])
    c1, _, _ = bert_self_attention(att, None, sd, prefix + ".crossattention.self1", H, enc=enc[1])
    d0 = linear(c0, sd, prefix + ".crossattention.output.dense0")
    d1 = linear(c1, sd, prefix + ".crossattention.output.dense1")
    if layer_num >= 6:
        mixed = linear(torch.cat([d0, d1], dim=-1), sd, prefix + ".crossattention.output.merge_layer")
    else:
        mixed = (d0 + d1) / 2
    co = layer_norm(mixed + att, sd, prefix + ".crossattention.output.LayerNorm", eps)
    inter = F.gelu(linear(co, sd, prefix + ".intermediate.dense"))                        # :359-372
    out = layer_norm(linear(inter, sd, prefix + ".output.dense") + co, sd, prefix + ".output.LayerNorm", eps)
    if trace is not None:
        trace.layer_output, trace.mask_out = out, ext_mask
    return out, ext_mask


def bert_embeddings(ids: Tensor, sd: SD, prefix: str, eps: float = 1e-12) -> Tensor:
    """nlvr_encoder.py:61-85."""
    L = ids.shape[1]
    e = sd[prefix + "word_embeddings.weight"][ids] + sd[prefix + "position_embeddings.weight"][:L].unsqueeze(0)
    return layer_norm(e, sd, prefix + "LayerNorm", eps)


def nlvr_text_encoder(ids: Tensor, attn_mask: Tensor, sd: SD, prefix: str, enc: List[Tensor], space_dict: Tensor,
                      temperature: float, depth: int = 12, traces: Optional[List[PruneTrace]] = None):
    """nlvr_encoder.py:874-1015 + :570-687 (mode='multimodal'). Returns (last_hidden_state, sd_txt_ft)."""
    ext_mask = (1.0 - attn_mask[:, None, None, :].to(torch.float32)) * -10000.0          # :870-871
    h = bert_embeddings(ids, sd, prefix + "embeddings.")
    sd_ft_all = None
    for i in range(depth):
        token_attn, sd_ft = query_model(h[:, 1:, :], space_dict, space_dict.shape[1])    # :602-608
        sd_ft_all = sd_ft if sd_ft_all is None else sd_ft_all + sd_ft
        tr = None
        if traces is not None:
            tr = PruneTrace()
            traces.append(tr)
        h, ext_mask = nlvr_layer(h, ext_mask, sd, f"{prefix}encoder.layer.{i}", i, enc, temperature, token_attn,
                                 trace=tr)
    return h, sd_ft_all


# ---------------------------------------------------------------------------------------------------------------
# BLIP-NLVR eval forward  (models/blip_nlvr.py:63-81,99-100)
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class NlvrTrace:
    vit: List[PruneTrace] = field(default_factory=list)
    text: List[PruneTrace] = field(default_factory=list)
    image_embeds: Optional[Tensor] = None
    last_hidden: Optional[Tensor] = None
    sd_img_ft: Optional[Tensor] = None
    sd_txt_ft: Optional[Tensor] = None


def blip_nlvr_forward(images: Tensor, input_ids: Tensor, attn_mask: Tensor, sd: SD, temperature: float,
                      enc_token_id: int = 30523, trace: Optional[NlvrTrace] = None) -> Tensor:
    """images [2P,3,H,W] = cat(image0, image1); input_ids/attn_mask [P,L]. Returns prediction [P,2]."""
    space_dict = sd["space_dict"]
    emb, sd_img = vit_forward(images, sd, "visual_encoder.", space_dict, temperature,
                              traces=None if trace is None else trace.vit)
    P = input_ids.shape[0]
    img0, img1 = emb[:P], emb[P:]                                                        # blip_nlvr.py:67
    ids = input_ids.clone()
    ids[:, 0] = enc_token_id                                                             # blip_nlvr.py:69
    hidden, sd_txt = nlvr_text_encoder(ids, attn_mask, sd, "text_encoder.", [img0, img1], space_dict, temperature,
                                       traces=None if trace is None else trace.text)
    cls = hidden[:, 0, :]
    pred = linear(torch.relu(linear(cls, sd, "cls_head.0")), sd, "cls_head.2")           # blip_nlvr.py:56-60,80-81
    if trace is not None:
        trace.image_embeds, trace.last_hidden, trace.sd_img_ft, trace.sd_txt_ft = emb, hidden, sd_img, sd_txt
    return pred


# ---------------------------------------------------------------------------------------------------------------
# CLIP  (clip/model.py ResidualAttentionBlock / Transformer / VisionTransformer / encode_text, clip/mock.py MHA)
# ---------------------------------------------------------------------------------------------------------------
def clip_block(x: Tensor, sd: SD, prefix: str, H: int, space_dict: Optional[Tensor], temperature: float,
               sd_ft_all: Optional[Tensor], max_keep: int = 1, causal: bool = False,
               trace: Optional[PruneTrace] = None, eps: float = 1e-5):
    """clip/model.py:236-261 on x [B,N,C] (the reference carries [N,B,C]; the math is layout-free).
    Returns (x', sd_ft_all)."""
    B, N, C = x.shape
    dh = C // H
    if trace is not None:
        trace.layer_input = x
    token_attn = None
    if space_dict is not None:                                                           # :239-245
        qm = {"q_map.0.weight": sd[prefix + ".query_model.q_map.0.weight"],
              "q_map.0.bias": sd[prefix + ".query_model.q_map.0.bias"]}
        token_attn, sd_ft = query_model(x[:, 1:, :], space_dict, space_dict.shape[1], q_map=qm)
        sd_ft_all = sd_ft if sd_ft_all is None else sd_ft_all + sd_ft
    y = layer_norm(x, sd, prefix + ".ln_1", eps)
    qkv = F.linear(y, sd[prefix + ".attn.in_proj_weight"], sd[prefix + ".attn.in_proj_bias"])
    q, k, v = (split_heads(t, H) for t in qkv.chunk(3, dim=-1))
    q = q / math.sqrt(dh)                                                                # torch 1.11 _scaled_dot_product_attention
    scores = q @ k.transpose(-1, -2)
    if causal:                                                                           # :452-457, mock.py:309-310
        scores = scores + torch.full((N, N), float("-inf")).triu_(1)
    probs = torch.softmax(scores, dim=-1)
    ctx = probs @ v
    x = x + linear(merge_heads(ctx), sd, prefix + ".attn.out_proj")
    if space_dict is not None and temperature > 0:                                       # :254-258
        cls_attn = cls_attention(probs, ctx)                                             # mock.py:225-232
        patches, _ = reduce_token(x[:, 1:, :], probs, cls_attn, token_attn, temperature, trace=trace,
                                  max_keep=int(max_keep))
        x = torch.cat([x[:, :1, :], patches], dim=1)
    h = linear(layer_norm(x, sd, prefix + ".ln_2", eps), sd, prefix + ".mlp.c_fc")
    x = x + linear(h * torch.sigmoid(1.702 * h), sd, prefix + ".mlp.c_proj")             # QuickGELU :168-170
    if trace is not None:
        trace.layer_output = x
    return x, sd_ft_all


def clip_vision_forward(img: Tensor, sd: SD, prefix: str, space_dict: Optional[Tensor], temperature: float,
                        layers: int, H: int, patch: int = 16, traces: Optional[List[PruneTrace]] = None):
    """clip/model.py:292-313. Returns (image embedding [B, output_dim], sd_img_ft)."""
    x = F.conv2d(img, sd[prefix + "conv1.weight"], None, stride=patch).flatten(2).transpose(1, 2)
    cls = sd[prefix + "class_embedding"] + torch.zeros(x.shape[0], 1, x.shape[-1])
    x = torch.cat([cls, x], dim=1) + sd[prefix + "positional_embedding"]
    x = layer_norm(x, sd, prefix + "ln_pre", 1e-5)
    sd_ft_all = None
    for i in range(layers):
        tr = None
        if traces is not None:
            tr = PruneTrace()
            traces.append(tr)
        x, sd_ft_all = clip_block(x, sd, f"{prefix}transformer.resblocks.{i}", H, space_dict, temperature, sd_ft_all,
                                  max_keep=1, causal=False, trace=tr)
    x = layer_norm(x[:, 0, :], sd, prefix + "ln_post", 1e-5)
    return x @ sd[prefix + "proj"], sd_ft_all


def clip_text_forward(text: Tensor, sd: SD, space_dict: Optional[Tensor], temperature: float, layers: int, H: int,
                      traces: Optional[List[PruneTrace]] = None):
    """clip/model.py:489-503 (encode_text). Returns (text embedding, sd_txt_ft). After a prune the EOT token is read at
    its ORIGINAL position in the pruned sequence, as the reference does (:501)."""
    x = sd["token_embedding.weight"][text] + sd["positional_embedding"]
    max_keep = int(text.argmax(dim=-1).max()) + 2                                        # :492
    sd_ft_all = None
    for i in range(layers):
        tr = None
        if traces is not None:
            tr = PruneTrace()
            traces.append(tr)
        x, sd_ft_all = clip_block(x, sd, f"transformer.resblocks.{i}", H, space_dict, temperature, sd_ft_all,
                                  max_keep=max_keep, causal=True, trace=tr)
    x = layer_norm(x, sd, "ln_final", 1e-5)
    x = x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ sd["text_projection"]
    return x, sd_ft_all


# ---------------------------------------------------------------------------------------------------------------
# Analytic MAC model and temperature calibration
# ---------------------------------------------------------------------------------------------------------------
def vit_layer_macs(n_in: int, n_out: int, d: int = 768, dff: int = 3072, T: int = 100) -> int:
    """MACs of one ViT layer per image: N_in tokens through attention, N_out through the FFN (both incl. CLS)."""
    return 4 * n_in * d * d + 2 * n_in * n_in * d + 2 * n_out * d * dff + 2 * (n_in - 1) * d * T


def vit_macs_from_traces(traces: List[PruneTrace], n0: int, d: int = 768, patch_k: int = 768) -> int:
    total = (n0 - 1) * patch_k * d
    n_in = n0
    for tr in traces:
        n_out = (tr.k + 2) if tr.pruned else n_in
        total += vit_layer_macs(n_in, n_out, d)
        n_in = n_out
    return total


def text_layer_macs(L_in: int, L_out: int, n_img: int, layer_num: int, d: int = 768, dff: int = 3072,
                    T: int = 100) -> int:
    """MACs of one NLVR text layer per sample: L_in tokens through self-attention, L_out
    through the twin cross-attention (keys/values = n_img image tokens, two images) and the FFN."""
    macs = 4 * L_in * d * d + 2 * L_in * L_in * d + 2 * (L_in - 1) * d * T
    macs += 2 * (2 * L_out * d * d + 2 * n_img * d * d + 2 * L_out * n_img * d)
    if layer_num >= 6:
        macs += 2 * L_out * d * d
    return macs + 2 * L_out * d * dff


def nlvr_macs_from_trace(trace: "NlvrTrace", n0: int, text_len: int, d: int = 768) -> int:
    """MACs of one BLIP-NLVR sample (two images + one sentence) along the oracle's pruning trajectory."""
    total = 2 * vit_macs_from_traces(trace.vit, n0, d)
    n_img = trace.image_embeds.shape[1]
    L = text_len
    for i, tr in enumerate(trace.text):
        L_out = (tr.k + 2) if tr.pruned else L
        total += text_layer_macs(L, L_out, n_img, i, d)
        L = L_out
    return total + d * d + 2 * d


def nlvr_macs_unpruned(n0: int, text_len: int, d: int = 768, depth: int = 12) -> int:
    total = 2 * ((n0 - 1) * 768 * d + depth * vit_layer_macs(n0, n0, d))
    for i in range(depth):
        total += text_layer_macs(text_len, text_len, n0, i, d)
    return total + d * d + 2 * d


# ---------------------------------------------------------------------------------------------------------------
# med.py text encoder (BLIP retrieval / VQA): single cross-attention, mode 'text' | 'multimodal'
# ---------------------------------------------------------------------------------------------------------------
def med_layer(h: Tensor, ext_mask: Tensor, sd: SD, prefix: str, enc: Optional[Tensor], temperature: float,
              token_attn: Optional[Tensor], mode: str, H: int = 12, eps: float = 1e-12,
              trace: Optional[PruneTrace] = None):
    """models/med.py:393-467. Returns (layer_output, pruned ext_mask)."""
    if trace is not None:
        trace.layer_input, trace.mask_in = h, ext_mask
    ctx, probs, cls_attn = bert_self_attention(h, ext_mask, sd, prefix + ".attention.self", H)
    att = layer_norm(linear(ctx, sd, prefix + ".attention.output.dense") + h, sd,
                     prefix + ".attention.output.LayerNorm", eps)
    if temperature > 0:                                                                  # :427-441
        tokens, pm = reduce_token(att[:, 1:, :], probs, cls_attn, token_attn, temperature,
                                  mask=ext_mask[:, 0, 0, 1:], variant="med", trace=trace)
        att = torch.cat([att[:, :1, :], tokens], dim=1)
        ext_mask = torch.cat([ext_mask[:, :, :, :1], pm[:, None, None, :]], dim=-1)
    if mode == "multimodal":                                                             # :443-455; no mask (:197)
        c, _, _ = bert_self_attention(att, None, sd, prefix + ".crossattention.self", H, enc=enc, enc_mask=None)
        att = layer_norm(linear(c, sd, prefix + ".crossattention.output.dense") + att, sd,
                         prefix + ".crossattention.output.LayerNorm", eps)
    inter = F.gelu(linear(att, sd, prefix + ".intermediate.dense"))
    out = layer_norm(linear(inter, sd, prefix + ".output.dense") + att, sd, prefix + ".output.LayerNorm", eps)
    if trace is not None:
        trace.layer_output, trace.mask_out = out, ext_mask
    return out, ext_mask


def med_text_encoder(ids: Tensor, attn_mask: Tensor, sd: SD, prefix: str, enc: Optional[Tensor],
                     space_dict: Optional[Tensor], temperature: float, mode: str, depth: int = 12,
                     traces: Optional[List[PruneTrace]] = None):
    """models/med.py:788-929 + :478-598. Returns (last_hidden_state, sd_txt_ft)."""
    ext_mask = (1.0 - attn_mask[:, None, None, :].to(torch.float32)) * -10000.0          # :784-785
    h = bert_embeddings(ids, sd, prefix + "embeddings.")
    sd_ft_all = None
    for i in range(depth):
        token_attn = None
        if space_dict is not None:                                                       # :513-524
            token_attn, sd_ft = query_model(h[:, 1:, :], space_dict, space_dict.shape[1])
            if sd_ft_all is None:
                sd_ft_all = sd_ft
            elif sd_ft_all.shape == sd_ft.shape:
                sd_ft_all = sd_ft_all + sd_ft
            else:
                sd_ft_all = None
        tr = None
        if traces is not None:
            tr = PruneTrace()
            traces.append(tr)
        h, ext_mask = med_layer(h, ext_mask, sd, f"{prefix}encoder.layer.{i}", enc, temperature, token_attn, mode,
                                trace=tr)
    return h, sd_ft_all


# ---------------------------------------------------------------------------------------------------------------
# VQA answer ranking  (models/blip_vqa.py:156-203 over models/med.py:BertLMHeadModel)
# ---------------------------------------------------------------------------------------------------------------
def med_decoder_forward(ids: Tensor, attn_mask: Tensor, sd: SD, prefix: str, enc: Tensor, depth: int = 12) -> Tensor:
    """BertModel(is_decoder=True) without pruning: causal x padding mask (models/med.py:749-771), cross-attention to
    `enc` with NO mask (:197). Returns the last hidden state [B, L, d]."""
    B, L = ids.shape
    seq = torch.arange(L)
    causal = (seq[None, None, :].repeat(B, L, 1) <= seq[None, :, None]).to(torch.float32)     # :753-756
    ext = causal[:, None, :, :] * attn_mask[:, None, None, :].to(torch.float32)               # :768
    ext = (1.0 - ext) * -10000.0                                                              # :784-785
    h = bert_embeddings(ids, sd, prefix + "embeddings.")
    for i in range(depth):
        h, _ = med_layer(h, ext, sd, f"{prefix}encoder.layer.{i}", enc, 0.0, None, "multimodal")
    return h


def lm_head(h: Tensor, sd: SD, prefix: str, eps: float = 1e-12) -> Tensor:
    """BertOnlyMLMHead (models/med.py:615-657): dense -> GELU(erf) -> LayerNorm -> tied decoder + bias."""
    t = layer_norm(F.gelu(linear(h, sd, prefix + ".transform.dense")), sd, prefix + ".transform.LayerNorm", eps)
    return F.linear(t, sd[prefix + ".decoder.weight"], sd[prefix + ".bias"])


def lm_sequence_loss(logits: Tensor, labels: Tensor) -> Tensor:
    """models/med.py:1040-1047 with reduction='none': label-smoothed (0.1) next-token cross entropy, ignore_index -100,
    summed over the positions of every sequence."""
    B = logits.shape[0]
    shifted = logits[:, :-1, :].contiguous()
    lab = labels[:, 1:].contiguous()
    loss = F.cross_entropy(shifted.view(-1, shifted.shape[-1]), lab.view(-1), reduction="none", label_smoothing=0.1)
    return loss.view(B, -1).sum(1)


def vqa_rank_answer(question_states: Tensor, answer_ids: Tensor, answer_atts: Tensor, k: int, sd: SD,
                    prefix: str = "text_decoder.", pad_id: int = 0):
    """models/blip_vqa.py:156-203. Returns (max_ids [Q], topk_ids [Q,k], log_probs_sum [Q,k], prob_first_token)."""
    Q = question_states.shape[0]
    start_ids = answer_ids[0, 0].repeat(Q, 1)
    h0 = med_decoder_forward(start_ids, torch.ones_like(start_ids), sd, prefix + "bert.", question_states)
    logits = lm_head(h0, sd, prefix + "cls.predictions")[:, 0, :]
    prob_first = torch.softmax(logits, dim=1).index_select(1, answer_ids[:, 1])
    topk_probs, topk_ids = prob_first.topk(k, dim=1)
    input_ids = torch.cat([answer_ids.index_select(0, t) for t in topk_ids], 0)
    input_atts = torch.cat([answer_atts.index_select(0, t) for t in topk_ids], 0)
    targets = input_ids.masked_fill(input_ids == pad_id, -100)
    qs = question_states.repeat_interleave(k, dim=0)                 # == tile(question_states, 0, k), :214-220
    h = med_decoder_forward(input_ids, input_atts, sd, prefix + "bert.", qs)
    loss = lm_sequence_loss(lm_head(h, sd, prefix + "cls.predictions"), targets)
    log_probs_sum = (-loss).view(Q, k)
    best = log_probs_sum.argmax(dim=1)
    max_ids = topk_ids[best >= 0, best]
    return max_ids, topk_ids, log_probs_sum, prob_first
