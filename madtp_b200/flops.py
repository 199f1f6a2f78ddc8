"""Closed-form MAC accounting from the pruning trajectory and the p -> temperature calibration built on it.

The reference spends a second, fvcore-traced forward on every evaluation batch just to count MACs
(compress_nlvr_dtp.py:93-98) and steers `temperature` towards `Ori_Gflops * (1 - p)` with a per-epoch controller
(compress_nlvr_dtp.py:162-201). The select kernel already returns the per-layer survivor count `k`, so the same number
is a closed form of the trajectory (fvcore counts one MAC as one "flop", so every GFLOPs
figure in the reference is GMACs):

    ViT layer      4 N_in d^2 + 2 N_in^2 d + 2 N_out d dff + 2 (N_in - 1) d T        (attention on N_in, FFN on N_out)
    NLVR text      4 L_in d^2 + 2 L_in^2 d + 2 (L_in - 1) d T
                   + 2 [2 L_out d^2 + 2 N_img d^2 + 2 L_out N_img d] (+ 2 L_out d^2 merge_layer, layers >= 6)
                   + 2 L_out d dff
    med.py text    the same with ONE cross-attention branch in mode 'multimodal', none in mode 'text'

Unpruned BLIP-NLVR at 384 x 384 with a 22-token sentence comes to 132.6 GMAC against the reference's hard-coded
132.54 (compress_nlvr_dtp.py:162).
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence, Tuple

D, DFF, T_BOOK = 768, 3072, 100


def trajectory(n0: int, ks: Iterable[int]) -> List[Tuple[int, int]]:
    """[(N_in, N_out)] per layer from the entry token count n0 (incl. CLS) and the per-layer topk_num (k < 0 or None:
    the layer did not prune). A pruned layer leaves k survivors + the merged token + CLS."""
    out, n = [], n0
    for k in ks:
        n_out = n if (k is None or k < 0) else k + 2
        out.append((n, n_out))
        n = n_out
    return out


def vit_layer_macs(n_in: int, n_out: int, d: int = D, dff: int = DFF, T: int = T_BOOK) -> int:
    return 4 * n_in * d * d + 2 * n_in * n_in * d + 2 * n_out * d * dff + 2 * (n_in - 1) * d * T


def vit_macs(n0: int, ks: Iterable[int], d: int = D, patch_k: int = 768, dff: int = DFF, T: int = T_BOOK) -> int:
    """One image through the pruned ViT: patch embedding + every layer along the trajectory."""
    return (n0 - 1) * patch_k * d + sum(vit_layer_macs(a, b, d, dff, T) for a, b in trajectory(n0, ks))


def text_layer_macs(l_in: int, l_out: int, n_img: int, layer_num: int, branches: int = 2, d: int = D, dff: int = DFF,
                    T: int = T_BOOK) -> int:
    """One text layer per sample: self-attention on l_in tokens, then `branches` cross-attentions over n_img image
    tokens (2 = the NLVR twin, 1 = med.py multimodal, 0 = med.py mode 'text') and the FFN on l_out tokens."""
    macs = 4 * l_in * d * d + 2 * l_in * l_in * d + 2 * (l_in - 1) * d * T
    macs += branches * (2 * l_out * d * d + 2 * n_img * d * d + 2 * l_out * n_img * d)
    if branches == 2 and layer_num >= 6:
        macs += 2 * l_out * d * d                      # merge_layer on the concatenation (nlvr_encoder.py:282)
    return macs + 2 * l_out * d * dff


def text_macs(l0: int, ks: Iterable[int], n_img: int, branches: int = 2, d: int = D) -> int:
    return sum(text_layer_macs(a, b, n_img, i, branches, d) for i, (a, b) in enumerate(trajectory(l0, ks)))


def nlvr_macs(n0: int, vit_ks: Sequence[int], text_len: int, text_ks: Sequence[int], d: int = D) -> int:
    """One BLIP-NLVR sample (two images + one sentence, models/blip_nlvr.py:63-81) along a pruning trajectory."""
    n_img = trajectory(n0, vit_ks)[-1][1] if len(vit_ks) else n0
    return 2 * vit_macs(n0, vit_ks, d) + text_macs(text_len, text_ks, n_img, 2, d) + d * d + 2 * d


def nlvr_macs_unpruned(n0: int, text_len: int, depth: int = 12, d: int = D) -> int:
    return nlvr_macs(n0, [-1] * depth, text_len, [-1] * depth, d)


def model_trajectory(blocks) -> List[int]:
    """Per-layer topk_num of the most recent forward from the `last_prune` records the module mirrors keep."""
    return [(b.last_prune.k if (b.last_prune is not None and b.last_prune.pruned) else -1) for b in blocks]


def nlvr_gmacs_of_last_forward(model, image_size: int, text_len: int) -> float:
    """GMACs per sample of the forward `model` (madtp_b200.blip_nlvr.BLIP_NLVR) just ran -- the number the reference
    obtains by re-tracing the model with fvcore on every batch (compress_nlvr_dtp.py:93-98)."""
    patch = model.visual_encoder.patch_embed.patch_size[0]
    n0 = (image_size // patch) ** 2 + 1
    return nlvr_macs(n0, model_trajectory(model.visual_encoder.blocks), text_len,
                     model_trajectory(model.text_encoder.encoder.layer)) / 1e9


def temperature_step(temperature: float, cur_gflops: float, target_gflops: float) -> float:
    """One step of the reference's per-epoch controller (compress_nlvr_dtp.py:174-201)."""
    gap = abs(cur_gflops - target_gflops)
    step = 1.0 if gap > 30 else 0.5 if gap > 10 else 0.25 if gap > 5 else 0.1 if gap > 1 else 0.01
    return temperature + step if cur_gflops > target_gflops else temperature - step


def calibrate_temperature(ratio_at: Callable[[float], float], p: float, lo: float = 0.25, hi: float = 64.0,
                          tol: float = 0.004, max_iter: int = 16) -> Tuple[float, float, int]:
    """Temperature at which MACs(pruned) / MACs(unpruned) = 1 - p on a fixed batch: geometric bisection over
    `ratio_at(temperature)` (one forward + the closed form above per probe, no tracing). Pruning is monotone in the
    temperature (SURVEY.md probe P1). Returns (temperature, ratio, probes)."""
    target = 1.0 - p
    best: Optional[Tuple[float, float]] = None
    for it in range(max_iter):
        mid = (lo * hi) ** 0.5
        r = ratio_at(mid)
        if best is None or abs(r - target) < abs(best[1] - target):
            best = (mid, r)
        if abs(r - target) < tol:
            return mid, r, it + 1
        if r > target:
            lo = mid
        else:
            hi = mid
    return best[0], best[1], max_iter


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configurations 3-5: BLIP retrieval evaluation path, CLIP ViT-B/16 towers, BLIP-VQA question encoder
# ---------------------------------------------------------------------------------------------------------------
def retrieval_macs(n0: int, vit_ks: Sequence[int], text_len: int, text_ks: Sequence[int], mm_ks: Sequence[int],
                   d: int = D) -> int:
    """One image-text pair through the evaluation path of compress_retrieval_dtp.py:104,120,170-177: pruned ViT, text
    encoder in mode 'text', one multimodal (ITM) pass over the pruned image tokens, itm_head."""
    n_img = trajectory(n0, vit_ks)[-1][1] if len(vit_ks) else n0
    return (vit_macs(n0, vit_ks, d) + text_macs(text_len, text_ks, 0, 0, d) + text_macs(text_len, mm_ks, n_img, 1, d)
            + 2 * d)


def vqa_encoder_macs(n0: int, vit_ks: Sequence[int], text_len: int, mm_ks: Sequence[int], d: int = D) -> int:
    """One (image, question) pair through models/blip_vqa.py:60,119-125: pruned ViT and the question encoder with
    cross-attention over the pruned image tokens (the answer decoder is not part of this count)."""
    n_img = trajectory(n0, vit_ks)[-1][1] if len(vit_ks) else n0
    return vit_macs(n0, vit_ks, d) + text_macs(text_len, mm_ks, n_img, 1, d)


def clip_layer_macs(n_in: int, n_out: int, d: int, sd_dim: int = 768, T: int = T_BOOK) -> int:
    """clip/model.py:236-261: q_map Linear(d -> sd_dim) and the codebook products on the prunable tokens, attention on
    n_in tokens, the 4x QuickGELU MLP on n_out."""
    return (n_in - 1) * (d * sd_dim + 2 * sd_dim * T) + 4 * n_in * d * d + 2 * n_in * n_in * d + 8 * n_out * d * d


def clip_macs(n0: int, vision_ks: Sequence[int], text_ks: Sequence[int], context: int = 77, vision_width: int = 768,
              text_width: int = 512, embed_dim: int = 512, patch: int = 16) -> int:
    """One (image, caption) pair through CLIP.encode_image + CLIP.encode_text (clip/model.py:482-503), ViT-B/16."""
    total = (n0 - 1) * 3 * patch * patch * vision_width + vision_width * embed_dim + text_width * embed_dim
    total += sum(clip_layer_macs(a, b, vision_width) for a, b in trajectory(n0, vision_ks))
    total += sum(clip_layer_macs(a, b, text_width) for a, b in trajectory(context, text_ks))
    return total
