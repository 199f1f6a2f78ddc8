"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) -- TEST INFRASTRUCTURE.

    python -m oracle.gen_golden [--only block|nlvr|calib] [--pairs 32]

Runs only in the build container (the reference tree does not exist on the GPU box). For every fixture it
  1. builds the reference modules under oracle/ref_shims.py, loads the seeded state dict of oracle/weights.py,
  2. runs the reference forward on seeded inputs (CPU fp32), recording per-layer inputs and top-k indices through
     forward hooks and an in-memory wrapper around models.utils.vector_gather (no reference file is edited),
  3. runs oracle/dtp_oracle.py teacher-forced on the same per-layer inputs and ASSERTS agreement (keep-masks equal,
     outputs within 1e-4 after canonical re-ordering), which is what pins the oracle to the reference,
  4. writes compact expected values (scores, thresholds, counts, k, packed keep-masks, strided outputs, logits).

Token order: the reference keeps survivors in `topk(sorted=False)` order (implementation-defined); the oracle and the
CUDA path keep ascending token order. Fixtures are stored in the canonical (ascending) order; the generator tracks
the permutation between the reference's running order and the canonical one layer by layer.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import dtp_oracle as O  # noqa: E402
from oracle import ref_shims, weights  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
BLOCK_TEMPS = (1.0, 5.0, 50.0)


def maxdiff(a, b) -> float:
    return float((a.double() - b.double()).abs().max())


def digest(*tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()[:16]


class GatherRecorder:
    """Wraps the reference's vector_gather to record the index tensors it is called with."""

    def __init__(self, module):
        self.module, self.orig, self.calls = module, module.vector_gather, []

    def __enter__(self):
        def rec(vectors, indices):
            self.calls.append(indices.detach().clone())
            return self.orig(vectors, indices)
        self.module.vector_gather = rec
        return self

    def __exit__(self, *a):
        self.module.vector_gather = self.orig


# ---------------------------------------------------------------------------------------------------------------
# BASELINE config 1: one vit.Block + Query_model, B=2, N=197
# ---------------------------------------------------------------------------------------------------------------
block_inputs = weights.block_inputs


def gen_block():
    ref_shims.install()
    import models.vit as rvit
    from models.utils import Query_model
    sd = weights.block_state_dict(1234)
    from functools import partial
    # norm_layer as VisionTransformer passes it (models/vit.py:239); a bare Block would default to eps=1e-5
    blk = rvit.Block(768, 12, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    blk.load_state_dict(sd, strict=True)
    blk.eval()
    qm = Query_model(768, 768)
    x, space = block_inputs()
    out = {"input_digest": np.array(digest(x, space)), "temps": np.array(BLOCK_TEMPS)}
    with torch.no_grad():
        for ti, temp in enumerate(BLOCK_TEMPS):
            token_attn, sd_ft, _ = qm(x[:, 1:, :], space, return_token_att=True)
            ta_oracle, sd_ft_o = O.query_model(x[:, 1:, :], space, 768)
            assert torch.equal(token_attn, ta_oracle)
            assert (sd_ft - sd_ft_o).abs().max() < 1e-5
            with GatherRecorder(rvit) as rec:
                y_ref = blk(x, False, 0, temp, token_attn.clone())          # Reduce_token divides token_attn in place
            tr = O.PruneTrace()
            y_or = O.vit_block(x, {"blk." + k: v for k, v in sd.items()}, "blk", 12, temp, ta_oracle.clone(), tr)
            assert tr.pruned and len(rec.calls) == 2
            idx = rec.calls[0]                                              # [B,k] unsorted top-k indices
            k = idx.shape[1]
            keep_ref = torch.zeros(x.shape[0], x.shape[1] - 1, dtype=torch.bool).scatter_(1, idx, True)
            assert k == tr.k and torch.equal(keep_ref, tr.keep), "oracle keep-mask differs from the reference"
            perm = torch.argsort(idx, dim=1)
            y_canon = torch.cat([y_ref[:, :1], torch.gather(y_ref[:, 1:1 + k], 1, perm[..., None].expand(-1, -1, 768)),
                                 y_ref[:, 1 + k:]], dim=1)
            err = (y_canon - y_or).abs().max().item()
            assert err < 2e-5, err
            print(f"block T={temp}: k={k} counts={tr.count.tolist()} |oracle-ref|max={err:.2e}")
            p = f"t{ti}_"
            out[p + "k"] = np.array(k)
            out[p + "score"] = tr.score.numpy()
            out[p + "threshold"] = tr.threshold.numpy()
            out[p + "count"] = tr.count.numpy()
            out[p + "keep"] = np.packbits(keep_ref.numpy(), axis=1)
            out[p + "out_s4"] = y_canon[:, :, ::4].contiguous().numpy()
            out[p + "sd_ft_s4"] = sd_ft[:, :, ::4].contiguous().numpy()
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / "block_cfg1.npz", **out)


# ---------------------------------------------------------------------------------------------------------------
# BLIP-NLVR forward through the reference, with per-layer capture
# ---------------------------------------------------------------------------------------------------------------
def run_reference_nlvr(model, tok, images, ids, mask, temperature):
    """Returns dict(pred, vit=[per-layer dict], text=[per-layer dict], image_embeds, last_hidden) from the reference."""
    import models.nlvr_encoder as rnl
    import models.vit as rvit
    cap = {"vit": [], "text": []}
    hooks = []

    def vit_pre(mod, args):
        cap["vit"].append({"x": args[0].detach().clone(), "token_attn": args[4].detach().clone(), "gather": None})

    def vit_post(mod, args, out):
        cap["vit"][-1]["out"] = out.detach().clone()

    def txt_pre(mod, args, kwargs):
        cap["text"].append({"h": args[0].detach().clone(), "mask": args[1].detach().clone(),
                            "token_attn": kwargs["token_attn"].detach().clone(), "gather": None})

    def txt_post(mod, args, kwargs, out):
        cap["text"][-1]["out"] = out[0].detach().clone()
        cap["text"][-1]["mask_out"] = out[-1].detach().clone()

    for blk in model.visual_encoder.blocks:
        hooks.append(blk.register_forward_pre_hook(vit_pre))
        hooks.append(blk.register_forward_hook(vit_post))
    for layer in model.text_encoder.encoder.layer:
        hooks.append(layer.register_forward_pre_hook(txt_pre, with_kwargs=True))
        hooks.append(layer.register_forward_hook(txt_post, with_kwargs=True))

    # record the first vector_gather of each Reduce_token (= the unsorted top-k indices)
    class LayerGather(GatherRecorder):
        def __init__(s, module, key):
            super().__init__(module)
            s.key = key

        def __enter__(s):
            def rec(vectors, indices):
                layer = cap[s.key][-1]
                if layer["gather"] is None and vectors.shape[-1] == 768:
                    layer["gather"] = indices.detach().clone()
                return s.orig(vectors, indices)
            s.module.vector_gather = rec
            return s

    tok.next_ids = (ids, mask)
    with torch.no_grad(), LayerGather(rvit, "vit"), LayerGather(rnl, "text"):
        pred = model(images, ["x"] * ids.shape[0], torch.zeros(ids.shape[0], dtype=torch.long), temperature, train=False)
    for h in hooks:
        h.remove()
    # the hooks cannot see image_embeds / last_hidden directly; recompute the cheap tails
    cap["pred"] = pred.detach()
    return cap


def canonicalise_stream(layers, key_x, key_out):
    """Adds canonical-order tensors to every captured layer: 'cx' (input), 'cta' (token_attn), 'ckeep', 'cout'."""
    B = layers[0][key_x].shape[0]
    n0 = layers[0][key_x].shape[1] - 1
    pos = torch.arange(n0).unsqueeze(0).expand(B, n0).clone()      # canonical position of each reference token
    for L in layers:
        x, ta = L[key_x], L["token_attn"]
        n = x.shape[1] - 1
        inv = torch.argsort(pos, dim=1)                            # reference index of canonical token i
        L["inv"] = inv
        L["cx"] = torch.cat([x[:, :1], torch.gather(x[:, 1:], 1, inv[..., None].expand(-1, -1, x.shape[-1]))], dim=1)
        L["cta"] = torch.gather(ta, 1, inv[..., None].expand(-1, -1, ta.shape[-1]))
        out = L[key_out]
        if L["gather"] is None or out.shape[1] == x.shape[1]:
            L["ckeep"] = torch.ones(B, n, dtype=torch.bool)
            L["k"] = n
            L["pruned"] = False
            L["cout"] = torch.cat([out[:, :1], torch.gather(out[:, 1:], 1, inv[..., None].expand(-1, -1, out.shape[-1]))],
                                  dim=1)
            continue
        idx = L["gather"]                                          # [B,k] reference indices of the survivors
        k = idx.shape[1]
        cpos = torch.gather(pos, 1, idx)                           # canonical positions of survivors, reference order
        L["ckeep"] = torch.zeros(B, n, dtype=torch.bool).scatter_(1, cpos, True)
        L["k"], L["pruned"] = k, True
        rank = torch.argsort(torch.argsort(cpos, dim=1), dim=1)    # new canonical position of each survivor
        perm = torch.argsort(cpos, dim=1)
        L["cout"] = torch.cat([out[:, :1], torch.gather(out[:, 1:1 + k], 1, perm[..., None].expand(-1, -1, out.shape[-1])),
                               out[:, 1 + k:]], dim=1)
        pos = torch.cat([rank, torch.full((B, 1), k, dtype=torch.long)], dim=1)
    return layers


def check_oracle_vit(layers, sd, temperature):
    """Teacher-forced oracle vs reference for every ViT layer; returns the oracle traces."""
    traces = []
    for i, L in enumerate(layers):
        tr = O.PruneTrace()
        y = O.vit_block(L["cx"], sd, f"visual_encoder.blocks.{i}", 12, temperature, L["cta"].clone(), tr)
        assert tr.pruned == L["pruned"] and (not tr.pruned or tr.k == L["k"]), (i, tr.k, L["k"])
        assert torch.equal(tr.keep, L["ckeep"]), f"ViT layer {i}: oracle keep-mask differs from the reference"
        err = (y - L["cout"]).abs().max().item()
        assert err < 2e-4, (i, err)
        traces.append(tr)
    return traces


def gen_nlvr(image_size: int, pairs: int, text_len: int, temps, name: str, pad_to: int = 0):
    model, tok = ref_shims.build_blip_nlvr(image_size)
    sd = weights.blip_nlvr_state_dict(1234, img_size=image_size)
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    assert all("position_ids" in k or "tokenizer" in k for k in missing.missing_keys), missing.missing_keys
    images, ids, mask = weights.nlvr_inputs(pairs, image_size, text_len, seed=0, pad_to=pad_to)
    out = {"input_digest": np.array(digest(images, ids, mask)), "temps": np.array(temps),
           "image_size": np.array(image_size), "pairs": np.array(pairs), "text_len": np.array(text_len)}
    for ti, temp in enumerate(temps):
        t0 = time.time()
        cap = run_reference_nlvr(model, tok, images, ids, mask, temp)
        t_ref = time.time() - t0
        vit = canonicalise_stream(cap["vit"], "x", "out")
        traces = check_oracle_vit(vit, sd, temp)
        # free-running oracle: logits and trajectories
        ntr = O.NlvrTrace()
        pred_or = O.blip_nlvr_forward(images, ids, mask, sd, temp, trace=ntr)
        perr = (pred_or - cap["pred"]).abs().max().item()
        ks_ref = [L["k"] if L["pruned"] else -1 for L in vit]
        ks_or = [t.k if t.pruned else -1 for t in ntr.vit]
        tks_ref = [(L["out"].shape[1] - 2) if L["out"].shape[1] != L["h"].shape[1] else -1 for L in cap["text"]]
        tks_or = [t.k if t.pruned else -1 for t in ntr.text]
        print(f"{name} T={temp}: ref {t_ref:.1f}s  vit k ref={ks_ref}\n    oracle(free)={ks_or}\n    text k ref={tks_ref} "
              f"oracle={tks_or}  |pred diff|={perr:.2e}")
        assert ks_ref == ks_or and tks_ref == tks_or, "free-running oracle trajectory differs from the reference"
        assert perr < 1e-4
        if pad_to == 0:
            for i, (t, L) in enumerate(zip(ntr.text, cap["text"])):
                # un-padded text: every later op is permutation-equivariant, compare as sorted multisets of rows
                a = t.layer_output.sum(-1).sort(dim=1)[0]
                b = L["out"].sum(-1).sort(dim=1)[0]
                assert (a - b).abs().max() < 1e-3, (i, (a - b).abs().max())
        p = f"t{ti}_"
        out[p + "pred"] = cap["pred"].numpy()
        out[p + "vit_k"] = np.array(ks_ref)
        out[p + "text_k"] = np.array(tks_ref)
        for i, (L, tr) in enumerate(zip(vit, traces)):
            out[p + f"vit{i}_keep"] = np.packbits(L["ckeep"].numpy(), axis=1)
            out[p + f"vit{i}_score"] = tr.score.numpy()
            out[p + f"vit{i}_count"] = tr.count.numpy()
            out[p + f"vit{i}_threshold"] = tr.threshold.numpy()
        for i, t in enumerate(ntr.text):
            out[p + f"text{i}_keep"] = np.packbits(t.keep.numpy(), axis=1)
            out[p + f"text{i}_score"] = t.score.numpy()
        out[p + "image_embeds_s8"] = ntr.image_embeds[:, :, ::8].contiguous().numpy()
        out[p + "last_hidden"] = ntr.last_hidden.numpy()
        macs_p = O.nlvr_macs_from_trace(ntr, n0=(image_size // 16) ** 2 + 1, text_len=ids.shape[1])
        out[p + "macs"] = np.array(macs_p)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / f"{name}.npz", **out)


# ---------------------------------------------------------------------------------------------------------------
# models/med.py text encoder (BLIP retrieval / VQA): mode 'text' with padded text and mode 'multimodal'
# ---------------------------------------------------------------------------------------------------------------
def gen_med():
    ref_shims.install()
    import models.med as rmed
    from transformers.models.bert.configuration_bert import BertConfig as HFBertConfig
    cfg = HFBertConfig.from_json_file(os.path.join(ref_shims.REFERENCE_ROOT, "configs/med_config.json"))
    cfg.encoder_width = 768
    cfg.evaluate = True
    model = rmed.BertModel(config=cfg, add_pooling_layer=False, sd_dim=768)
    g = torch.Generator().manual_seed(4321)
    sd = weights.med_text_state_dict(g, "")
    space = torch.randn(100, 768, generator=g)
    msg = model.load_state_dict(sd, strict=False)
    assert not msg.unexpected_keys and all("position_ids" in k for k in msg.missing_keys), msg
    model.eval()
    _, ids, mask = weights.retrieval_inputs(4, img_size=32, max_len=35, seed=0)
    enc = torch.randn(4, 60, 768, generator=g)
    out = {"input_digest": np.array(digest(ids, mask, enc, space))}
    cases = [("text", 3.0), ("text", 10.0), ("multimodal", 10.0)]
    out["modes"] = np.array([c[0] for c in cases])
    out["temps"] = np.array([c[1] for c in cases])
    for ci, (mode, temp) in enumerate(cases):
        cap = []
        hooks = []

        def pre(mod, args, kwargs):
            cap.append({"h": args[0].detach().clone(), "mask": args[1].detach().clone(),
                        "token_attn": kwargs["token_attn"].detach().clone(), "gather": None})

        def post(mod, args, kwargs, o):
            cap[-1]["out"], cap[-1]["mask_out"] = o[0].detach().clone(), o[-1].detach().clone()
        for layer in model.encoder.layer:
            hooks.append(layer.register_forward_pre_hook(pre, with_kwargs=True))
            hooks.append(layer.register_forward_hook(post, with_kwargs=True))

        class Rec(GatherRecorder):
            def __enter__(s):
                def rec(vectors, indices):
                    if cap[-1]["gather"] is None and vectors.shape[-1] == 768:
                        cap[-1]["gather"] = indices.detach().clone()
                    return s.orig(vectors, indices)
                s.module.vector_gather = rec
                return s
        with torch.no_grad(), Rec(rmed):
            o, sd_txt = model(ids, attention_mask=mask, encoder_hidden_states=enc if mode == "multimodal" else None,
                              return_dict=True, mode=mode, space_dict=space, temperature=temp)
        for h in hooks:
            h.remove()
        # canonical order bookkeeping; med keeps the first k of topk(k+1) and gathers masks with the same indices
        B, n0 = ids.shape[0], ids.shape[1] - 1
        pos = torch.arange(n0).unsqueeze(0).expand(B, n0).clone()
        ks, n_pruned = [], 0
        for i, L in enumerate(cap):
            x, ta, m = L["h"], L["token_attn"], L["mask"]
            n = x.shape[1] - 1
            inv = torch.argsort(pos, dim=1)
            cx = torch.cat([x[:, :1], torch.gather(x[:, 1:], 1, inv[..., None].expand(-1, -1, 768))], dim=1)
            cta = torch.gather(ta, 1, inv[..., None].expand(-1, -1, ta.shape[-1]))
            cm = torch.cat([m[..., :1], torch.gather(m[:, 0, 0, 1:], 1, inv)[:, None, None, :]], dim=-1)
            tr = O.PruneTrace()
            y, mo = O.med_layer(cx, cm, sd, f"encoder.layer.{i}", enc if mode == "multimodal" else None, temp,
                                cta.clone(), mode, trace=tr)
            o_ref, mo_ref = L["out"], L["mask_out"]
            if L["gather"] is None or o_ref.shape[1] == x.shape[1]:
                assert not tr.pruned, i
                c_out = torch.cat([o_ref[:, :1], torch.gather(o_ref[:, 1:], 1, inv[..., None].expand(-1, -1, 768))], 1)
                c_mo = torch.cat([mo_ref[..., :1], torch.gather(mo_ref[:, 0, 0, 1:], 1, inv)[:, None, None, :]], -1)
                ks.append(-1)
            else:
                idx = L["gather"]
                k = idx.shape[1] - 1
                cpos = torch.gather(pos, 1, idx[:, :k])
                ckeep = torch.zeros(B, n, dtype=torch.bool).scatter_(1, cpos, True)
                assert tr.pruned and tr.k == k and torch.equal(tr.keep, ckeep), f"med layer {i}: keep-mask differs"
                perm = torch.argsort(cpos, dim=1)
                c_out = torch.cat([o_ref[:, :1], torch.gather(o_ref[:, 1:1 + k], 1, perm[..., None].expand(-1, -1, 768)),
                                   o_ref[:, 1 + k:]], dim=1)
                mrow = mo_ref[:, 0, 0, :]
                c_mo = torch.cat([mrow[:, :1], torch.gather(mrow[:, 1:1 + k], 1, perm), mrow[:, 1 + k:]], 1)[:, None, None, :]
                rank = torch.argsort(torch.argsort(cpos, dim=1), dim=1)
                pos = torch.cat([rank, torch.full((B, 1), k, dtype=torch.long)], dim=1)
                ks.append(k)
                n_pruned += 1
                out[f"c{ci}_l{i}_keep"] = np.packbits(ckeep.numpy(), axis=1)
                out[f"c{ci}_l{i}_score"] = tr.score.numpy()
            err = (y - c_out).abs().max().item()
            assert err < 2e-4, (i, err)
            assert torch.equal(mo, c_mo), f"med layer {i}: pruned mask differs"
        h_or, sd_or = O.med_text_encoder(ids, mask, sd, "", enc if mode == "multimodal" else None, space, temp, mode)
        a = h_or.sum(-1).sort(dim=1)[0]
        b = o.last_hidden_state.sum(-1).sort(dim=1)[0]
        print(f"med {mode} T={temp}: k per layer {ks}; free-running |row-sum diff| {(a - b).abs().max():.2e}")
        assert n_pruned > 0
        out[f"c{ci}_k"] = np.array(ks)
        out[f"c{ci}_cls"] = o.last_hidden_state[:, 0, :].numpy()
        out[f"c{ci}_sd_txt_s4"] = sd_txt[:, :, ::4].contiguous().numpy()
        assert (h_or[:, 0] - o.last_hidden_state[:, 0]).abs().max() < 1e-3
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / "med_text.npz", **out)


# ---------------------------------------------------------------------------------------------------------------
# CLIP (clip/model.py + clip/mock.py): vision tower (free-running + per block) and causal text blocks (per block)
# ---------------------------------------------------------------------------------------------------------------
CLIP_LAYERS = 4   # fixture depth (every block has identical structure; full-depth shapes are covered on the GPU)


def gen_clip():
    ref_shims.install_clip()
    import clip.model as cm
    sd = weights.clip_state_dict(777, vision_layers=CLIP_LAYERS, text_layers=CLIP_LAYERS)
    space = sd["space_dict"]
    images, text = weights.clip_inputs(2)
    out = {"input_digest": np.array(digest(images, text, space)), "layers": np.array(CLIP_LAYERS)}

    def run_blocks(blocks, x_lnd, temp, max_keep):
        cap = []
        hooks = []

        def pre(mod, args):
            cap.append({"x": args[0][0].detach().permute(1, 0, 2).clone(), "gather": None})

        def post(mod, args, o):
            cap[-1]["out"] = o[0].detach().permute(1, 0, 2).clone()
        for blk in blocks:
            hooks.append(blk.register_forward_pre_hook(pre))
            hooks.append(blk.register_forward_hook(post))

        class Rec(GatherRecorder):
            def __enter__(s):
                def rec(vectors, indices):
                    if cap[-1]["gather"] is None:
                        cap[-1]["gather"] = indices.detach().clone()
                    return s.orig(vectors, indices)
                s.module.vector_gather = rec
                return s
        with torch.no_grad(), Rec(cm):
            y = torch.nn.Sequential(*blocks)((x_lnd, space, temp, None, max_keep))
        for h in hooks:
            h.remove()
        return cap, y

    # ---- vision tower ----
    vis = cm.VisionTransformer(input_resolution=224, patch_size=16, width=768, layers=CLIP_LAYERS, heads=12,
                               output_dim=512, sd_dim=768)
    msg = vis.load_state_dict({k[len("visual."):]: v for k, v in sd.items() if k.startswith("visual.")}, strict=True)
    vis.eval()
    temp_v = 5.0
    with torch.no_grad():
        emb_ref, sd_img_ref = vis(images, space_dict=space, temperature=temp_v)
        x0 = vis.ln_pre(torch.cat([vis.class_embedding + torch.zeros(2, 1, 768),
                                   vis.conv1(images).reshape(2, 768, -1).permute(0, 2, 1)], 1) + vis.positional_embedding)
    cap, _ = run_blocks(list(vis.transformer.resblocks), x0.permute(1, 0, 2), temp_v, 1)
    B, n0 = 2, x0.shape[1] - 1
    pos = torch.arange(n0).unsqueeze(0).expand(B, n0).clone()
    ks = []
    for i, L in enumerate(cap):
        x = L["x"]
        n = x.shape[1] - 1
        inv = torch.argsort(pos, dim=1)
        cx = torch.cat([x[:, :1], torch.gather(x[:, 1:], 1, inv[..., None].expand(-1, -1, 768))], dim=1)
        tr = O.PruneTrace()
        y, _ = O.clip_block(cx, sd, f"visual.transformer.resblocks.{i}", 12, space, temp_v, None, 1, False, tr)
        o_ref = L["out"]
        assert L["gather"] is not None and o_ref.shape[1] != x.shape[1], "vision blocks are expected to prune"
        idx = L["gather"]
        k = idx.shape[1]
        cpos = torch.gather(pos, 1, idx)
        ckeep = torch.zeros(B, n, dtype=torch.bool).scatter_(1, cpos, True)
        assert tr.pruned and tr.k == k and torch.equal(tr.keep, ckeep), f"CLIP vision block {i}: keep-mask differs"
        perm = torch.argsort(cpos, dim=1)
        c_out = torch.cat([o_ref[:, :1], torch.gather(o_ref[:, 1:1 + k], 1, perm[..., None].expand(-1, -1, 768)),
                           o_ref[:, 1 + k:]], dim=1)
        err = (y - c_out).abs().max().item()
        assert err < 2e-4, (i, err)
        rank = torch.argsort(torch.argsort(cpos, dim=1), dim=1)
        pos = torch.cat([rank, torch.full((B, 1), k, dtype=torch.long)], dim=1)
        ks.append(k)
        out[f"v{i}_keep"] = np.packbits(ckeep.numpy(), axis=1)
        out[f"v{i}_score"] = tr.score.numpy()
    emb_or, sd_img_or = O.clip_vision_forward(images, sd, "visual.", space, temp_v, CLIP_LAYERS, 12)
    print(f"clip vision T={temp_v}: k {ks}; |emb oracle - ref| {(emb_or - emb_ref).abs().max():.2e}")
    assert (emb_or - emb_ref).abs().max() < 1e-4 and (sd_img_or - sd_img_ref).abs().max() < 1e-3
    out["v_k"], out["v_temp"] = np.array(ks), np.array(temp_v)
    out["v_emb"] = emb_ref.numpy()

    # ---- text transformer blocks (causal, max_keep guard) ----
    mask = torch.empty(77, 77).fill_(float("-inf")).triu_(1)
    txt = cm.Transformer(width=512, layers=CLIP_LAYERS, heads=8, attn_mask=mask, sd_dim=768)
    txt.load_state_dict({k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")},
                        strict=True)
    txt.eval()
    temp_t = 50.0
    max_keep = int(text.argmax(dim=-1).max()) + 2
    x0 = sd["token_embedding.weight"][text] + sd["positional_embedding"]
    cap, _ = run_blocks(list(txt.resblocks), x0.permute(1, 0, 2), temp_t, max_keep)
    tks = []
    for i, L in enumerate(cap):
        x = L["x"]                                   # reference order: the causal mask makes order observable
        tr = O.PruneTrace()
        y, _ = O.clip_block(x, sd, f"transformer.resblocks.{i}", 8, space, temp_t, None, max_keep, True, tr)
        o_ref = L["out"]
        out[f"t{i}_x"] = x.numpy()
        if L["gather"] is None or o_ref.shape[1] == x.shape[1]:
            assert not tr.pruned, i
            c_out = o_ref
            tks.append(-1)
        else:
            idx = L["gather"]
            k = idx.shape[1]
            keep = torch.zeros(2, x.shape[1] - 1, dtype=torch.bool).scatter_(1, idx, True)
            assert tr.pruned and tr.k == k and torch.equal(tr.keep, keep), f"CLIP text block {i}: keep-mask differs"
            perm = torch.argsort(idx, dim=1)
            c_out = torch.cat([o_ref[:, :1], torch.gather(o_ref[:, 1:1 + k], 1, perm[..., None].expand(-1, -1, 512)),
                               o_ref[:, 1 + k:]], dim=1)
            tks.append(k)
            out[f"t{i}_keep"] = np.packbits(keep.numpy(), axis=1)
            out[f"t{i}_score"] = tr.score.numpy()
        err = (y - c_out).abs().max().item()
        assert err < 2e-4, (i, err)
        out[f"t{i}_out_s4"] = c_out[:, :, ::4].contiguous().numpy()
    print(f"clip text T={temp_t} max_keep={max_keep}: k {tks}")
    assert any(k > 0 for k in tks) and any(k < 0 for k in tks), "fixture must exercise both the prune and the guard"
    out["t_k"], out["t_temp"], out["t_max_keep"] = np.array(tks), np.array(temp_t), np.array(max_keep)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / "clip_blocks.npz", **out)


# ---------------------------------------------------------------------------------------------------------------
# Temperature calibration for "p = 0.5" on the bench batch (BASELINE config 2), on the oracle
# ---------------------------------------------------------------------------------------------------------------
def gen_calibration(pairs: int, image_size: int = 384, text_len: int = 20, p: float = 0.5):
    sd = weights.blip_nlvr_state_dict(1234, img_size=image_size)
    images, ids, mask = weights.nlvr_inputs(pairs, image_size, text_len, seed=0)
    n0 = (image_size // 16) ** 2 + 1

    def ratio(temp):
        ntr = O.NlvrTrace()
        with torch.no_grad():
            pred = O.blip_nlvr_forward(images, ids, mask, sd, temp, trace=ntr)
        full = O.nlvr_macs_unpruned(n0, ids.shape[1])
        return O.nlvr_macs_from_trace(ntr, n0, ids.shape[1]) / full, ntr, pred

    lo, hi = 0.25, 16.0
    best = None
    for it in range(12):
        mid = (lo * hi) ** 0.5
        t0 = time.time()
        r, ntr, pred = ratio(mid)
        print(f"calib it{it}: T={mid:.4f} ratio={r:.4f} ({time.time() - t0:.1f}s) vit k={[t.k for t in ntr.vit]}")
        best = (mid, r, ntr, pred)
        if abs(r - (1 - p)) < 0.004:
            break
        if r > 1 - p:
            lo = mid
        else:
            hi = mid
    temp, r, ntr, pred = best
    assert abs(r - (1 - p)) < 0.01
    out = {"temperature": np.array(temp), "ratio": np.array(r), "pairs": np.array(pairs), "p": np.array(p),
           "image_size": np.array(image_size), "text_len": np.array(text_len),
           "input_digest": np.array(digest(images, ids, mask)),
           "vit_k": np.array([t.k if t.pruned else -1 for t in ntr.vit]),
           "text_k": np.array([t.k if t.pruned else -1 for t in ntr.text]),
           "pred": pred.numpy(),
           "macs_pruned": np.array(O.nlvr_macs_from_trace(ntr, n0, ids.shape[1])),
           "macs_unpruned": np.array(O.nlvr_macs_unpruned(n0, ids.shape[1]))}
    for i, t in enumerate(ntr.vit):
        out[f"vit{i}_keep"] = np.packbits(t.keep.numpy(), axis=1)
        out[f"vit{i}_count"] = t.count.numpy()
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / f"calib_nlvr_p{int(p * 100)}_b{pairs}.npz", **out)


def gen_vqa_rank():
    """BLIP_VQA(inference='rank') of the unmodified reference (models/blip_vqa.py:117-203): pruned image + question
    encoders, then the answer decoder ranks k_test candidates. The oracle must reproduce the chosen answers, the top-k
    candidate sets and the summed log-probabilities."""
    ref_shims.install()
    import models.blip as blip
    tok = ref_shims.FakeTokenizer()
    blip.init_tokenizer = lambda: tok
    import models.blip_vqa as bv
    bv.init_tokenizer = lambda: tok
    size, B, k_test, temp = 224, 3, 3, 6.0
    model = bv.BLIP_VQA(med_config=os.path.join(ref_shims.REFERENCE_ROOT, "configs/med_config.json"), image_size=size,
                        vit="base", evaluate=True)
    sd = weights.vqa_state_dict(99, img_size=size)
    msg = model.load_state_dict(sd, strict=False)
    assert not msg.unexpected_keys and all("position_ids" in k for k in msg.missing_keys), msg
    model.eval()
    images, ids, mask = weights.retrieval_inputs(B, size, 20, seed=3)
    L = int(mask.sum(1).max())                       # padding='longest'
    ids, mask = ids[:, :L].contiguous(), mask[:, :L].contiguous()
    ans_ids, ans_mask = weights.vqa_answer_candidates(6, 5, seed=1, bos_id=tok.bos_token_id)
    tok.next_ids = (ids, mask)

    class Answer:
        input_ids, attention_mask = ans_ids, ans_mask
    cap = {}
    orig_rank = model.rank_answer

    def rank_spy(question_states, question_atts, answer_ids, answer_atts, k):
        cap["question_states"] = question_states.detach().clone()
        return orig_rank(question_states, question_atts, answer_ids, answer_atts, k)
    model.rank_answer = rank_spy
    dec_out = []
    hook = model.text_decoder.register_forward_hook(lambda m, a, o: dec_out.append(o))
    with torch.no_grad():
        max_ids = model(images, ["q"] * B, Answer, temperature=temp, train=False, inference="rank", k_test=k_test)
    hook.remove()
    ref_first_logits = dec_out[0].logits[:, 0, :]
    ref_prob_first = torch.softmax(ref_first_logits, dim=1).index_select(1, ans_ids[:, 1])
    ref_logp = (-dec_out[1].loss).view(B, k_test)

    # the oracle, end to end from the same inputs
    with torch.no_grad():
        feat, _ = O.vit_forward(images, sd, "visual_encoder.", sd["space_dict"], temp)
        ids2 = ids.clone()
        ids2[:, 0] = tok.enc_token_id
        q_states, _ = O.med_text_encoder(ids2, mask, sd, "text_encoder.", feat, sd["space_dict"], temp, "multimodal")
        o_max, o_topk, o_logp, o_prob = O.vqa_rank_answer(q_states, ans_ids, ans_mask, k_test, sd)
    assert q_states.shape == cap["question_states"].shape, (q_states.shape, cap["question_states"].shape)
    print("vqa: question states", maxdiff(q_states, cap["question_states"]), "prob_first", maxdiff(o_prob, ref_prob_first),
          "log-probs", maxdiff(o_logp, ref_logp))
    assert maxdiff(q_states, cap["question_states"]) < 2e-4
    assert maxdiff(o_prob, ref_prob_first) < 1e-6 and maxdiff(o_logp, ref_logp) < 2e-3
    assert torch.equal(o_max, max_ids), (o_max, max_ids)
    assert torch.equal(o_topk.sort(1)[0], ref_prob_first.topk(k_test, dim=1)[1].sort(1)[0])
    out = {"input_digest": np.array(digest(images, ids, mask, ans_ids, ans_mask)), "temperature": np.array(temp),
           "k_test": np.array(k_test), "image_size": np.array(size), "ids": ids.numpy(), "mask": mask.numpy(),
           "answer_ids": ans_ids.numpy(), "answer_mask": ans_mask.numpy(),
           "question_states": cap["question_states"].numpy(), "prob_first": ref_prob_first.numpy(),
           "topk_ids": ref_prob_first.topk(k_test, dim=1)[1].numpy(), "log_probs_sum": ref_logp.numpy(),
           "max_ids": max_ids.numpy()}
    np.savez_compressed(GOLDEN / "vqa_rank.npz", **out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="all")
    ap.add_argument("--pairs", type=int, default=32)
    a = ap.parse_args()
    assert ref_shims.available() or a.only == "calib", "the reference tree is not mounted"
    torch.manual_seed(0)
    if a.only in ("all", "block"):
        gen_block()
    if a.only in ("all", "nlvr"):
        gen_nlvr(224, 2, 20, (1.0, 8.0), "nlvr_small224")
    if a.only in ("all", "clip"):
        gen_clip()
    if a.only in ("all", "med"):
        gen_med()
    if a.only in ("all", "vqa"):
        gen_vqa_rank()
    if a.only in ("all", "nlvr384"):
        gen_nlvr(384, 2, 20, (1.5,), "nlvr_small384")
    if a.only in ("all", "calib"):
        gen_calibration(a.pairs)
