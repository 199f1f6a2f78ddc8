"""Drop-in proof -- TEST INFRASTRUCTURE. Executes the re-exports INTEGRATION.md section 2 tells a maintainer of the
reference to add, then runs the REFERENCE's own task-model wiring (models/blip_nlvr.py, models/blip_retrieval.py,
models/blip_vqa.py, clip/model.py -- unmodified, from /root/reference or its staging oracle/_ref) on top of the
madtp_b200 mirrors on cuda:0, and checks the results against the golden fixtures the unmodified reference produced.

    python -m oracle.dropin [nlvr] [checkpoint] [retrieval] [vqa] [clip]        # prints one JSON line per case

It runs in its own process (tests/test_dropin_gpu.py spawns it) because it replaces `models.vit`, `models.utils`,
`models.nlvr_encoder` and `models.med` in sys.modules -- exactly what the integration does.
"""
from __future__ import annotations

import importlib
import importlib.util
import json
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import ref_shims  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def _reexport(name: str, src, names):
    m = types.ModuleType(name)
    for n in names:
        setattr(m, n, getattr(src, n))
    m.__all__ = list(names)
    sys.modules[name] = m
    return m


def install_mirrors():
    """INTEGRATION.md section 2, executed: the reference's module paths re-export the madtp_b200 classes."""
    ref_shims.install()
    import models  # the reference's package (empty __init__)
    from madtp_b200 import med, nlvr_encoder, utils, vit
    models.vit = _reexport("models.vit", vit, ["Mlp", "Attention", "Block", "VisionTransformer", "interpolate_pos_embed"])
    models.utils = _reexport("models.utils", utils, ["vector_gather", "Query_model"])
    models.nlvr_encoder = _reexport("models.nlvr_encoder", nlvr_encoder, [
        "BertEmbeddings", "BertSelfAttention", "BertSelfOutput", "BertAttention", "BertIntermediate", "BertOutput",
        "BertLayer", "BertEncoder", "BertModel"])
    models.med = _reexport("models.med", med, [
        "BertConfig", "BertEmbeddings", "BertSelfAttention", "BertSelfOutput", "BertAttention", "BertIntermediate",
        "BertOutput", "BertLayer", "BertEncoder", "BertModel", "BertOnlyMLMHead", "BertLMHeadModel"])
    tok = ref_shims.FakeTokenizer()
    import models.blip as blip
    blip.init_tokenizer = lambda: tok
    return tok


def _real_reference_module(rel: str, name: str):
    """Import an UNSWAPPED reference file under a private name (to compare a helper against the original)."""
    spec = importlib.util.spec_from_file_location(name, os.path.join(ref_shims.REFERENCE_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _mirror_classes(model):
    return sorted({type(m).__module__ for m in model.modules() if type(m).__module__.startswith(("madtp_b200", "models"))})


# ---------------------------------------------------------------------------------------------------------------
def case_nlvr(tok):
    """Reference models/blip_nlvr.py:BLIP_NLVR (wiring, tokenizer call, cls_head) over the mirrored encoders, against
    nlvr_small224.npz (the unmodified reference's logits and pruning trajectory)."""
    from madtp_b200 import synthetic
    import models.blip_nlvr as bn
    bn.init_tokenizer = lambda: tok
    model = bn.BLIP_NLVR(med_config=os.path.join(ref_shims.REFERENCE_ROOT, "configs/med_config.json"), image_size=224,
                         vit="base", evaluate=True)
    assert type(model.visual_encoder).__module__ == "madtp_b200.vit", type(model.visual_encoder)
    assert type(model.text_encoder).__module__ == "madtp_b200.nlvr_encoder", type(model.text_encoder)
    msg = model.load_state_dict(synthetic.blip_nlvr_state_dict(1234, img_size=224), strict=False)
    assert not msg.unexpected_keys and not msg.missing_keys, msg
    model = model.cuda().eval()
    gold = np.load(GOLDEN / "nlvr_small224.npz")
    images, ids, mask = synthetic.nlvr_inputs(2, 224, 20, seed=0)
    assert synthetic.tensor_digest(images, ids, mask) == str(gold["input_digest"])
    out = {"case": "nlvr", "class_modules": _mirror_classes(model), "temps": []}
    for ti, temp in enumerate(gold["temps"].tolist()):
        tok.next_ids = (ids, mask)
        with torch.no_grad():
            pred = model(images.cuda(), ["x"] * 2, torch.zeros(2, dtype=torch.long, device="cuda"), temp, train=False)
        ks = [(b.last_prune.k if b.last_prune is not None and b.last_prune.pruned else -1)
              for b in model.visual_encoder.blocks]
        tks = [(l.last_prune.k if l.last_prune is not None and l.last_prune.pruned else -1)
               for l in model.text_encoder.encoder.layer]
        err = float((pred.cpu() - torch.from_numpy(gold[f"t{ti}_pred"])).abs().max())
        assert err < 5e-3, err
        assert ks == gold[f"t{ti}_vit_k"].tolist(), (ks, gold[f"t{ti}_vit_k"].tolist())
        assert tks == gold[f"t{ti}_text_k"].tolist(), (tks, gold[f"t{ti}_text_k"].tolist())
        out["temps"].append({"temperature": temp, "logit_max_abs_err": err, "vit_k": ks, "text_k": tks})
    return out


def case_checkpoint(tok):
    """The reference's own load_checkpoint (models/blip_nlvr.py:131-160) and blip_nlvr(pretrained=...) factory over the
    mirrors: a pretrained-BLIP-style checkpoint (ONE cross-attention, 224 x 224 position grid) loads into a 384 x 384
    BLIP-NLVR; the mirrored interpolate_pos_embed must equal the reference's function bit for bit."""
    from madtp_b200 import synthetic
    import models.blip_nlvr as bn
    bn.init_tokenizer = lambda: tok
    full = synthetic.blip_nlvr_state_dict(77, img_size=224)
    pre = {}
    for k, v in full.items():
        if "crossattention.self1." in k or "crossattention.output.dense1." in k or "merge_layer" in k:
            continue
        pre[k.replace("crossattention.self0.", "crossattention.self.").replace("crossattention.output.dense0.",
                                                                               "crossattention.output.dense.")] = v
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "pretrained.pth")
        torch.save({"model": pre, "epoch": 3, "temperature": 2.5}, path)
        model = bn.blip_nlvr(pretrained=path, med_config=os.path.join(ref_shims.REFERENCE_ROOT, "configs/med_config.json"),
                             image_size=384, vit="base", evaluate=True)
    sd = model.state_dict()
    n_fan = 0
    for k, v in pre.items():
        if "crossattention.self." in k:
            for br in ("self0", "self1"):
                assert torch.equal(sd[k.replace("self", br)], v), k
                n_fan += 1
        elif "crossattention.output.dense." in k:
            for br in ("dense0", "dense1"):
                assert torch.equal(sd[k.replace("dense", br)], v), k
                n_fan += 1
    assert n_fan == 12 * (6 + 2) * 2, n_fan
    real_vit = _real_reference_module("models/vit.py", "_ref_models_vit_real")
    want = real_vit.interpolate_pos_embed(pre["visual_encoder.pos_embed"].clone(), model.visual_encoder)
    assert want.shape == (1, 577, 768) and torch.equal(sd["visual_encoder.pos_embed"], want)
    # and the loaded model runs
    model = model.cuda().eval()
    images, ids, mask = synthetic.nlvr_inputs(1, 384, 12, seed=4)
    tok.next_ids = (ids, mask)
    with torch.no_grad():
        pred = model(images.cuda(), ["x"], torch.zeros(1, dtype=torch.long, device="cuda"), 2.0, train=False)
    assert pred.shape == (1, 2) and bool(torch.isfinite(pred).all())
    return {"case": "checkpoint", "fanned_out_tensors": n_fan, "pos_embed": list(want.shape)}


def case_retrieval(tok):
    """Reference models/blip_retrieval.py:BLIP_Retrieval constructed over the mirrors (incl. its momentum twins) and
    driven exactly like compress_retrieval_dtp.py:104-122,166-177 (text encoder in mode 'text', image encoder, ITM pass)
    against the oracle on the same seeded inputs."""
    from madtp_b200 import synthetic
    from oracle import dtp_oracle as O
    import models.blip_retrieval as br
    br.init_tokenizer = lambda: tok
    size, temp = 224, 8.0
    model = br.BLIP_Retrieval(med_config=os.path.join(ref_shims.REFERENCE_ROOT, "configs/med_config.json"),
                              image_size=size, vit="base", evaluate=True)
    assert type(model.visual_encoder).__module__ == "madtp_b200.vit"
    assert type(model.text_encoder).__module__ == "madtp_b200.med"
    sd = synthetic.retrieval_state_dict(4321, img_size=size)
    msg = model.load_state_dict(sd, strict=False)
    assert not msg.unexpected_keys, msg.unexpected_keys
    model = model.cuda().eval()
    images, ids, mask = synthetic.retrieval_inputs(3, size, 35, seed=1)
    space = sd["space_dict"]
    with torch.no_grad():
        t_out, _ = model.text_encoder(ids.cuda(), attention_mask=mask.cuda(), mode='text', space_dict=model.space_dict,
                                      temperature=temp)                                  # compress_retrieval_dtp.py:104
        feat, _ = model.visual_encoder(images.cuda(), space_dict=model.space_dict, temperature=temp)   # :120
        enc_ids = ids.clone()
        enc_ids[:, 0] = tok.enc_token_id                                                 # :112
        atts = torch.ones(feat.size()[:-1], dtype=torch.long, device="cuda")
        mm = model.text_encoder(enc_ids.cuda(), attention_mask=mask.cuda(), encoder_hidden_states=feat,
                                encoder_attention_mask=atts, return_dict=True, space_dict=model.space_dict,
                                temperature=temp)[0]                                     # :170-176
        itm = model.itm_head(mm.last_hidden_state[:, 0, :])[:, 1]                        # :177
        feat_o, _ = O.vit_forward(images, sd, "visual_encoder.", space, temp)
        txt_o, _ = O.med_text_encoder(ids, mask, sd, "text_encoder.", None, space, temp, "text")
        mm_o, _ = O.med_text_encoder(enc_ids, mask, sd, "text_encoder.", feat_o, space, temp, "multimodal")
        itm_o = O.linear(mm_o[:, 0, :], sd, "itm_head")[:, 1]
    assert abs(feat.shape[1] - feat_o.shape[1]) <= 2, (feat.shape, feat_o.shape)
    e_txt = _rel(t_out.last_hidden_state[:, 0, :], txt_o[:, 0, :])
    e_itm = float((itm.cpu() - itm_o).abs().max())
    assert e_txt < 5e-3 and e_itm < 2e-2, (e_txt, e_itm)
    return {"case": "retrieval", "image_tokens": [int(feat.shape[1]), int(feat_o.shape[1])], "text_cls_rel": e_txt,
            "itm_abs": e_itm}


def case_vqa(tok):
    """Reference models/blip_vqa.py:BLIP_VQA.forward(train=False, inference='rank') (its own rank_answer / tile wiring)
    over the mirrored encoders and answer decoder, against vqa_rank.npz from the unmodified reference."""
    from madtp_b200 import synthetic
    import models.blip_vqa as bv
    bv.init_tokenizer = lambda: tok
    fx = np.load(GOLDEN / "vqa_rank.npz")
    size, k, temp = int(fx["image_size"]), int(fx["k_test"]), float(fx["temperature"])
    model = bv.BLIP_VQA(med_config=os.path.join(ref_shims.REFERENCE_ROOT, "configs/med_config.json"), image_size=size,
                        vit="base", evaluate=True)
    assert type(model.text_decoder).__module__ == "madtp_b200.med"
    sd = synthetic.vqa_state_dict(99, img_size=size)
    msg = model.load_state_dict(sd, strict=False)
    assert set(msg.unexpected_keys) <= {"text_decoder.cls.predictions.decoder.bias"} and not msg.missing_keys, msg
    model = model.cuda().eval()
    ids, mask = torch.from_numpy(fx["ids"]), torch.from_numpy(fx["mask"])
    images, _, _ = synthetic.retrieval_inputs(ids.shape[0], size, 20, seed=3)

    class Enc:
        def __init__(s, i, m):
            s.input_ids, s.attention_mask = i.cuda(), m.cuda()

        def to(s, device):
            return s
    answer = Enc(torch.from_numpy(fx["answer_ids"]), torch.from_numpy(fx["answer_mask"]))
    tok.next_ids = (ids, mask)
    with torch.no_grad():
        max_ids = model(images.cuda(), ["q"] * ids.shape[0], answer, temperature=temp, train=False, inference='rank',
                        k_test=k)
    assert max_ids.cpu().tolist() == fx["max_ids"].tolist(), (max_ids.cpu().tolist(), fx["max_ids"].tolist())
    return {"case": "vqa", "max_ids": max_ids.cpu().tolist()}


def case_clip(_tok):
    """Reference clip/model.py:CLIP (its own __init__, encode_image / encode_text and build_model) with the tower
    classes swapped for the mirrors, against clip_blocks.npz / the oracle."""
    ref_shims.install_clip()
    import clip.model as cm
    from madtp_b200 import clip_model as mirror, synthetic
    from oracle import dtp_oracle as O
    for name in ("LayerNorm", "QuickGELU", "ResidualAttentionBlock", "Transformer", "VisionTransformer"):
        setattr(cm, name, getattr(mirror, name))
    layers = 4
    sd = synthetic.clip_state_dict(777, img_size=224, vision_layers=layers, text_layers=layers)
    model = cm.CLIP(512, 224, layers, 768, 16, 77, 49408, 512, 8, layers, True, None)
    assert type(model.visual).__module__ == "madtp_b200.clip_model"
    msg = model.load_state_dict(sd, strict=False)
    assert not msg.unexpected_keys, msg.unexpected_keys
    model = model.cuda().eval().float()
    images, text = synthetic.clip_inputs(2, 224, seed=0)
    space = sd["space_dict"]
    with torch.no_grad():
        img, _ = model.encode_image(images.cuda(), model.space_dict, 5.0)
        # text without pruning: after a prune the reference reads the EOT token at its ORIGINAL index of the pruned
        # sequence (clip/model.py:501), which depends on the implementation-defined order of topk(sorted=False)
        txt, _ = model.encode_text(text.cuda(), model.space_dict, 0)
        img_o, _ = O.clip_vision_forward(images, sd, "visual.", space, 5.0, layers, 12)
        txt_o, _ = O.clip_text_forward(text, sd, space, 0.0, layers, 8)
    gold = np.load(GOLDEN / "clip_blocks.npz")
    assert synthetic.tensor_digest(images, text, space) == str(gold["input_digest"])
    assert _rel(img, torch.from_numpy(gold["v_emb"])) < 5e-3          # the unmodified reference's embedding
    e_i, e_t = _rel(img, img_o), _rel(txt, txt_o)
    assert e_i < 5e-3 and e_t < 5e-3, (e_i, e_t)
    # the mirror's own build_model derives the same architecture from the checkpoint's shapes
    m2 = mirror.build_model(dict(sd), evaluate=True)
    assert m2.visual.transformer.layers == layers and m2.transformer.layers == layers and m2.context_length == 77
    assert m2.visual.conv1.weight.dtype == torch.float16          # convert_weights, as clip/model.py:713
    return {"case": "clip", "image_rel": e_i, "text_rel": e_t}


CASES = {"nlvr": case_nlvr, "checkpoint": case_checkpoint, "retrieval": case_retrieval, "vqa": case_vqa, "clip": case_clip}


def main(argv):
    if not torch.cuda.is_available():
        raise RuntimeError("oracle.dropin needs cuda:0 -- madtp_b200 has no CPU fallback")
    if not ref_shims.available():
        raise RuntimeError(f"reference tree not found at {ref_shims.REFERENCE_ROOT} (run python -m oracle.make_ref)")
    names = argv or list(CASES)
    tok = install_mirrors()
    for n in names:
        print(json.dumps(CASES[n](tok)), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
