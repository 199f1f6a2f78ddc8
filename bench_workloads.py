"""The five BASELINE.json configurations as bench.py workloads.

    1  single ViT-B/16 Block + DTP head, batch 2, 197 tokens            (models/vit.py:183-207, models/utils.py:147-183)
    2  BLIP-NLVR forward, 32 pairs of 384 x 384, p = 0.5                 (models/blip_nlvr.py:63-81)   <- the headline
    3  BLIP retrieval evaluation path, 64 x 384 x 384 + 35-token text, p = 0.75
                                                                         (compress_retrieval_dtp.py:104,120,170-177)
    4  CLIP ViT-B/16 towers, 64 x 336 x 336 + 77-token text per GPU, p = 0.5   (clip/model.py:482-503)
    5  BLIP-VQA image + question encoders, 64 x 480 x 480 per GPU, p = 0.5     (models/blip_vqa.py:60,119-125)

Every workload provides: the GPU model (module mirrors -> C ABI), pinned-host inputs, one step on device inputs, the
algorithmic FLOPs of a step (closed form of the ORACLE's pruning trajectory, madtp_b200/flops.py -- extra passes are
not credited), the agreement of the GPU run with the oracle fixture, and the CPU arm: the UNMODIFIED reference modules
(oracle/_ref staging) on the host cores, or the oracle port when the reference tree is absent.
Only the `cpu_reference` methods touch `oracle/` (the checker / CPU-baseline leg bench.py is allowed to run).
"""
from __future__ import annotations

import os
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
GOLDEN = ROOT / "tests" / "golden"


def _time_cpu(fwd, steps, warmup):
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            fwd()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def _ks(blocks):
    return [(b.last_prune.k if (b.last_prune is not None and b.last_prune.pruned) else -1) for b in blocks]


def _keeps(blocks):
    return [(b.last_prune.keep.cpu().numpy() if (b.last_prune is not None and b.last_prune.pruned) else None)
            for b in blocks]


def mask_agreement(fix, prefix, ks_oracle, keeps_gpu, B, n0):
    """Per layer: fraction of the oracle's surviving ORIGINAL patches that the (free-running) GPU pass also kept. Tokens
    are tracked back to their patch index through every prune; merged tokens are not counted."""
    ids_o = np.tile(np.arange(n0), (B, 1))
    ids_g = ids_o.copy()
    agree = []
    for i, k_o in enumerate(ks_oracle):
        if k_o >= 0 and f"{prefix}{i}_keep" in fix:
            ko = np.unpackbits(fix[f"{prefix}{i}_keep"], axis=1)[:, :ids_o.shape[1]].astype(bool)
            ids_o = np.stack([np.concatenate([ids_o[b][ko[b]], [-1]]) for b in range(B)])
        kg = keeps_gpu[i]
        if kg is not None:
            kg = kg.astype(bool)
            ids_g = np.stack([np.concatenate([ids_g[b][kg[b]], [-1]]) for b in range(B)])
        fr = []
        for b in range(B):
            so, sg = set(ids_o[b].tolist()) - {-1}, set(ids_g[b].tolist()) - {-1}
            fr.append(len(so & sg) / max(len(so), 1))
        agree.append(round(float(np.mean(fr)), 5))
    return agree


class Workload:
    config = 0
    metric = ""
    workload = ""
    units = 0                      # images per step and rank
    graphable = False

    def __init__(self):
        self.model = None
        self.temperature = 0.0

    # --- to override ------------------------------------------------------------------------------------------
    def build(self, dev, rank):
        raise NotImplementedError

    def host_inputs(self, rank):
        raise NotImplementedError

    def step(self, inputs):
        raise NotImplementedError

    def step_flops(self):
        raise NotImplementedError

    def config_extra(self):
        return {}

    def parity(self, result):
        return None

    def cpu_reference(self, sample, steps, warmup):
        raise NotImplementedError

    def enable_graphs(self, flag):
        if self.graphable:
            self.model.enable_cuda_graphs(flag)
            return flag
        return False

    def suspend_graphs(self):
        """Python-issued launches for a while WITHOUT dropping the captured graphs; returns a token for resume_graphs."""
        if not self.graphable:
            return None
        held, self.model._graphs = self.model._graphs, None
        return held

    def resume_graphs(self, token):
        if self.graphable:
            self.model._graphs = token

    def graph_launches(self):
        if not self.graphable:
            return 0
        return sum(g.launches * g.replays for g in (self.model._graphs or {}).values())


# ---------------------------------------------------------------------------------------------------------------
# 2: BLIP-NLVR (the headline)
# ---------------------------------------------------------------------------------------------------------------
class NlvrWorkload(Workload):
    config = 2
    PAIRS, IMAGE, TEXT_LEN = 32, 384, 20
    metric = "BLIP-NLVR p=0.5 forward images/sec"
    workload = "BLIP-NLVR forward (vit.py + nlvr_encoder.py), 384x384 synthetic pairs, p=0.5, batch=32 pairs/GPU"
    units = 64
    graphable = True
    CALIB = GOLDEN / "calib_nlvr_p50_b32.npz"

    def __init__(self):
        super().__init__()
        self.cal = np.load(self.CALIB)
        self.temperature = float(self.cal["temperature"])

    def build(self, dev, rank):
        from madtp_b200 import synthetic
        from madtp_b200.blip_nlvr import BLIP_NLVR
        model = BLIP_NLVR(image_size=self.IMAGE, evaluate=True)
        if rank == 0:
            msg = model.load_state_dict(synthetic.blip_nlvr_state_dict(1234, img_size=self.IMAGE), strict=False)
            assert not msg.missing_keys and not msg.unexpected_keys
        self.model = model.to(dev).eval()
        return self.model

    def host_inputs(self, rank):
        from madtp_b200 import synthetic
        return synthetic.nlvr_inputs(self.PAIRS, self.IMAGE, self.TEXT_LEN, seed=rank)      # this rank's own pairs

    def step(self, inputs):
        from madtp_b200.blip_nlvr import TokenizedText
        images, ids, mask = inputs
        return self.model(images, TokenizedText(ids, mask), self.PAIRS, self.temperature, train=False)

    def step_flops(self):
        return 2.0 * int(self.cal["macs_pruned"]) * self.PAIRS

    def config_extra(self):
        return {"pairs_per_gpu": self.PAIRS, "image_size": self.IMAGE, "text_len": self.TEXT_LEN,
                "mac_ratio_oracle": float(self.cal["ratio"]), "vit_topk_per_layer": _ks(self.model.visual_encoder.blocks),
                "vit_topk_oracle": self.cal["vit_k"].tolist()}

    def parity(self, result):
        c = self.cal
        blocks = self.model.visual_encoder.blocks
        pred = result[:self.PAIRS].float().cpu().numpy()
        return {"mask_agreement_per_layer": mask_agreement(c, "vit", c["vit_k"].tolist(), _keeps(blocks), 2 * self.PAIRS,
                                                           (self.IMAGE // 16) ** 2),
                "k_gpu": _ks(blocks), "k_oracle": c["vit_k"].tolist(),
                "text_k_gpu": _ks(self.model.text_encoder.encoder.layer), "text_k_oracle": c["text_k"].tolist(),
                "logit_max_abs": float(np.abs(pred - c["pred"]).max()),
                "argmax_agreement": float((pred.argmax(1) == c["pred"].argmax(1)).mean()),
                "what": "free-running (no teacher forcing) vs tests/golden/calib_nlvr_p50_b32.npz; layer 0 sees "
                        "identical inputs and must read 1.0; the bit-exact teacher-forced gates are tests/test_parity_gpu.py"}

    def cpu_reference(self, sample, steps, warmup):
        from madtp_b200 import synthetic
        from oracle import ref_shims
        pairs = sample or self.PAIRS
        sd = synthetic.blip_nlvr_state_dict(1234, img_size=self.IMAGE)
        images, ids, mask = synthetic.nlvr_inputs(pairs, self.IMAGE, self.TEXT_LEN, seed=0)
        if ref_shims.available():
            kind = "reference"
            model, tok = ref_shims.build_blip_nlvr(self.IMAGE)
            msg = model.load_state_dict(sd, strict=False)
            assert not msg.unexpected_keys, msg.unexpected_keys
            targets = torch.zeros(pairs, dtype=torch.long)

            def fwd():
                tok.next_ids = (ids, mask)
                return model(images, ["x"] * pairs, targets, self.temperature, train=False)
        else:
            kind = "port"
            from oracle import dtp_oracle as O

            def fwd():
                return O.blip_nlvr_forward(images, ids, mask, sd, self.temperature)
        sec = _time_cpu(fwd, steps, warmup)
        what = ("the unmodified reference BLIP_NLVR.forward(train=False) (models/blip_nlvr.py:63-100)" if kind == "reference"
                else "oracle/dtp_oracle.py blip_nlvr_forward (reference tree absent)")
        return 2 * pairs / sec, sec, kind, f"{pairs} pairs ({2 * pairs} images) of the {self.PAIRS}-pair batch per step, {what}"


# ---------------------------------------------------------------------------------------------------------------
# 1: single ViT-B/16 Block + DTP head
# ---------------------------------------------------------------------------------------------------------------
class BlockWorkload(Workload):
    config = 1
    metric = "ViT-B/16 Block + DTP head forward images/sec (batch 2, 197 tokens)"
    workload = "single ViT-B/16 Block + DTP head (vit.py Block + utils.py Query_model), batch=2, 197 tokens"
    units = 2
    TI = 1                          # tests/golden/block_cfg1.npz case: temperature 5.0

    def __init__(self):
        super().__init__()
        self.gold = np.load(GOLDEN / "block_cfg1.npz")
        self.temperature = float(self.gold["temps"][self.TI])

    def build(self, dev, rank):
        from functools import partial
        from madtp_b200 import synthetic
        from madtp_b200.utils import Query_model
        from madtp_b200.vit import Block
        blk = Block(768, 12, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
        if rank == 0:
            blk.load_state_dict(synthetic.block_state_dict(1234), strict=True)
        self.model = torch.nn.ModuleDict({"blk": blk, "qm": Query_model(768, 768)}).to(dev).eval()
        return self.model

    def host_inputs(self, rank):
        from madtp_b200 import synthetic
        return synthetic.block_inputs()

    def step(self, inputs):
        x, space = inputs
        token_attn, _, _ = self.model["qm"](x[:, 1:, :], space, return_token_att=True)
        return self.model["blk"](x, False, 0, self.temperature, token_attn)

    def step_flops(self):
        from madtp_b200 import flops
        return 2.0 * 2 * flops.vit_layer_macs(197, int(self.gold[f"t{self.TI}_k"]) + 2)

    def config_extra(self):
        r = self.model["blk"].last_prune
        return {"tokens": 197, "k": None if r is None else r.k, "k_reference": int(self.gold[f"t{self.TI}_k"])}

    def parity(self, result):
        g, p = self.gold, f"t{self.TI}_"
        r = self.model["blk"].last_prune
        keep_ref = np.unpackbits(g[p + "keep"], axis=1)[:, :196].astype(bool)
        out_ref = torch.from_numpy(g[p + "out_s4"])
        out = result.float().cpu()[:, :, ::4]
        return {"keep_mask_bit_exact": bool(np.array_equal(r.keep.cpu().numpy().astype(bool), keep_ref)),
                "k_gpu": r.k, "k_reference": int(g[p + "k"]),
                "hidden_rel_err": float(((out.double() - out_ref.double()).norm() / out_ref.double().norm())),
                "what": "vs tests/golden/block_cfg1.npz, generated from the unmodified reference vit.Block"}

    def cpu_reference(self, sample, steps, warmup):
        from functools import partial
        from madtp_b200 import synthetic
        from oracle import ref_shims
        x, space = synthetic.block_inputs()
        sd = synthetic.block_state_dict(1234)
        if ref_shims.available():
            kind = "reference"
            ref_shims.install()
            import models.vit as rvit
            from models.utils import Query_model
            blk = rvit.Block(768, 12, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
            blk.load_state_dict(sd, strict=True)
            blk.eval()
            qm = Query_model(768, 768)

            def fwd():
                ta, _, _ = qm(x[:, 1:, :], space, return_token_att=True)
                return blk(x, False, 0, self.temperature, ta.clone())
        else:
            kind = "port"
            from oracle import dtp_oracle as O
            sdb = {"b." + k: v for k, v in sd.items()}

            def fwd():
                ta, _ = O.query_model(x[:, 1:, :], space, 768)
                return O.vit_block(x, sdb, "b", 12, self.temperature, ta)
        sec = _time_cpu(fwd, max(steps, 20), max(warmup, 3))
        return 2 / sec, sec, kind, ("the whole workload (batch 2), " +
                                    ("unmodified reference vit.Block + Query_model" if kind == "reference" else "oracle port"))


# ---------------------------------------------------------------------------------------------------------------
# 3: BLIP retrieval evaluation path
# ---------------------------------------------------------------------------------------------------------------
class RetrievalWorkload(Workload):
    config = 3
    B, IMAGE, TEXT_LEN = 64, 384, 35
    metric = "BLIP retrieval p=0.75 evaluation-path images/sec"
    workload = ("BLIP retrieval evaluation path (blip_retrieval.py encoders: ViT + text encoder mode 'text' + one multimodal "
                "ITM pass + itm_head), 384x384, text padded to 35, p=0.75, batch=64/GPU")
    units = 64
    graphable = True
    CALIB = GOLDEN / "calib_retrieval_p75_b64.npz"

    def __init__(self):
        super().__init__()
        self.cal = np.load(self.CALIB)
        self.temperature = float(self.cal["temperature"])

    def build(self, dev, rank):
        from madtp_b200 import synthetic
        from madtp_b200.blip_retrieval import BLIP_Retrieval
        model = BLIP_Retrieval(image_size=self.IMAGE, evaluate=True)
        if rank == 0:
            msg = model.load_state_dict(synthetic.retrieval_state_dict(4321, img_size=self.IMAGE), strict=False)
            assert not msg.unexpected_keys
        self.model = model.to(dev).eval()
        return self.model

    def host_inputs(self, rank):
        from madtp_b200 import synthetic
        return synthetic.retrieval_inputs(self.B, self.IMAGE, self.TEXT_LEN, seed=rank)

    def step(self, inputs):
        images, ids, mask = inputs
        _, itm = self.model(images, (ids, mask), 0.0, None, self.temperature, train=False)
        return itm

    def step_flops(self):
        return 2.0 * int(self.cal["macs_pruned"]) * self.B

    def config_extra(self):
        return {"batch_per_gpu": self.B, "image_size": self.IMAGE, "text_len": self.TEXT_LEN,
                "mac_ratio_oracle": float(self.cal["ratio"]), "vit_topk_per_layer": _ks(self.model.visual_encoder.blocks),
                "vit_topk_oracle": self.cal["vit_k"].tolist()}

    def parity(self, result):
        c = self.cal
        blocks = self.model.visual_encoder.blocks
        itm = result[:self.B].float().cpu().numpy()
        return {"mask_agreement_per_layer": mask_agreement(c, "vit", c["vit_k"].tolist(), _keeps(blocks), self.B,
                                                           (self.IMAGE // 16) ** 2),
                "k_gpu": _ks(blocks), "k_oracle": c["vit_k"].tolist(),
                "mm_k_gpu": _ks(self.model.text_encoder.encoder.layer), "mm_k_oracle": c["mm_k"].tolist(),
                "itm_logit_max_abs": float(np.abs(itm - c["itm"]).max()),
                "what": "free-running vs tests/golden/calib_retrieval_p75_b64.npz (oracle trajectory)"}

    def cpu_reference(self, sample, steps, warmup):
        from madtp_b200 import synthetic
        from oracle import ref_shims
        B = sample or self.B
        sd = synthetic.retrieval_state_dict(4321, img_size=self.IMAGE)
        images, ids, mask = synthetic.retrieval_inputs(B, self.IMAGE, self.TEXT_LEN, seed=0)
        ids2 = ids.clone()
        ids2[:, 0] = 30523
        t = self.temperature
        if ref_shims.available():
            kind = "reference"
            ref_shims.install()
            import models.blip as blip
            tok = ref_shims.FakeTokenizer()
            blip.init_tokenizer = lambda: tok
            import models.blip_retrieval as br
            br.init_tokenizer = lambda: tok
            model = br.BLIP_Retrieval(med_config=os.path.join(ref_shims.REFERENCE_ROOT, "configs/med_config.json"),
                                      image_size=self.IMAGE, vit="base", evaluate=True)
            model.load_state_dict(sd, strict=False)
            model.eval()

            def fwd():   # compress_retrieval_dtp.py:104,120,170-177
                model.text_encoder(ids, attention_mask=mask, mode='text', space_dict=model.space_dict, temperature=t)
                feat, _ = model.visual_encoder(images, space_dict=model.space_dict, temperature=t)
                atts = torch.ones(feat.size()[:-1], dtype=torch.long)
                mm = model.text_encoder(ids2, attention_mask=mask, encoder_hidden_states=feat, encoder_attention_mask=atts,
                                        return_dict=True, space_dict=model.space_dict, temperature=t)[0]
                return model.itm_head(mm.last_hidden_state[:, 0, :])
        else:
            kind = "port"
            from oracle import dtp_oracle as O
            space = sd["space_dict"]

            def fwd():
                feat, _ = O.vit_forward(images, sd, "visual_encoder.", space, t)
                O.med_text_encoder(ids, mask, sd, "text_encoder.", None, space, t, "text")
                mm, _ = O.med_text_encoder(ids2, mask, sd, "text_encoder.", feat, space, t, "multimodal")
                return O.linear(mm[:, 0, :], sd, "itm_head")
        sec = _time_cpu(fwd, steps, warmup)
        return B / sec, sec, kind, (f"{B} of the {self.B} image-text pairs per step, " +
                                    ("unmodified reference BLIP_Retrieval encoders + itm_head" if kind == "reference"
                                     else "oracle port"))


# ---------------------------------------------------------------------------------------------------------------
# 5: BLIP-VQA image + question encoders
# ---------------------------------------------------------------------------------------------------------------
class VqaWorkload(Workload):
    config = 5
    B, IMAGE = 64, 480
    metric = "BLIP-VQA p=0.5 image+question encoder images/sec"
    workload = ("BLIP-VQA encoders (blip_vqa.py: ViT at 480x480 = 901 tokens + question encoder with med.py cross-attention "
                "over the pruned image tokens), p=0.5, batch=64/GPU")
    units = 64
    graphable = True
    CALIB = GOLDEN / "calib_vqa_p50_b64.npz"

    def __init__(self):
        super().__init__()
        self.cal = np.load(self.CALIB)
        self.temperature = float(self.cal["temperature"])
        self.text_len = int(self.cal["text_len"])

    def _inputs(self, B, seed):
        from madtp_b200 import synthetic
        images, ids, mask = synthetic.retrieval_inputs(B, self.IMAGE, 20, seed=seed)
        return images, ids[:, :self.text_len].contiguous(), mask[:, :self.text_len].contiguous()

    def build(self, dev, rank):
        from madtp_b200 import synthetic
        from madtp_b200.blip_retrieval import BLIP_VQA
        model = BLIP_VQA(image_size=self.IMAGE, evaluate=True)
        if rank == 0:
            model.load_state_dict(synthetic.vqa_state_dict(99, img_size=self.IMAGE), strict=False)
        self.model = model.to(dev).eval()
        return self.model

    def host_inputs(self, rank):
        return self._inputs(self.B, 2 + 100 * rank)

    def step(self, inputs):
        from madtp_b200.vit import device_lengths_enabled
        images, ids, mask = inputs
        if device_lengths_enabled():     # device-resident lengths end to end (a CUDA graph replay when enabled)
            return self.model.encode_question_packed(images, ids, mask, self.temperature)[2]
        q, _ = self.model.encode_question(images, ids, mask, self.temperature)
        return q[:, 0, :].contiguous()

    def step_flops(self):
        return 2.0 * int(self.cal["macs_pruned"]) * self.B

    def config_extra(self):
        return {"batch_per_gpu": self.B, "image_size": self.IMAGE, "text_len": self.text_len,
                "mac_ratio_oracle": float(self.cal["ratio"]), "vit_topk_per_layer": _ks(self.model.visual_encoder.blocks),
                "vit_topk_oracle": self.cal["vit_k"].tolist()}

    def parity(self, result):
        c = self.cal
        blocks = self.model.visual_encoder.blocks
        q = result[:self.B].float().cpu()
        ref = torch.from_numpy(c["question_cls"])
        return {"mask_agreement_per_layer": mask_agreement(c, "vit", c["vit_k"].tolist(), _keeps(blocks), self.B,
                                                           (self.IMAGE // 16) ** 2),
                "k_gpu": _ks(blocks), "k_oracle": c["vit_k"].tolist(),
                "mm_k_gpu": _ks(self.model.text_encoder.encoder.layer), "mm_k_oracle": c["mm_k"].tolist(),
                "question_cls_rel_err": float((q.double() - ref.double()).norm() / ref.double().norm()),
                "what": "free-running vs tests/golden/calib_vqa_p50_b64.npz (oracle trajectory)"}

    def cpu_reference(self, sample, steps, warmup):
        from madtp_b200 import synthetic
        from oracle import ref_shims
        B = sample or self.B
        sd = synthetic.vqa_state_dict(99, img_size=self.IMAGE)
        images, ids, mask = self._inputs(B, 2)
        ids2 = ids.clone()
        ids2[:, 0] = 30523
        t = self.temperature
        if ref_shims.available():
            kind = "reference"
            ref_shims.install()
            import models.blip as blip
            tok = ref_shims.FakeTokenizer()
            blip.init_tokenizer = lambda: tok
            import models.blip_vqa as bv
            bv.init_tokenizer = lambda: tok
            model = bv.BLIP_VQA(med_config=os.path.join(ref_shims.REFERENCE_ROOT, "configs/med_config.json"),
                                image_size=self.IMAGE, vit="base", evaluate=True)
            model.load_state_dict(sd, strict=False)
            model.eval()

            def fwd():   # models/blip_vqa.py:60,119-125
                emb, _ = model.visual_encoder(images, space_dict=model.space_dict, temperature=t)
                atts = torch.ones(emb.size()[:-1], dtype=torch.long)
                return model.text_encoder(ids2, attention_mask=mask, encoder_hidden_states=emb,
                                          encoder_attention_mask=atts, return_dict=True, space_dict=model.space_dict,
                                          temperature=t)[0]
        else:
            kind = "port"
            from oracle import dtp_oracle as O
            space = sd["space_dict"]

            def fwd():
                feat, _ = O.vit_forward(images, sd, "visual_encoder.", space, t)
                return O.med_text_encoder(ids2, mask, sd, "text_encoder.", feat, space, t, "multimodal")
        sec = _time_cpu(fwd, steps, warmup)
        return B / sec, sec, kind, (f"{B} of the {self.B} (image, question) pairs per step, " +
                                    ("unmodified reference BLIP_VQA encoders" if kind == "reference" else "oracle port"))


# ---------------------------------------------------------------------------------------------------------------
# 4: CLIP ViT-B/16 towers
# ---------------------------------------------------------------------------------------------------------------
class ClipWorkload(Workload):
    config = 4
    B, IMAGE = 64, 336
    metric = "CLIP ViT-B/16 p=0.5 encode_image+encode_text images/sec"
    workload = ("CLIP ViT-B/16 towers (clip/model.py encode_image + encode_text, retrieval_flickr_clip.yaml image size 336 = 442 "
                "tokens, context 77), p=0.5, batch=64/GPU")
    units = 64
    graphable = True
    CALIB = GOLDEN / "calib_clip_p50_b64_r336.npz"

    def __init__(self):
        super().__init__()
        self.cal = np.load(self.CALIB)
        self.temperature = float(self.cal["temperature"])

    def build(self, dev, rank):
        from madtp_b200 import synthetic
        from madtp_b200.clip_model import CLIP
        model = CLIP(512, self.IMAGE, 12, 768, 16, 77, 49408, 512, 8, 12, True, None)
        if rank == 0:
            msg = model.load_state_dict(synthetic.clip_state_dict(777, img_size=self.IMAGE), strict=False)
            assert not msg.unexpected_keys
        self.model = model.to(dev).eval()
        return self.model

    def host_inputs(self, rank):
        from madtp_b200 import synthetic
        return synthetic.clip_inputs(self.B, self.IMAGE, seed=rank)

    def step(self, inputs):
        images, text = inputs
        img, _ = self.model.encode_image(images, self.model.space_dict, self.temperature)
        txt, _ = self.model.encode_text(text, self.model.space_dict, self.temperature)
        return torch.cat([img, txt], dim=1)

    def step_flops(self):
        return 2.0 * int(self.cal["macs_pruned"]) * self.B

    def config_extra(self):
        return {"batch_per_gpu": self.B, "image_size": self.IMAGE, "context": 77,
                "mac_ratio_oracle": float(self.cal["ratio"]),
                "vision_topk_per_layer": _ks(self.model.visual.transformer.resblocks),
                "vision_topk_oracle": self.cal["vision_k"].tolist()}

    def parity(self, result):
        c = self.cal
        blocks = self.model.visual.transformer.resblocks
        img = result[:self.B, :512].float().cpu()
        ref = torch.from_numpy(c["image_emb"])
        return {"mask_agreement_per_layer": mask_agreement(c, "vit", c["vision_k"].tolist(), _keeps(blocks), self.B,
                                                           (self.IMAGE // 16) ** 2),
                "k_gpu": _ks(blocks), "k_oracle": c["vision_k"].tolist(),
                "text_k_gpu": _ks(self.model.transformer.resblocks), "text_k_oracle": c["text_k"].tolist(),
                "image_emb_rel_err": float((img.double() - ref.double()).norm() / ref.double().norm()),
                "what": "free-running vs tests/golden/calib_clip_p50_b64_r336.npz (oracle trajectory); the text embedding "
                        "is not compared: after a prune the reference reads the EOT token at an index that depends on the "
                        "implementation-defined order of topk(sorted=False) (clip/model.py:501)"}

    def cpu_reference(self, sample, steps, warmup):
        from madtp_b200 import synthetic
        from oracle import ref_shims
        B = sample or self.B
        sd = synthetic.clip_state_dict(777, img_size=self.IMAGE)
        images, text = synthetic.clip_inputs(B, self.IMAGE, seed=0)
        t = self.temperature
        if ref_shims.available():
            kind = "reference"
            ref_shims.install_clip()
            import clip.model as cm
            model = cm.CLIP(512, self.IMAGE, 12, 768, 16, 77, 49408, 512, 8, 12, True, None)
            model.load_state_dict(sd, strict=False)
            model.eval().float()

            def fwd():
                model.encode_image(images, model.space_dict, t)
                return model.encode_text(text, model.space_dict, t)
        else:
            kind = "port"
            from oracle import dtp_oracle as O
            space = sd["space_dict"]

            def fwd():
                O.clip_vision_forward(images, sd, "visual.", space, t, 12, 12)
                return O.clip_text_forward(text, sd, space, t, 12, 8)
        sec = _time_cpu(fwd, steps, warmup)
        return B / sec, sec, kind, (f"{B} of the {self.B} (image, caption) pairs per step, " +
                                    ("unmodified reference clip.model.CLIP towers" if kind == "reference" else "oracle port"))


WORKLOADS = {1: BlockWorkload, 2: NlvrWorkload, 3: RetrievalWorkload, 4: ClipWorkload, 5: VqaWorkload}
