"""Host-side logic that needs no GPU: the closed-form MAC accounting against the oracle's counter and the reference's
hard-coded total, the temperature controller / calibration, checkpoint loaders (key fan-out, position-grid resize
against the reference's own function), CLIP build_model, and the prepared-weight cache invalidation rules."""
import hashlib
import json
import os
import random
from pathlib import Path

import pytest
import torch

from madtp_b200 import flops, synthetic
from oracle import dtp_oracle as O
from oracle import ref_shims

ROOT = Path(__file__).resolve().parent.parent


# ---------------------------------------------------------------------------------------------------------------
# f3: analytic MACs + calibration
# ---------------------------------------------------------------------------------------------------------------
def _fake_traces(n0, rng, depth=12):
    traces, n = [], n0
    for _ in range(depth):
        t = O.PruneTrace()
        if rng.random() < 0.8 and n > 8:
            t.pruned, t.k = True, rng.randint(2, n - 3)
            n = t.k + 2
        else:
            t.pruned, t.k = False, n - 1
        traces.append(t)
    return traces


def test_macs_match_the_oracle_counter_on_random_trajectories():
    rng = random.Random(7)
    for _ in range(50):
        n0 = rng.choice([197, 442, 577, 901])
        L0 = rng.randint(8, 40)
        vt, tt = _fake_traces(n0, rng), _fake_traces(L0, rng)
        ks_v = [t.k if t.pruned else -1 for t in vt]
        ks_t = [t.k if t.pruned else -1 for t in tt]
        assert flops.vit_macs(n0, ks_v) == O.vit_macs_from_traces(vt, n0)
        ntr = O.NlvrTrace()
        ntr.vit, ntr.text = vt, tt
        n_img = flops.trajectory(n0, ks_v)[-1][1]
        ntr.image_embeds = torch.empty(1, n_img, 1)
        assert flops.nlvr_macs(n0, ks_v, L0, ks_t) == O.nlvr_macs_from_trace(ntr, n0, L0)
        assert flops.nlvr_macs_unpruned(n0, L0) == O.nlvr_macs_unpruned(n0, L0)


def test_unpruned_nlvr_total_matches_the_reference_constant():
    """compress_nlvr_dtp.py:162 hard-codes Ori_Gflops = 132.54 (fvcore GMACs, 2 x 384^2 images + the probe sentence of
    utils.py:297-299, ~22 word pieces)."""
    g = flops.nlvr_macs_unpruned(577, 22) / 1e9
    assert abs(g - 132.54) / 132.54 < 2e-3, g


def test_calibration_fixture_is_reproduced_by_the_closed_form():
    import numpy as np
    cal = np.load(ROOT / "tests" / "golden" / "calib_nlvr_p50_b32.npz")
    macs = flops.nlvr_macs(577, cal["vit_k"].tolist(), int(cal["text_len"]), cal["text_k"].tolist())
    assert macs == int(cal["macs_pruned"])
    assert flops.nlvr_macs_unpruned(577, int(cal["text_len"])) == int(cal["macs_unpruned"])


def test_temperature_controller_steps_like_the_reference():
    # compress_nlvr_dtp.py:174-201
    assert flops.temperature_step(1.0, 132.0, 66.0) == 2.0          # gap > 30
    assert flops.temperature_step(1.0, 80.0, 66.0) == 1.5           # gap > 10
    assert flops.temperature_step(1.0, 72.0, 66.0) == 1.25          # gap > 5
    assert abs(flops.temperature_step(1.0, 68.0, 66.0) - 1.1) < 1e-12
    assert abs(flops.temperature_step(1.0, 66.5, 66.0) - 1.01) < 1e-12
    assert flops.temperature_step(3.0, 30.0, 66.0) == 2.0
    assert abs(flops.temperature_step(3.0, 65.5, 66.0) - 2.99) < 1e-12


def test_calibrate_temperature_bisection():
    import math
    ratio = lambda t: 1.0 / (1.0 + 0.3 * math.log1p(t))            # monotone decreasing stand-in for a forward
    t, r, probes = flops.calibrate_temperature(ratio, p=0.25, tol=1e-4, max_iter=40)
    assert abs(r - 0.75) < 1e-4 and abs(ratio(t) - r) < 1e-12 and probes <= 40


# ---------------------------------------------------------------------------------------------------------------
# f4: checkpoint loaders
# ---------------------------------------------------------------------------------------------------------------
def _pretrained_blip_like(seed, img):
    full = synthetic.blip_nlvr_state_dict(seed, img_size=img)
    pre = {}
    for k, v in full.items():
        if "crossattention.self1." in k or "crossattention.output.dense1." in k or "merge_layer" in k:
            continue
        pre[k.replace("crossattention.self0.", "crossattention.self.")
             .replace("crossattention.output.dense0.", "crossattention.output.dense.")] = v
    return pre


def test_nlvr_checkpoint_fan_out_and_pos_embed_resize(tmp_path):
    from madtp_b200.blip_nlvr import BLIP_NLVR, blip_nlvr
    from madtp_b200.checkpoint import load_compressed_checkpoint
    pre = _pretrained_blip_like(5, 224)
    path = tmp_path / "pre.pth"
    torch.save({"model": pre, "epoch": 1, "temperature": 1.75}, path)
    model = blip_nlvr(pretrained=str(path), image_size=384, evaluate=True)
    sd = model.state_dict()
    for i in range(12):
        p = f"text_encoder.encoder.layer.{i}.crossattention."
        for leaf in ("query.weight", "key.bias", "value.weight"):
            src = pre[p + "self." + leaf]
            assert torch.equal(sd[p + "self0." + leaf], src) and torch.equal(sd[p + "self1." + leaf], src)
        src = pre[p + "output.dense.weight"]
        assert torch.equal(sd[p + "output.dense0.weight"], src) and torch.equal(sd[p + "output.dense1.weight"], src)
    assert sd["visual_encoder.pos_embed"].shape == (1, 577, 768)
    assert torch.equal(sd["visual_encoder.pos_embed"][:, 0], pre["visual_encoder.pos_embed"][:, 0])     # CLS row untouched
    if ref_shims.available():       # against the reference's own function (models/vit.py:398-422)
        import importlib.util
        ref_shims.install()
        spec = importlib.util.spec_from_file_location("_ref_vit_real", os.path.join(ref_shims.REFERENCE_ROOT, "models/vit.py"))
        rv = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(rv)
        want = rv.interpolate_pos_embed(pre["visual_encoder.pos_embed"].clone(), model.visual_encoder)
        assert torch.equal(sd["visual_encoder.pos_embed"], want)
    # compressed checkpoint round trip: {'model', 'epoch', 'temperature'} (compress_nlvr_dtp.py:229-236, :153-158)
    path2 = tmp_path / "best.pth"
    torch.save({"model": model.state_dict(), "epoch": 4, "temperature": 3.25}, path2)
    m2 = BLIP_NLVR(image_size=384, evaluate=True)
    msg, temp = load_compressed_checkpoint(m2, str(path2))
    assert not msg.missing_keys and not msg.unexpected_keys and temp == 3.25
    assert all(torch.equal(a, b) for a, b in zip(m2.state_dict().values(), model.state_dict().values()))


def test_blip_checkpoint_loader_drops_mismatched_shapes(tmp_path):
    from madtp_b200.blip_retrieval import BLIP_Retrieval, BLIP_VQA, blip_retrieval, load_checkpoint
    sd = synthetic.retrieval_state_dict(11, img_size=224)
    sd["itm_head.weight"] = torch.zeros(3, 768)                      # wrong shape: must be dropped, not raise (:274-276)
    sd["visual_encoder_m.pos_embed"] = sd["visual_encoder.pos_embed"].clone()     # momentum twin key: unexpected here
    path = tmp_path / "ret.pth"
    torch.save({"model": sd}, path)
    model = blip_retrieval(pretrained=str(path), image_size=384, evaluate=True)
    assert model.visual_encoder.pos_embed.shape == (1, 577, 768)
    assert model.itm_head.weight.shape == (2, 768)
    m2 = BLIP_Retrieval(image_size=384, evaluate=True)
    _, msg = load_checkpoint(m2, {"model": sd})
    assert "itm_head.weight" in msg.missing_keys and "visual_encoder_m.pos_embed" in msg.unexpected_keys
    with pytest.raises(RuntimeError):
        load_checkpoint(m2, str(tmp_path / "missing.pth"))
    vq = synthetic.vqa_state_dict(3, img_size=224)
    m3 = BLIP_VQA(image_size=480, evaluate=True)
    _, msg = load_checkpoint(m3, vq)
    assert m3.visual_encoder.pos_embed.shape == (1, 901, 768) and not msg.missing_keys


def test_clip_build_model_and_constructor_order():
    from madtp_b200.clip_model import CLIP, build_model
    sd = synthetic.clip_state_dict(9, img_size=224, vision_layers=2, text_layers=3)
    m = build_model(dict(sd), evaluate=True)
    assert (m.visual.input_resolution, m.visual.transformer.layers, m.transformer.layers, m.context_length,
            m.vocab_size) == (224, 2, 3, 77, 49408)
    # convert_weights (clip/model.py:655-676): linear / conv / in_proj / projections in fp16, LayerNorm and embeddings not
    blk = m.visual.transformer.resblocks[0]
    assert blk.attn.in_proj_weight.dtype == torch.float16 and blk.mlp.c_fc.weight.dtype == torch.float16
    assert m.visual.proj.dtype == torch.float16 and m.text_projection.dtype == torch.float16
    assert blk.ln_1.weight.dtype == torch.float32 and m.token_embedding.weight.dtype == torch.float32
    m = m.float()                                                    # clip/clip.py:148
    assert torch.equal(blk.attn.in_proj_weight, sd["visual.transformer.resblocks.0.attn.in_proj_weight"].half().float())
    # the reference's positional order: (..., transformer_layers, evaluate, config)  (clip/model.py:317-332)
    m2 = CLIP(512, 224, 1, 768, 16, 77, 49408, 512, 8, 1, True, {"sd_num": 50, "sd_dim": 768})
    assert m2.space_dict.shape == (50, 768)
    with pytest.raises(TypeError):
        CLIP(512, 224, 1, 768, 16, 77, 49408, 512, 8, 1, {"sd_num": 50, "sd_dim": 768})


# ---------------------------------------------------------------------------------------------------------------
# prepared-weight cache
# ---------------------------------------------------------------------------------------------------------------
def test_weight_cache_invalidation_rules():
    from madtp_b200.functional import WeightCache, clear_caches
    builds = [0]

    def build():
        builds[0] += 1
        return builds[0]
    p = torch.nn.Parameter(torch.zeros(8))
    c = WeightCache()
    assert c.get("w", [p], build) == 1 and c.get("w", [p], build) == 1
    with torch.no_grad():
        p.copy_(torch.ones(8))                       # in-place through the parameter: version moves
    assert c.get("w", [p], build) == 2
    with torch.no_grad():
        p.detach().copy_(torch.zeros(8))             # detach() shares the version counter (dist.broadcast_parameters)
    assert c.get("w", [p], build) == 3
    # a different tensor object with identical (pointer, version, shape) must miss: simulate by aliasing storage
    q = p.detach().view(8)
    assert q.data_ptr() == p.data_ptr() and q._version == p._version
    assert c.get("w", [q], build) == 4
    holder = torch.nn.Linear(2, 2)
    holder._cache = c
    assert clear_caches(holder) == 1 and c.get("w", [q], build) == 5


def test_reference_staging_manifest_is_byte_exact():
    """oracle/_ref is a byte-for-byte staging of the reference's hot-path files (oracle/make_ref.py)."""
    man = ROOT / "oracle" / "_ref" / "MANIFEST.json"
    if not man.exists():
        pytest.skip("oracle/_ref not staged (run python -m oracle.make_ref where /root/reference exists)")
    files = json.loads(man.read_text())["files"]
    assert "models/vit.py" in files and "models/nlvr_encoder.py" in files and "clip/model.py" in files
    for rel, digest in files.items():
        assert hashlib.sha256((man.parent / rel).read_bytes()).hexdigest() == digest, rel
        src = Path("/root/reference") / rel
        if src.exists():
            assert src.read_bytes() == (man.parent / rel).read_bytes(), rel
