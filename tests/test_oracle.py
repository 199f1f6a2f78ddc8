"""CPU tests: the oracle (oracle/dtp_oracle.py) against the golden fixtures generated from the UNMODIFIED reference
(oracle/gen_golden.py), plus the restated quirks of the reference the fixtures cannot cover."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import dtp_oracle as O
from oracle import weights

GOLDEN = Path(__file__).resolve().parent / "golden"


def unpack(bits, n):
    return torch.from_numpy(np.unpackbits(bits, axis=1)[:, :n].astype(bool))


def ambiguous_only(keep_a, keep_b, score, k, margin=1e-8):
    """True if every disagreement between two keep-masks sits on a top-k boundary narrower than `margin`."""
    if torch.equal(keep_a, keep_b):
        return True
    srt = score.sort(dim=1, descending=True)[0]
    for b in (keep_a != keep_b).any(dim=1).nonzero().flatten().tolist():
        if (srt[b, k - 1] - srt[b, k]).item() > margin:
            return False
    return True


@pytest.mark.parametrize("ti", [0, 1, 2])
def test_oracle_block_config1_matches_reference_fixture(ti):
    gold = np.load(GOLDEN / "block_cfg1.npz")
    temp = float(gold["temps"][ti])
    x, space = weights.block_inputs()
    assert weights.tensor_digest(x, space) == str(gold["input_digest"]), "seeded inputs drifted from the fixture"
    sd = {"b." + k: v for k, v in weights.block_state_dict(1234).items()}
    torch.set_num_threads(1)
    with torch.no_grad():
        ta, sd_ft = O.query_model(x[:, 1:, :], space, 768)
        tr = O.PruneTrace()
        y = O.vit_block(x, sd, "b", 12, temp, ta, tr)
    k = int(gold[f"t{ti}_k"])
    assert tr.pruned and tr.k == k
    assert torch.equal(tr.count, torch.from_numpy(gold[f"t{ti}_count"]))
    score = torch.from_numpy(gold[f"t{ti}_score"])
    assert (tr.score - score).abs().max().item() < 1e-8
    assert ambiguous_only(tr.keep, unpack(gold[f"t{ti}_keep"], 196), score, k)
    assert (y[:, :, ::4] - torch.from_numpy(gold[f"t{ti}_out_s4"])).abs().max().item() < 1e-4
    assert (sd_ft[:, :, ::4] - torch.from_numpy(gold[f"t{ti}_sd_ft_s4"])).abs().max().item() < 1e-4


@pytest.mark.parametrize("ti", [0, 1])
def test_oracle_nlvr_small_matches_reference_fixture(ti):
    gold = np.load(GOLDEN / "nlvr_small224.npz")
    temp = float(gold["temps"][ti])
    images, ids, mask = weights.nlvr_inputs(2, 224, 20, seed=0)
    assert weights.tensor_digest(images, ids, mask) == str(gold["input_digest"])
    sd = weights.blip_nlvr_state_dict(1234, img_size=224)
    tr = O.NlvrTrace()
    with torch.no_grad():
        pred = O.blip_nlvr_forward(images, ids, mask, sd, temp, trace=tr)
    assert (pred - torch.from_numpy(gold[f"t{ti}_pred"])).abs().max().item() < 1e-4
    assert [t.k if t.pruned else -1 for t in tr.vit] == gold[f"t{ti}_vit_k"].tolist()
    assert [t.k if t.pruned else -1 for t in tr.text] == gold[f"t{ti}_text_k"].tolist()
    for i, t in enumerate(tr.vit):
        n = t.score.shape[1]
        score = torch.from_numpy(gold[f"t{ti}_vit{i}_score"])
        assert (t.score - score).abs().max().item() < 1e-7, i       # free-running: upstream fp32 noise accumulates
        assert ambiguous_only(t.keep, unpack(gold[f"t{ti}_vit{i}_keep"], n), score, t.k if t.pruned else n), i
    for i, t in enumerate(tr.text):
        assert ambiguous_only(t.keep, unpack(gold[f"t{ti}_text{i}_keep"], t.score.shape[1]),
                              torch.from_numpy(gold[f"t{ti}_text{i}_score"]), t.k if t.pruned else t.score.shape[1]), i
    assert (tr.last_hidden - torch.from_numpy(gold[f"t{ti}_last_hidden"])).abs().max().item() < 1e-3
    assert int(gold[f"t{ti}_macs"]) == O.nlvr_macs_from_trace(tr, 197, 20)


def test_calibration_fixture_is_p50():
    c = np.load(GOLDEN / "calib_nlvr_p50_b32.npz")
    assert abs(float(c["ratio"]) - 0.5) < 0.01
    assert int(c["macs_unpruned"]) == O.nlvr_macs_unpruned(577, 20)
    # the reference hard-codes 132.54 GMACs for the unpruned model (compress_nlvr_dtp.py:162); the analytic count
    # must land within a few percent of it (fvcore counts a slightly different op set)
    assert abs(int(c["macs_unpruned"]) / 1e9 - 132.54) / 132.54 < 0.05


def test_select_and_merge_properties():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 50, 16, generator=g)
    score = torch.rand(3, 50, generator=g)
    out, keep, order = O.select_and_merge(x, score, 20)
    assert out.shape == (3, 21, 16) and keep.sum(1).tolist() == [20, 20, 20]
    for b in range(3):
        thr = score[b].sort(descending=True)[0][19]
        assert torch.equal(keep[b], score[b] >= thr)                    # exact top-k set
        assert torch.equal(out[b, :20], x[b][keep[b]])                  # survivors in ascending token order
        w = score[b][~keep[b]]
        merged = ((w / (w.sum() + 1e-8))[:, None] * x[b][~keep[b]]).sum(0)
        assert (out[b, 20] - merged).abs().max().item() < 1e-6


def test_reduce_token_early_out_and_mask_variants():
    g = torch.Generator().manual_seed(1)
    B, n, T = 2, 12, 100
    probs = torch.softmax(torch.randn(B, 4, n + 1, n + 1, generator=g), dim=-1)
    cls_attn = torch.rand(B, n, generator=g) / n
    ta = torch.randn(B, n, T, generator=g) * 5
    x = torch.randn(B, n, 8, generator=g)
    mask = torch.where(torch.rand(B, n, generator=g) < 0.3, -10000.0, 0.0)
    # temperature -> 0+: the softmax over tokens is a one-hot at each column's arg-max, the threshold is the smallest
    # of those scores; everything above it survives. A huge temperature gives threshold = mean score.
    for variant in ("nlvr", "med"):
        tr = O.PruneTrace()
        xo, mo = O.reduce_token(x, probs, cls_attn, ta.clone(), 1e6, mask=mask, variant=variant, trace=tr)
        assert tr.pruned and xo.shape[1] == tr.k + 1 and mo.shape[1] == tr.k + 1
        order = tr.score.sort(dim=1, descending=True, stable=True)[1]
        if variant == "nlvr":      # slot r <- mask of the r-th ranked token (models/nlvr_encoder.py:451-452)
            assert torch.equal(mo, torch.gather(mask, 1, order[:, :tr.k + 1]))
        else:                      # masks travel with tokens, merged slot <- rank k (models/med.py:377-390)
            for b in range(B):
                assert torch.equal(mo[b, :tr.k], mask[b][tr.keep[b]])
                assert mo[b, tr.k] == mask[b, order[b, tr.k]]
    # early-out: when all but one token pass the threshold nothing is pruned (models/vit.py:148-149)
    flat = torch.zeros(B, n, T)
    flat[:, 0, :] = -50.0
    tr = O.PruneTrace()
    xo, _ = O.reduce_token(x, probs, cls_attn, flat, 1.0, trace=tr)
    assert xo.shape[1] in (n, tr.k + 1)


def test_mac_model_matches_survey_numbers():
    # unpruned ViT-B/16 @ 384 = 12 x 4.59 GMAC = 55.1 GMAC per image
    per_layer = O.vit_layer_macs(577, 577) - 2 * 576 * 768 * 100
    assert abs(per_layer / 1e9 - 4.595) < 0.01


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_oracle_med_matches_reference_fixture(ci):
    gold = np.load(GOLDEN / "med_text.npz")
    mode, temp = str(gold["modes"][ci]), float(gold["temps"][ci])
    g = torch.Generator().manual_seed(4321)
    sd = weights.med_text_state_dict(g, "")
    space = torch.randn(100, 768, generator=g)
    _, ids, mask = weights.retrieval_inputs(4, img_size=32, max_len=35, seed=0)
    enc = torch.randn(4, 60, 768, generator=g)
    assert weights.tensor_digest(ids, mask, enc, space) == str(gold["input_digest"])
    traces = []
    with torch.no_grad():
        h, sd_txt = O.med_text_encoder(ids, mask, sd, "", enc if mode == "multimodal" else None, space, temp, mode,
                                       traces=traces)
    assert [t.k if t.pruned else -1 for t in traces] == gold[f"c{ci}_k"].tolist()
    assert (h[:, 0, :] - torch.from_numpy(gold[f"c{ci}_cls"])).abs().max().item() < 1e-3
    assert (sd_txt[:, :, ::4] - torch.from_numpy(gold[f"c{ci}_sd_txt_s4"])).abs().max().item() < 1e-3


def test_oracle_clip_matches_reference_fixture():
    gold = np.load(GOLDEN / "clip_blocks.npz")
    layers = int(gold["layers"])
    sd = weights.clip_state_dict(777, vision_layers=layers, text_layers=layers)
    images, text = weights.clip_inputs(2)
    assert weights.tensor_digest(images, text, sd["space_dict"]) == str(gold["input_digest"])
    traces = []
    with torch.no_grad():
        emb, _ = O.clip_vision_forward(images, sd, "visual.", sd["space_dict"], float(gold["v_temp"]), layers, 12,
                                       traces=traces)
    assert [t.k for t in traces] == gold["v_k"].tolist()
    assert (emb - torch.from_numpy(gold["v_emb"])).abs().max().item() < 1e-4
    for i, t in enumerate(traces):
        score = torch.from_numpy(gold[f"v{i}_score"])
        assert ambiguous_only(t.keep, unpack(gold[f"v{i}_keep"], score.shape[1]), score, t.k), i
    max_keep, temp = int(gold["t_max_keep"]), float(gold["t_temp"])
    for i in range(layers):                       # text blocks: teacher-forced on the reference's own block inputs
        x = torch.from_numpy(gold[f"t{i}_x"])
        tr = O.PruneTrace()
        with torch.no_grad():
            y, _ = O.clip_block(x, sd, f"transformer.resblocks.{i}", 8, sd["space_dict"], temp, None, max_keep, True, tr)
        k_ref = int(gold["t_k"][i])
        assert tr.pruned == (k_ref > 0)
        if k_ref > 0:
            assert tr.k == k_ref and torch.equal(tr.keep, unpack(gold[f"t{i}_keep"], x.shape[1] - 1))
        assert (y[:, :, ::4] - torch.from_numpy(gold[f"t{i}_out_s4"])).abs().max().item() < 2e-4


def test_vqa_rank_oracle_against_reference_fixture():
    """The answer decoder + ranking restatement (oracle.vqa_rank_answer) against what the unmodified reference produced
    (tests/golden/vqa_rank.npz, oracle/gen_golden.py:gen_vqa_rank), teacher-forced with the reference's question states."""
    from oracle import weights
    fx = np.load(GOLDEN / "vqa_rank.npz")
    sd = weights.vqa_state_dict(99, img_size=int(fx["image_size"]))
    dec = {k: v for k, v in sd.items() if k.startswith("text_decoder.")}
    q = torch.from_numpy(fx["question_states"])
    with torch.no_grad():
        max_ids, topk_ids, logp, prob = O.vqa_rank_answer(q, torch.from_numpy(fx["answer_ids"]),
                                                          torch.from_numpy(fx["answer_mask"]), int(fx["k_test"]), dec)
    assert torch.equal(max_ids, torch.from_numpy(fx["max_ids"]))
    assert torch.equal(topk_ids.sort(1)[0], torch.from_numpy(fx["topk_ids"]).sort(1)[0])
    assert (prob - torch.from_numpy(fx["prob_first"])).abs().max().item() < 1e-8
    want = torch.from_numpy(fx["log_probs_sum"]).gather(1, torch.from_numpy(fx["topk_ids"]).argsort(1))
    assert (logp.gather(1, topk_ids.argsort(1)) - want).abs().max().item() < 1e-3
