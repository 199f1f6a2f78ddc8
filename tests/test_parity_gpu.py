"""Parity of the CUDA path (through the C ABI and the module mirrors) with the oracle and the golden fixtures.

Keep-masks must be bit-exact (teacher-forced per layer: both sides consume the same layer input); hidden states within
1e-3 relative (fp16 value lane), as BASELINE.json's north_star states."""
import math
from functools import partial
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import dtp_oracle as O
from oracle import weights

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
REL_TOL = 1e-3          # hidden states: relative L2 error, fp16 value lane (north_star)
THR_RTOL = 2e-6         # threshold: the reference's own fp32 softmax+bmm rounding (a few ulp); counts must still match
SCORE_TOL = 2e-9        # Importance_score: absolute (scores are ~1e-3; fp32 noise floor ~5e-10)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def unpack(bits, n):
    return torch.from_numpy(np.unpackbits(bits, axis=1)[:, :n].astype(bool))


def assert_masks_equal(keep_cuda, keep_ref, score_ref, k, what):
    """Bit-exact keep-mask check that names the decision margin of any flipped token."""
    keep_cuda, keep_ref = keep_cuda.cpu().bool(), keep_ref.cpu().bool()
    if torch.equal(keep_cuda, keep_ref):
        return
    srt = score_ref.sort(dim=1, descending=True)[0]
    msgs = []
    for b, j in (keep_cuda != keep_ref).nonzero().tolist():
        gap = (srt[b, k - 1] - srt[b, k]).item()
        msgs.append(f"row {b} token {j}: score {score_ref[b, j]:.9e}, boundary gap {gap:.3e}")
    raise AssertionError(f"{what}: keep-mask differs from the oracle\n" + "\n".join(msgs[:20]))


@pytest.fixture(scope="module")
def dev(lib):
    return torch.device("cuda:0")


# ---------------------------------------------------------------------------------------------------------------
# BASELINE config 1: single ViT-B/16 Block + DTP head, batch 2, 197 tokens
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ti", [0, 1, 2])
def test_block_config1_against_golden(dev, ti):
    from madtp_b200.utils import Query_model
    from madtp_b200.vit import Block
    gold = np.load(GOLDEN / "block_cfg1.npz")
    temp = float(gold["temps"][ti])
    x, space = weights.block_inputs()
    assert weights.tensor_digest(x, space) == str(gold["input_digest"]), "seeded inputs drifted from the fixture"
    sd = weights.block_state_dict(1234)
    blk = Block(768, 12, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    blk.load_state_dict(sd, strict=True)
    blk = blk.to(dev).eval()
    qm = Query_model(768, 768).to(dev)
    xg, sg = x.to(dev), space.to(dev)
    with torch.no_grad():
        token_attn, sd_ft, _ = qm(xg[:, 1:, :], sg, return_token_att=True)
        ta_ref, sd_ft_ref = O.query_model(x[:, 1:, :], space, 768)
        assert (token_attn.cpu() - ta_ref).abs().max().item() < 2e-4         # |dots| ~ 1e2, fp32 rounding
        assert rel(sd_ft[:, :, ::4], torch.from_numpy(gold[f"t{ti}_sd_ft_s4"])) < 1e-5
        y = blk(xg, False, 0, temp, token_attn)
    res = blk.last_prune
    k = int(gold[f"t{ti}_k"])
    n = x.shape[1] - 1
    score_ref = torch.from_numpy(gold[f"t{ti}_score"])
    assert res.pruned and res.k == k
    assert torch.equal(res.count.cpu().long(), torch.from_numpy(gold[f"t{ti}_count"]).long())
    assert (res.score.cpu() - score_ref).abs().max().item() < SCORE_TOL
    thr_ref = torch.from_numpy(gold[f"t{ti}_threshold"])
    assert ((res.threshold.cpu() - thr_ref).abs() / thr_ref.abs()).max().item() < THR_RTOL
    assert_masks_equal(res.keep, unpack(gold[f"t{ti}_keep"], n), score_ref, k, f"config 1, T={temp}")
    assert y.shape == (2, k + 2, 768)
    assert rel(y[:, :, ::4], torch.from_numpy(gold[f"t{ti}_out_s4"])) < REL_TOL


def test_block_unpruned_matches_oracle(dev):
    from madtp_b200.vit import Block
    x, _ = weights.block_inputs(seed=3)
    sd = weights.block_state_dict(1234)
    blk = Block(768, 12, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    blk.load_state_dict(sd)
    blk = blk.to(dev).eval()
    with torch.no_grad():
        y = blk(x.to(dev))
        y_ref = O.vit_block(x, {"b." + k: v for k, v in sd.items()}, "b", 12)
    assert y.shape == y_ref.shape
    assert rel(y, y_ref) < REL_TOL
