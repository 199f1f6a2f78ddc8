"""Device-resident token counts (ABI 3) and CUDA-graph replay.

The reference reads `topk_num` back once per pruned layer (models/vit.py:145); the default path here keeps every count on
the device. Gate: the device-length forward must be BIT-IDENTICAL to the host-length forward (the round-1 path, itself
pinned to the oracle / reference by tests/test_parity_gpu.py) -- logits, k trajectories, keep-masks, pruned attention
masks -- and a captured CUDA graph must reproduce it on fresh inputs of the same shape."""
import math

import numpy as np
import pytest
import torch

from oracle import weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(lib):
    return torch.device("cuda:0")


_CACHE = {}


def nlvr_model(dev, image_size):
    if image_size not in _CACHE:
        from madtp_b200.blip_nlvr import BLIP_NLVR
        model = BLIP_NLVR(image_size=image_size, evaluate=True)
        msg = model.load_state_dict(weights.blip_nlvr_state_dict(1234, img_size=image_size), strict=False)
        assert not msg.unexpected_keys and not msg.missing_keys
        _CACHE[image_size] = model.to(dev).eval()
    return _CACHE[image_size]


def _records(model):
    out = []
    for m in list(model.visual_encoder.blocks) + list(model.text_encoder.encoder.layer):
        r = m.last_prune
        if r is None or not r.pruned:
            out.append((-1, None, None, None))
        else:
            out.append((r.k, r.keep.cpu().clone(), r.count.cpu().clone(), r.score.cpu().clone()))
    return out


@pytest.mark.parametrize("image_size,pairs,text_len,temp,pad_to", [(224, 2, 20, 1.0, 0), (224, 2, 20, 8.0, 0),
                                                                   (224, 3, 12, 8.0, 20), (384, 4, 20, 3.5894, 0)])
def test_device_lengths_match_host_lengths_bit_for_bit(dev, image_size, pairs, text_len, temp, pad_to):
    from madtp_b200 import vit
    from madtp_b200.blip_nlvr import TokenizedText
    model = nlvr_model(dev, image_size)
    model.record_states = True
    images, ids, mask = weights.nlvr_inputs(pairs, image_size, text_len, seed=3, pad_to=pad_to)
    text = TokenizedText(ids.to(dev), mask.to(dev))
    try:
        vit.device_lengths_enabled(False)
        with torch.no_grad():
            ref = model(images.to(dev), text, pairs, temp, train=False)
        ref_rec = _records(model)
        ref_img, ref_h = model.last["image_embeds"].clone(), model.last["last_hidden_state"].clone()
        vit.device_lengths_enabled(True)
        with torch.no_grad():
            out = model(images.to(dev), text, pairs, temp, train=False)
        rec = _records(model)
        img, hid = model.last["image_embeds"].clone(), model.last["last_hidden_state"].clone()
    finally:
        vit.device_lengths_enabled(True)
        model.record_states = False
    assert [r[0] for r in rec] == [r[0] for r in ref_rec], "k trajectories differ"
    assert any(r[0] >= 0 for r in rec[12:]) or temp < 4, "the text encoder was expected to prune"
    for i, (a, b) in enumerate(zip(rec, ref_rec)):
        if a[0] >= 0:
            assert torch.equal(a[1], b[1]), f"layer {i}: keep-mask"
            assert torch.equal(a[2], b[2]), f"layer {i}: counts"
            assert torch.equal(a[3], b[3]), f"layer {i}: scores"
    assert img.shape == ref_img.shape and torch.equal(img, ref_img), "image_embeds"
    assert hid.shape == ref_h.shape and torch.equal(hid, ref_h), "last_hidden_state"
    assert torch.equal(out, ref), f"logits differ by {(out - ref).abs().max().item():.3e}"


def test_public_encoder_calls_use_one_readback_and_match(dev):
    """VisionTransformer.forward / med.BertModel.forward (the calls compress_retrieval_dtp.py:104,120,170 makes) on the
    device-length path against the host-length path."""
    from madtp_b200 import vit
    from madtp_b200.blip_retrieval import BLIP_Retrieval
    sd = weights.retrieval_state_dict(4321, img_size=224)
    model = BLIP_Retrieval(image_size=224, evaluate=True)
    model.load_state_dict(sd, strict=False)
    model = model.to(dev).eval()
    images, ids, mask = weights.retrieval_inputs(3, 224, 35, seed=1)
    outs = []
    try:
        for flag in (False, True):
            vit.device_lengths_enabled(flag)
            with torch.no_grad():
                feat, sd_img = model.visual_encoder(images.to(dev), space_dict=model.space_dict, temperature=8.0)
                txt, sd_txt = model.text_encoder(ids.to(dev), attention_mask=mask.to(dev), mode='text',
                                                 space_dict=model.space_dict, temperature=8.0)
                itm = model.itm_score(ids.to(dev), mask.to(dev), feat, 8.0)
            outs.append((feat.clone(), sd_img.clone(), txt.last_hidden_state.clone(), sd_txt.clone(), itm.clone()))
    finally:
        vit.device_lengths_enabled(True)
    names = ["image_feat", "sd_img_ft", "text_states", "sd_txt_ft", "itm"]
    for nm, a, b in zip(names, *outs):
        assert a.shape == b.shape, nm
        if nm.startswith("sd_"):
            # the aggregated codebook feature is a model OUTPUT, not part of the scoring lane: with device-resident
            # lengths every layer takes the tensor-core kernel (the host-length path switches to the FFMA kernel below
            # 64 tokens), so it agrees to rounding, not bit for bit
            assert ((a - b).norm() / b.norm()).item() < 1e-5, nm
        else:
            assert torch.equal(a, b), f"{nm}: max diff {(a - b).abs().max().item():.3e}"
    assert outs[0][0].shape[1] < 197 and outs[0][2].shape[1] < 35, "both encoders were expected to prune"


def test_cuda_graph_replay_matches_eager_on_fresh_inputs(dev, lib):
    from madtp_b200.blip_nlvr import TokenizedText
    model = nlvr_model(dev, 224)
    temp, pairs = 8.0, 2
    batches = [weights.nlvr_inputs(pairs, 224, 20, seed=s) for s in (0, 5, 9)]
    eager = []
    with torch.no_grad():
        for images, ids, mask in batches:
            eager.append(model(images.to(dev), TokenizedText(ids.to(dev), mask.to(dev)), pairs, temp, train=False).clone())
    ks_eager = [b.last_prune.k for b in model.visual_encoder.blocks]
    model.enable_cuda_graphs(True)
    try:
        with torch.no_grad():
            for rep in range(2):
                for (images, ids, mask), want in zip(batches, eager):
                    n0 = lib.launch_count()
                    got = model(images.to(dev), TokenizedText(ids.to(dev), mask.to(dev)), pairs, temp, train=False)
                    assert torch.equal(got, want), f"graph replay differs by {(got - want).abs().max().item():.3e}"
            # a replay launches nothing through the C ABI: the kernels are inside the graph
            assert lib.launch_count() == n0
            assert [b.last_prune.k for b in model.visual_encoder.blocks] == ks_eager
        assert len(model._graphs) == 1
    finally:
        model.enable_cuda_graphs(False)


# ---------------------------------------------------------------------------------------------------------------
# kernels behind it
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,Lq,Nk,H", [(3, 20, 255, 12), (2, 17, 577, 12), (2, 35, 901, 12), (4, 128, 130, 2),
                                       (2, 5, 64, 1), (1, 33, 1, 3)])
def test_cross_attention_any_key_count(lib, dev, B, Lq, Nk, H):
    """The key-block loop (online softmax over blocks of 128 keys) against fp64 attention, incl. Nk > 256 (VQA at
    480 x 480: 901 image tokens, models/blip_vqa.py:119-125) and a per-sequence key mask."""
    g = torch.Generator().manual_seed(B * 1000 + Nk)
    C = H * 64
    q = torch.randn(B, Lq, C, generator=g) * 0.7
    P = (Nk + 7) // 8 * 8
    k = torch.zeros(B, P, C)
    k[:, :Nk] = torch.randn(B, Nk, C, generator=g) * 0.7
    v = torch.zeros(B, P, C)
    v[:, :Nk] = torch.randn(B, Nk, C, generator=g)
    vb = torch.randn(C, generator=g) * 0.1
    mask = torch.zeros(B, Nk)
    mask[:, Nk // 2:] = torch.where(torch.rand(B, Nk - Nk // 2, generator=g) < 0.3, -10000.0, 0.0)
    mask[:, 0] = 0.0
    q16, k16 = q.half().to(dev), k.half().to(dev)
    vt16 = v.half().reshape(B * P, C).t().contiguous().to(dev)             # [C, B*P]
    out = torch.empty(B, Lq, C, dtype=torch.float16, device=dev)
    scale = 1.0 / math.sqrt(64)
    for km in (None, mask):
        lib.attn_cross_tc(q16, k16[:, :Nk], vt16, H, scale, out, keys_per_batch=P, v_bias=vb.to(dev),
                          key_mask=None if km is None else km.to(dev))
        qd = q.half().double().view(B, Lq, H, 64).transpose(1, 2)
        kd = k.half().double()[:, :Nk].view(B, Nk, H, 64).transpose(1, 2)
        vd = v.half().double()[:, :Nk].view(B, Nk, H, 64).transpose(1, 2)
        s = qd @ kd.transpose(-1, -2) * scale
        if km is not None:
            s = s + km.double()[:, None, None, :]
        ref = (torch.softmax(s, -1) @ vd).transpose(1, 2).reshape(B, Lq, C) + vb.double()
        err = (out.double().cpu() - ref).abs().max().item()
        assert err < 4e-3, f"Nk={Nk} mask={'yes' if km is not None else 'no'}: {err:.2e}"


def test_cross_attention_device_resident_lengths(lib, dev):
    """lq_dev / nk_dev: capacity-sized packed buffers, lengths read on the device -> same result as exact shapes."""
    g = torch.Generator().manual_seed(7)
    B, H, Lq, Nk, Lcap, Ncap = 3, 12, 14, 255, 20, 577
    C = H * 64
    P, Pcap = (Nk + 7) // 8 * 8, (Ncap + 7) // 8 * 8
    q = (torch.randn(B, Lq, C, generator=g) * 0.7).half()
    k = torch.zeros(B, P, C, dtype=torch.float16)
    k[:, :Nk] = (torch.randn(B, Nk, C, generator=g) * 0.7).half()
    v = torch.zeros(B, P, C, dtype=torch.float16)
    v[:, :Nk] = torch.randn(B, Nk, C, generator=g).half()
    want = torch.empty(B, Lq, C, dtype=torch.float16, device=dev)
    lib.attn_cross_tc(q.to(dev), k.to(dev)[:, :Nk], v.reshape(B * P, C).t().contiguous().to(dev), H, 0.125, want,
                      keys_per_batch=P)
    # the same data packed into capacity-sized buffers
    qc = torch.zeros(B * Lcap, C, dtype=torch.float16)
    qc[:B * Lq] = q.reshape(B * Lq, C)
    kc = torch.zeros(B * Pcap, C, dtype=torch.float16)
    kc[:B * P] = k.reshape(B * P, C)
    vtc = torch.zeros(C, B * Pcap, dtype=torch.float16)
    vtc[:, :B * P] = v.reshape(B * P, C).t()
    got = torch.zeros(B * Lcap, C, dtype=torch.float16, device=dev)
    lq_dev = torch.tensor([Lq], dtype=torch.int32, device=dev)
    nk_dev = torch.tensor([Nk], dtype=torch.int32, device=dev)
    lib.attn_cross_tc(qc.to(dev).view(B, Lcap, C), kc.to(dev).view(B, Pcap, C)[:, :Ncap], vtc.to(dev), H, 0.125,
                      got.view(B, Lcap, C), keys_per_batch=Pcap, lq_dev=lq_dev, nk_dev=nk_dev)
    assert torch.equal(got[:B * Lq].view(B, Lq, C), want)
    assert not got[B * Lq:].any(), "rows beyond the dynamic extent must not be written"


def test_gemm_dynamic_extents_and_layernorm_pack(lib, dev):
    g = torch.Generator().manual_seed(3)
    M, Mcap, N, Ncap, K = 300, 640, 200, 256, 256
    a = torch.randn(Mcap, K, generator=g).half().to(dev)
    b = torch.randn(Ncap, K, generator=g).half().to(dev)
    want = torch.empty(M, N, dtype=torch.float32, device=dev)
    lib.gemm(lib.GEMM_F16, a[:M], b[:N], want)
    got = torch.full((Mcap, Ncap), 7.0, dtype=torch.float32, device=dev)
    m_dev = torch.tensor([M // 4], dtype=torch.int32, device=dev)
    n_dev = torch.tensor([N // 8], dtype=torch.int32, device=dev)
    lib.gemm(lib.GEMM_F16, a, b, got, m_dev=m_dev, m_mult=4, n_dev=n_dev, n_mult=8)
    assert torch.equal(got[:M, :N], want)
    assert bool((got[M:] == 7.0).all()) and bool((got[:, N:] == 7.0).all()), "nothing outside the dynamic extent is written"
    # LayerNorm -> fp16 cross-attention operand layout, two groups, dynamic N
    B, Ncap2, Nd, d = 6, 21, 13, 768
    x = torch.randn(B * Ncap2, d, generator=g)
    xp = torch.zeros(B * Ncap2, d)
    xp[:B * Nd] = x[:B * Nd]                                     # packed with the dynamic length
    gamma, beta = torch.randn(d, generator=g), torch.randn(d, generator=g)
    Pcap, P = (Ncap2 + 7) // 8 * 8, (Nd + 7) // 8 * 8
    y16 = torch.full((2, 3 * Pcap, d), 9.0, dtype=torch.float16, device=dev)
    y32 = torch.zeros(B * Ncap2, d, device=dev)
    p_out = torch.zeros(1, dtype=torch.int32, device=dev)
    lib.layernorm_pack(xp.to(dev), B, Ncap2, gamma.to(dev), beta.to(dev), 1e-6, y16, 3, 3 * Pcap * d, y_f32=y32,
                       p_out=p_out, n_dev=torch.tensor([Nd], dtype=torch.int32, device=dev))
    ref = torch.nn.functional.layer_norm(x[:B * Nd].double(), (d,), gamma.double(), beta.double(), 1e-6).view(B, Nd, d)
    assert int(p_out) == P
    assert (y32[:B * Nd].double().cpu().view(B, Nd, d) - ref).abs().max().item() < 1e-4
    for grp in range(2):
        blk = y16[grp].view(-1)[:3 * P * d].view(3, P, d).float().cpu()
        assert (blk[:, :Nd].double() - ref[grp * 3:(grp + 1) * 3]).abs().max().item() < 2e-2
        assert not blk[:, Nd:].any(), "padding rows are zero"
    tok = lib.take_token(y32, B, Ncap2, 2, n_dev=torch.tensor([Nd], dtype=torch.int32, device=dev))
    assert torch.equal(tok, y32[:B * Nd].view(B, Nd, d)[:, 2])


def test_two_batches_in_flight_match_sequential(dev):
    """pipeline.StreamPool: consecutive forwards on alternating streams (own graph + buffers per stream) give exactly the
    logits of the sequential run, whatever the interleaving."""
    from madtp_b200.blip_nlvr import TokenizedText
    from madtp_b200.pipeline import StreamPool
    model = nlvr_model(dev, 224)
    temp, pairs = 8.0, 2
    batches = [weights.nlvr_inputs(pairs, 224, 20, seed=s) for s in (11, 12, 13, 14, 15)]
    dev_batches = [(im.to(dev), TokenizedText(ids.to(dev), mask.to(dev))) for im, ids, mask in batches]
    with torch.no_grad():
        want = [model(im, tx, pairs, temp, train=False).clone() for im, tx in dev_batches]
    model.enable_cuda_graphs(True)
    try:
        pool = StreamPool(dev, 2)
        got = [None] * len(batches)
        with torch.no_grad():
            for rep in range(3):
                for i, (im, tx) in enumerate(dev_batches):
                    with pool.stream(i):
                        got[i] = model(im, tx, pairs, temp, train=False).clone()
                pool.join()
                torch.cuda.synchronize()
                for i in range(len(batches)):
                    assert torch.equal(got[i], want[i]), f"rep {rep}, batch {i}"
        assert len(model._graphs) == 2       # one captured graph per stream
    finally:
        model.enable_cuda_graphs(False)


def test_retrieval_forward_device_and_graph_match_encoder_calls(dev):
    """BLIP_Retrieval.forward(train=False): the device-length chain (image encoder -> text-only pass -> multimodal ITM
    pass with nothing read back) and its CUDA-graph replay are bit-identical to the host-length path built from the
    public encoder calls (compress_retrieval_dtp.py:104-122,166-176)."""
    from madtp_b200 import vit
    from madtp_b200.blip_retrieval import BLIP_Retrieval
    model = BLIP_Retrieval(image_size=224, evaluate=True)
    msg = model.load_state_dict(weights.retrieval_state_dict(4321, img_size=224), strict=False)
    assert not msg.unexpected_keys
    model = model.to(dev).eval()
    temp = 6.0
    outs = []
    try:
        for seed in (0, 1):
            images, ids, mask = (t.to(dev) for t in weights.retrieval_inputs(4, 224, 35, seed=seed))
            vit.device_lengths_enabled(False)
            ref_sim, ref_itm = model(images, (ids, mask), 0.0, None, temp, train=False)
            ref_k = [b.last_prune.k if b.last_prune is not None and b.last_prune.pruned else -1
                     for b in model.visual_encoder.blocks]
            vit.device_lengths_enabled(True)
            sim, itm = model(images, (ids, mask), 0.0, None, temp, train=False)
            k = [b.last_prune.k if b.last_prune.pruned else -1 for b in model.visual_encoder.blocks]
            assert k == ref_k and any(v >= 0 for v in k)
            assert torch.equal(itm, ref_itm), f"ITM logits differ by {(itm - ref_itm).abs().max().item():.3e}"
            assert torch.allclose(sim, ref_sim, atol=1e-6)
            outs.append((images, ids, mask, sim.clone(), itm.clone()))
        model.enable_cuda_graphs(True)
        for images, ids, mask, sim, itm in outs + outs[:1]:      # capture on the first input, replay on fresh ones
            g_sim, g_itm = model(images, (ids, mask), 0.0, None, temp, train=False)
            assert torch.equal(g_itm, itm) and torch.equal(g_sim, sim)
        assert len(model._graphs) == 1 and next(iter(model._graphs.values())).replays == 3
    finally:
        vit.device_lengths_enabled(True)
        model.enable_cuda_graphs(False)


def test_vqa_encode_question_packed_and_graph_match_encode_question(dev):
    """BLIP_VQA.encode_question_packed (device-resident lengths end to end, optionally a graph replay) against the
    public encode_question (models/blip_vqa.py:60,119-125): same question states, bit for bit."""
    from madtp_b200 import vit
    from madtp_b200.blip_retrieval import BLIP_VQA
    model = BLIP_VQA(image_size=224, evaluate=True)
    model.load_state_dict(weights.vqa_state_dict(99, img_size=224), strict=False)
    model = model.to(dev).eval()
    temp = 10.0
    images, ids, mask = (t.to(dev) for t in weights.retrieval_inputs(3, 224, 20, seed=5))
    try:
        vit.device_lengths_enabled(False)
        ref, _ = model.encode_question(images, ids, mask, temp)
        vit.device_lengths_enabled(True)
        h, l_dev, cls = model.encode_question_packed(images, ids, mask, temp)
        n = int(l_dev.item())
        B, _, d = h.shape
        assert n == ref.shape[1]
        assert any(b.last_prune.pruned for b in model.visual_encoder.blocks), "the image encoder was expected to prune"
        assert torch.equal(h.reshape(-1)[:B * n * d].view(B, n, d), ref)
        assert torch.equal(cls, ref[:, 0, :])
        model.enable_cuda_graphs(True)
        for _ in range(2):
            _, l2, cls2 = model.encode_question_packed(images, ids, mask, temp)
            assert int(l2.item()) == n and torch.equal(cls2, ref[:, 0, :])
        images2, ids2, mask2 = (t.to(dev) for t in weights.retrieval_inputs(3, 224, 20, seed=6))
        vit.device_lengths_enabled(False)
        ref2, _ = model.encode_question(images2, ids2, mask2, temp)
        vit.device_lengths_enabled(True)
        _, _, cls3 = model.encode_question_packed(images2, ids2, mask2, temp)     # a replay on fresh inputs
        assert torch.equal(cls3, ref2[:, 0, :])
    finally:
        vit.device_lengths_enabled(True)
        model.enable_cuda_graphs(False)


def test_clip_towers_device_lengths_and_graph_match_host_lengths(dev):
    """CLIP.encode_image / encode_text (clip/model.py:482-503): device-resident lengths and the CUDA-graph replay against
    the host-length path (one read-back per block, the reference's own control flow): embeddings, sd_ft and k
    trajectories bit for bit -- including the causal text tower with its `max_keep` guard and the EOT gather."""
    from madtp_b200 import vit
    from madtp_b200.clip_model import CLIP
    layers = 4
    sd = weights.clip_state_dict(777, vision_layers=layers, text_layers=layers)
    model = CLIP(512, 224, layers, 768, 16, 77, 49408, 512, 8, layers, True, None)
    msg = model.load_state_dict(sd, strict=False)
    assert not msg.unexpected_keys
    model = model.to(dev).eval()
    space = model.space_dict

    def ks(blocks):
        return [b.last_prune.k if b.last_prune is not None and b.last_prune.pruned else -1 for b in blocks]

    try:
        cases = []
        for seed, t_img, t_txt in ((0, 5.0, 50.0), (1, 5.0, 50.0)):
            images, text = (t.to(dev) for t in weights.clip_inputs(3, seed=seed))
            vit.device_lengths_enabled(False)
            ri, rs = model.encode_image(images, space, t_img)
            rk = ks(model.visual.transformer.resblocks)
            rt, rts = model.encode_text(text, space, t_txt)
            rtk = ks(model.transformer.resblocks)
            vit.device_lengths_enabled(True)
            ei, es = model.encode_image(images, space, t_img)
            assert ks(model.visual.transformer.resblocks) == rk and any(k >= 0 for k in rk)
            et, ets = model.encode_text(text, space, t_txt)
            assert ks(model.transformer.resblocks) == rtk
            assert torch.equal(ei, ri) and torch.equal(es, rs), "vision tower"
            # sd_ft is a by-product off the scoring lane: below 64 tokens the host-length path aggregates it on the CUDA
            # cores, the capacity-sized device path always on the tensor cores
            assert torch.equal(et, rt) and torch.allclose(ets, rts, rtol=2e-5, atol=2e-6), "text tower"
            cases.append((images, text, t_img, t_txt, ri.clone(), rt.clone()))
        model.enable_cuda_graphs(True)
        for images, text, t_img, t_txt, ri, rt in cases + cases[:1]:
            gi, _ = model.encode_image(images, space, t_img)
            gt, _ = model.encode_text(text, space, t_txt)
            assert torch.equal(gi, ri) and torch.equal(gt, rt)
    finally:
        vit.device_lengths_enabled(True)
        model.enable_cuda_graphs(False)
