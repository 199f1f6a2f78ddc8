"""Kernel-level checks of the C-ABI entry points on a real GPU, each against a straightforward fp64 PyTorch statement
of the same arithmetic (this file checks kernels in isolation; parity with the reference algorithm is in
test_parity_gpu.py, which goes through the oracle and the golden fixtures)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


# ---------------------------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------------------------
GEMM_SHAPES = [
    (128, 128, 64), (128, 256, 128), (256, 128, 32), (300, 200, 96), (1000, 768, 768), (77, 100, 768),
    (4099, 2304, 768), (640, 3072, 768), (640, 768, 3072), (2, 8, 8), (36928, 128, 768),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_simt(lib, dev, M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(dev)
    b = torch.randn(N, K, generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    out = torch.empty(M, N, device=dev)
    lib.gemm(lib.GEMM_SIMT, a, b, out, bias=bias)
    ref = a.double() @ b.double().T + bias.double()
    assert _rel(out, ref) < 2e-6


@pytest.mark.parametrize("M,N,K,act", [(32, 768, 768, 2), (32, 2, 768, 0), (64, 5, 132, 1), (1, 1, 4, 0)])
def test_gemm_simt_few_rows(lib, dev, M, N, K, act):
    """Classification heads (cls_head, itm_head): the one-warp-per-output kernel behind MADTP_GEMM_SIMT for M <= 64."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dev)
    b = torch.randn(N, K, generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    out = torch.full((M, N), float("nan"), device=dev)
    lib.gemm(lib.GEMM_SIMT, a, b, out, bias=bias, act=act)
    ref = a.double() @ b.double().T + bias.double()
    ref = {0: ref, 1: torch.nn.functional.gelu(ref), 2: torch.relu(ref)}[act]
    assert _rel(out, ref) < 2e-6


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_tf32x3(lib, dev, M, N, K):
    if K % 4:
        pytest.skip("row pitch must be a multiple of 16 bytes")
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + 1)
    a = torch.randn(M, K, generator=g).to(dev)
    b = torch.randn(N, K, generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    a_hi, a_lo = lib.split_tf32(a)
    b_hi, b_lo = lib.split_tf32(b)
    assert torch.equal(a_hi + a_lo, a), "tf32 split must be exact"
    assert int((a_hi.view(torch.int32) & 0x1FFF).abs().max()) == 0, "hi must have the low 13 mantissa bits clear"
    out = torch.full((M, N), float("nan"), device=dev)
    lib.gemm(lib.GEMM_TF32X3, a_hi, b_hi, out, a_lo=a_lo, b_lo=b_lo, bias=bias)
    ref = a.double() @ b.double().T + bias.double()
    err = _rel(out, ref)
    # an fp32 FFMA GEMM of this depth sits around 1e-7; plain TF32 would be ~5e-4, and TF32x3 accumulated over the
    # whole K inside the tensor core measured 5e-6 (K=768) .. 2e-5 (K=3072) because its accumulator truncates
    simt = torch.empty(M, N, device=dev)
    lib.gemm(lib.GEMM_SIMT, a, b, simt, bias=bias)
    print(f"tf32x3 M={M} N={N} K={K}: rel err {err:.3e} (simt fp32 {_rel(simt, ref):.3e})")
    assert err < 6e-7, f"TF32x3 relative error {err:.3e}"


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("w_std", [1.0, 0.02])
def test_gemm_f16x3(lib, dev, M, N, K, w_std):
    """The product path's scoring-lane GEMM: fp16 hi/lo planes, weights scaled by a power of two, alpha removes it."""
    if K % 8:
        pytest.skip("row pitch must be a multiple of 16 bytes")
    from madtp_b200.functional import pow2_scale
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + 2)
    a = torch.randn(M, K, generator=g).to(dev)
    b = (torch.randn(N, K, generator=g) * w_std).to(dev)
    bias = (torch.randn(N, generator=g) * w_std).to(dev)
    s = pow2_scale(b)
    a_hi, a_lo = lib.split_f16(a)
    b_hi, b_lo = lib.split_f16(b, s)
    assert torch.equal(a_hi, a.half())
    assert bool(((b_hi.double() + b_lo.double()) / s - b.double()).abs().max() <= b.abs().max().double() * 2.0 ** -22)
    out = torch.full((M, N), float("nan"), device=dev)
    lib.gemm(lib.GEMM_F16X3, a_hi, b_hi, out, a_lo=a_lo, b_lo=b_lo, bias=bias, alpha=1.0 / s)
    ref = a.double() @ b.double().T + bias.double()
    err = _rel(out, ref)
    print(f"f16x3 M={M} N={N} K={K} w_std={w_std}: rel err {err:.3e}")
    assert err < 6e-7, f"F16x3 relative error {err:.3e}"


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_f16(lib, dev, M, N, K):
    if K % 8:
        pytest.skip("row pitch must be a multiple of 16 bytes")
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + 2)
    a = torch.randn(M, K, generator=g).to(dev).half()
    b = torch.randn(N, K, generator=g).to(dev).half()
    out = torch.full((M, N), float("nan"), device=dev)
    lib.gemm(lib.GEMM_F16, a, b, out)
    ref = a.double() @ b.double().T
    # operands are exactly representable; the error is the tensor core's truncating fp32 accumulation (grows ~K)
    assert _rel(out, ref) < 1e-5


def test_gemm_epilogues(lib, dev):
    M, N, K = 333, 768, 256
    g = torch.Generator(device="cpu").manual_seed(5)
    a = torch.randn(M, K, generator=g).to(dev).half()
    b = (torch.randn(N, K, generator=g) * 0.05).to(dev).half()
    bias = torch.randn(N, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev)
    acc = a.double() @ b.double().T
    for act, fn in ((lib.ACT_NONE, lambda x: x), (lib.ACT_GELU, lambda x: torch.nn.functional.gelu(x)),
                    (lib.ACT_RELU, torch.relu), (lib.ACT_QUICKGELU, lambda x: x * torch.sigmoid(1.702 * x))):
        out = torch.empty(M, N, device=dev)
        lib.gemm(lib.GEMM_F16, a, b, out, bias=bias, residual=res, act=act, alpha=0.5)
        ref = fn(0.5 * acc + bias.double()) + res.double()
        assert _rel(out, ref) < 5e-6, f"act {act}"
    # fp16 output into a column slice of a wider buffer (ldc != N)
    wide = torch.zeros(M, 2 * N, device=dev, dtype=torch.float16)
    lib.gemm(lib.GEMM_F16, a, b, wide[:, N:], bias=bias)
    ref = (acc + bias.double())
    assert _rel(wide[:, N:], ref) < 1e-3
    assert float(wide[:, :N].abs().max()) == 0.0


def test_gemm_rejects_cpu_tensors(lib):
    a = torch.randn(8, 8)
    with pytest.raises(RuntimeError):
        lib.gemm(lib.GEMM_SIMT, a, a, torch.empty(8, 8))


# ---------------------------------------------------------------------------------------------------------------
# LayerNorm / row ops
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,d,eps", [(1154, 768, 1e-6), (37, 768, 1e-12), (64, 1024, 1e-6)])
def test_layernorm(lib, dev, rows, d, eps):
    g = torch.Generator(device="cpu").manual_seed(rows)
    x = (torch.randn(rows, d, generator=g) * 3 + 0.5).to(dev)
    gamma = torch.randn(d, generator=g).to(dev)
    beta = torch.randn(d, generator=g).to(dev)
    outs = {k: torch.empty(rows, d, device=dev, dtype=torch.float32 if k == "y_f32" else torch.float16)
            for k in ("y_f32", "y_hi", "y_lo", "x_hi", "x_lo")}
    y16 = torch.empty(rows, d, device=dev, dtype=torch.float16)
    lib.layernorm(x, gamma, beta, eps, y_f16=y16, **outs)
    ref = torch.nn.functional.layer_norm(x.double(), (d,), gamma.double(), beta.double(), eps)
    assert (outs["y_f32"].double() - ref).abs().max().item() < 5e-6
    # fp16 hi/lo planes: hi = fp16(v), lo = fp16(v - hi) => |hi + lo - v| <= 2^-23 |v| (+ half a subnormal step)
    for planes, v in ((("y_hi", "y_lo"), outs["y_f32"]), (("x_hi", "x_lo"), x)):
        hi, lo = outs[planes[0]], outs[planes[1]]
        assert torch.equal(hi, v.half())
        err = (hi.double() + lo.double() - v.double()).abs()
        assert bool((err <= v.double().abs() * 2.0 ** -22 + 2.0 ** -25).all())
    assert torch.equal(y16, outs["y_f32"].half())


def test_patchify_matches_conv(lib, dev):
    g = torch.Generator(device="cpu").manual_seed(3)
    B, C, H, W, P, D = 3, 3, 64, 48, 16, 768
    img = torch.randn(B, C, H, W, generator=g).to(dev)
    w = (torch.randn(D, C, P, P, generator=g) * 0.02).to(dev)
    bias = torch.randn(D, generator=g).to(dev)
    hi, lo = lib.patchify(img, P)
    want = img.view(B, C, H // P, P, W // P, P).permute(0, 2, 4, 1, 3, 5).double()
    got = (hi.double() + lo.double()).view(B, H // P, W // P, C, P, P)
    assert bool(((got - want).abs() <= want.abs() * 2.0 ** -22 + 2.0 ** -25).all())
    w_hi, w_lo = lib.split_f16(w.view(D, -1), 2.0 ** 16)
    out = torch.empty(hi.shape[0], D, device=dev)
    lib.gemm(lib.GEMM_F16X3, hi, w_hi, out, a_lo=lo, b_lo=w_lo, bias=bias, alpha=2.0 ** -16)
    ref = torch.nn.functional.conv2d(img.double(), w.double(), bias.double(), stride=P).flatten(2).transpose(1, 2)
    assert _rel(out.view(B, -1, D), ref) < 6e-7


# ---------------------------------------------------------------------------------------------------------------
# attention + statistics
# ---------------------------------------------------------------------------------------------------------------
def _attn_ref(q, k, v, scale, mask):
    s = q @ k.transpose(-1, -2) * scale
    if mask is not None:
        s = s + mask[:, None, None, :]
    p = torch.softmax(s, dim=-1)
    o = p @ v
    return p, o


@pytest.mark.parametrize("B,H,N,masked", [(2, 12, 197, False), (3, 12, 20, True), (2, 12, 577, False), (1, 4, 64, False),
                                          (2, 3, 65, True)])
def test_attention_and_stats(lib, dev, B, H, N, masked):
    g = torch.Generator(device="cpu").manual_seed(N + B)
    qkv = torch.randn(B, N, 3 * H * 64, generator=g).to(dev)
    q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
    mask = None
    if masked:
        lens = torch.randint(N // 2, N + 1, (B,), generator=g)
        mask = torch.zeros(B, N)
        for b in range(B):
            mask[b, lens[b]:] = -10000.0
        mask = mask.to(dev)
    scale = 0.125
    out = torch.empty(B, N, H * 64, device=dev, dtype=torch.float16)
    stats = tuple(torch.empty(B, H, N, device=dev) for _ in range(3))
    lib.attn_fwd(q, k, v, H, scale, out, key_mask=mask, stats=stats)
    n_parts = (N + 63) // 64
    col_part = torch.zeros(B, n_parts, N, device=dev)
    cls_attn = torch.zeros(B, N, device=dev)
    lib.attn_stats(q, k, H, scale, stats, col_part, cls_attn, key_mask=mask)

    def heads(t):
        return t.reshape(B, N, H, 64).permute(0, 2, 1, 3).double()

    p, o = _attn_ref(heads(q), heads(k), heads(v), scale, None if mask is None else mask.double())
    o_merged = o.permute(0, 2, 1, 3).reshape(B, N, H * 64)
    assert _rel(out, o_merged) < 6e-4  # fp16 output rounding
    assert (stats[2].double() - o.norm(dim=-1)).abs().max().item() < 1e-5
    hi = o[..., 1:, :].norm(dim=-1)
    hi = hi / (hi.sum(dim=1, keepdim=True) + 1e-8)
    cls_ref = (p[:, :, 0, 1:] * hi).sum(dim=1)
    a_ref = p[:, :, 1:, 1:].max(dim=1)[0].sum(dim=1)
    assert (cls_attn[:, 1:].double() - cls_ref).abs().max().item() < 1e-6
    a = col_part.double().sum(dim=1)[:, 1:]
    assert ((a - a_ref).abs() / a_ref.abs().clamp_min(1e-3)).max().item() < 2e-6


def test_cross_attention(lib, dev):
    g = torch.Generator(device="cpu").manual_seed(11)
    B, H, Lq, Nk = 4, 12, 19, 283
    q = torch.randn(B, Lq, H * 64, generator=g).to(dev)
    kv = torch.randn(B, Nk, 2 * H * 64, generator=g).to(dev)
    k, v = kv[..., :H * 64], kv[..., H * 64:]
    wide = torch.zeros(B, Lq, 2 * H * 64, device=dev, dtype=torch.float16)
    lib.attn_fwd(q, k, v, H, 0.125, wide[..., H * 64:])

    def heads(t, n):
        return t.reshape(B, n, H, 64).permute(0, 2, 1, 3).double()

    _, o = _attn_ref(heads(q, Lq), heads(k, Nk), heads(v, Nk), 0.125, None)
    assert _rel(wide[..., H * 64:], o.permute(0, 2, 1, 3).reshape(B, Lq, H * 64)) < 6e-4
    assert float(wide[..., :H * 64].abs().max()) == 0.0


# ---------------------------------------------------------------------------------------------------------------
# Query_model pieces and DTP kernels
# ---------------------------------------------------------------------------------------------------------------
def test_query_model_kernels(lib, dev):
    g = torch.Generator(device="cpu").manual_seed(21)
    B, n, T, d = 3, 196, 100, 768
    x = torch.randn(B, n + 1, d, generator=g).to(dev)
    ta = (torch.randn(B, n + 1, 128, generator=g) * 20).to(dev)
    ta_p = ta[:, 1:, :]
    div = math.sqrt(768)
    cm, cs = lib.token_colstats(ta_p, n, T, div)
    xs = ta_p[..., :T].double() / div
    assert (cm.double() - xs.max(dim=1)[0]).abs().max().item() < 1e-6
    w = torch.softmax(xs, dim=1)
    sd = torch.zeros(B, T, d, device=dev)
    lib.query_sdft(ta_p, cm, cs, x[:, 1:, :], n, T, div, sd, False)
    ref = w.transpose(1, 2) @ x[:, 1:, :].double()
    assert _rel(sd, ref) < 2e-6
    lib.query_sdft(ta_p, cm, cs, x[:, 1:, :], n, T, div, sd, True)
    assert _rel(sd, 2 * ref) < 2e-6


@pytest.mark.parametrize("B,n,temp", [(4, 196, 1.0), (2, 576, 5.0), (5, 19, 2.0), (3, 900, 50.0)])
def test_dtp_score_select_gather(lib, dev, B, n, temp):
    g = torch.Generator(device="cpu").manual_seed(n)
    T, d, N = 100, 768, n + 1
    n_parts = (N + 63) // 64
    col_part = torch.rand(B, n_parts, N, generator=g).to(dev)
    cls_attn = (torch.rand(B, N, generator=g) / n).to(dev)
    ta = (torch.randn(B, N, 128, generator=g) * 10).to(dev)
    x = torch.randn(B, N, d, generator=g).to(dev)
    score, thr, cnt, topk = lib.dtp_score(col_part, cls_attn, ta[:, 1:, :], n, T, temp)

    a = col_part.double().sum(1)[:, 1:]
    a = a / (a.sum(1, keepdim=True) + 1e-8)
    tb = ta[:, 1:, :T].double()
    bm = tb.max(2)[0]
    bm = bm / (bm.sum(1, keepdim=True) + 1e-8)
    s_ref = (a + bm + cls_attn[:, 1:].double()) / 3.0
    assert (score.double() - s_ref).abs().max().item() < 1e-8
    w = torch.softmax(tb / temp, dim=1)
    thr_ref = (w * score.double()[..., None]).sum(1).min(1)[0]
    assert (thr.double() - thr_ref).abs().max().item() < 1e-9
    cnt_ref = (score > thr[:, None]).sum(1)
    assert torch.equal(cnt.long(), cnt_ref)
    k = int(topk.item())
    assert k == int(cnt_ref.max())

    mask_in = torch.where(torch.rand(B, N, generator=g) < 0.3, -10000.0, 0.0).to(dev)
    for mode in (0, 1, 2):
        keep, dst, tail_w, tail_idx, mask_out = lib.dtp_select(score, topk, mask_mode=mode,
                                                               mask_in=mask_in if mode else None)
        if k < 1 or n - k <= 1:
            assert bool(keep.all())
            continue
        order = torch.sort(score, dim=1, descending=True, stable=True)[1]
        keep_ref = torch.zeros(B, n, dtype=torch.bool, device=dev)
        keep_ref.scatter_(1, order[:, :k], True)
        assert torch.equal(keep.bool(), keep_ref)
        out = lib.dtp_gather(x, topk, dst, tail_w, tail_idx, k)
        out_b, out16 = lib.dtp_gather(x, topk, dst, tail_w, tail_idx, k, want_f16=True)
        assert torch.equal(out_b, out) and torch.equal(out16, out.half())
        for b in range(B):
            idx = keep_ref[b].nonzero().flatten()
            assert torch.equal(out[b, 0], x[b, 0])
            assert torch.equal(out[b, 1:1 + k], x[b, 1 + idx])
            tail = (~keep_ref[b]).nonzero().flatten()
            wts = score[b, tail].double()
            wts = wts / (wts.sum() + 1e-8)
            merged = (wts[:, None] * x[b, 1 + tail].double()).sum(0)
            assert (out[b, 1 + k].double() - merged).abs().max().item() < 1e-5
            if mode == 1:
                exp = torch.cat([mask_in[b, :1], mask_in[b, 1 + order[b, :k + 1]]])
                assert torch.equal(mask_out[b, :k + 2], exp)
            if mode == 2:
                exp = torch.cat([mask_in[b, :1], mask_in[b, 1 + idx], mask_in[b, 1 + order[b, k:k + 1]]])
                assert torch.equal(mask_out[b, :k + 2], exp)


@pytest.mark.parametrize("B,n,T,ld,off", [(3, 1024, 100, 128, 0),     # the longest sequence the slab kernels take
                                         (2, 333, 97, 128, 0),      # T not a multiple of 4: scalar slab loads
                                         (2, 50, 128, 128, 0),      # every codebook column in use
                                         (5, 1, 100, 128, 0),       # a single prunable token
                                         (2, 77, 33, 131, 1),       # odd row pitch and a view that starts off 16 bytes
                                         (3, 19, 100, 2432, 2304)])  # the text encoder's fused q|k|v + codebook buffer
def test_token_statistics_edge_shapes(lib, dev, B, n, T, ld, off):
    """token_colstats (slab in shared memory) and dtp_score (4-CTA clusters): both load paths (16-byte cp.async and the
    scalar fallback), the maximum length, odd pitches -- against fp64 (models/vit.py:126-145, models/utils.py:174-178)."""
    g = torch.Generator(device="cpu").manual_seed(1000 * n + T)
    N = n + 1
    buf = (torch.randn(B, N, ld, generator=g) * 8).to(dev)
    ta = buf[:, 1:, off:off + T] if off + T <= ld else buf[:, 1:, :T]
    div, temp = math.sqrt(768), 3.0
    cm, cs = lib.token_colstats(ta, n, T, div)
    xs = ta.double() / div
    assert (cm.double() - xs.max(dim=1)[0]).abs().max().item() < 1e-6
    ref_sum = torch.exp(xs - xs.max(dim=1, keepdim=True)[0]).sum(1)
    assert ((cs.double() - ref_sum).abs() / ref_sum).max().item() < 2e-6
    n_parts = (N + 127) // 128
    col_part = torch.rand(B, n_parts, N, generator=g).to(dev)
    cls_attn = (torch.rand(B, N, generator=g) / n).to(dev)
    score, thr, cnt, topk = lib.dtp_score(col_part, cls_attn, ta, n, T, temp)
    a = col_part.double().sum(1)[:, 1:]
    a = a / (a.sum(1, keepdim=True) + 1e-8)
    bm = ta.double().max(2)[0]
    bm = bm / (bm.sum(1, keepdim=True) + 1e-8)
    s_ref = (a + bm + cls_attn[:, 1:].double()) / 3.0
    assert ((score.double() - s_ref).abs() / s_ref.abs().clamp_min(1e-12)).max().item() < 2e-6
    w = torch.softmax(ta.double() / temp, dim=1)
    thr_ref = (w * score.double()[..., None]).sum(1).min(1)[0]
    assert ((thr.double() - thr_ref).abs() / thr_ref.abs().clamp_min(1e-12)).max().item() < 2e-6
    assert torch.equal(cnt.long(), (score > thr[:, None]).sum(1))
    assert int(topk.item()) == int(cnt.max())


# ---------------------------------------------------------------------------------------------------------------
# tensor-core attention path: fused q|k|v projection with split / transposed epilogue, fwd + statistics
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,N,masked", [(2, 12, 197, False), (2, 12, 577, False), (1, 2, 128, False), (3, 4, 65, True),
                                          (2, 12, 346, False), (1, 12, 901, False)])
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_attention_tensor_core_path(lib, dev, monkeypatch, B, H, N, masked, variant):
    monkeypatch.setenv("MADTP_ATTN_VARIANT", str(variant))   # softmax warp geometry / CTAs per SM (attn_tc.cu)
    g = torch.Generator(device="cpu").manual_seed(N * 3 + B)
    K = 256
    HD = H * 64
    x = torch.randn(B * N, K, generator=g).to(dev)
    w = (torch.randn(3 * HD, K, generator=g) * 0.08).to(dev)
    bias = (torch.randn(3 * HD, generator=g) * 0.1).to(dev)
    x_hi, x_lo = lib.split_f16(x)
    w_hi, w_lo = lib.split_f16(w, 2.0 ** 14)
    qk_hi, qk_lo, vt_hi, vt_lo = lib.gemm_qkv(x_hi, x_lo, w_hi, w_lo, bias, N, H, alpha=2.0 ** -14)
    ref = x.double() @ w.double().T + bias.double()
    assert qk_hi.dtype == torch.float16 and vt_lo.dtype == torch.float16
    qk = (qk_hi.double() + qk_lo.double()) / lib.QK_PLANE_SCALE
    assert _rel(qk, ref[:, :2 * HD]) < 6e-7
    vt = ((vt_hi.double() + vt_lo.double()) / lib.V_PLANE_SCALE)[:, :N].reshape(B, H, 64, N)
    v_ref = ref[:, 2 * HD:].reshape(B, N, H, 64).permute(0, 2, 3, 1)
    assert _rel(vt, v_ref) < 6e-7

    mask = None
    if masked:
        lens = torch.randint(N // 2, N + 1, (B,), generator=g)
        mask = torch.zeros(B, N)
        for b in range(B):
            mask[b, lens[b]:] = -10000.0
        mask = mask.to(dev)
    scale = 0.125
    out = torch.empty(B, N, HD, device=dev, dtype=torch.float16)
    lse = torch.empty(B, H, N, device=dev)
    norm = torch.empty(B, H, N, device=dev)
    cls_p = torch.zeros(B, H, N, device=dev)
    cls_m = torch.zeros(B, H, (N + 63) // 64, device=dev)
    lib.attn_tc_fwd(qk_hi, qk_lo, vt_hi, vt_lo, B, H, N, scale, out, lse, norm, key_mask=mask, cls_p=cls_p, cls_tile_max=cls_m)
    n_parts = (N + 127) // 128
    col_part = torch.zeros(B, n_parts, N, device=dev)
    cls_attn = torch.zeros(B, N, device=dev)
    lib.attn_tc_stats(qk_hi, qk_lo, B, H, N, scale, lse, norm, col_part, cls_attn, cls_p, cls_m, key_mask=mask)

    # fp64 reference from the very operands the kernels consumed (so projection rounding does not enter)
    q = qk[:, :HD].reshape(B, N, H, 64).permute(0, 2, 1, 3).double()
    k = qk[:, HD:].reshape(B, N, H, 64).permute(0, 2, 1, 3).double()
    v = vt.permute(0, 1, 3, 2).double()
    p, o = _attn_ref(q, k, v, scale, None if mask is None else mask.double())
    o_merged = o.permute(0, 2, 1, 3).reshape(B, N, HD)
    assert _rel(out, o_merged) < 6e-4
    assert (norm.double() - o.norm(dim=-1)).abs().max().item() < 1e-5 * max(1.0, o.norm(dim=-1).max().item())
    s = q @ k.transpose(-1, -2) * scale
    if mask is not None:
        s = s + mask.double()[:, None, None, :]
    assert (lse.double() - torch.logsumexp(s, dim=-1)).abs().max().item() < 2e-5
    hi = o[..., 1:, :].norm(dim=-1)
    hi = hi / (hi.sum(dim=1, keepdim=True) + 1e-8)
    cls_ref = (p[:, :, 0, 1:] * hi).sum(dim=1)
    a_ref = p[:, :, 1:, 1:].max(dim=1)[0].sum(dim=1)
    assert (cls_attn[:, 1:].double() - cls_ref).abs().max().item() < 1e-6
    a = col_part.double().sum(dim=1)[:, 1:]
    assert ((a - a_ref).abs() / a_ref.abs().clamp_min(1e-3)).max().item() < 3e-6


@pytest.mark.parametrize("B,H,N", [(3, 8, 77), (2, 8, 27), (2, 4, 130), (1, 2, 300)])
def test_attention_tensor_core_causal(lib, dev, B, H, N):
    """causal = 1: the CLIP text tower (clip/model.py:452-457, mock.py:309-310; 77 tokens, 8 heads) on the tensor-core
    scoring-lane kernels -- context, log-sum-exp, norms and both pruning statistics against fp64."""
    g = torch.Generator(device="cpu").manual_seed(N * 7 + B)
    K, HD = 128, H * 64
    x = torch.randn(B * N, K, generator=g).to(dev)
    w = (torch.randn(3 * HD, K, generator=g) * 0.1).to(dev)
    bias = (torch.randn(3 * HD, generator=g) * 0.1).to(dev)
    x_hi, x_lo = lib.split_f16(x)
    w_hi, w_lo = lib.split_f16(w, 2.0 ** 14)
    qk_hi, qk_lo, vt_hi, vt_lo = lib.gemm_qkv(x_hi, x_lo, w_hi, w_lo, bias, N, H, alpha=2.0 ** -14)
    qk = (qk_hi.double() + qk_lo.double()) / lib.QK_PLANE_SCALE
    vt = ((vt_hi.double() + vt_lo.double()) / lib.V_PLANE_SCALE)[:, :N].reshape(B, H, 64, N)
    scale = 0.125
    out = torch.empty(B, N, HD, device=dev, dtype=torch.float16)
    lse = torch.empty(B, H, N, device=dev)
    norm = torch.empty(B, H, N, device=dev)
    cls_p = torch.zeros(B, H, N, device=dev)
    cls_m = torch.zeros(B, H, (N + 63) // 64, device=dev)
    lib.attn_tc_fwd(qk_hi, qk_lo, vt_hi, vt_lo, B, H, N, scale, out, lse, norm, cls_p=cls_p, cls_tile_max=cls_m, causal=True)
    col_part = torch.zeros(B, (N + 127) // 128, N, device=dev)
    cls_attn = torch.zeros(B, N, device=dev)
    lib.attn_tc_stats(qk_hi, qk_lo, B, H, N, scale, lse, norm, col_part, cls_attn, cls_p, cls_m, causal=True)
    q = qk[:, :HD].reshape(B, N, H, 64).permute(0, 2, 1, 3).double()
    k = qk[:, HD:].reshape(B, N, H, 64).permute(0, 2, 1, 3).double()
    v = vt.permute(0, 1, 3, 2).double()
    s = q @ k.transpose(-1, -2) * scale + torch.full((N, N), float("-inf"), dtype=torch.float64, device=dev).triu_(1)
    p = torch.softmax(s, dim=-1)
    o = p @ v
    assert _rel(out, o.permute(0, 2, 1, 3).reshape(B, N, HD)) < 6e-4
    assert (lse.double() - torch.logsumexp(s, dim=-1)).abs().max().item() < 2e-5
    assert (norm.double() - o.norm(dim=-1)).abs().max().item() < 1e-5 * max(1.0, o.norm(dim=-1).max().item())
    hi = o[..., 1:, :].norm(dim=-1)
    hi = hi / (hi.sum(dim=1, keepdim=True) + 1e-8)
    cls_ref = (p[:, :, 0, 1:] * hi).sum(dim=1)               # identically zero: the CLS query only sees itself
    a_ref = p[:, :, 1:, 1:].max(dim=1)[0].sum(dim=1)
    assert (cls_attn[:, 1:].double() - cls_ref).abs().max().item() < 1e-6
    a = col_part.double().sum(dim=1)[:, 1:]
    assert ((a - a_ref).abs() / a_ref.abs().clamp_min(1e-3)).max().item() < 3e-6


@pytest.mark.parametrize("B,n,T,d", [(3, 196, 100, 768), (2, 576, 100, 768), (4, 19, 100, 768), (2, 76, 100, 512)])
def test_query_sdft_tensor_core(lib, dev, B, n, T, d):
    """sd_ft = softmax_tokens(token_att / sqrt(d))^T . x on tcgen05 (operands re-laid out K-major in shared memory)."""
    g = torch.Generator(device="cpu").manual_seed(n + d)
    N = n + 1
    x = torch.randn(B, N, d, generator=g).to(dev)
    ta = (torch.randn(B, N, 128, generator=g) * 20).to(dev)
    ta_p = ta[:, 1:, :]
    div = math.sqrt(d)
    cm, cs = lib.token_colstats(ta_p, n, T, div)
    sd = torch.full((B, T, d), float("nan"), device=dev)
    lib.query_sdft_tc(ta_p, cm, cs, x.view(B * N, d), N, 1, n, T, div, sd, False)
    w = torch.softmax(ta_p[..., :T].double() / div, dim=1)
    ref = w.transpose(1, 2) @ x[:, 1:, :].double()
    assert _rel(sd, ref) < 5e-6
    lib.query_sdft_tc(ta_p, cm, cs, x.view(B * N, d), N, 1, n, T, div, sd, True)
    assert _rel(sd, 2 * ref) < 5e-6


@pytest.mark.parametrize("B,n,T,d", [(3, 196, 100, 768), (2, 576, 100, 768), (4, 19, 100, 768), (2, 76, 100, 512),
                                     (2, 345, 100, 768), (1, 64, 128, 384)])
def test_query_sdft_planes(lib, dev, B, n, T, d):
    """The same aggregation from the fp16 hi/lo planes of x with MN-major tensor-core operands (no transposition)."""
    g = torch.Generator(device="cpu").manual_seed(n + d + 1)
    N = n + 1
    x = (torch.randn(B, N, d, generator=g) * 3).to(dev)
    ta = (torch.randn(B, N, 128, generator=g) * 20).to(dev)
    ta_p = ta[:, 1:, :]
    div = math.sqrt(d)
    cm, cs = lib.token_colstats(ta_p, n, T, div)
    x_hi, x_lo = lib.split_f16(x.view(B * N, d))
    sd = torch.full((B, T, d), float("nan"), device=dev)
    lib.query_sdft_planes(ta_p, cm, cs, x_hi, x_lo, N, 1, n, T, div, sd, False)
    w = torch.softmax(ta_p[..., :T].double() / div, dim=1)
    ref = w.transpose(1, 2) @ x[:, 1:, :].double()
    assert _rel(sd, ref) < 5e-6
    lib.query_sdft_planes(ta_p, cm, cs, x_hi, x_lo, N, 1, n, T, div, sd, True)
    assert _rel(sd, 2 * ref) < 5e-6
    # device-resident length: a capacity-sized buffer holding B packed sequences of N_dyn tokens
    Nd = N - 5
    if Nd - 1 >= 8:
        xp = torch.zeros(B * N, d, device=dev)
        xp[:B * Nd] = x[:, :Nd].reshape(B * Nd, d)
        tap = torch.zeros(B * N, 128, device=dev)
        tap[:B * Nd] = ta[:, :Nd].reshape(B * Nd, 128)
        n_dev = torch.tensor([Nd], dtype=torch.int32, device=dev)
        ta_v = tap.view(B, N, 128)[:, 1:, :]
        cm2, cs2 = lib.token_colstats(ta_v, n, T, div, n_dev=n_dev, n_sub=1)
        h2, l2 = lib.split_f16(xp)
        sd2 = torch.full((B, T, d), float("nan"), device=dev)
        lib.query_sdft_planes(ta_v, cm2, cs2, h2, l2, N, 1, n, T, div, sd2, False, n_dev=n_dev)
        w2 = torch.softmax(ta[:, 1:Nd, :T].double() / div, dim=1)
        ref2 = w2.transpose(1, 2) @ x[:, 1:Nd, :].double()
        assert _rel(sd2, ref2) < 5e-6


@pytest.mark.parametrize("B,H,L,masked,causal", [(3, 12, 20, True, False), (2, 12, 35, True, False), (2, 8, 64, False, True),
                                                 (4, 12, 7, False, False), (1, 12, 1, False, False)])
def test_small_self_attention_fused_stats(lib, dev, B, H, L, masked, causal):
    g = torch.Generator(device="cpu").manual_seed(L * 5 + B)
    qkv = torch.randn(B, L, 3 * H * 64, generator=g).to(dev)
    q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
    mask = None
    if masked:
        lens = torch.randint(max(1, L // 2), L + 1, (B,), generator=g)
        mask = torch.zeros(B, L)
        for b in range(B):
            mask[b, lens[b]:] = -10000.0
        mask = mask.to(dev)
    out = torch.empty(B, L, H * 64, device=dev, dtype=torch.float16)
    col = torch.full((B, L), float("nan"), device=dev)
    cls_attn = torch.full((B, L), float("nan"), device=dev)
    lib.attn_small_self(q, k, v, H, 0.125, out, key_mask=mask, col_sum=col, cls_attn=cls_attn, causal=causal)

    def heads(t):
        return t.reshape(B, L, H, 64).permute(0, 2, 1, 3).double()
    s = heads(q) @ heads(k).transpose(-1, -2) * 0.125
    if mask is not None:
        s = s + mask.double()[:, None, None, :]
    if causal:
        s = s + torch.full((L, L), float("-inf"), device=dev, dtype=torch.float64).triu_(1)
    p = torch.softmax(s, dim=-1)
    o = p @ heads(v)
    assert _rel(out, o.permute(0, 2, 1, 3).reshape(B, L, H * 64)) < 6e-4
    if L > 1:
        hi = o[..., 1:, :].norm(dim=-1)
        hi = hi / (hi.sum(dim=1, keepdim=True) + 1e-8)
        cls_ref = (p[:, :, 0, 1:] * hi).sum(dim=1)
        a_ref = p[:, :, 1:, 1:].max(dim=1)[0].sum(dim=1)
        assert (cls_attn[:, 1:].double() - cls_ref).abs().max().item() < 1e-6
        assert ((col[:, 1:].double() - a_ref).abs() / a_ref.abs().clamp_min(1e-3)).max().item() < 2e-6
    out2 = torch.empty_like(out)
    lib.attn_small_self(q, k, v, H, 0.125, out2, key_mask=mask, causal=causal)
    assert torch.equal(out, out2)


# ---------------------------------------------------------------------------------------------------------------
# tensor-core cross-attention (value lane): text queries over image tokens
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,Lq,Nk,masked,broadcast", [(4, 12, 20, 255, False, False), (3, 12, 35, 197, True, False),
                                                        (2, 4, 128, 256, False, False), (5, 12, 20, 130, False, True),
                                                        (2, 12, 7, 64, True, False), (1, 2, 1, 3, False, False)])
def test_cross_attention_tensor_core(lib, dev, B, H, Lq, Nk, masked, broadcast):
    g = torch.Generator(device="cpu").manual_seed(Lq * 31 + Nk)
    C = H * 64
    # q and k are column slices of wider row-major buffers, as in the encoders (merged q01 / all-layer K projections);
    # every sequence owns P = Nk rounded up to 8 key rows / V^T columns (TMA box origins are 16-byte aligned)
    q_buf = torch.randn(B * Lq, 2 * C, generator=g).to(dev).half()
    q16 = q_buf.view(B, Lq, 2 * C)[..., C:]
    Bk = 1 if broadcast else B
    P = (Nk + 7) // 8 * 8
    k_buf = torch.randn(Bk, P, 3 * C, generator=g).to(dev).half()
    k16 = k_buf[:, :Nk, C:2 * C]
    v_pad = torch.randn(Bk, P, C, generator=g).to(dev).half()
    v = v_pad[:, :Nk]
    vt = torch.zeros(2 * C, Bk * P, device=dev, dtype=torch.float16)
    vt[C:] = v_pad.reshape(Bk * P, C).t()
    v_bias = torch.randn(C, generator=g).to(dev)
    mask = None
    if masked:
        lens = torch.randint(max(1, Nk // 2), Nk + 1, (B,), generator=g)
        mask = torch.zeros(B, Nk)
        for b in range(B):
            mask[b, lens[b]:] = -10000.0
        mask = mask.to(dev)
    out = torch.full((B, Lq, C), float("nan"), device=dev, dtype=torch.float16)
    scale = 0.125
    lib.attn_cross_tc(q16, k16[0] if broadcast else k16, vt[C:], H, scale, out, keys_per_batch=0 if broadcast else P,
                      v_bias=v_bias, key_mask=mask)
    qd = q16.double().view(B, Lq, H, 64).permute(0, 2, 1, 3)
    kd = k16.double().expand(B, Nk, C).reshape(B, Nk, H, 64).permute(0, 2, 1, 3)
    vd = v.double().expand(B, Nk, C).reshape(B, Nk, H, 64).permute(0, 2, 1, 3)
    s = qd @ kd.transpose(-1, -2) * scale
    if mask is not None:
        s = s + mask.double()[:, None, None, :]
    ref = (torch.softmax(s, dim=-1) @ vd).permute(0, 2, 1, 3).reshape(B, Lq, C) + v_bias.double()
    assert not torch.isnan(out).any()
    assert _rel(out, ref) < 1.5e-3


# ---------------------------------------------------------------------------------------------------------------
# fused DTP apply (radix select + compaction + merged token + LayerNorm) against the three kernels it replaces
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,n,d,mode", [(64, 576, 768, 0), (3, 196, 768, 0), (5, 1023, 512, 0), (4, 34, 768, 1),
                                        (4, 34, 768, 2), (2, 5, 128, 0), (32, 19, 768, 1)])
def test_dtp_apply_matches_select_gather_layernorm(lib, dev, B, n, d, mode):
    g = torch.Generator(device="cpu").manual_seed(n * 11 + B + mode)
    x = torch.randn(B, n + 1, d, generator=g).to(dev)
    score = (torch.rand(B, n, generator=g) * 1e-3 + 1e-3)
    score[:, ::7] = score[:, :1]                      # exact ties: the lower token index wins
    if n > 20:
        score[1 % B, 3] = -score[1 % B, 3]            # a negative score (vit.py:131-132 can produce them)
    score = score.to(dev)
    gamma, beta = torch.randn(d, generator=g).to(dev), torch.randn(d, generator=g).to(dev)
    mask_in = None
    if mode:
        mask_in = torch.where(torch.rand(B, n + 1, generator=g) < 0.3, -10000.0, 0.0).to(dev)
    for k in sorted({1, max(1, n // 3), max(1, n - 2)}):
        if n - k <= 1:
            continue
        topk = torch.tensor([k], dtype=torch.int32, device=dev)
        keep, dst, tw, ti, mo = lib.dtp_select(score, topk, mask_mode=mode, mask_in=mask_in)
        ref = lib.dtp_gather(x, topk, dst, tw, ti, k, want_f16=True)
        ref_out, ref16 = ref
        ref_ln = torch.empty(B * (k + 2), d, dtype=torch.float16, device=dev)
        lib.layernorm(ref_out.view(B * (k + 2), d), gamma, beta, 1e-6, y_f16=ref_ln)
        out, out16, ln16, keep2, mo2 = lib.dtp_apply(x, score, topk, k, mask_mode=mode, mask_in=mask_in, want_f16=True,
                                                     ln=(gamma, beta, 1e-6))
        assert torch.equal(keep2, keep), f"k={k}: keep flags"
        assert torch.equal(out, ref_out), f"k={k}: survivors / merged token differ by {(out - ref_out).abs().max().item():.3e}"
        assert torch.equal(out16, ref16)
        assert torch.equal(ln16.view(B * (k + 2), d), ref_ln), f"k={k}: fused LayerNorm"
        if mode:
            assert torch.equal(mo2[:, :k + 2], mo[:, :k + 2]), f"k={k}: pruned mask (mode {mode})"
        # exact top-k set against torch (ties -> lower index first = a stable descending sort)
        order = torch.sort(score.cpu(), dim=1, descending=True, stable=True)[1][:, :k]
        want = torch.zeros(B, n, dtype=torch.bool).scatter_(1, order, True)
        assert torch.equal(keep2.cpu().bool(), want), f"k={k}: not the exact top-k set"
