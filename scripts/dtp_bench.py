"""Micro-benchmark of the small DTP kernels at BLIP-NLVR shapes (development aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import math
import torch
from madtp_b200 import _lib as lib
dev = torch.device("cuda:0")


def t(fn, n=20):
    """us per launch inside a CUDA graph of n back-to-back launches (no host launch cost in the number)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
            for _ in range(n):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * n) * 1e3


for B, n in ((64, 576), (64, 345), (64, 255), (32, 19)):
    g = torch.Generator().manual_seed(0)
    N, T = n + 1, 100
    ta = (torch.randn(B, N, 128, generator=g) * 5).to(dev)
    ta_p = ta[:, 1:, :]
    n_parts = (N + 127) // 128
    col_part = torch.rand(B, n_parts, N, generator=g).to(dev)
    cls_attn = torch.rand(B, N, generator=g).to(dev)
    c = t(lambda: lib.token_colstats(ta_p, n, T, math.sqrt(768)))
    s = t(lambda: lib.dtp_score(col_part, cls_attn, ta_p[:, :, :T], n, T, 3.5894))
    print(f"B={B} n={n}: token_colstats {c:.1f} us, dtp_score {s:.1f} us")

# Query_model aggregation: fp32-x kernel (builder warps transpose) against the MN-major plane kernel
for B, n in ((64, 576), (64, 345), (64, 255)):
    g = torch.Generator().manual_seed(1)
    N, T, d = n + 1, 100, 768
    x = torch.randn(B, N, d, generator=g).to(dev)
    ta = (torch.randn(B, N, 128, generator=g) * 5).to(dev)
    ta_p = ta[:, 1:, :]
    div = math.sqrt(d)
    cm, cs = lib.token_colstats(ta_p, n, T, div)
    hi, lo = lib.split_f16(x.view(B * N, d))
    sd = torch.zeros(B, T, d, device=dev)
    a = t(lambda: lib.query_sdft_tc(ta_p, cm, cs, x.view(B * N, d), N, 1, n, T, div, sd, False))
    p = t(lambda: lib.query_sdft_planes(ta_p, cm, cs, hi, lo, N, 1, n, T, div, sd, False))
    print(f"B={B} n={n}: query_sdft_tc {a:.1f} us, query_sdft_planes {p:.1f} us")

# dtp_apply at ViT shapes (select + gather + merged token + LayerNorm)
for B, n, k in ((64, 576, 410), (64, 411, 345), (64, 346, 319), (64, 260, 257)):
    g = torch.Generator().manual_seed(2)
    d = 768
    x = torch.randn(B, n + 1, d, generator=g).to(dev)
    score = torch.rand(B, n, generator=g).to(dev)
    topk = torch.tensor([k], dtype=torch.int32, device=dev)
    gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    a = t(lambda: lib.dtp_apply(x, score, topk, k, ln=(gamma, beta, 1e-6)))
    mb = (B * (n + 1) * d * 4 + B * (k + 2) * d * 6) / 1e6
    print(f"B={B} n={n} k={k}: dtp_apply {a:.1f} us, {mb / a:.2f} TB/s algorithmic")
