"""Micro-benchmark of the small DTP kernels at BLIP-NLVR shapes (development aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import math
import torch
from madtp_b200 import _lib as lib
dev = torch.device("cuda:0")


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


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
