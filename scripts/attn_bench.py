"""Micro-benchmark of the tensor-core attention kernels at BLIP-NLVR layer shapes (development aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os
import torch
from madtp_b200 import _lib as lib
dev = torch.device("cuda:0")
B, H = 64, 12
for N in (577, 346, 256):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B * N, 768, generator=g).to(dev)
    w = (torch.randn(3 * H * 64, 768, generator=g) * 0.03).to(dev)
    bias = torch.zeros(3 * H * 64, device=dev)
    xh, xl = lib.split_f16(x); wh, wl = lib.split_f16(w, 2.0 ** 14)
    qk_hi, qk_lo, vt_hi, vt_lo = lib.gemm_qkv(xh, xl, wh, wl, bias, N, H, alpha=2.0 ** -14)
    out = torch.empty(B, N, H * 64, device=dev, dtype=torch.float16)
    lse = torch.empty(B, H, N, device=dev); norm = torch.empty(B, H, N, device=dev)
    n_parts = (N + 127) // 128
    col = torch.empty(B, n_parts, N, device=dev); cls = torch.empty(B, N, device=dev)
    cls_p = torch.zeros(B, H, N, device=dev); cls_m = torch.zeros(B, H, (N + 63) // 64, device=dev)
    def t(fn, n=10):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    fv = []
    for var in ("0", "1", "2"):
        os.environ["MADTP_ATTN_VARIANT"] = var
        fv.append(t(lambda: lib.attn_tc_fwd(qk_hi, qk_lo, vt_hi, vt_lo, B, H, N, 0.125, out, lse, norm, cls_p=cls_p, cls_tile_max=cls_m)))
    os.environ.pop("MADTP_ATTN_VARIANT", None)
    f = min(fv)
    s = t(lambda: lib.attn_tc_stats(qk_hi, qk_lo, B, H, N, 0.125, lse, norm, col, cls, cls_p, cls_m))
    q = t(lambda: lib.gemm_qkv(xh, xl, wh, wl, bias, N, H, alpha=2.0 ** -14))
    print(f"N={N}: attn_tc_fwd variants {fv[0]:.1f} / {fv[1]:.1f} / {fv[2]:.1f} us ({4.0*B*H*N*N*64/f/1e6:.0f} TFLOP/s alg), attn_tc_stats {s:.1f} us, gemm_qkv {q:.1f} us")
