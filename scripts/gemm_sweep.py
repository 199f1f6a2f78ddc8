"""GEMM micro-benchmark over the shapes of the BLIP-NLVR forward (development aid, not part of the bench contract)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from madtp_b200 import _lib as L

dev = torch.device("cuda:0")
L.load()


def timeit(fn, n=20):
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


def f16case(M, N, K, out_f16, bias, res, act, name):
    a = torch.randn(M, K, device=dev).half()
    b = torch.randn(N, K, device=dev).half()
    out = torch.empty(M, N, device=dev, dtype=torch.float16 if out_f16 else torch.float32)
    bs = torch.randn(N, device=dev) if bias else None
    rs = torch.randn(M, N, device=dev) if res else None
    us = timeit(lambda: L.gemm(L.GEMM_F16, a, b, out, bias=bs, residual=rs, act=act))
    print(f"f16 {name:28s} M={M} N={N} K={K}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")


def tf32case(M, N, K, name):
    a = torch.randn(M, K, device=dev)
    b = torch.randn(N, K, device=dev)
    ah, al = L.split_tf32(a)
    bh, bl = L.split_tf32(b)
    out = torch.empty(M, N, device=dev)
    bs = torch.randn(N, device=dev)
    us = timeit(lambda: L.gemm(L.GEMM_TF32X3, ah, bh, out, a_lo=al, b_lo=bl, bias=bs))
    print(f"tf32x3 {name:25s} M={M} N={N} K={K}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s (algorithmic)")


M = 36928
f16case(M, 768, 768, False, False, False, 0, "plain f32 out")
f16case(M, 768, 768, False, True, False, 0, "bias")
f16case(M, 768, 768, False, True, True, 0, "bias+residual (proj)")
f16case(M, 768, 768, True, True, False, 0, "bias, f16 out")
f16case(M, 3072, 768, True, True, False, 0, "bias, f16 out")
f16case(M, 3072, 768, True, True, False, 1, "bias+GELU, f16 out (fc1)")
f16case(M, 768, 3072, False, True, True, 0, "bias+residual (fc2)")
f16case(M, 2304, 768, False, True, False, 0, "qkv-shaped f16")
f16case(8192, 18432, 768, False, True, False, 0, "cross K/V all layers")
f16case(640, 768, 768, False, True, True, 0, "text dense")
f16case(640, 3072, 768, True, True, False, 1, "text fc1")
tf32case(M, 2304, 768, "qkv")
tf32case(M, 128, 768, "token_att")
tf32case(M, 768, 768, "patch embed")
tf32case(640, 2304, 768, "text qkv")
