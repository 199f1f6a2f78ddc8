"""Accuracy and speed of the fp16 hi/lo split GEMM against the tf32 hi/lo one (both 3 MMAs per product) and fp64.
python scripts/f16x3_probe.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import math
import torch
from madtp_b200 import _lib as L

dev = torch.device("cuda:0")
torch.manual_seed(0)


def pow2_scale(w, target=2.0 ** 14):
    m = float(w.abs().max())
    return 2.0 ** math.floor(math.log2(target / m)) if m > 0 else 1.0


def run(M, N, K, w_std, x_kind):
    if x_kind == "ln":
        x = torch.randn(M, K, device=dev)
    else:   # residual stream with outlier channels
        x = torch.randn(M, K, device=dev) * 0.7
        x[:, ::97] *= 40.0
    w = torch.randn(N, K, device=dev) * w_std
    bias = torch.randn(N, device=dev) * 0.1
    ref = (x.double() @ w.double().t() + bias.double())
    out_t = torch.empty(M, N, device=dev)
    out_h = torch.empty(M, N, device=dev)
    ah, al = L.split_tf32(x)
    bh, bl = L.split_tf32(w)
    s = pow2_scale(w)
    xh, xl = L.split_f16(x, 1.0)
    wh, wl = L.split_f16(w, s)

    def t_tf32():
        L.gemm(L.GEMM_TF32X3, ah, bh, out_t, a_lo=al, b_lo=bl, bias=bias)

    def t_f16():
        L.gemm(L.GEMM_F16X3, xh, wh, out_h, a_lo=xl, b_lo=wl, bias=bias, alpha=1.0 / s)

    res = {}
    for nm, fn in (("tf32x3", t_tf32), ("f16x3", t_f16)):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[nm] = e0.elapsed_time(e1) / 10 * 1e3
    torch.backends.cuda.matmul.allow_tf32 = False
    out_c = x @ w.t() + bias
    den = ref.abs().mean()
    errs = {nm: (float((o.double() - ref).abs().max() / den), float((o.double() - ref).abs().mean() / den))
            for nm, o in (("tf32x3", out_t), ("f16x3", out_h), ("cublas_fp32", out_c))}
    print(f"M={M} N={N} K={K} w_std={w_std} x={x_kind} scale=2^{int(math.log2(s))}")
    for nm in errs:
        print(f"   {nm:12s} max/mean|ref| {errs[nm][0]:.3e}  mean {errs[nm][1]:.3e}   {res.get(nm, float('nan')):8.1f} us")


run(36928, 2304, 768, 0.02, "ln")
run(22208, 2304, 768, 0.02, "ln")
run(22208, 2304, 768, 0.05, "resid")
run(640, 2304 + 128, 768, 0.02, "ln")
run(12544, 768, 768, 1.0, "ln")
run(4096, 256, 4096, 0.02, "ln")
