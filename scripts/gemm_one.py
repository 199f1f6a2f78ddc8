"""One GEMM launch pattern for ncu captures: python scripts/gemm_one.py f16|tf32 M N K [res] [act]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from madtp_b200 import _lib as L
dev = torch.device("cuda:0")
kind, M, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
res = len(sys.argv) > 5 and sys.argv[5] == "1"
act = int(sys.argv[6]) if len(sys.argv) > 6 else 0
f16out = len(sys.argv) > 7 and sys.argv[7] == "1"
a = torch.randn(M, K, device=dev)
b = torch.randn(N, K, device=dev)
bias = torch.randn(N, device=dev)
rs = torch.randn(M, N, device=dev) if res else None
out = torch.empty(M, N, device=dev, dtype=torch.float16 if f16out else torch.float32)
if kind == "f16":
    a, b = a.half(), b.half()
    for _ in range(3):
        L.gemm(L.GEMM_F16, a, b, out, bias=bias, residual=rs, act=act)
else:
    ah, al = L.split_tf32(a)
    bh, bl = L.split_tf32(b)
    for _ in range(3):
        L.gemm(L.GEMM_TF32X3, ah, bh, out, a_lo=al, b_lo=bl, bias=bias, residual=rs, act=act)
torch.cuda.synchronize()
