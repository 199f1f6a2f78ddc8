"""One launch pattern of the tensor-core attention kernels for ncu captures: python scripts/attn_one.py [N]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from madtp_b200 import _lib as lib
dev = torch.device("cuda:0")
B, H = 64, 12
N = int(sys.argv[1]) if len(sys.argv) > 1 else 346
g = torch.Generator().manual_seed(0)
x = torch.randn(B * N, 768, generator=g).to(dev)
w = (torch.randn(3 * H * 64, 768, generator=g) * 0.03).to(dev)
bias = torch.zeros(3 * H * 64, device=dev)
xh, xl = lib.split_f16(x)
wh, wl = lib.split_f16(w, 2.0 ** 14)
out = torch.empty(B, N, H * 64, device=dev, dtype=torch.float16)
lse = torch.empty(B, H, N, device=dev)
norm = torch.empty(B, H, N, device=dev)
n_parts = (N + 127) // 128
col = torch.empty(B, n_parts, N, device=dev)
cls = torch.empty(B, N, device=dev)
cls_p = torch.zeros(B, H, N, device=dev)
cls_m = torch.zeros(B, H, (N + 63) // 64, device=dev)
for _ in range(2):
    qk_hi, qk_lo, vt_hi, vt_lo = lib.gemm_qkv(xh, xl, wh, wl, bias, N, H, alpha=2.0 ** -14)
    lib.attn_tc_fwd(qk_hi, qk_lo, vt_hi, vt_lo, B, H, N, 0.125, out, lse, norm, cls_p=cls_p, cls_tile_max=cls_m)
    lib.attn_tc_stats(qk_hi, qk_lo, B, H, N, 0.125, lse, norm, col, cls, cls_p, cls_m)
torch.cuda.synchronize()
