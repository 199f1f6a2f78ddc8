"""Timeline of one CTA of attn_fwd_tc_kernel (library built with MADTP_NVCC_EXTRA=-DMADTP_ATTN_TRACE).
python scripts/attn_trace.py [N] [variant]"""
import ctypes as C
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from madtp_b200 import _lib as lib
dev = torch.device("cuda:0")
B, H = 64, 12
N = int(sys.argv[1]) if len(sys.argv) > 1 else 346
os.environ["MADTP_ATTN_VARIANT"] = sys.argv[2] if len(sys.argv) > 2 else "1"
g = torch.Generator().manual_seed(0)
x = torch.randn(B * N, 768, generator=g).to(dev)
w = (torch.randn(3 * H * 64, 768, generator=g) * 0.03).to(dev)
bias = torch.zeros(3 * H * 64, device=dev)
xh, xl = lib.split_f16(x)
wh, wl = lib.split_f16(w, 2.0 ** 14)
out = torch.empty(B, N, H * 64, device=dev, dtype=torch.float16)
lse = torch.empty(B, H, N, device=dev)
norm = torch.empty(B, H, N, device=dev)
qk_hi, qk_lo, vt_hi, vt_lo = lib.gemm_qkv(xh, xl, wh, wl, bias, N, H, alpha=2.0 ** -14)
for _ in range(3):
    lib.attn_tc_fwd(qk_hi, qk_lo, vt_hi, vt_lo, B, H, N, 0.125, out, lse, norm)
torch.cuda.synchronize()
buf = np.zeros(4 * 64 * 8, dtype=np.uint64)
st = lib.load().madtp_debug_read_attn_trace(buf.ctypes.data_as(C.c_void_p))
assert st == 0, st
tr = buf.reshape(4, 64, 8).astype(np.int64)
t0 = tr[0, 63, 0]
T = min(3 * ((N + 63) // 64), 60)
print(f"N={N} variant={os.environ['MADTP_ATTN_VARIANT']} tiles={T}; cycles relative to kernel entry; CTA end {tr[0,63,1]-t0}")
print("tile | producer K,V issue | mma: qk_begin qk_sempty qk_issued v_full p_full pv_issued | softmax: begin s_full s_read max_exch exp_done drained p_stored")
for t in range(T):
    pr = [int(v - t0) if v else -1 for v in tr[0, t, :2]]
    mm = [int(v - t0) if v else -1 for v in tr[1, t, :6]]
    sm = [int(v - t0) if v else -1 for v in tr[2, t, :7]]
    print(t, "|", pr, "|", mm, "|", sm)
print("S in registers per softmax warp (warps 2..9), relative to warp 2's s_full:")
for t in range(T):
    print(t, [int(v - tr[2, t, 1]) if v else -1 for v in tr[3, t, :8]])
