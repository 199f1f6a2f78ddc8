"""Development aid: dtp_select / dtp_gather with device-resident lengths against the exact-shape launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from madtp_b200 import _lib as L

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
bad = 0
for trial in range(40):
    B, d = 4, 768
    ncap = 196
    n = int(torch.randint(20, ncap + 1, (1,), generator=g))
    k = int(torch.randint(1, n - 2, (1,), generator=g))
    score = torch.rand(B, n, generator=g)
    x = torch.randn(B, n + 1, d, generator=g)
    topk = torch.tensor([k], dtype=torch.int32, device=dev)
    # exact shapes
    keep, dst, tw, ti, _ = L.dtp_select(score.to(dev), topk)
    out = L.dtp_gather(x.to(dev), topk, dst, tw, ti, k)
    # capacity shapes, packed
    sc = torch.full((B, ncap), float("nan"))
    sc.view(-1)[:B * n] = score.reshape(-1)
    xc = torch.full((B, ncap + 1, d), float("nan"))
    xc.view(-1)[:B * (n + 1) * d] = x.reshape(-1)
    lens = torch.tensor([n + 1, -7, -9], dtype=torch.int32, device=dev)
    keep2, dst2, tw2, ti2, _ = L.dtp_select(sc.to(dev), topk, n_dev=lens[0:1], n_out=lens[1:2], k_out=lens[2:3])
    out2 = L.dtp_gather(xc.to(dev), topk, dst2, tw2, ti2, ncap - 1, n_dev=lens[0:1])
    torch.cuda.synchronize()
    ok = (torch.equal(keep2.view(-1)[:B * n], keep.view(-1)) and torch.equal(dst2.view(-1)[:B * n], dst.view(-1))
          and torch.equal(out2.view(-1)[:B * (k + 2) * d], out.view(-1)) and lens.tolist() == [n + 1, k + 2, k])
    if not ok:
        bad += 1
        print("MISMATCH trial", trial, "n", n, "k", k, "lens", lens.tolist(),
              "keep", torch.equal(keep2.view(-1)[:B * n], keep.view(-1)),
              "dst", torch.equal(dst2.view(-1)[:B * n], dst.view(-1)),
              "out", torch.equal(out2.view(-1)[:B * (k + 2) * d], out.view(-1)))
print("bad", bad)
