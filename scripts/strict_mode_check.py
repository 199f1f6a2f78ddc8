"""Multi-GPU strict mode (madtp_b200.dist.global_topk): every rank runs its shard of a batch and must reproduce, bit for
bit, the rows a single process computes for the whole batch. Launch with torchrun on 2+ GPUs:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/strict_mode_check.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
from madtp_b200 import dist as mdist, synthetic
from madtp_b200.blip_nlvr import BLIP_NLVR, TokenizedText

rank, local_rank, world = mdist.init("nccl")
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
size, pairs, temp = 224, 8, 3.0
model = BLIP_NLVR(image_size=size, evaluate=True)
model.load_state_dict(synthetic.blip_nlvr_state_dict(1234, img_size=size), strict=False)
model = model.to(dev).eval()
images, ids, mask = synthetic.nlvr_inputs(pairs, size, 20, seed=0)


def run(lo, hi):
    img = torch.cat([images[:pairs][lo:hi], images[pairs:][lo:hi]], 0).to(dev)
    return model(img, TokenizedText(ids[lo:hi].to(dev), mask[lo:hi].to(dev)), hi - lo, temp, train=False)


lo, hi = mdist.shard_range(pairs, rank, world)
with torch.no_grad():
    mdist.global_topk(False)
    local_default = run(lo, hi)
    ks_default = [b.last_prune.k for b in model.visual_encoder.blocks if b.last_prune is not None]
    mdist.global_topk(True)
    local_strict = run(lo, hi)
    ks_strict = [b.last_prune.k for b in model.visual_encoder.blocks if b.last_prune is not None]
    mdist.global_topk(False)
    full = run(0, pairs)                      # every rank also runs the whole batch on its own (no collective)
    ks_full = [b.last_prune.k for b in model.visual_encoder.blocks if b.last_prune is not None]
ok_strict = torch.equal(local_strict, full[lo:hi]) and ks_strict == ks_full
same_default = torch.equal(local_default, full[lo:hi])
flags = torch.tensor([int(ok_strict), int(same_default)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world}: strict shard == whole batch on every rank: {bool(flags[0])}; "
          f"default (local topk_num) == whole batch: {bool(flags[1])}; k trajectory full {ks_full[:4]} "
          f"rank0 default {ks_default[:4]}")
    assert bool(flags[0]), "strict mode must reproduce the single-process result"
dist.destroy_process_group()
