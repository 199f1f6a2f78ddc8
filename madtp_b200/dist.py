"""Batch-only sharding of the pruned forward over the GPUs of one node (one process per GPU, torch.distributed).

The path shards over samples only: weights are replicated with ONE broadcast of a flat blob
from rank 0, every rank runs the forward on its own pairs (topk_num is the max over the LOCAL batch, which is the
reference's own multi-GPU evaluation semantics, compress_nlvr_dtp.py:131,210-211), and the logits are all-gathered.
There is no collective on the data path between those two points unless the strict mode below is switched on.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def init(backend: str = "nccl"):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `total` samples owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_parameters(module: torch.nn.Module, src: int = 0):
    """Replicates rank `src`'s parameters and buffers with a single broadcast of one flat fp32 blob. The copies go
    through `param.detach()`, which shares the parameter's version counter (unlike `.data`), so the prepared-weight
    caches (functional.WeightCache) see the update; they are cleared as well for good measure."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    tensors: List[torch.Tensor] = [p.detach() for p in module.parameters()] + \
        [b.detach() for b in module.buffers() if b.is_floating_point()]
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in tensors])
    dist.broadcast(flat, src=src)
    off = 0
    with torch.no_grad():
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
    from .functional import clear_caches
    clear_caches(module)
    return flat.numel() * 4


def all_gather_rows(x: torch.Tensor) -> torch.Tensor:
    """Concatenates every rank's [B_local, ...] tensor along dim 0 (equal B_local on all ranks)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    out = torch.empty((dist.get_world_size() * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous())
    return out


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_over_ranks(value: float, device) -> List[float]:
    """Every rank's scalar, in rank order (diagnostics: which GPU of the node set the max-over-ranks time)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [value]
    t = torch.tensor([value], dtype=torch.float64, device=device)
    out = torch.empty(dist.get_world_size(), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, t)
    return [float(v) for v in out.tolist()]


# ---- optional strict mode: the pruning of a sharded batch matches the single-process run on the whole batch ---------
# topk_num is a maximum over the batch (models/vit.py:145), so sharding the batch changes how many tokens each sample
# keeps. The reference's multi-GPU evaluation lives with that (every rank takes the max over its LOCAL batch); with
# global_topk(True) one all-reduce(MAX) of the int32 scalar per pruned layer and modality (24 per forward, latency-bound)
# runs before the read-back, and every rank prunes exactly as one process holding all samples would.
_GLOBAL_TOPK = bool(int(os.environ.get("MADTP_GLOBAL_TOPK", "0")))


def global_topk(enable: bool = None) -> bool:
    global _GLOBAL_TOPK
    if enable is not None:
        _GLOBAL_TOPK = bool(enable)
    return _GLOBAL_TOPK


def allreduce_topk_(topk: torch.Tensor) -> torch.Tensor:
    """In-place all-reduce(MAX) of the per-layer survivor count when the strict mode is on; otherwise a no-op."""
    if _GLOBAL_TOPK and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(topk, op=dist.ReduceOp.MAX)
    return topk


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
