"""The N>1 plumbing (shard ranges, one-blob weight broadcast, logit all-gather, max-over-ranks timing) with
world_size 2 over gloo on the CPU -- the data path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from madtp_b200 import dist as mdist
    r, lr, w = mdist.init("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(rank)                       # every rank starts from different weights
    model = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.LayerNorm(8))
    nbytes = mdist.broadcast_parameters(model, src=0)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    lo, hi = mdist.shard_range(10, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32).unsqueeze(1).repeat(1, 2)      # this rank's "logits"
    gathered = mdist.all_gather_rows(local)
    t = mdist.max_over_ranks(float(rank + 1), torch.device("cpu"))
    per_rank = mdist.gather_over_ranks(float(10 * rank + 1), torch.device("cpu"))
    # strict mode: the per-layer survivor count becomes the maximum over all ranks; off by default (reference semantics)
    k_local = torch.tensor([100 + 7 * rank], dtype=torch.int32)
    k_off = int(mdist.allreduce_topk_(k_local.clone()))
    mdist.global_topk(True)
    k_on = int(mdist.allreduce_topk_(k_local.clone()))
    mdist.global_topk(False)
    mdist.barrier()
    q.put((rank, nbytes, flat.tolist(), gathered.tolist(), t, k_off, k_on, per_rank))
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from madtp_b200.dist import shard_range
    for total in (0, 1, 7, 32, 512):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_broadcast_and_gather_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, n0, w0, g0, t0, koff0, kon0, pr0), (r1, n1, w1, g1, t1, koff1, kon1, pr1) = res
    assert pr0 == pr1 == [1.0, 11.0], "every rank sees every rank's own step time, in rank order"
    assert (koff0, koff1) == (100, 107), "default: every rank keeps its local topk_num"
    assert kon0 == kon1 == 107, "strict mode: the maximum over all ranks"
    assert n0 == n1 == (8 * 8 + 8 + 8 + 8) * 4
    assert w0 == w1, "weights must be identical after the broadcast from rank 0"
    assert g0 == g1 == [[float(i), float(i)] for i in range(10)], "all-gather must concatenate the shards in rank order"
    assert t0 == t1 == 2.0
