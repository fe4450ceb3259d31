"""CPU, world_size 2 over gloo: the N>1 host logic (episode sharding + the logits all-gather) of vima_b200.dist."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vima_b200.dist import shard_range


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vima_b200.dist import all_gather_logits

    n_total = 6
    a, b = shard_range(n_total, rank, world)
    full = torch.arange(n_total * 700, dtype=torch.float32).view(n_total, 700)
    got = all_gather_logits(full[a:b].clone())
    ok = torch.equal(got, full)
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the bench's max-over-ranks timing reduction
    q.put((rank, bool(ok), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert res == [(0, True, 2.0), (1, True, 2.0)]


def test_shard_range_partitions():
    for n in (0, 1, 7, 2048):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
