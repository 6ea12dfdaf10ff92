"""world_size-2 gloo test of the multi-GPU host logic (sample sharding, count reduction, logits gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from peekvit_b200 import sharding


def test_shard_range_partitions():
    for total in (0, 1, 7, 2048, 2049):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, n_cls):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        logits = torch.randn(total, n_cls, generator=g)            # what a single process would compute
        labels = torch.randint(0, n_cls, (total,), generator=g)
        b, e = sharding.shard_range(total, rank, world)
        local = logits[b:e]
        counts = sharding.reduce_counts((local.argmax(1) == labels[b:e]).sum(), torch.tensor(e - b))
        assert counts.tolist() == [int((logits.argmax(1) == labels).sum()), total]
        assert sharding.max_over_ranks(1.0 + rank) == float(world)  # pass time of the slowest rank
        full = sharding.gather_logits(local, total)
        assert torch.equal(full, logits)                           # identical per-sample logits in global order
    finally:
        dist.destroy_process_group()


def test_two_rank_counts_and_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 37, 10), nprocs=2, join=True)
