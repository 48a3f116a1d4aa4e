"""CPU, world_size 2, gloo: the host-side sharding logic of attentive_dfprior_b200.dist."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from attentive_dfprior_b200 import dist as D


def test_shard_ranges_partition():
    for n in (0, 1, 7, 1000, 816000):
        for world in (1, 2, 3, 8):
            r = [D.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    for world in (1, 2, 4, 8):
        r = [D.chunk_aligned_range(816000, 100000, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == 816000
        assert all(a % 100000 == 0 for a, _ in r)
        assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))


def _worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = dist.group.WORLD
        # global depth maximum: every rank sees the maximum over all shards
        full = torch.arange(10, dtype=torch.float32) * 0.37
        lo, hi = D.shard_range(10, rank, world)
        m = D.global_depth_max(full[lo:hi], g)
        assert float(m) == float(full.max())
        # ragged gather: shards of different lengths come back in rank order
        n = 11
        lo, hi = D.shard_range(n, rank, world)
        x = torch.arange(n, dtype=torch.float64).reshape(n, 1).repeat(1, 3)
        sizes = [b - a for a, b in (D.shard_range(n, r, world) for r in range(world))]
        got = D.gather_concat(x[lo:hi].contiguous(), sizes, g)
        assert torch.equal(got, x)
        only0 = D.gather_concat(x[lo:hi].contiguous(), sizes, g, dst=0)
        assert (only0 is not None) == (rank == 0)
        # gradient all-reduce semantics used by the sharded backward: sum of per-rank partial sums
        part = torch.full((5,), float(rank + 1))
        dist.all_reduce(part, group=g)
        assert torch.equal(part, torch.full((5,), 3.0))
    finally:
        dist.destroy_process_group()


def test_gloo_world2():
    mp.spawn(_worker, args=(2, 29533), nprocs=2, join=True)


def test_interleaved_shard_partitions_the_batch():
    from attentive_dfprior_b200 import dist as D
    n, world = 1003, 8
    shards = [D.interleaved_shard(n, r, world) for r in range(world)]
    allidx = torch.cat(shards).sort().values
    assert torch.equal(allidx, torch.arange(n))
    assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    # every rank sees every contiguous block (frame) of the batch in the same proportion
    block = n // 4
    for s in shards:
        counts = [int(((s >= b * block) & (s < (b + 1) * block)).sum()) for b in range(4)]
        assert max(counts) - min(counts) <= 1
