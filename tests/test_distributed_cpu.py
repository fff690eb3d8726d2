"""N > 1 host logic on CPU: gloo, world_size 2 - sharding of primaries and the tally all-reduce."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from petite_b200.distributed import shard, allreduce_sum


def test_shard_partitions_all_primaries():
    for n in (0, 1, 7, 100000, 100003):
        for w in (1, 2, 4, 8):
            parts = [shard(n, r, w) for r in range(w)]
            assert sum(c for _, c in parts) == n
            assert parts[0][0] == 0 and all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def _fake_tally(first, count):
    """Deterministic per-shower 'tally' that depends only on the global shower id."""
    t = torch.zeros(1024, dtype=torch.float64)
    for i in range(first, first + count):
        g = np.random.default_rng(i)
        t.index_add_(0, torch.from_numpy(g.integers(1, 1024, size=16)), torch.from_numpy(g.random(16)))
        t[0] += 1
    return t


def _worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard(n, rank, world)
    t = _fake_tally(first, count)
    allreduce_sum(t)
    if rank == 0:
        torch.save(t, out)
    dist.destroy_process_group()


def test_two_rank_tallies_equal_single_process(tmp_path):
    n = 101
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "t.pt")
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    got = torch.load(out)
    want = _fake_tally(0, n)
    assert got[0].item() == n
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)
