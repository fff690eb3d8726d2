"""N > 1 host logic on CPU: gloo, world_size 2 - sharding of primaries and the tally all-reduce."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from petite_b200.distributed import shard, allreduce_sum, gather_history


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


def _fake_history(rank):
    """A rank's to_host()-style dictionary: 3 + rank primaries, two daughters each; contents encode (rank, local slot)."""
    n0 = 3 + rank
    n = 3 * n0
    h = {"p0": np.zeros((n, 4)), "pid": np.full(n, 22, dtype=np.int32), "parent": np.full(n, -1, dtype=np.int32),
         "shower": np.zeros(n, dtype=np.int32), "weight": np.ones(n)}
    h["p0"][:, 0] = 1000 * rank + np.arange(n)
    h["shower"][:n0] = np.arange(n0)
    for k in range(n0):
        for b in range(2):
            s = n0 + 2 * k + b
            h["parent"][s] = k
            h["shower"][s] = k
    return h, n0


def _history_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    h, n0 = _fake_history(rank)
    g = gather_history(h, n0, dst=0)
    assert (g is None) == (rank != 0)
    if rank == 0:
        np.savez(out, **g)
    dist.destroy_process_group()


def test_two_rank_history_gather_keeps_parent_links(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "h.npz")
    mp.spawn(_history_worker, args=(2, port, out), nprocs=2, join=True)
    g = np.load(out)
    (h0, a0), (h1, a1) = _fake_history(0), _fake_history(1)
    n_0, n_1 = len(h0["pid"]), len(h1["pid"])
    assert g["record_offsets"].tolist() == [0, n_0, n_0 + n_1] and g["shower_offsets"].tolist() == [0, a0, a0 + a1]
    assert np.array_equal(g["p0"], np.concatenate([h0["p0"], h1["p0"]])) and g["p0"].dtype == np.float64 and g["pid"].dtype == np.int32
    par = g["parent"]
    assert np.array_equal(par[:n_0], h0["parent"]) and np.array_equal(par[n_0:][h1["parent"] < 0], h1["parent"][h1["parent"] < 0])
    kids = np.nonzero(par >= 0)[0]
    # a daughter's parent is a record of the same rank and the same shower, also after the shift
    assert np.all((par[kids] >= n_0) == (kids >= n_0)) and np.array_equal(g["shower"][par[kids]], g["shower"][kids])
    assert np.array_equal(np.unique(g["shower"]), np.arange(a0 + a1))
    single = gather_history(h0, a0)              # no process group: the rank's own history plus the offsets
    assert np.array_equal(single["p0"], h0["p0"]) and single["record_offsets"].tolist() == [0, n_0]
