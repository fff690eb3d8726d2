"""Multi-GPU plumbing: showers are independent, so primaries are sharded across ranks and only the tally buffer is
all-reduced (one NCCL all-reduce over NVLink per step; ``gloo`` on CPU for the tests)."""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard(n_total, rank, world):
    """Contiguous block of primaries for ``rank``: (first_index, count).  Shower i keeps its global id, and therefore
    its Philox root key, whatever the number of ranks."""
    base, rem = divmod(int(n_total), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def allreduce_sum(tensor):
    """Sum a tally tensor over all ranks in place (no-op for a single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def gather_history(history, n_primaries, dst=0, group=None):
    """Full-history mode across ranks ("no collective; host concatenates per-GPU stacks").

    ``history`` is this rank's ``ShowerBatch.to_host()`` dictionary (NumPy arrays, one row per record, the rank's ``n_primaries``
    primaries first).  Nothing touches the stepping path: after the showers are done every rank ships its host arrays to rank
    ``dst`` point to point (``group`` must be able to move CPU tensors - a ``gloo`` group next to the NCCL one:
    ``dist.new_group(backend="gloo")``; 8 x 25 GB of config-2 history does not fit one GPU, host memory is where it belongs).
    On ``dst`` the arrays are concatenated in rank order with the rank-local indices made global: ``parent`` (stack slot of the
    parent, -1 for primaries) is shifted by the rank's record offset and ``shower`` by its primary offset, so
    ``out["p0"][out["parent"][k]]`` is still record k's parent.  -> on ``dst``: the dictionary plus ``record_offsets`` and
    ``shower_offsets`` (world + 1 entries each: rank r owns records [record_offsets[r], record_offsets[r + 1])); ``None`` elsewhere.
    A single process gets its own history back (with the two offset arrays)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    keys = sorted(history)
    n_local = int(len(history[keys[0]])) if keys else 0
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = dict(history)
        out["record_offsets"] = np.array([0, n_local], dtype=np.int64)
        out["shower_offsets"] = np.array([0, int(n_primaries)], dtype=np.int64)
        return out
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sizes = [None] * world
    dist.all_gather_object(sizes, (n_local, int(n_primaries), [(k, str(history[k].dtype), tuple(history[k].shape[1:])) for k in keys]), group=group)
    if any(s[2] != sizes[0][2] for s in sizes):
        raise ValueError("gather_history: the ranks' histories have different columns")
    if rank != dst:
        for k in keys:
            if n_local:
                dist.send(torch.from_numpy(np.ascontiguousarray(history[k])), dst=dst, group=group)
        return None
    rec = np.concatenate([[0], np.cumsum([s[0] for s in sizes])]).astype(np.int64)
    shw = np.concatenate([[0], np.cumsum([s[1] for s in sizes])]).astype(np.int64)
    out = {k: np.empty((int(rec[-1]),) + tuple(history[k].shape[1:]), dtype=history[k].dtype) for k in keys}
    for r in range(world):
        lo, hi = int(rec[r]), int(rec[r + 1])
        for k in keys:
            if hi == lo:
                continue
            if r == dst:
                out[k][lo:hi] = history[k]
            else:
                dist.recv(torch.from_numpy(out[k][lo:hi]), src=r, group=group)      # a contiguous row slice: received in place
        if "parent" in out:
            sl = out["parent"][lo:hi]
            sl[sl >= 0] += lo
        if "shower" in out:
            out["shower"][lo:hi] += shw[r]
    out["record_offsets"], out["shower_offsets"] = rec, shw
    return out
