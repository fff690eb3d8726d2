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
