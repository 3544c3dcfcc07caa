"""Host-side sharding logic for one 8xB200 box.  Pure bookkeeping + torch.distributed plumbing:
the data-path collectives are one all-gather of coefficient columns and one all-gather of cap slices."""
import numpy as np


def column_range(num_cols, rank, world):
    """Columns whose iNTT rank `rank` computes; every rank gets q = ceil(C / world) slots, the tail is padding."""
    q = (num_cols + world - 1) // world
    lo = min(rank * q, num_cols)
    return lo, min(lo + q, num_cols), q


def coset_range(rate_bits, rank, world):
    """Leaf blocks (cosets, in leaf order) rank `rank` extends and hashes: (begin, count)."""
    cosets = 1 << rate_bits
    if world > cosets or cosets % world:
        raise ValueError("world size must divide 2^rate_bits")
    per = cosets // world
    return rank * per, per


def cap_slice(cap_height, rate_bits, rank, world):
    """Indices of the full Merkle cap owned by rank `rank` (requires cap_height >= log2(world))."""
    begin, count = coset_range(rate_bits, rank, world)
    per_coset = (1 << cap_height) >> rate_bits
    if per_coset == 0:
        raise ValueError("cap_height < rate_bits: a cap entry would span several cosets")
    return begin * per_coset, (begin + count) * per_coset


def shard_range(num_items, rank, world):
    """Independent proofs (replicas): contiguous, balanced split of [0, num_items)."""
    base, extra = divmod(num_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_cap(local_cap, group=None):
    """All-gather the per-rank cap slices (numpy (k,4) uint64) into the full cap.  Works on any backend."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(local_cap).view(np.int64).reshape(-1))
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = t.to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return np.concatenate([o.cpu().numpy().view(np.uint64).reshape(-1, 4) for o in out])


def gather_columns(local_cols, num_cols, group=None):
    """All-gather column slices (numpy (q, n) uint64, zero padded) into the first num_cols columns."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(local_cols).view(np.int64))
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = t.to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return np.concatenate([o.cpu().numpy().view(np.uint64) for o in out])[:num_cols]
