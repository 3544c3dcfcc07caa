"""One commitment over the GPUs of an 8xB200 box, and independent proofs as replicas.

`CosetSplitCommitter` is the multi-GPU form of `PolynomialBatch::from_values` (one process per GPU, torch.distributed for
the plumbing): the iNTT is split by columns, the coefficient exchange is either fused into the first LDE pass (every
rank reads its peers' iNTT outputs over NVLink through CUDA IPC mappings -- `exchange="p2p"`) or one NCCL all-gather
(`"nccl"`), every rank extends and hashes 2^rate_bits / world cosets = whole Merkle subtrees, and the cap slices are
all-gathered (64 B per rank).  bench.py --gpus N drives it; the helpers below it are the index bookkeeping (tested on CPU
with gloo) and the replica split of configs 4/5."""
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


class CosetSplitCommitter:
    """One `PolynomialBatch::from_values(values, rate_bits, false, cap_height)` over `world` ranks (1, 2, 4 or 8).

        cs = CosetSplitCommitter(ctx, log_n, num_cols, rate_bits, cap_height, rank, world, device, exchange="p2p")
        cs.intt(values_ptr)              # this rank's column slice (cs.q columns, zero padded), values -> coefficients
        batch = cs.commit()              # this rank's cosets: LDE + leaf hashes + subtrees (a qpzk.PolynomialBatch slice)
        cap = cs.gather_cap(batch)       # (world, 2^cap_height / world * 4) int64 tensor = the full cap, in order

    Everything is enqueued on the context's stream; `gather_cap` reads this rank's cap slice (a tiny D2H) and so
    synchronises.  The iNTT output buffer is reused by the next commitment: `gather_cap` is the barrier that keeps a fast
    rank from overwriting its slice while a slow one still reads it."""

    def __init__(self, ctx, log_n, num_cols, rate_bits, cap_height, rank, world, device, exchange="p2p"):
        import torch
        import torch.distributed as dist
        import qpzk
        self.ctx, self.log_n, self.num_cols, self.rate_bits, self.cap_height = ctx, log_n, num_cols, rate_bits, cap_height
        self.rank, self.world, self.device = rank, world, device
        self.n = 1 << log_n
        self.c0, self.c1, self.q = column_range(num_cols, rank, world)
        self.coset_begin, self.cosets_per = coset_range(rate_bits, rank, world)
        self.p2p = world > 1 and exchange == "p2p"
        self.peers = self.peer_ptrs = self.tick = self.coeff_mine = self.coeff_all = self.cap_all = None
        if world > 1:
            if (1 << cap_height) % world:
                raise ValueError("cap_height too small for this many ranks")
            self.cap_all = torch.empty((world, (1 << cap_height) // world * 4), dtype=torch.int64, device=device)
            if self.p2p:
                # this rank's iNTT output lives in an IPC-exported buffer; every peer's buffer is mapped into this process
                self.peers = qpzk.PeerBuffers(ctx, self.q * self.n)
                handles = [None] * world
                dist.all_gather_object(handles, self.peers.handle)
                self.peer_ptrs = self.peers.exchange(handles, rank)
                self.tick = torch.zeros(1, device=device)
            else:
                self.coeff_mine = torch.empty((self.q, self.n), dtype=torch.int64, device=device)
                self.coeff_all = torch.empty((world * self.q, self.n), dtype=torch.int64, device=device)

    @property
    def coeff_ptr(self):
        """Device address of this rank's coefficient slice (q x n words)."""
        return self.peers.ptr if self.p2p else self.coeff_mine.data_ptr()

    def intt(self, values_ptr, first_col=0, cols=None):
        """values -> coefficients for columns [first_col, first_col + cols) of this rank's slice (default: all q)."""
        import qpzk
        cols = self.q - first_col if cols is None else cols
        if cols:
            qpzk._check(self.ctx.L.qpzk_intt_columns_dev(self.ctx.h, values_ptr, self.coeff_ptr + first_col * self.n * 8, self.log_n, cols))

    def commit(self):
        import torch.distributed as dist
        import qpzk
        if self.p2p:
            # every rank's iNTT has finished when this 1-element all-reduce has (stream ordered)
            dist.all_reduce(self.tick)
            return qpzk.PolynomialBatch.cosets_from_coeff_slices_device(self.ctx, self.peer_ptrs, self.q, self.log_n, self.num_cols,
                                                                        self.rate_bits, self.cap_height, self.coset_begin, self.cosets_per)
        dist.all_gather_into_tensor(self.coeff_all, self.coeff_mine)
        return qpzk.PolynomialBatch.cosets_from_coeffs_device(self.ctx, self.coeff_all.data_ptr(), self.log_n, self.num_cols, self.rate_bits,
                                                              self.cap_height, self.coset_begin, self.cosets_per)

    def gather_cap(self, batch):
        import torch
        import torch.distributed as dist
        cap = batch.cap                                   # tiny D2H read of this rank's cap slice
        dist.all_gather_into_tensor(self.cap_all, torch.from_numpy(cap.view(np.int64).reshape(1, -1)).to(self.device))
        return self.cap_all

    def close(self):
        if self.peers is not None:
            self.peers.close()
            self.peers = None
