"""World-size-2 gloo test of the multi-GPU host logic: column split for the iNTT, all-gather of the
coefficient matrix, per-rank coset blocks hashed as whole Merkle subtrees, all-gather of the cap slices.  The oracle
stands in for the kernels (no GPU here); the point is that the sharded result equals the unsharded commitment."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 0xFFFFFFFF00000001


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits_b200"))
    import torch.distributed as dist
    from oracle import oracle as orc
    from qpzk import dist as qd
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    log_n, C_, rate_bits, h = 6, 7, 3, 4
    n = 1 << log_n
    cols = np.random.default_rng(3).integers(0, P, size=(C_, n), dtype=np.uint64)  # same on every rank
    full = orc.PolynomialBatch.from_values(cols, rate_bits, h)
    # 1. iNTT of this rank's columns, zero padded to q slots
    lo, hi, q = qd.column_range(C_, rank, world)
    mine = np.zeros((q, n), np.uint64)
    for k, c in enumerate(range(lo, hi)):
        mine[k] = orc.fft(cols[c], inverse=True)
    # 2. all-gather the coefficient matrix
    coeffs = qd.gather_columns(mine, C_)
    assert (coeffs == full.coeffs()).all()
    # 3. this rank's cosets = contiguous leaf blocks = whole subtrees
    begin, count = qd.coset_range(rate_bits, rank, world)
    leaves = full.leaves()[begin * n:(begin + count) * n]
    sub = orc.MerkleTree(leaves, h - (rate_bits - (count.bit_length() - 1)))
    a, b = qd.cap_slice(h, rate_bits, rank, world)
    assert (sub.cap() == full.cap()[a:b]).all()
    # 4. all-gather the cap slices
    cap = qd.gather_cap(sub.cap())
    ok = (cap == full.cap()).all()
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_coset_split_world2_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29517, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_sharding_bookkeeping():
    sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits_b200"))
    from qpzk import dist as qd
    for world in (1, 2, 4, 8):
        cols = [qd.column_range(135, r, world) for r in range(world)]
        assert cols[0][0] == 0 and cols[-1][1] == 135 and all(c[2] == cols[0][2] for c in cols)
        assert all(cols[i][1] == cols[i + 1][0] for i in range(world - 1))
        cos = [qd.coset_range(3, r, world) for r in range(world)]
        assert sum(c[1] for c in cos) == 8
        caps = [qd.cap_slice(4, 3, r, world) for r in range(world)]
        assert caps[0][0] == 0 and caps[-1][1] == 16
    assert [qd.shard_range(1024, r, 8) for r in range(8)][-1] == (896, 1024)
    assert sum(b - a for a, b in (qd.shard_range(10, r, 4) for r in range(4))) == 10
