#!/usr/bin/env python3
"""Small-size pass over every kernel family of libqpzk.so, meant to run under compute-sanitizer
(memcheck / racecheck / synccheck / initcheck; SURVEY.md 5 "race detection / sanitizers"):

    compute-sanitizer --tool racecheck python tools/sanitize_subset.py

Sizes are chosen so that each code path runs (multi-pass iNTT, register-resident strided LDE passes, the
cp.async.bulk + mbarrier last pass, warp-cooperative Merkle levels, pp_scan, gate interpreter, FRI fold / PoW / gather,
the single-call prover) while staying within minutes under the sanitizer.  Results are only sanity-checked here
(self-consistency of caps); bit-exactness is the job of tests/."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import qpzk  # noqa: E402
from qpzk import prover  # noqa: E402

P = qpzk.P
rng = np.random.default_rng(7)
ctx = qpzk.Context(0)
rnd = lambda *shape: rng.integers(0, P, size=shape, dtype=np.uint64)

# commits: single pass, two-pass iNTT + strided LDE + TMA last pass, salted, tiny
caps = []
for log_n, C_, salted in [(3, 5, False), (8, 135, False), (13, 20, True), (14, 9, False), (6, 135, True)]:
    cols = rnd(C_, 1 << log_n)
    salt = rnd(4, 1 << (log_n + 3)) if salted else None
    b = qpzk.PolynomialBatch.from_values(ctx, cols, 3, salt, min(4, log_n + 3))
    b2 = qpzk.PolynomialBatch.from_coeffs(ctx, b.polynomials, 3, salt, min(4, log_n + 3))
    assert (b.cap == b2.cap).all()
    rows, sib = b.merkle_tree.open([0, (1 << (log_n + 3)) - 1])
    b.eval_ext([3, 5])
    b.free(); b2.free()
# bare Merkle trees incl. the warp-cooperative small levels
for nl, ll, h in [(1, 5, 0), (64, 8, 4), (4096, 20, 0), (2048, 143, 8)]:
    qpzk.MerkleTree.new(ctx, rnd(nl, ll), h).free()
ctx.poseidon_permute(rnd(1000, 12))
# challenger, FRI (salted oracle, several rounds, PoW, queries)
for log_n, Cs, salted in [(8, [6, 10, 3, 2], False), (12, [5, 9, 4, 8], True)]:
    cols = [rnd(c, 1 << log_n) for c in Cs]
    salts = [rnd(4, 1 << (log_n + 3)) if (salted and k == 1) else None for k in range(4)]
    gb = [qpzk.PolynomialBatch.from_values(ctx, c, 3, s, 4) for c, s in zip(cols, salts)]
    ch = qpzk.Challenger(ctx)
    for g in gb:
        ch.observe_cap(g)
    zeta = ch.get_extension_challenge()
    inst = qpzk.FriInstanceInfo([([int(zeta[0]), int(zeta[1])], [(o, i) for o in range(4) for i in range(Cs[o])]),
                                 ([int(zeta[1]), int(zeta[0])], [(2, 0), (2, 1)])])
    qpzk.PolynomialBatch.prove_openings(inst, gb, ch, qpzk.fri_params(log_n, 3, 4, 8, 6, 4, 3))
    for g in gb:
        g.free()
    ch.free()
# permutation argument + quotient with a gate program, device resident
log_n, R, W, K = 9, 19, 24, 3
n = 1 << log_n
ctx.set_keep_values(True)
cs = qpzk.PolynomialBatch.from_values(ctx, rnd(K + R, n), 3, None, 4)
wb = qpzk.PolynomialBatch.from_values(ctx, rnd(W, n), 3, None, 4)
ctx.set_keep_values(False)
ch3 = rnd(3, 2)
zb = qpzk.PolynomialBatch.commit_zs_partial_products(cs, wb, K, R, ch3[0], ch3[1], 4)
from qpzk import gates  # noqa: E402
rec = gates.Recorder()
gates.arithmetic_gate(rec, num_ops=4, selector=rec.const(0))
prog = rec.program()
qb = qpzk.PolynomialBatch.commit_quotient_prog(cs, wb, zb, K, R, ch3[0], ch3[1], ch3[2], prog, 4)
for b in (qb, zb, wb, cs):
    b.free()
# the leaf-shaped synthetic prove at a reduced size (every prove() stage in sequence)
shape = dict(prover.LEAF_SHAPE, log_n=9)
circuit = prover.SyntheticCircuit(ctx, rnd(shape["constants_sigmas"], 1 << 9), shape)
prover.prove_synthetic(circuit, **prover.synthetic_inputs(1, shape))
if hasattr(prover, "sanitize_hook"):
    prover.sanitize_hook(ctx, rnd)
ctx.synchronize()
print("sanitize subset done: %d launches, peak %.1f MB" % (ctx.launch_count(), ctx.peak_bytes() / 1e6))
