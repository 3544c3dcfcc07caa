#!/usr/bin/env python3
"""Writes tests/golden/vectors.json: small known-answer vectors for every entry point on the path.

    python tools/gen_golden.py            (re)generate
    python tools/gen_golden.py --check    fail if the committed file differs from a fresh run

Provenance.  `poseidon_kats` are the three upstream plonky2 permutation vectors: they pin the oracle.
Everything else is produced BY the oracle (oracle/, the CPU restatement) on seeded inputs, because the reference tree
holds no golden vector for the path and the reference cannot run here: these vectors freeze the
restatement (a later edit of the oracle that changes any of them is caught by tests/test_golden.py without a GPU) and
give the CUDA path fixed values to reproduce on the GPU box, where the oracle is rebuilt from source.  Large outputs
are stored as SHA-256 of their little-endian words.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden", "vectors.json")
P = 0xFFFFFFFF00000001


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def inputs(seed, rows, n):
    """The seeded inputs every vector is computed from (kept out of the file: regenerated identically by the tests)."""
    return np.random.default_rng(seed).integers(0, P, size=(rows, n), dtype=np.uint64)


def cases():
    """name -> (description of the inputs, callable(backend) -> dict of results); backend = the oracle module or an
    adapter with the same calls over the CUDA library (tests/test_golden.py)."""
    return {
        "commit_2p5x7": dict(seed=101, log_n=5, cols=7, rate_bits=3, cap_height=2),
        "commit_2p8x135": dict(seed=102, log_n=8, cols=135, rate_bits=3, cap_height=4),
        "commit_salted_2p4x3": dict(seed=103, log_n=4, cols=3, rate_bits=3, cap_height=1, salted=True),
        "perm_argument_2p6": dict(seed=104, log_n=6, routed=19, extra_wires=2, consts=2, chunk=8),
        "fri_2p6": dict(seed=105, log_n=6, oracle_cols=[4, 9, 3, 2], pow_bits=6, queries=5, final_poly_bits=2),
    }


def run_oracle():
    from oracle import oracle as orc
    import gen_poseidon_constants as gen
    from perm_util import satisfied_permutation
    orc.build()
    v = {"poseidon_kats": [{"input": [int(x) for x in inp], "output": [int(w, 16) for w in want.split()]} for inp, want in gen.KATS]}
    C = cases()
    for name in ("commit_2p5x7", "commit_2p8x135"):
        c = C[name]
        b = orc.PolynomialBatch.from_values(inputs(c["seed"], c["cols"], 1 << c["log_n"]), c["rate_bits"], c["cap_height"])
        v[name] = dict(c, cap=[int(x) for x in b.cap().reshape(-1)], coeffs_sha256=sha(b.coeffs()),
                       leaf0=[int(x) for x in b.merkle_tree.leaf(0)], leaf_last=[int(x) for x in b.merkle_tree.leaf((8 << c["log_n"]) - 1)])
    c = C["commit_salted_2p4x3"]
    salt = inputs(c["seed"] + 1000, 4, 8 << c["log_n"])
    b = orc.PolynomialBatch.from_values(inputs(c["seed"], c["cols"], 1 << c["log_n"]), c["rate_bits"], c["cap_height"], salt)
    v["commit_salted_2p4x3"] = dict(c, cap=[int(x) for x in b.cap().reshape(-1)], leaf5=[int(x) for x in b.merkle_tree.leaf(5)])
    c = C["perm_argument_2p6"]
    n = 1 << c["log_n"]
    routed, sigmas, _ = satisfied_permutation(c["log_n"], c["routed"], seed=c["seed"])
    wires = np.concatenate([routed, inputs(c["seed"] + 1, c["extra_wires"], n)])
    cs = np.concatenate([inputs(c["seed"] + 2, c["consts"], n), sigmas])
    ch = inputs(c["seed"] + 3, 3, 2)
    zs = orc.partial_products(routed, sigmas, ch[0], ch[1], c["chunk"])
    bs = [orc.PolynomialBatch.from_values(a, 3, 2) for a in (cs, wires, zs)]
    q = orc.quotient_polys(bs[0], bs[1], bs[2], c["consts"], c["routed"], ch[0], ch[1], ch[2], c["chunk"])
    v["perm_argument_2p6"] = dict(c, partial_products_sha256=sha(zs), z_row1=[int(zs[0, 1]), int(zs[1, 1])],
                                  quotient_sha256=sha(q), quotient_first=[int(x) for x in q[:, 0]])
    c = C["fri_2p6"]
    cols = [inputs(c["seed"] + 10 * i, k, 1 << c["log_n"]) for i, k in enumerate(c["oracle_cols"])]
    ob = [orc.PolynomialBatch.from_values(a, 3, 4) for a in cols]
    och = orc.Challenger()
    for o in ob:
        och.observe(o.cap())
    zeta = och.get(2)
    batches = [([int(zeta[0]), int(zeta[1])], [(o, i) for o in range(4) for i in range(c["oracle_cols"][o])])]
    op = np.concatenate([b.eval_ext(zeta) for b in ob])
    och.observe(op)
    par = orc.fri_params(c["log_n"], 3, 4, c["pow_bits"], c["queries"], 4, c["final_poly_bits"])
    proof = orc.fri_prove(ob, orc.Instance(batches), par, och)
    v["fri_2p6"] = dict(c, zeta=[int(zeta[0]), int(zeta[1])], openings_sha256=sha(op), proof_words=int(proof.size),
                        proof_sha256=sha(proof), pow_witness=int(proof[-1]))
    return v


if __name__ == "__main__":
    fresh = json.dumps(run_oracle(), indent=1, sort_keys=True) + "\n"
    if "--check" in sys.argv:
        sys.exit(0 if open(OUT).read() == fresh else 1)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    open(OUT, "w").write(fresh)
    print("wrote", OUT, len(fresh), "bytes")
