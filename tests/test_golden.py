"""Committed known-answer vectors (tests/golden/vectors.json, written by tools/gen_golden.py -- see its docstring for
the provenance: the Poseidon vectors are upstream's, the rest freezes the CPU restatement).
CPU: the oracle reproduces the file.  GPU: the CUDA path, through the C ABI, reproduces the same values."""
import json
import os

import numpy as np
import pytest

import gen_golden as gg

V = json.load(open(gg.OUT))


def test_oracle_reproduces_the_committed_vectors(orc):
    fresh = json.loads(json.dumps(gg.run_oracle(), sort_keys=True))
    assert sorted(fresh) == sorted(V)
    for name in V:
        assert fresh[name] == V[name], name


def test_poseidon_kats_are_upstreams(orc):
    assert len(V["poseidon_kats"]) == 3
    assert V["poseidon_kats"][0]["input"] == [0] * 12 and V["poseidon_kats"][0]["output"][0] == 0x3c18a9786cb0b359
    for k in V["poseidon_kats"]:
        assert [int(x) for x in orc.poseidon(np.array(k["input"], np.uint64))] == k["output"]


# ---- the CUDA path against the same file ---------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_poseidon_kats(ctx):
    st = np.array([k["input"] for k in V["poseidon_kats"]], np.uint64)
    assert ctx.poseidon_permute(st).tolist() == [k["output"] for k in V["poseidon_kats"]]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["commit_2p5x7", "commit_2p8x135"])
def test_gpu_commit_vectors(ctx, qpzk, name):
    c = V[name]
    b = qpzk.PolynomialBatch.from_values(ctx, gg.inputs(c["seed"], c["cols"], 1 << c["log_n"]), c["rate_bits"], None, c["cap_height"])
    assert [int(x) for x in b.cap.reshape(-1)] == c["cap"]
    assert gg.sha(b.polynomials) == c["coeffs_sha256"]
    assert [int(x) for x in b.merkle_tree.get(0)] == c["leaf0"]
    assert [int(x) for x in b.merkle_tree.get((8 << c["log_n"]) - 1)] == c["leaf_last"]


@pytest.mark.gpu
def test_gpu_salted_commit_vector(ctx, qpzk):
    c = V["commit_salted_2p4x3"]
    salt = gg.inputs(c["seed"] + 1000, 4, 8 << c["log_n"])
    b = qpzk.PolynomialBatch.from_values(ctx, gg.inputs(c["seed"], c["cols"], 1 << c["log_n"]), c["rate_bits"], salt, c["cap_height"])
    assert [int(x) for x in b.cap.reshape(-1)] == c["cap"]
    assert [int(x) for x in b.merkle_tree.get(5)] == c["leaf5"]


@pytest.mark.gpu
def test_gpu_permutation_argument_vectors(ctx, qpzk):
    from perm_util import satisfied_permutation
    c = V["perm_argument_2p6"]
    n = 1 << c["log_n"]
    routed, sigmas, _ = satisfied_permutation(c["log_n"], c["routed"], seed=c["seed"])
    wires = np.concatenate([routed, gg.inputs(c["seed"] + 1, c["extra_wires"], n)])
    cs = np.concatenate([gg.inputs(c["seed"] + 2, c["consts"], n), sigmas])
    ch = gg.inputs(c["seed"] + 3, 3, 2)
    zs = ctx.partial_products(routed, sigmas, ch[0], ch[1], c["chunk"])
    assert gg.sha(zs) == c["partial_products_sha256"] and [int(zs[0, 1]), int(zs[1, 1])] == c["z_row1"]
    bs = [qpzk.PolynomialBatch.from_values(ctx, a, 3, None, 2) for a in (cs, wires, zs)]
    q = qpzk.PolynomialBatch.quotient_polys(bs[0], bs[1], bs[2], c["consts"], c["routed"], ch[0], ch[1], ch[2], c["chunk"])
    assert gg.sha(q) == c["quotient_sha256"] and [int(x) for x in q[:, 0]] == c["quotient_first"]


@pytest.mark.gpu
def test_gpu_fri_vector(ctx, qpzk):
    c = V["fri_2p6"]
    cols = [gg.inputs(c["seed"] + 10 * i, k, 1 << c["log_n"]) for i, k in enumerate(c["oracle_cols"])]
    gb = [qpzk.PolynomialBatch.from_values(ctx, a, 3, None, 4) for a in cols]
    ch = qpzk.Challenger(ctx)
    for b in gb:
        ch.observe_cap(b)
    zeta = ch.get_extension_challenge()
    assert [int(zeta[0]), int(zeta[1])] == c["zeta"]
    batches = [([int(zeta[0]), int(zeta[1])], [(o, i) for o in range(4) for i in range(c["oracle_cols"][o])])]
    op = np.concatenate([b.eval_ext(zeta) for b in gb])
    assert gg.sha(op) == c["openings_sha256"]
    ch.observe_extension_elements(op)
    par = qpzk.fri_params(c["log_n"], 3, 4, c["pow_bits"], c["queries"], 4, c["final_poly_bits"])
    proof = qpzk.PolynomialBatch.prove_openings(qpzk.FriInstanceInfo(batches), gb, ch, par)
    assert proof.size == c["proof_words"] and int(proof[-1]) == c["pow_witness"]
    assert gg.sha(proof) == c["proof_sha256"]
