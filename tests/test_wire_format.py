"""Wire format: qpzk_proof_to_bytes / qpzk_proof_from_bytes are host functions of the C ABI, so they
are tested without a GPU -- against an independent field-by-field serializer over the STRUCTURED proof (oracle/oracle.py:
parse_fri_proof + proof_to_bytes), on a real proof produced by the CPU restatement, plus the failure modes upstream's
reader has (wrong length, wrong sibling count, non-canonical element)."""
import numpy as np
import pytest

from conftest import P


def _oracle_proof(orc, log_n=5, R=12, num_consts=2, num_wires=15, npi=3):
    from perm_util import satisfied_permutation
    n = 1 << log_n
    rng = np.random.default_rng(55)
    routed, sigmas, _ = satisfied_permutation(log_n, R, seed=6)
    wires = np.concatenate([routed, rng.integers(0, P, size=(num_wires - R, n), dtype=np.uint64)])
    cs = np.concatenate([rng.integers(0, P, size=(num_consts, n), dtype=np.uint64), sigmas])
    ocs = orc.PolynomialBatch.from_values(cs, 3, 4)
    pih = rng.integers(0, P, size=4, dtype=np.uint64)
    out = orc.prove_synthetic(ocs, wires, None, None, pih, log_n, sigmas=sigmas, sigma_col0=num_consts, pow_bits=6,
                              num_queries=5, final_poly_bits=2)
    out["public_inputs"] = rng.integers(0, P, size=npi, dtype=np.uint64)
    nchunks = (R + 7) // 8
    out["counts"] = dict(log_n=log_n, num_constants=num_consts, num_sigmas=R, num_wires=num_wires, num_zs=2,
                         num_partial_products=2 * (nchunks - 1), num_quotient_polys=16, num_public_inputs=npi)
    return out


def _structured_bytes(orc, pr, params):
    c = pr["counts"]
    op = [(int(a), int(b)) for a, b in pr["openings"]]
    names = [("constants", c["num_constants"]), ("plonk_sigmas", c["num_sigmas"]), ("wires", c["num_wires"]),
             ("plonk_zs", c["num_zs"]), ("partial_products", c["num_partial_products"]),
             ("quotient_polys", c["num_quotient_polys"]), ("plonk_zs_next", c["num_zs"])]   # FRI-instance order
    os_, o = {}, 0
    for name, k in names:
        os_[name] = op[o:o + k]
        o += k
    assert o == len(op)
    fri = orc.parse_fri_proof(pr["fri_proof"], params, c["log_n"], pr["leaf_lens"])
    caps3 = [[[int(v) for v in h] for h in cap.reshape(-1, 4)] for cap in pr["caps"]]
    return orc.proof_to_bytes(caps3, os_, fri, [int(v) for v in pr["public_inputs"]])


def test_to_bytes_matches_the_structured_serializer_and_round_trips(qpzk, orc):
    pr = _oracle_proof(orc)
    c = pr["counts"]
    gp = qpzk.fri_params(c["log_n"], 3, 4, 6, 5, 4, 2)
    op = orc.fri_params(c["log_n"], 3, 4, 6, 5, 4, 2)
    shape = qpzk.proof_shape(**c)
    assert list(shape.leaf_lens) == list(pr["leaf_lens"])
    got = qpzk.proof_to_bytes(gp, shape, pr["caps"][0], pr["caps"][1], pr["caps"][2], pr["openings"], pr["fri_proof"],
                              pr["public_inputs"])
    want = _structured_bytes(orc, pr, op)
    assert len(got) == len(want) and got == want
    back = qpzk.proof_from_bytes(gp, shape, got)
    assert (back["wires_cap"].reshape(-1) == pr["caps"][0].reshape(-1)).all()
    assert (back["zs_cap"].reshape(-1) == pr["caps"][1].reshape(-1)).all()
    assert (back["quotient_cap"].reshape(-1) == pr["caps"][2].reshape(-1)).all()
    assert (back["openings"] == pr["openings"]).all()
    assert (back["fri_proof"] == pr["fri_proof"]).all()
    assert (back["public_inputs"] == pr["public_inputs"]).all()
    # the parsed proof still verifies (FRI) -- nothing was reordered on the way
    assert orc.fri_verify(c["log_n"], pr["leaf_lens"], pr["all_caps"], pr["instance"], back["openings"], pr["params"],
                          pr["verifier_challenger"], back["fri_proof"]) == 0


def test_reference_leaf_shape_length(qpzk):
    """Length of a leaf-circuit proof with the reference's FRI parameters (wormhole/circuit/src/profile.rs:168-174):
    3 caps + 276 openings + FRI proof (2 commit-phase caps, 28 query rounds, final poly of 32, pow witness) + inputs."""
    gp = qpzk.fri_params(13)
    shape = qpzk.proof_shape(13, 5, 80, 135, 2, 18, 16, num_public_inputs=21)
    n = qpzk.load_library().qpzk_proof_bytes_len(gp, shape)
    caps = 3 * 16 * 32
    openings = 16 * (5 + 80 + 135 + 2 + 2 + 18 + 16)
    per_query = sum(8 * ll + 1 + 32 * (16 - 4) for ll in (85, 135, 20, 16)) + (16 * 16 + 1 + 32 * (12 - 4)) + (16 * 16 + 1 + 32 * (8 - 4))
    fri = 2 * 16 * 32 + 28 * per_query + 16 * 32 + 8
    assert n == caps + openings + fri + 8 + 8 * 21


def test_from_bytes_rejects_malformed_proofs(qpzk, orc):
    pr = _oracle_proof(orc, npi=1)
    c = pr["counts"]
    gp = qpzk.fri_params(c["log_n"], 3, 4, 6, 5, 4, 2)
    shape = qpzk.proof_shape(**c)
    data = bytearray(qpzk.proof_to_bytes(gp, shape, pr["caps"][0], pr["caps"][1], pr["caps"][2], pr["openings"],
                                         pr["fri_proof"], pr["public_inputs"]))
    with pytest.raises(qpzk.QpzkError):
        qpzk.proof_from_bytes(gp, shape, data[:-1])                       # truncated
    with pytest.raises(qpzk.QpzkError):
        qpzk.proof_from_bytes(gp, shape, data + b"\\0")                    # trailing byte
    bad = bytearray(data)
    bad[0:8] = (P).to_bytes(8, "little")                                  # first cap word = p: not canonical
    with pytest.raises(qpzk.QpzkError):
        qpzk.proof_from_bytes(gp, shape, bad)
    # the first Merkle proof's sibling count sits right after the caps, the openings and the commit-phase caps + leaf
    H = 16 * 32
    k = c["num_constants"] + c["num_sigmas"] + c["num_wires"] + 2 * c["num_zs"] + c["num_partial_products"] + c["num_quotient_polys"]
    off = 3 * H + 16 * k + gp.num_rounds * H + 8 * shape.leaf_lens[0]
    assert data[off] == c["log_n"] + 3 - 4
    bad = bytearray(data)
    bad[off] += 1
    with pytest.raises(qpzk.QpzkError):
        qpzk.proof_from_bytes(gp, shape, bad)
    bad = bytearray(data)
    bad[-16:-8] = (7).to_bytes(8, "little")                               # public input count != circuit's
    with pytest.raises(qpzk.QpzkError):
        qpzk.proof_from_bytes(gp, shape, bad)
    with pytest.raises(qpzk.QpzkError):                                   # pieces of the wrong size
        qpzk.proof_to_bytes(gp, shape, pr["caps"][0], pr["caps"][1], pr["caps"][2], pr["openings"][:-1], pr["fri_proof"],
                            pr["public_inputs"])
    bad_open = pr["openings"].copy()
    bad_open[0, 0] = np.uint64(P)
    with pytest.raises(qpzk.QpzkError):                                   # non-canonical element on the way out
        qpzk.proof_to_bytes(gp, shape, pr["caps"][0], pr["caps"][1], pr["caps"][2], bad_open, pr["fri_proof"], pr["public_inputs"])
