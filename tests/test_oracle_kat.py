"""CPU tests: pin the oracle against every known-answer vector available for this path.

The reference tree holds no golden LDE / Merkle / FRI vectors; what exists is pinned here:
the three upstream Poseidon-12 permutation KATs, the round-constant self-checks, the Goldilocks generator
facts and the reference-side constants that fix the configuration
(p at wormhole/inputs/src/lib.rs:23; FRI parameters at wormhole/circuit/src/profile.rs:168-174).
"""
import os

import numpy as np
import pytest

import gen_poseidon_constants as gen

P = 0xFFFFFFFF00000001


def test_round_constants_self_checks():
    rc = gen.round_constants()
    gen.self_check(rc)
    assert len(rc) == 360 and all(c < P for c in rc)


def test_permutation_kats_python_twin():
    rc = gen.round_constants()
    fast = gen.derive_fast_partial(rc)
    for inp, want in gen.KATS:
        want = [int(w, 16) for w in want.split()]
        assert gen.poseidon_naive(inp, rc) == want
        assert gen.poseidon_fast(inp, rc, fast) == want


def test_permutation_kats_c_oracle(orc):
    for inp, want in gen.KATS:
        want = [int(w, 16) for w in want.split()]
        assert [int(x) for x in orc.poseidon(inp)] == want
        assert [int(x) for x in orc.poseidon(inp, naive=True)] == want


def test_naive_equals_sparse_random(orc):
    rng = np.random.default_rng(11)
    for _ in range(200):
        x = rng.integers(0, P, size=12, dtype=np.uint64)
        assert (orc.poseidon(x) == orc.poseidon(x, naive=True)).all()


def test_goldilocks_generators(orc):
    L = orc.lib()
    g, w = 14293326489335486720, 7277203076849721926
    assert L.orc_gl_exp(g, (P - 1) >> 32) == w
    assert L.orc_gl_exp(w, 1 << 31) == P - 1 and L.orc_gl_exp(w, 1 << 32) == 1
    for q in (2, 3, 5, 17, 257, 65537):  # g is a generator: non-trivial at every prime factor of p-1
        assert L.orc_gl_exp(g, (P - 1) // q) != 1
    assert L.orc_gl_exp(7, (P - 1) // 2) == P - 1  # 7 is a non-residue => X^2 - 7 irreducible
    for k in range(0, 33):
        assert L.orc_gl_root_of_unity(k) == pow(w, 1 << (32 - k), P)


def test_field_mul_against_bigint(orc):
    rng = np.random.default_rng(5)
    L = orc.lib()
    edge = [0, 1, 2, P - 1, P - 2, 0xFFFFFFFF, 0x100000000, 0xFFFFFFFF00000000, 1 << 63]
    vals = edge + [int(x) for x in rng.integers(0, P, size=64, dtype=np.uint64)]
    for a in vals:
        for b in vals[:24]:
            assert L.orc_gl_mul(a, b) == a * b % P
    for a in vals[1:20]:
        assert L.orc_gl_mul(a, L.orc_gl_inv(a)) == 1


def test_fft_against_definition(orc):
    rng = np.random.default_rng(2)
    for log_n in (0, 1, 3, 5):
        n = 1 << log_n
        a = rng.integers(0, P, size=n, dtype=np.uint64)
        w = orc.lib().orc_gl_root_of_unity(log_n)
        want = [sum(int(a[j]) * pow(w, i * j, P) for j in range(n)) % P for i in range(n)]
        assert [int(x) for x in orc.fft(a)] == want
        assert (orc.fft(orc.fft(a), inverse=True) == a).all()
        shift = orc.GENERATOR
        want = [sum(int(a[j]) * pow(shift * pow(w, i, P) % P, j, P) for j in range(n)) % P for i in range(n)]
        assert [int(x) for x in orc.coset_fft(a)] == want


def test_hashing_rules(orc):
    # hash_or_noop: <= 4 elements are copied and zero padded; longer inputs go through the sponge
    assert list(orc.hash_or_noop([5, 6])) == [5, 6, 0, 0]
    assert list(orc.hash_or_noop([1, 2, 3, 4])) == [1, 2, 3, 4]
    x = np.arange(1, 6, dtype=np.uint64)
    st = np.zeros(12, np.uint64); st[:5] = x
    assert (orc.hash_or_noop(x) == orc.poseidon(st)[:4]).all()
    # overwrite mode: second chunk overwrites state[0..len) only, no padding
    x = np.arange(1, 12, dtype=np.uint64)
    st = np.zeros(12, np.uint64); st[:8] = x[:8]
    st = orc.poseidon(st); st[:3] = x[8:]
    assert (orc.hash_no_pad(x) == orc.poseidon(st)[:4]).all()
    l, r = np.arange(4, dtype=np.uint64), np.arange(4, 8, dtype=np.uint64)
    st = np.zeros(12, np.uint64); st[:4] = l; st[4:8] = r
    assert (orc.two_to_one(l, r) == orc.poseidon(st)[:4]).all()


def test_merkle_tree(orc):
    rng = np.random.default_rng(3)
    for nl, ll, h in [(1, 5, 0), (8, 3, 0), (8, 4, 3), (16, 7, 2), (64, 135, 4)]:
        leaves = rng.integers(0, P, size=(nl, ll), dtype=np.uint64)
        t = orc.MerkleTree(leaves, h)
        cap = t.cap()
        assert cap.shape == (1 << h, 4)
        if h == nl.bit_length() - 1:  # cap == leaf digests
            assert all((cap[i] == orc.hash_or_noop(leaves[i])).all() for i in range(nl))
        for i in range(nl):
            assert orc.merkle_verify(leaves[i], i, cap, t.prove(i))
        if nl > 1:
            assert not orc.merkle_verify(leaves[0], 1, cap, t.prove(0))
    with pytest.raises(ValueError):
        orc.MerkleTree(rng.integers(0, P, size=(8, 3), dtype=np.uint64), 4)  # cap_height > log2(leaves)


def test_challenger_duplex(orc):
    c = orc.Challenger()
    c.observe([1, 2, 3])
    st = np.zeros(12, np.uint64); st[:3] = [1, 2, 3]
    out = orc.poseidon(st)
    got = c.get(3)
    assert list(got) == [int(out[7]), int(out[6]), int(out[5])]  # popped from the back of state[0..8]
    c.observe([9])  # observing clears the output buffer
    st = out.copy(); st[0] = 9
    assert int(c.get(1)[0]) == int(orc.poseidon(st)[7])
    c2 = orc.Challenger(); c2.observe(np.arange(8))  # a full rate block triggers the duplex immediately
    st = np.zeros(12, np.uint64); st[:8] = np.arange(8)
    assert list(c2.export()[0]) == [int(x) for x in orc.poseidon(st)]


def test_fri_reduction_strategy(orc):
    # ConstantArityBits(4, 5) with rate_bits 3, cap_height 4 (reference wormhole/circuit/src/profile.rs:168-174)
    for d, want in [(13, [4, 4]), (15, [4, 4, 4]), (16, [4, 4, 4]), (5, []), (12, [4, 4]), (20, [4, 4, 4, 4])]:
        p = orc.fri_params(d)
        assert list(p.arity_bits)[:p.num_rounds] == want


def _instance(orc, bs, Cs, log_n, ch):
    zeta = ch.get(2)
    wn = orc.lib().orc_gl_root_of_unity(log_n)
    gz = [orc.lib().orc_gl_mul(int(zeta[0]), wn), orc.lib().orc_gl_mul(int(zeta[1]), wn)]
    allp = [(o, i) for o in range(len(Cs)) for i in range(Cs[o])]
    nxt = [(2, 0), (2, 1)]
    batches = [([int(zeta[0]), int(zeta[1])], allp), (gz, nxt)]
    op = np.concatenate([np.concatenate([bs[o].eval_ext(zeta) for o in range(len(Cs))]),
                         np.stack([bs[o].eval_ext(gz)[i] for o, i in nxt])])
    return orc.Instance(batches), op


@pytest.mark.parametrize("log_n,pow_bits,final_bits,salted", [(6, 8, 2, False), (5, 4, 5, False), (8, 6, 3, True)])
def test_fri_prove_verify_roundtrip(orc, log_n, pow_bits, final_bits, salted):
    rng = np.random.default_rng(log_n)
    Cs = [5, 9, 3, 2]
    n = 1 << log_n
    salt = lambda: rng.integers(0, P, size=(4, n << 3), dtype=np.uint64) if salted else None
    bs = [orc.PolynomialBatch.from_values(rng.integers(0, P, size=(c, n), dtype=np.uint64), 3, 4, salt()) for c in Cs]
    ch = orc.Challenger()
    for b in bs:
        ch.observe(b.cap())
    inst, op = _instance(orc, bs, Cs, log_n, ch)
    ch.observe(op)
    params = orc.fri_params(log_n, pow_bits=pow_bits, num_queries=6, final_poly_bits=final_bits)
    chv = ch.clone()
    proof = orc.fri_prove(bs, inst, params, ch)
    lens, caps = [b.leaf_len for b in bs], [b.cap() for b in bs]
    assert orc.fri_verify(log_n, lens, caps, inst, op, params, chv, proof) == 0
    assert chv.export() == ch.export()  # prover and verifier leave the transcript in the same state
    # minimum-witness policy: no smaller witness satisfies the difficulty
    for pos in (0, proof.size // 2, proof.size - 2, proof.size - 1):
        bad = proof.copy(); bad[pos] ^= np.uint64(1)
        c3 = orc.Challenger()
        for b in bs:
            c3.observe(b.cap())
        c3.get(2); c3.observe(op)
        assert orc.fri_verify(log_n, lens, caps, inst, op, params, c3, bad) != 0
    op_bad = op.copy(); op_bad[3, 0] ^= np.uint64(1)
    c4 = orc.Challenger()
    for b in bs:
        c4.observe(b.cap())
    c4.get(2); c4.observe(op)
    assert orc.fri_verify(log_n, lens, caps, inst, op_bad, params, c4, proof) != 0


def test_lde_is_polynomial_evaluation(orc):
    rng = np.random.default_rng(9)
    log_n, C_ = 4, 3
    n, L = 1 << log_n, 1 << (log_n + 3)
    cols = rng.integers(0, P, size=(C_, n), dtype=np.uint64)
    b = orc.PolynomialBatch.from_values(cols, 3, 2)
    co, lv = b.coeffs(), b.leaves()
    wn, wl = orc.lib().orc_gl_root_of_unity(log_n), orc.lib().orc_gl_root_of_unity(log_n + 3)
    br = lambda x, bits: int(format(x, "0%db" % bits)[::-1], 2)
    for c in range(C_):
        for i in range(n):  # coefficients interpolate the values on the subgroup
            assert sum(int(co[c][k]) * pow(wn, i * k, P) for k in range(n)) % P == int(cols[c][i])
        for j in range(L):  # leaf row j holds P(g * w_L^bitrev(j))
            x = orc.GENERATOR * pow(wl, br(j, log_n + 3), P) % P
            assert sum(int(co[c][k]) * pow(x, k, P) for k in range(n)) % P == int(lv[j][c])


def test_oracle_prove_synthetic_is_accepted_by_the_oracle_verifier(orc):
    """The CPU mirror of the commitment/opening half of prove() (used as the checker of qpzk.prover and as the CPU
    baseline of bench.py's prove line) produces proofs its own verifier restatement accepts, and rejects tampering."""
    log_n = 7
    rng = np.random.default_rng(3)
    mk = lambda c: rng.integers(0, orc.P, size=(c, 1 << log_n), dtype=np.uint64)
    cs = orc.PolynomialBatch.from_values(mk(9), 3, 4)
    out = orc.prove_synthetic(cs, mk(12), mk(4), mk(3), rng.integers(0, orc.P, size=4, dtype=np.uint64), log_n,
                              pow_bits=6, num_queries=5)
    args = (log_n, out["leaf_lens"], out["all_caps"], out["instance"], out["openings"], out["params"])
    assert orc.fri_verify(*args, out["verifier_challenger"].clone(), out["fri_proof"]) == 0
    bad = out["fri_proof"].copy()
    bad[len(bad) // 2] ^= 1
    assert orc.fri_verify(*args, out["verifier_challenger"].clone(), bad) != 0


# ---- permutation argument ------------------------------------------------------------------
def _partial_products_bigint(wires, sigmas, betas, gammas, chunk):
    """Literal big-integer restatement of upstream `wires_permutation_partial_products_and_zs` + the Z-first reorder
    done in `prove()`: every quotient inverted on its own with pow(d, p-2, p)."""
    from perm_util import P as PP, G, root_of_unity
    R, n = wires.shape
    w = root_of_unity(n.bit_length() - 1)
    nch = (R + chunk - 1) // chunk
    zs, pps = [], []
    for beta, gamma in zip(betas, gammas):
        z_x, x = 1, 1
        z_col, pp_cols = [], [[] for _ in range(nch - 1)]
        for i in range(n):
            q = []
            for j in range(R):
                num = (int(wires[j, i]) + int(beta) * pow(G, j, PP) * x + int(gamma)) % PP
                den = (int(wires[j, i]) + int(beta) * int(sigmas[j, i]) + int(gamma)) % PP
                q.append(num * pow(den, PP - 2, PP) % PP)
            acc = z_x
            row = []
            for k in range(nch):
                for v in q[k * chunk:(k + 1) * chunk]:
                    acc = acc * v % PP
                row.append(acc)
            z_col.append(z_x)
            for k in range(nch - 1):
                pp_cols[k].append(row[k])
            z_x = row[-1]
            x = x * w % PP
        zs.append(z_col)
        pps.extend(pp_cols)
    return np.array(zs + pps, dtype=np.uint64), z_x


@pytest.mark.parametrize("log_n,R,chunk", [(3, 10, 8), (4, 7, 8), (2, 13, 5), (0, 3, 2), (5, 16, 8)])
def test_partial_products_against_bigint(orc, log_n, R, chunk):
    rng = np.random.default_rng(100 + log_n)
    n = 1 << log_n
    wires = rng.integers(0, P, size=(R, n), dtype=np.uint64)
    sigmas = rng.integers(0, P, size=(R, n), dtype=np.uint64)
    betas = rng.integers(0, P, size=2, dtype=np.uint64)
    gammas = rng.integers(0, P, size=2, dtype=np.uint64)
    want, _ = _partial_products_bigint(wires, sigmas, betas, gammas, chunk)
    got = orc.partial_products(wires, sigmas, betas, gammas, chunk)
    assert got.shape == want.shape and (got == want).all()


def test_partial_products_close_for_satisfied_copy_constraints(orc):
    """The defining property: when the wires are constant on the cycles of sigma, Z returns to 1 after the last row,
    i.e. Z(w^(n-1)) times the last row's quotient product is 1 -- and it does not when one wire is changed."""
    from perm_util import satisfied_permutation
    R, chunk, log_n = 20, 8, 5
    wires, sigmas, _ = satisfied_permutation(log_n, R, seed=3)
    betas = np.array([0x1234567890abcdef % P, 77], np.uint64)
    gammas = np.array([0xfedcba0987654321 % P, 5], np.uint64)
    _, z_end = _partial_products_bigint(wires, sigmas, betas, gammas, chunk)
    assert z_end == 1
    out = orc.partial_products(wires, sigmas, betas, gammas, chunk)
    assert (out[0:2, 0] == 1).all()                       # Z(1) = 1 for both challenges
    bad = wires.copy()
    bad[3, 7] = (int(bad[3, 7]) + 1) % P
    _, z_bad = _partial_products_bigint(bad, sigmas, betas, gammas, chunk)
    assert z_bad != 1


# ---- quotient polynomials of the permutation argument ---------------------------------------
def _gateless_instance(orc, log_n, R, extra_wires, num_consts, seed, break_wire=False):
    from perm_util import satisfied_permutation
    rng = np.random.default_rng(seed)
    n = 1 << log_n
    routed, sigmas, _ = satisfied_permutation(log_n, R, seed=seed)
    if break_wire:
        routed = routed.copy()
        routed[1, 3] = (int(routed[1, 3]) + 1) % P
    wires = np.concatenate([routed, rng.integers(0, P, size=(extra_wires, n), dtype=np.uint64)])
    consts = rng.integers(0, P, size=(num_consts, n), dtype=np.uint64)
    cs_b = orc.PolynomialBatch.from_values(np.concatenate([consts, sigmas]), 3, 2)
    w_b = orc.PolynomialBatch.from_values(wires, 3, 2)
    ch = rng.integers(0, P, size=(3, 2), dtype=np.uint64)
    zs = orc.partial_products(routed, sigmas, ch[0], ch[1], 8)
    z_b = orc.PolynomialBatch.from_values(zs, 3, 2)
    return dict(cs=cs_b, wires=w_b, zs=z_b, betas=ch[0], gammas=ch[1], alphas=ch[2], R=R, sigma_col0=num_consts, rng=rng)


def test_quotient_satisfies_the_verifier_identity(orc):
    """For wires that satisfy the copy constraints the quotient chunks pass the verifier's identity at a random point;
    with one wire changed they do not (the vanishing polynomial is no longer divisible by Z_H)."""
    from perm_util import quotient_identity_holds
    log_n, R = 5, 19
    for broken in (False, True):
        inst = _gateless_instance(orc, log_n, R, extra_wires=3, num_consts=2, seed=21, break_wire=broken)
        q = orc.quotient_polys(inst["cs"], inst["wires"], inst["zs"], inst["sigma_col0"], R, inst["betas"], inst["gammas"],
                               inst["alphas"])
        assert q.shape == (16, 1 << log_n)
        wc, sc, zc = inst["wires"].coeffs()[:R], inst["cs"].coeffs()[2:2 + R], inst["zs"].coeffs()
        for zeta in (0x123456789abcdef % P, 3):
            ok = quotient_identity_holds(log_n, wc, sc, zc, q, inst["betas"], inst["gammas"], inst["alphas"], 8, zeta)
            assert ok == (not broken)


def test_quotient_gate_terms_enter_the_alpha_combination(orc):
    """Gate terms of the form Z_H(x) r_g(x) (they vanish on the subgroup, as satisfied gate constraints do): the
    quotient still passes the verifier's identity, with the gate values Z_H(zeta) r_g(zeta) appended to the terms."""
    from perm_util import quotient_identity_holds, horner
    log_n, R, G = 4, 9, 3
    n, Lq = 1 << log_n, 8 << log_n
    inst = _gateless_instance(orc, log_n, R, extra_wires=1, num_consts=1, seed=5)
    rng = inst["rng"]
    r_coeffs = rng.integers(0, P, size=(G, n), dtype=np.uint64)
    g_pow_n = pow(orc.GENERATOR, n, P)
    w8 = pow(7277203076849721926, 1 << (32 - 3), P)
    gate = np.zeros((Lq, G), np.uint64)
    for g in range(G):
        padded = np.zeros(Lq, np.uint64)
        padded[:n] = r_coeffs[g]
        vals = orc.coset_fft(padded)                               # r_g on the quotient coset, natural order
        for i in range(Lq):
            zh = (g_pow_n * pow(w8, i % 8, P) - 1) % P
            gate[int(format(i, "0%db" % (log_n + 3))[::-1], 2), g] = int(vals[i]) * zh % P
    q = orc.quotient_polys(inst["cs"], inst["wires"], inst["zs"], inst["sigma_col0"], R, inst["betas"], inst["gammas"],
                           inst["alphas"], gate_terms=gate)
    wc, sc, zc = inst["wires"].coeffs()[:R], inst["cs"].coeffs()[1:1 + R], inst["zs"].coeffs()
    zeta = 0xabcdef0123456789 % P
    zh_z = (pow(zeta, n, P) - 1) % P
    gv = [zh_z * horner(r_coeffs[g], zeta) % P for g in range(G)]
    assert quotient_identity_holds(log_n, wc, sc, zc, q, inst["betas"], inst["gammas"], inst["alphas"], 8, zeta, gv)
    assert not quotient_identity_holds(log_n, wc, sc, zc, q, inst["betas"], inst["gammas"], inst["alphas"], 8, zeta)


def test_oracle_gateless_proof_passes_the_vanishing_check_on_opened_values(orc):
    """The CPU mirror of the whole flow (wires -> partial products -> quotient -> openings -> FRI): FRI verifies and the
    plonk vanishing identity holds on the opened extension-field values."""
    from perm_util import satisfied_permutation, verify_gateless_openings
    log_n, R, num_consts, num_wires = 6, 20, 2, 24
    n = 1 << log_n
    rng = np.random.default_rng(31)
    routed, sigmas, _ = satisfied_permutation(log_n, R, seed=4)
    wires = np.concatenate([routed, rng.integers(0, P, size=(num_wires - R, n), dtype=np.uint64)])
    cs = np.concatenate([rng.integers(0, P, size=(num_consts, n), dtype=np.uint64), sigmas])
    ocs = orc.PolynomialBatch.from_values(cs, 3, 4)
    pih = rng.integers(0, P, size=4, dtype=np.uint64)
    out = orc.prove_synthetic(ocs, wires, None, None, pih, log_n, sigmas=sigmas, sigma_col0=num_consts, pow_bits=8)
    assert orc.fri_verify(log_n, out["leaf_lens"], out["all_caps"], out["instance"], out["openings"], out["params"],
                          out["verifier_challenger"], out["fri_proof"]) == 0
    bg, alphas, zeta = out["challenges"]
    assert verify_gateless_openings(log_n, out["openings"], num_consts, R, num_wires, 2, 8, bg[:2], bg[2:], alphas, zeta)
    tampered = out["openings"].copy()
    tampered[num_consts + R + 1, 0] ^= np.uint64(1)
    assert not verify_gateless_openings(log_n, tampered, num_consts, R, num_wires, 2, 8, bg[:2], bg[2:], alphas, zeta)


# ---- gate constraints handed over as a program -------------------------------------------------------------------------
def _arith_instance(orc, log_n, seed, break_gate=False, break_copy=False):
    from perm_util import arithmetic_circuit
    wires, consts, sigmas = arithmetic_circuit(log_n, seed)
    if break_gate:
        wires = wires.copy(); wires[7, 2] = (int(wires[7, 2]) + 1) % P          # an output no gate produces
    if break_copy:
        wires = wires.copy(); wires[9, 1] = (int(wires[9, 1]) + 1) % P          # m1_2 != m0_3 (and its gate re-satisfied)
        wires[11, 1] = (int(consts[1, 1]) * int(wires[8, 1]) % P * int(wires[9, 1]) + int(consts[2, 1]) * int(wires[10, 1])) % P
    rng = np.random.default_rng(seed + 1)
    ch = rng.integers(0, P, size=(3, 2), dtype=np.uint64)
    cs_b = orc.PolynomialBatch.from_values(np.concatenate([consts, sigmas]), 3, 2)
    w_b = orc.PolynomialBatch.from_values(wires, 3, 2)
    zs = orc.partial_products(wires, sigmas, ch[0], ch[1], 8)
    z_b = orc.PolynomialBatch.from_values(zs, 3, 2)
    return dict(cs=cs_b, wires=w_b, zs=z_b, ch=ch, consts=consts, arrays=(wires, consts, sigmas))


def test_gate_program_quotient_passes_the_verifier_identity(orc):
    """An arithmetic-gate circuit whose gates arrive as a program: the oracle interprets it at every quotient point, the
    quotient passes the verifier's identity with the gate values computed by a big-integer interpreter from the
    polynomials at zeta -- and fails when a gate or a copy constraint is violated."""
    from perm_util import arithmetic_program, interpret_program, quotient_identity_holds, horner
    log_n = 5
    code, imm = arithmetic_program().arrays()
    for kind in ("ok", "gate", "copy"):
        inst = _arith_instance(orc, log_n, 77, break_gate=kind == "gate", break_copy=kind == "copy")
        ch = inst["ch"]
        gt = orc.gate_terms(inst["cs"], inst["wires"], code, imm)
        assert gt.shape == (8 << log_n, 4)
        q = orc.quotient_polys(inst["cs"], inst["wires"], inst["zs"], 3, 16, ch[0], ch[1], ch[2], gate_terms=gt)
        wc, cc, zc = inst["wires"].coeffs(), inst["cs"].coeffs(), inst["zs"].coeffs()
        zeta = 0x0123456789abcdef % P
        gv = interpret_program(code, imm, [horner(c, zeta) for c in wc], [horner(c, zeta) for c in cc[:3]])
        ok = quotient_identity_holds(log_n, wc, cc[3:19], zc, q, ch[0], ch[1], ch[2], 8, zeta, gv)
        assert ok == (kind == "ok"), kind
    # on the subgroup itself a satisfied gate evaluates to zero in every row
    wires, consts, _ = _arith_instance(orc, log_n, 77)["arrays"]
    for i in (0, 5, (1 << log_n) - 1):
        assert interpret_program(code, imm, wires[:, i], consts[:, i]) == [0, 0, 0, 0]


def test_oracle_proof_with_gates_passes_the_vanishing_check_on_opened_values(orc):
    from perm_util import arithmetic_circuit, arithmetic_program, verify_gateless_openings
    log_n = 6
    wires, consts, sigmas = arithmetic_circuit(log_n, 33)
    code, imm = arithmetic_program().arrays()
    cs = np.concatenate([consts, sigmas])
    ocs = orc.PolynomialBatch.from_values(cs, 3, 4)
    pih = np.arange(4, dtype=np.uint64)
    out = orc.prove_synthetic(ocs, wires, None, None, pih, log_n, sigmas=sigmas, sigma_col0=3, gate_program=(code, imm), pow_bits=8)
    assert orc.fri_verify(log_n, out["leaf_lens"], out["all_caps"], out["instance"], out["openings"], out["params"],
                          out["verifier_challenger"], out["fri_proof"]) == 0
    bg, alphas, zeta = out["challenges"]
    assert verify_gateless_openings(log_n, out["openings"], 3, 16, 16, 2, 8, bg[:2], bg[2:], alphas, zeta, program=(code, imm))
    assert not verify_gateless_openings(log_n, out["openings"], 3, 16, 16, 2, 8, bg[:2], bg[2:], alphas, zeta)


# ---- PoseidonGate as a gate program ------------------------------------------------------------------------------------
def test_poseidon_gate_program_vanishes_on_a_correct_trace(orc):
    """The reference's dominant gate (135 wires, 123 constraints of degree 7) recorded as a gate program: every
    constraint is zero on the wires of a real permutation trace -- whose outputs are the KAT-pinned permutation -- with
    and without the swap, and some constraint is non-zero when any single wire is changed."""
    from qpzk import gates
    from perm_util import interpret_program
    rc = gen.round_constants()
    fast = gen.derive_fast_partial(rc)
    rec = gates.Recorder()
    gates.poseidon_gate(rec, rc, fast)
    code, imm, regs = rec.finish()
    assert (code[:, 0] == 6).sum() == 123 and regs <= 256
    assert code.shape[0] < 3200 and rec.finish(fuse=False)[0].shape[0] > 5000   # multiply-adds by immediates fused
    rng = np.random.default_rng(8)
    for swap in (0, 1):
        inp = [int(x) for x in rng.integers(0, P, size=12, dtype=np.uint64)]
        w = gates.poseidon_gate_witness(inp, swap, rc)
        perm_in = inp if not swap else inp[4:8] + inp[0:4] + inp[8:]
        assert w[12:24] == [int(x) for x in orc.poseidon(np.array(perm_in, np.uint64))]
        assert interpret_program(code, imm, w, []) == [0] * 123
        for j in (0, 5, 24, 26, 30, 64, 65, 86, 87, 134, 12, 23):     # one wire of every group
            bad = list(w)
            bad[j] = (bad[j] + 1) % P
            assert any(interpret_program(code, imm, bad, []))
    # the three upstream vectors through the gate
    for inp, want in gen.KATS:
        w = gates.poseidon_gate_witness(inp, 0, rc)
        assert w[12:24] == [int(x, 16) for x in want.split()]


def test_poseidon2_gate_program_vanishes_on_a_correct_trace(orc):
    """The gate behind the reference's in-circuit hashing (`hash_n_to_hash_no_pad_p2::<Poseidon2Hash>`,
    wormhole/circuit/src/nullifier.rs:298-299) over the permutation that is pinned to the reference's own vectors: all 123
    constraints vanish on a real trace (with and without the swap), whose output wires are the oracle's -- pinned --
    permutation; any changed wire breaks some constraint; and a trace chained the way the sponge chains permutations
    reproduces the reference's first address vector."""
    import json
    import gen_poseidon2_constants as gp
    from qpzk import gates
    from perm_util import interpret_program
    ext, internal, diag = gp.parameters()
    flat = [c for row in ext for c in row]
    rec = gates.Recorder()
    gates.poseidon2_gate(rec, flat, internal, diag)
    code, imm, regs = rec.finish()
    assert (code[:, 0] == 6).sum() == 123 and regs <= 64
    rng = np.random.default_rng(28)
    for swap in (0, 1):
        inp = [int(x) for x in rng.integers(0, P, size=12, dtype=np.uint64)]
        w = gates.poseidon2_gate_witness(inp, swap, flat, internal, diag)
        perm_in = inp if not swap else inp[4:8] + inp[0:4] + inp[8:]
        assert w[12:24] == [int(x) for x in orc.poseidon2(np.array(perm_in, np.uint64))]
        assert interpret_program(code, imm, w, []) == [0] * 123
        for j in (0, 5, 24, 26, 30, 64, 65, 86, 87, 134, 12, 23):
            bad = list(w)
            bad[j] = (bad[j] + 1) % P
            assert any(interpret_program(code, imm, bad, []))
    # hash_no_pad as a chain of gate rows (what the circuit builder lays down): the reference's first address vector,
    # secret -> H(H(salt || secret)), wormhole/tests/src/circuit/unspendable_account_tests.rs:9-24
    G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "poseidon2_kats.json")))
    vec = {x["name"]: x for x in G["vectors"]}

    def hash_rows(felts):
        felts = [int(x) for x in felts] + [1]
        felts += [0] * (-len(felts) % 8)
        st = [0] * 12
        for off in range(0, len(felts), 8):
            w = gates.poseidon2_gate_witness([(st[i] + felts[off + i]) % P for i in range(8)] + st[8:], 0, flat, internal, diag)
            assert interpret_program(code, imm, w, []) == [0] * 123
            st = w[12:24]
        return st[:4]

    inner = hash_rows(vec["unspendable_account_0_inner"]["in"])
    assert inner == [int(x) for x in vec["unspendable_account_0_inner"]["out"]]
    outer = hash_rows(inner)
    assert b"".join(v.to_bytes(8, "little") for v in outer).hex() == "de68c6fcb3e38d6736b79a010e4504b98c6321f1e4d11cd8484f67c187ca090e"


def test_recorder_register_allocation(orc):
    from qpzk import gates
    from perm_util import interpret_program
    rec = gates.Recorder()
    a, b = rec.wire(0), rec.wire(1)
    t = rec.add(a, b)
    for _ in range(500):                                              # a long chain needs only a handful of registers
        t = rec.add(rec.mul(t, a), b)
    rec.emit(t)
    code, imm, regs = rec.finish()
    assert regs <= 4
    x, y, acc = 3, 5, 8
    for _ in range(500):
        acc = (acc * x + y) % P
    assert interpret_program(code, imm, [x, y], []) == [acc]
    rec = gates.Recorder()
    for v in [rec.mul(rec.wire(i % 7), rec.wire((i + 1) % 7)) for i in range(300)][::-1]:   # 300 values alive at once
        rec.emit(v)
    with pytest.raises(ValueError):
        rec.finish(256)


def test_oracle_poseidon_hash_chain_passes_the_vanishing_check(orc):
    """CPU mirror of the hash-chain proof (PoseidonGate program, 135 wires, 80 routed) at a small size."""
    from qpzk import gates
    from perm_util import poseidon_chain_circuit, verify_gateless_openings
    log_n = 4
    rc = gen.round_constants()
    rec = gates.Recorder()
    gates.poseidon_gate(rec, rc, gen.derive_fast_partial(rc), selector=rec.const(0))
    code, imm, _ = rec.finish()
    wires, consts, sigmas = poseidon_chain_circuit(log_n, 9, rc)
    cs = np.concatenate([consts, sigmas])
    ocs = orc.PolynomialBatch.from_values(cs, 3, 4)
    out = orc.prove_synthetic(ocs, wires, None, None, np.arange(4, dtype=np.uint64), log_n, sigmas=sigmas, sigma_col0=1,
                              gate_program=(code, imm), pow_bits=6, final_poly_bits=2)
    assert orc.fri_verify(log_n, out["leaf_lens"], out["all_caps"], out["instance"], out["openings"], out["params"],
                          out["verifier_challenger"], out["fri_proof"]) == 0
    bg, alphas, zeta = out["challenges"]
    assert verify_gateless_openings(log_n, out["openings"], 1, 80, 135, 2, 8, bg[:2], bg[2:], alphas, zeta, program=(code, imm))


def test_several_gates_share_constraint_slots_through_selector_filters(orc):
    """upstream's evaluate_gate_constraints: every gate's constraints are multiplied by its selector filter and summed
    slot by slot.  A circuit with Poseidon, arithmetic, constant and public-input rows: the combined program vanishes on
    a correct row of every kind (selected by the row's selector value), not on a row checked against the wrong gate."""
    from qpzk import gates
    from perm_util import interpret_program
    rc = gen.round_constants()
    fast = gen.derive_fast_partial(rc)
    pih = [11, 22, 33, 44]
    rec = gates.Recorder()
    s = rec.const(0)                                                   # selector polynomial: the gate index of the row
    group = range(4)
    n = gates.evaluate_gate_constraints(rec, [
        (lambda r: gates.poseidon_gate(r, rc, fast), gates.selector_filter(rec, s, 0, group, many_selectors=True)),
        (lambda r: gates.arithmetic_gate(r, 20, const_base=1), gates.selector_filter(rec, s, 1, group, many_selectors=True)),
        (lambda r: gates.constant_gate(r, 2, const_base=1), gates.selector_filter(rec, s, 2, group, many_selectors=True)),
        (lambda r: gates.public_input_gate(r, pih), gates.selector_filter(rec, s, 3, group, many_selectors=True)),
    ])
    assert n == 123
    code, imm, regs = rec.finish()
    assert (code[:, 0] & 0xff == 6).sum() == 123 and regs <= 256
    rng = np.random.default_rng(3)
    rnd = lambda k: [int(x) for x in rng.integers(0, P, size=k, dtype=np.uint64)]
    # row kinds: wires (135) and constants [selector, c0, c1]
    pos = gates.poseidon_gate_witness(rnd(12), 1, rc)
    c0, c1 = rnd(1)[0], rnd(1)[0]
    ar = rnd(135)
    for k in range(20):
        ar[4 * k + 3] = (c0 * ar[4 * k] % P * ar[4 * k + 1] + c1 * ar[4 * k + 2]) % P
    co = rnd(135); co[0], co[1] = c0, c1
    pi = rnd(135); pi[:4] = pih
    rows = {0: (pos, [0, 5, 6]), 1: (ar, [1, c0, c1]), 2: (co, [2, c0, c1]), 3: (pi, [3, 7, 8])}
    for kind, (w, consts) in rows.items():
        assert interpret_program(code, imm, w, consts) == [0] * 123, kind
        wrong = [(kind + 1) % 4] + consts[1:]                           # the same wires under another gate's selector
        assert any(interpret_program(code, imm, w, wrong)), kind
    unused = [gates.UNUSED_SELECTOR, 1, 2]                              # a row of another selector group: every filter is 0
    assert interpret_program(code, imm, rnd(135), unused) == [0] * 123
