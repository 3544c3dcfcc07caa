"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the CPU oracle
on the same seeded inputs -- bit-exact, since everything on this path is integer arithmetic."""
import numpy as np
import pytest

from conftest import P, splitmix_columns

pytestmark = pytest.mark.gpu


def brev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


# ---- Goldilocks arithmetic on the device, operation by operation -------------------------------
def test_field_ops_directed_and_random(ctx):
    """Every device formulation of the field arithmetic against Python integers.  The PTX carry chains have paths that
    uniform operands take with probability ~2^-32 (a wrapped sum that needs +EPS, a borrow that needs -EPS, both at
    once, a 128-bit product whose top word is all ones ...): the directed operands below are all pairs and triples of
    boundary values, which reach them; 2^17 uniform operands cover the bulk."""
    rng = np.random.default_rng(77)
    E = 2**32 - 1
    edge = [0, 1, 2, E - 1, E, E + 1, E + 2, 2**33 - 1, 2**48, 2**63, 2**63 + 1, P - 2**32, P - E - 1, P - 2, P - 1, P, P + 1,
            2**64 - 2**32, 2**64 - E, 2**64 - 2, 2**64 - 1, 0xFFFFFFFE00000001, 0xFFFFFFFF00000000, 0x00000001FFFFFFFF,
            0xFFFFFFFEFFFFFFFF, 0x7FFFFFFF80000000, 0x8000000080000000]
    n = len(edge)
    ta = np.array([edge[i] for i in range(n) for j in range(n) for k in range(n)], dtype=np.uint64)
    tb = np.array([edge[j] for i in range(n) for j in range(n) for k in range(n)], dtype=np.uint64)
    tc = np.array([edge[k] for i in range(n) for j in range(n) for k in range(n)], dtype=np.uint64)
    ra, rb, rc = (rng.integers(0, 2**64, size=1 << 17, dtype=np.uint64) for _ in range(3))
    A, B, Cc = np.concatenate([ta, ra]), np.concatenate([tb, rb]), np.concatenate([tc, rc])
    ai, bi, ci = [int(x) for x in A], [int(x) for x in B], [int(x) for x in Cc]
    u = lambda vals: np.array(vals, dtype=np.uint64)
    # any-representative operations, canonical result
    for op, f in ((2, lambda x, y, z: x * y % P), (3, lambda x, y, z: x * y % P), (4, lambda x, y, z: x * x % P),
                  (5, lambda x, y, z: (x * y + z) % P), (6, lambda x, y, z: x * y % P),
                  (7, lambda x, y, z: (x + (y << 64)) % P), (8, lambda x, y, z: (x + (y << 64)) % P),
                  (9, lambda x, y, z: (x + (y << 64)) % P),
                  (12, lambda x, y, z: (30 * x * y + y * z + z * x) % P)):
        got = ctx.selftest_field_ops(op, A, B, Cc)
        want = u([f(x, y, z) for x, y, z in zip(ai, bi, ci)])
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, (op, hex(ai[bad[0]]), hex(bi[bad[0]]), hex(ci[bad[0]]), hex(int(got[bad[0]])), hex(int(want[bad[0]])))
    # canonical-operand operations
    Ac, Bc, Zc = A % np.uint64(P), B % np.uint64(P), Cc % np.uint64(P)
    ac, bc, zc = [x % P for x in ai], [x % P for x in bi], [x % P for x in ci]
    assert (ctx.selftest_field_ops(0, Ac, Bc) == u([(x + y) % P for x, y in zip(ac, bc)])).all()
    assert (ctx.selftest_field_ops(1, Ac, Bc) == u([(x - y) % P for x, y in zip(ac, bc)])).all()
    assert (ctx.selftest_field_ops(10, A, Bc) == u([(x + y) % P for x, y in zip(ai, bc)])).all()
    ext = ctx.selftest_field_ops(13, Ac, Bc, Zc)                      # (a + b X)(c + a X), X^2 = 7
    want = u([((x * z + 7 * y * x) % P) if not (i & 1) else ((x * x + y * z) % P) for i, (x, y, z) in enumerate(zip(ac, bc, zc))])
    assert (ext == want).all()
    # x * 2^(6 j), j < 16, by shifts: every twiddle of the in-register transforms (w_32 = 2^6); all boundary values x every j
    ea = np.array([e % P for e in edge for j in range(16)], dtype=np.uint64)
    ej = np.array([j for e in edge for j in range(16)], dtype=np.uint64)
    pa, pj = np.concatenate([ea, Ac]), np.concatenate([ej, B % np.uint64(16)])
    got = ctx.selftest_field_ops(14, pa, pj)
    want = u([int(x) * pow(2, 6 * int(j), P) % P for x, j in zip(pa, pj)])
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, (hex(int(pa[bad[0]])), int(pj[bad[0]]), hex(int(got[bad[0]])), hex(int(want[bad[0]])))
    nz = Ac[:4096].copy()
    nz[nz == 0] = 1
    inv = ctx.selftest_field_ops(11, nz)
    assert all(int(x) * int(y) % P == 1 for x, y in zip(nz, inv))


# ---- PoseidonPermutation::permute and its linear layer -----------------------------------------------------------
MDS_CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]


def test_poseidon_permute_matches_oracle_and_kats(ctx, orc):
    rng = np.random.default_rng(5)
    states = rng.integers(0, P, size=(300, 12), dtype=np.uint64)
    states[0] = 0
    states[1] = np.arange(12)
    states[2] = P - 1
    got = ctx.poseidon_permute(states)
    want = np.stack([orc.poseidon(s) for s in states])
    assert (got == want).all()
    assert got[0][0] == 0x3c18a9786cb0b359 and got[1][0] == 0xd64e1e3efc5b8e9e and got[2][0] == 0xbe0085cfc57a8357
    # non-canonical representatives (x and x + p name the same field element) must give the same permutation
    nc = states[:64].copy()
    small = nc < np.uint64(2**32 - 1)
    nc[small] += np.uint64(P)
    assert small.any() or True
    edge = np.full((4, 12), 2**64 - 1, np.uint64)           # == 2^32 - 2 mod p
    assert (ctx.poseidon_permute(edge) == np.stack([orc.poseidon(np.full(12, (2**64 - 1) % P, np.uint64))] * 4)).all()
    assert (ctx.poseidon_permute(nc) == want[:64]).all()


def test_poseidon_permute_bulk_and_extreme_states(ctx, orc):
    """60 k permutations against the oracle: uniform states, states of extreme words (all limbs saturated: the limb-domain
    partial rounds carry between limbs and wrap the top limb in every round), sparse states and non-canonical words."""
    import ctypes as C
    rng = np.random.default_rng(2024)
    ext = np.array([0, 1, P - 1, P - 2, 2**32 - 1, 2**32, 2**44 - 1, 2**44, 2**22 - 1, 2**22, 0xFFFFFFFF00000000,
                    0xFFFFF00000000000, 0x00000FFFFFFFFFFF, 0x003FFFFF003FFFFF & (2**64 - 1), 0x8000000000000000], dtype=np.uint64)
    a = rng.integers(0, P, size=(40000, 12), dtype=np.uint64)
    b = ext[rng.integers(0, ext.size, size=(15000, 12))]
    c = np.zeros((5000, 12), np.uint64)
    c[np.arange(5000), rng.integers(0, 12, 5000)] = ext[rng.integers(0, ext.size, 5000)]
    states = np.concatenate([a, b, c])
    got = ctx.poseidon_permute(states)
    want = states.copy()
    L = orc.lib()
    for row in want:
        L.orc_poseidon(row.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert (got == want).all()
    nc = np.full((64, 12), 2**64 - 1, np.uint64)              # 2^64 - 1 = 2^32 - 2 (mod p)
    nc[:, ::2] = np.uint64(P)                                  # p = 0 (mod p)
    canon = nc % np.uint64(P)
    assert (ctx.poseidon_permute(nc) == np.stack([orc.poseidon(r) for r in canon])).all()


def test_mds_layer_carry_paths(ctx, orc):
    """One linear layer (limb-domain MDS + next round's constants) on states chosen to hit every carry/wrap path:
    all-ones words, values just below/above p and 2^64, single hot lanes, limb boundaries (2^22, 2^44)."""
    rc = orc.round_constants()
    rng = np.random.default_rng(9)
    specials = [0, 1, P - 1, P, P + 1, 2**64 - 1, 2**64 - 2**32, 2**32 - 1, 2**32, 2**22 - 1, 2**22, 2**44 - 1, 2**44,
                0xFFFFF00000000000, 0xFFFFFFFFFFC00000, 0x00000FFFFFFFFFFF, 0x8000000000000000]
    states = [[v] * 12 for v in specials]
    for lane in range(12):
        for v in (2**64 - 1, P - 1):
            st = [0] * 12
            st[lane] = v
            states.append(st)
    for _ in range(400):
        states.append([int(specials[i]) for i in rng.integers(0, len(specials), 12)])
    for _ in range(400):
        hi = rng.integers(0, 2**20, 12)
        states.append([(2**64 - 1) - int(h) for h in hi])        # top limb all ones: sums that cross 2^64
    arr = np.array(states, dtype=np.uint64)
    for nxt in (0, 1, 4, 17, 29, 30):
        got = ctx.poseidon_permute(arr, mds_only_round=nxt)
        for st, g in zip(states, got):
            for r in range(12):
                v = sum(MDS_CIRC[i] * st[(i + r) % 12] for i in range(12)) + (8 * st[0] if r == 0 else 0)
                if 1 <= nxt <= 29:
                    v += int(rc[12 * nxt + r])
                assert int(g[r]) == v % P, (nxt, st, r)


# ---- MerkleTree::new -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nl,ll,h", [(1, 5, 0), (1, 3, 0), (2, 9, 1), (8, 3, 0), (8, 4, 3), (16, 7, 2), (64, 8, 4),
                                     (512, 135, 4), (1024, 32, 4), (4096, 20, 0), (2048, 143, 8), (256, 1, 2)])
def test_merkle_tree_matches_oracle(ctx, qpzk, orc, nl, ll, h):
    rng = np.random.default_rng(nl * 1000 + ll)
    leaves = rng.integers(0, P, size=(nl, ll), dtype=np.uint64)
    t = qpzk.MerkleTree.new(ctx, leaves, h)
    o = orc.MerkleTree(leaves, h)
    assert (t.cap == o.cap()).all()
    k = min(nl, 64)
    assert (t.digests(0)[:k] == np.stack([orc.hash_or_noop(l) for l in leaves[:k]])).all()
    idx = sorted({0, nl - 1, nl // 3, nl // 2})
    rows, sib = t.open(idx)
    for k, i in enumerate(idx):
        assert (rows[k] == leaves[i]).all()
        assert (sib[k] == o.prove(i)).all()
        assert orc.merkle_verify(rows[k], i, t.cap, sib[k])


def test_merkle_edge_values(ctx, qpzk, orc):
    # all-zero and all-(p-1) leaves exercise the reduction corner cases of the permutation
    for v in (0, P - 1, 0xFFFFFFFF, 0xFFFFFFFF00000000):
        leaves = np.full((16, 12), v, np.uint64)
        assert (qpzk.MerkleTree.new(ctx, leaves, 1).cap == orc.MerkleTree(leaves, 1).cap()).all()


def test_merkle_errors(ctx, qpzk):
    rng = np.random.default_rng(0)
    with pytest.raises(qpzk.QpzkError):
        qpzk.MerkleTree.new(ctx, rng.integers(0, P, size=(8, 3), dtype=np.uint64), 4)   # cap_height > log2(leaves)
    with pytest.raises(qpzk.QpzkError):
        qpzk.MerkleTree.new(ctx, rng.integers(0, P, size=(6, 3), dtype=np.uint64), 0)   # not a power of two


# ---- PolynomialBatch::from_values / from_coeffs -----------------------------------------------------------------
@pytest.mark.parametrize("log_n,C_", [(0, 3), (1, 2), (2, 1), (3, 5), (5, 9), (6, 135), (8, 135), (10, 20), (11, 3),
                                      (12, 16), (13, 135), (13, 85), (14, 7), (15, 4), (16, 5), (17, 2)])
def test_commit_from_values_matches_oracle(ctx, qpzk, orc, log_n, C_):
    cols = splitmix_columns(0x51575A4B00000000 + log_n, C_, 1 << log_n)
    h = min(4, log_n + 3)
    b = qpzk.PolynomialBatch.from_values(ctx, cols, 3, None, h)
    o = orc.PolynomialBatch.from_values(cols, 3, h)
    assert (b.polynomials == o.coeffs()).all()
    assert (b.merkle_tree.rows(0, 1 << (log_n + 3)) == o.leaves()).all()
    assert (b.cap == o.cap()).all()
    b2 = qpzk.PolynomialBatch.from_coeffs(ctx, o.coeffs(), 3, None, h)
    assert (b2.cap == o.cap()).all()
    L = 1 << (log_n + 3)
    i = (3 * L) // 7
    assert (b.get_lde_values(i) == o.leaves()[brev(i, log_n + 3)]).all()


@pytest.mark.parametrize("rate_bits,h", [(0, 0), (1, 1), (2, 3), (4, 5), (3, 0), (3, 8)])
def test_commit_other_rates_and_caps(ctx, qpzk, orc, rate_bits, h):
    cols = splitmix_columns(42 + rate_bits, 6, 1 << 7)
    b = qpzk.PolynomialBatch.from_values(ctx, cols, rate_bits, None, h)
    o = orc.PolynomialBatch.from_values(cols, rate_bits, h)
    assert (b.cap == o.cap()).all()
    assert (b.merkle_tree.rows(0, 1 << (7 + rate_bits)) == o.leaves()).all()


def test_commit_with_salt(ctx, qpzk, orc):
    rng = np.random.default_rng(21)
    log_n, C_ = 7, 10
    cols = splitmix_columns(5, C_, 1 << log_n)
    salt = rng.integers(0, P, size=(4, 1 << (log_n + 3)), dtype=np.uint64)
    b = qpzk.PolynomialBatch.from_values(ctx, cols, 3, salt, 4)
    o = orc.PolynomialBatch.from_values(cols, 3, 4, salt)
    assert b.leaf_len == C_ + 4
    assert (b.merkle_tree.rows(0, 1 << (log_n + 3)) == o.leaves()).all()
    assert (b.cap == o.cap()).all()


def test_commit_edge_values(ctx, qpzk, orc):
    n = 64
    cols = np.stack([np.zeros(n, np.uint64), np.full(n, P - 1, np.uint64), np.arange(n, dtype=np.uint64),
                     np.full(n, 0xFFFFFFFF00000000, np.uint64), np.full(n, 0xFFFFFFFF, np.uint64)])
    b = qpzk.PolynomialBatch.from_values(ctx, cols, 3, None, 4)
    o = orc.PolynomialBatch.from_values(cols, 3, 4)
    assert (b.polynomials == o.coeffs()).all() and (b.cap == o.cap()).all()


def test_commit_errors(ctx, qpzk):
    cols = splitmix_columns(1, 2, 8)
    with pytest.raises(qpzk.QpzkError):
        qpzk.PolynomialBatch.from_values(ctx, cols, 3, None, 7)       # cap_height > log2(8*8)
    with pytest.raises(qpzk.QpzkError):
        qpzk.PolynomialBatch.from_values(ctx, cols[:, :6], 3, None, 1)  # length not a power of two
    with pytest.raises(qpzk.QpzkError):
        qpzk.PolynomialBatch.from_values(ctx, cols, 3, np.zeros((4, 8), np.uint64), 1)  # wrong salt shape


def test_linearity_of_the_lde(ctx, qpzk):
    # size-independent property: LDE(a + b) = LDE(a) + LDE(b) (mod p), checked on rows
    log_n = 12
    a = splitmix_columns(7, 4, 1 << log_n)
    b = splitmix_columns(8, 4, 1 << log_n)
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    ra = qpzk.PolynomialBatch.from_values(ctx, a, 3, None, 4).merkle_tree.rows(100, 50).astype(object)
    rb = qpzk.PolynomialBatch.from_values(ctx, b, 3, None, 4).merkle_tree.rows(100, 50).astype(object)
    rs = qpzk.PolynomialBatch.from_values(ctx, s, 3, None, 4).merkle_tree.rows(100, 50).astype(object)
    assert ((ra + rb) % P == rs).all()


def test_coset_split_equals_full_commit(ctx, qpzk, orc):
    """leaf block j is one Merkle subtree; per-coset commits reproduce slices of the full cap."""
    import torch
    log_n, C_ = 9, 11
    cols = splitmix_columns(99, C_, 1 << log_n)
    full = qpzk.PolynomialBatch.from_values(ctx, cols, 3, None, 4)
    coeffs = torch.from_numpy(full.polynomials.view(np.int64)).cuda()
    torch.cuda.synchronize()
    cap = full.cap
    for count in (1, 2, 4, 8):
        got = []
        for begin in range(0, 8, count):
            part = qpzk.PolynomialBatch.cosets_from_coeffs_device(ctx, coeffs.data_ptr(), log_n, C_, 3, 4, begin, count)
            got.append(part.cap)
            n = 1 << log_n
            assert (part.merkle_tree.rows(0, 16) == full.merkle_tree.rows(begin * n, 16)).all()
        assert (np.concatenate(got) == cap).all()


# ---- openings, challenger ------------------------------------------------------------------------------------------
def test_eval_ext_matches_oracle(ctx, qpzk, orc):
    for log_n, C_ in [(0, 2), (3, 4), (7, 33), (10, 135), (13, 20)]:
        cols = splitmix_columns(log_n + 300, C_, 1 << log_n)
        b = qpzk.PolynomialBatch.from_values(ctx, cols, 3, None, min(4, log_n + 3))
        o = orc.PolynomialBatch.from_values(cols, 3, min(4, log_n + 3))
        for z in ([3, 5], [P - 1, 0], [0, 0], [12345678901234567, 98765432109876543]):
            assert (b.eval_ext(z) == o.eval_ext(z)).all()


def test_eval_ext_many_matches_single_calls(ctx, qpzk, orc):
    rng = np.random.default_rng(12)
    batches = [qpzk.PolynomialBatch.from_values(ctx, rng.integers(0, P, size=(c, 256), dtype=np.uint64), 3, None, 4)
               for c in (5, 1, 17)]
    z = [int(x) for x in rng.integers(0, P, size=2, dtype=np.uint64)]
    many = qpzk.PolynomialBatch.eval_ext_many(batches, z)
    assert (many == np.concatenate([b.eval_ext(z) for b in batches])).all()
    with pytest.raises(qpzk.QpzkError):
        qpzk.PolynomialBatch.eval_ext_many(batches, [P, 0])          # point not canonical
    other = qpzk.PolynomialBatch.from_values(ctx, rng.integers(0, P, size=(2, 128), dtype=np.uint64), 3, None, 4)
    with pytest.raises(qpzk.QpzkError):
        qpzk.PolynomialBatch.eval_ext_many([batches[0], other], z)   # different degrees


def test_challenger_matches_oracle(ctx, qpzk, orc):
    g, o = qpzk.Challenger(ctx), orc.Challenger()
    rng = np.random.default_rng(1)
    for n_obs, n_get in [(3, 2), (8, 1), (13, 9), (0, 3), (1, 1), (64, 4), (31, 1), (33, 8), (100, 5), (516, 2), (32, 17)]:
        x = rng.integers(0, P, size=n_obs, dtype=np.uint64)
        g.observe_elements(x); o.observe(x)
        assert (g.get_n_challenges(n_get) == o.get(n_get)).all()
        assert g.export_state() == o.export()
    st = g.export_state()
    g2 = qpzk.Challenger(ctx)
    g2.import_state(*st)
    assert g2.get_challenge() == int(o.get(1)[0])
    with pytest.raises(qpzk.QpzkError):
        g.observe_elements([P])  # non-canonical input is rejected at the boundary


# ---- FRI -----------------------------------------------------------------------------------------------------------
def _setup(ctx, qpzk, orc, log_n, Cs, salted=()):
    n = 1 << log_n
    rng = np.random.default_rng(5 + log_n)
    gb, ob = [], []
    for i, c in enumerate(Cs):
        cols = splitmix_columns(77 + i + log_n, c, n)
        salt = rng.integers(0, P, size=(4, n << 3), dtype=np.uint64) if i in salted else None
        if i == len(Cs) - 1:  # the quotient oracle enters as coefficients (from_coeffs)
            gb.append(qpzk.PolynomialBatch.from_coeffs(ctx, cols, 3, salt, 4))
            ob.append(orc.PolynomialBatch.from_coeffs(cols, 3, 4, salt))
        else:
            gb.append(qpzk.PolynomialBatch.from_values(ctx, cols, 3, salt, 4))
            ob.append(orc.PolynomialBatch.from_values(cols, 3, 4, salt))
    gch, och = qpzk.Challenger(ctx), orc.Challenger()
    digest = np.arange(4, dtype=np.uint64) + 11  # stands for circuit_digest / public_inputs_hash
    gch.observe_hash(digest); och.observe(digest)
    for g, o in zip(gb, ob):
        gch.observe_cap(g); och.observe(o.cap())
    zg, zo = gch.get_extension_challenge(), och.get(2)
    assert (zg == zo).all()
    wn = orc.lib().orc_gl_root_of_unity(log_n)
    gz = [orc.lib().orc_gl_mul(int(zo[0]), wn), orc.lib().orc_gl_mul(int(zo[1]), wn)]
    allp = [(o, i) for o in range(len(Cs)) for i in range(Cs[o])]
    nxt = [(2, i) for i in range(min(2, Cs[2]))]
    batches = [([int(zo[0]), int(zo[1])], allp), (gz, nxt)]
    op = np.concatenate([np.concatenate([ob[o].eval_ext(zo) for o in range(len(Cs))]),
                         np.stack([ob[o].eval_ext(gz)[i] for o, i in nxt])])
    gch.observe_extension_elements(op); och.observe(op)
    return gb, ob, gch, och, qpzk.FriInstanceInfo(batches), orc.Instance(batches), op


@pytest.mark.parametrize("log_n,Cs,pow_bits,nq,fpb,salted", [
    (6, [5, 9, 3, 2], 8, 5, 2, ()),
    (5, [3, 4, 2, 2], 4, 3, 5, ()),           # no reduction rounds at all
    (8, [4, 6, 3, 2], 10, 7, 3, (1, 2, 3)),   # salted (blinding) oracles
    (10, [8, 13, 4, 4], 16, 28, 5, ()),
    (13, [85, 135, 20, 16], 16, 28, 5, ()),   # leaf-circuit shape, reference FRI parameters
])
def test_fri_proof_matches_oracle_and_verifies(ctx, qpzk, orc, log_n, Cs, pow_bits, nq, fpb, salted):
    gb, ob, gch, och, gi, oi, op = _setup(ctx, qpzk, orc, log_n, Cs, salted)
    gp = qpzk.fri_params(log_n, 3, 4, pow_bits, nq, 4, fpb)
    opar = orc.fri_params(log_n, 3, 4, pow_bits, nq, 4, fpb)
    ochv = och.clone()
    gproof, gopen = qpzk.PolynomialBatch.prove_openings(gi, gb, gch, gp, return_openings=True)
    oproof = orc.fri_prove(ob, oi, opar, och)   # coefficient-domain prover (upstream's algorithm)
    assert (gopen == op).all()
    assert (gproof == oproof).all()            # caps, query rounds, final poly, minimum PoW witness: bit-exact
    assert gch.export_state() == och.export()  # transcript left in the same state
    lens, caps = [b.leaf_len for b in gb], [b.cap for b in gb]
    assert orc.fri_verify(log_n, lens, caps, oi, op, opar, ochv, gproof) == 0


def test_fri_caller_supplied_pow_witness(ctx, qpzk, orc):
    gb, ob, gch, och, gi, oi, op = _setup(ctx, qpzk, orc, 6, [5, 9, 3, 2])
    gp = qpzk.fri_params(6, 3, 4, 8, 4, 4, 2)
    st = gch.export_state()
    proof = qpzk.PolynomialBatch.prove_openings(gi, gb, gch, gp)
    w = int(proof[-1])
    g2 = qpzk.Challenger(ctx); g2.import_state(*st)
    assert (qpzk.PolynomialBatch.prove_openings(gi, gb, g2, gp, pow_witness=w) == proof).all()
    if w > 0:  # the ground witness is the minimum: nothing below it is accepted
        g3 = qpzk.Challenger(ctx); g3.import_state(*st)
        with pytest.raises(qpzk.QpzkError) as e:
            qpzk.PolynomialBatch.prove_openings(gi, gb, g3, gp, pow_witness=w - 1)
        assert e.value.code == -4


def test_fri_argument_errors(ctx, qpzk, orc):
    gb, ob, gch, och, gi, oi, op = _setup(ctx, qpzk, orc, 6, [5, 9, 3, 2])
    gp = qpzk.fri_params(6, 3, 4, 8, 4, 4, 2)
    bad = qpzk.FriInstanceInfo([([1, 2], [(0, 99)])])
    with pytest.raises(qpzk.QpzkError):
        qpzk.PolynomialBatch.prove_openings(bad, gb, gch, gp)
    bad = qpzk.FriInstanceInfo([([P, 2], [(0, 0)])])
    with pytest.raises(qpzk.QpzkError):
        qpzk.PolynomialBatch.prove_openings(bad, gb, gch, gp)


# ---- BASELINE.json full-size shape through size-independent properties ---------------------------------------------
def test_full_size_commit_properties(ctx, qpzk, orc):
    """2^18 x 135 (config 3 of BASELINE.json is 2^16..2^22; the oracle needs minutes beyond 2^16, so the larger
    sizes are checked through properties): sampled LDE rows equal direct polynomial evaluation of the returned
    coefficients, sampled Merkle paths verify against the cap with the oracle hasher, first coset equals the
    oracle's LDE restricted to that coset."""
    import torch
    log_n, C_ = 18, 135
    n, log_L = 1 << log_n, log_n + 3
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randint(0, 2**62, (C_, n), generator=g, device="cuda", dtype=torch.int64)  # < p, canonical
    torch.cuda.synchronize()
    b = qpzk.PolynomialBatch.from_device(ctx, x.data_ptr(), log_n, C_, 3, 4, values=True)
    cap = b.cap
    cols = x[:3].cpu().numpy().view(np.uint64)
    idx = [0, 1, n - 1, n, 5 * n + 12345, (1 << log_L) - 1, 777777]
    rows, sib = b.merkle_tree.open(idx)
    wl = orc.lib().orc_gl_root_of_unity(log_L)
    for k, j in enumerate(idx):
        assert orc.merkle_verify(rows[k], j, cap, sib[k])
        pt = orc.GENERATOR * pow(wl, brev(j, log_L), P) % P
        for c in range(3):  # Horner on the oracle's coefficients of the same column
            co = orc.fft(cols[c], inverse=True)
            acc = 0
            for v in co[::-1]:
                acc = (acc * pt + int(v)) % P
            assert acc == int(rows[k][c])
    assert (b.polynomials[:3] == np.stack([orc.fft(cols[c], inverse=True) for c in range(3)])).all()


# ---- the commitment / opening half of prove() (leaf-shaped synthetic prove, BASELINE.json configs 1/2) -------------
@pytest.mark.parametrize("log_n", [8, 13])
def test_prove_synthetic_matches_oracle_and_verifies(ctx, qpzk, orc, log_n):
    from qpzk import prover
    shape = dict(prover.LEAF_SHAPE, log_n=log_n)
    rng = np.random.default_rng(40 + log_n)
    cs = rng.integers(0, P, size=(shape["constants_sigmas"], 1 << log_n), dtype=np.uint64)
    circuit = prover.SyntheticCircuit(ctx, cs, shape)
    inp = prover.synthetic_inputs(7, shape)
    got = prover.prove_synthetic(circuit, **inp)
    ocs = orc.PolynomialBatch.from_values(cs, 3, 4)
    want = orc.prove_synthetic(ocs, inp["wires"], inp["zs_partial_products"], inp["quotient_coeffs"],
                               inp["public_inputs_hash"], log_n)
    for a, b in zip(got["caps"], want["caps"]):
        assert (a == b).all()
    assert (got["openings"] == want["openings"]).all()
    assert (got["fri_proof"] == want["fri_proof"]).all()
    rc = orc.fri_verify(log_n, want["leaf_lens"], want["all_caps"], want["instance"], want["openings"], want["params"],
                        want["verifier_challenger"], got["fri_proof"])
    assert rc == 0
    # steady state: a second proof of the same shape allocates nothing new and frees everything it used
    import gc
    gc.collect()          # batches of earlier tests that are still waiting for the collector would join the cache mid-measurement
    cached = ctx.cached_bytes()
    prover.prove_synthetic(circuit, **prover.synthetic_inputs(8, shape))
    assert ctx.cached_bytes() == cached


# ---- permutation argument: all_wires_permutation_partial_products (row a16) ---------------------------------------
@pytest.mark.parametrize("log_n,R,chunk,nch", [(13, 80, 8, 2), (15, 60, 8, 2), (6, 7, 8, 1), (0, 3, 2, 2), (1, 13, 5, 3),
                                              (11, 135, 8, 2), (16, 80, 8, 2), (10, 256, 8, 1)])
def test_partial_products_match_oracle(ctx, orc, log_n, R, chunk, nch):
    rng = np.random.default_rng(1000 + log_n)
    n = 1 << log_n
    wires = rng.integers(0, P, size=(R, n), dtype=np.uint64)
    sigmas = rng.integers(0, P, size=(R, n), dtype=np.uint64)
    if n > 4:
        wires[0, :4] = [0, P - 1, 1, P - 2]
        sigmas[0, :4] = [P - 1, 0, P - 2, 1]
    betas = rng.integers(0, P, size=nch, dtype=np.uint64)
    gammas = rng.integers(0, P, size=nch, dtype=np.uint64)
    got = ctx.partial_products(wires, sigmas, betas, gammas, chunk)
    want = orc.partial_products(wires, sigmas, betas, gammas, chunk)
    assert got.shape == want.shape == (nch * ((R + chunk - 1) // chunk), n)
    assert (got == want).all()


def test_partial_products_close_and_feed_the_commitment(ctx, orc, qpzk):
    """Wires that satisfy a random copy-constraint system: Z(1) = 1 and the running product closes (the last row's
    Z times its quotient product is 1, checked through the oracle's row quotients), and the output, committed with
    from_values, gives the cap the oracle gets from its own partial products."""
    from perm_util import root_of_unity, satisfied_permutation
    log_n, R = 9, 24
    wires, sigmas, _ = satisfied_permutation(log_n, R, seed=11)
    betas = np.array([3, P - 5], np.uint64)
    gammas = np.array([P - 1, 12345678901234567], np.uint64)
    got = ctx.partial_products(wires, sigmas, betas, gammas, 8)
    want = orc.partial_products(wires, sigmas, betas, gammas, 8)
    assert (got == want).all() and (got[:2, 0] == 1).all()
    n = 1 << log_n
    for c in range(2):      # last row: Z(w^(n-1)) * prod of all quotients of that row == Z(w^n) == 1
        num = den = 1
        x = pow(root_of_unity(log_n), n - 1, P)
        for j in range(R):
            k = pow(14293326489335486720, j, P)
            num = num * ((int(wires[j, n - 1]) + int(betas[c]) * k * x + int(gammas[c])) % P) % P
            den = den * ((int(wires[j, n - 1]) + int(betas[c]) * int(sigmas[j, n - 1]) + int(gammas[c])) % P) % P
        assert int(got[c, n - 1]) * num % P == den
    gb = qpzk.PolynomialBatch.from_values(ctx, got, 3, None, 4)
    ob = orc.PolynomialBatch.from_values(want, 3, 4)
    assert (gb.cap == ob.cap()).all()


def test_partial_products_rejects_bad_arguments(ctx, qpzk):
    w = np.zeros((4, 8), np.uint64)
    with pytest.raises(qpzk.QpzkError):
        ctx.partial_products(w, w, [P], [0])                 # non-canonical challenge
    with pytest.raises(qpzk.QpzkError):
        ctx.partial_products(w, w[:, :4], [1], [0])          # shape mismatch
    with pytest.raises(qpzk.QpzkError):
        ctx.partial_products(np.zeros((300, 8), np.uint64), np.zeros((300, 8), np.uint64), [1], [0], chunk=8)  # 38 chunks


def test_prove_synthetic_with_device_partial_products(ctx, qpzk, orc):
    """Same transcript, but the zs / partial-product columns come from the routed wires and the sigma values through
    the permutation-argument kernels, with the betas and gammas the transcript produced."""
    from qpzk import prover
    log_n, routed = 9, 80
    shape = dict(prover.LEAF_SHAPE, log_n=log_n)
    rng = np.random.default_rng(77)
    cs = rng.integers(0, P, size=(shape["constants_sigmas"], 1 << log_n), dtype=np.uint64)
    circuit = prover.SyntheticCircuit(ctx, cs, shape)
    inp = prover.synthetic_inputs(9, shape)
    inp["zs_partial_products"] = None
    sigmas = rng.integers(0, P, size=(routed, 1 << log_n), dtype=np.uint64)
    got = prover.prove_synthetic(circuit, sigmas=sigmas, **inp)
    ocs = orc.PolynomialBatch.from_values(cs, 3, 4)
    want = orc.prove_synthetic(ocs, inp["wires"], None, inp["quotient_coeffs"], inp["public_inputs_hash"], log_n,
                               sigmas=sigmas)
    assert want["leaf_lens"][2] == 2 * ((routed + 7) // 8) == shape["zs_partial_products"]
    for a, b in zip(got["caps"], want["caps"]):
        assert (a == b).all()
    assert (got["openings"] == want["openings"]).all()
    assert (got["fri_proof"] == want["fri_proof"]).all()
    assert orc.fri_verify(log_n, want["leaf_lens"], want["all_caps"], want["instance"], want["openings"], want["params"],
                          want["verifier_challenger"], got["fri_proof"]) == 0


# ---- quotient polynomials of the permutation argument (row a17) ----------------------------------------------------
def _gateless_batches(ctx, qpzk, orc, log_n, R, num_wires, num_consts, seed, rate_bits=3):
    from perm_util import satisfied_permutation
    rng = np.random.default_rng(seed)
    n = 1 << log_n
    routed, sigmas, _ = satisfied_permutation(log_n, R, seed=seed)
    wires = np.concatenate([routed, rng.integers(0, P, size=(num_wires - R, n), dtype=np.uint64)])
    cs = np.concatenate([rng.integers(0, P, size=(num_consts, n), dtype=np.uint64), sigmas])
    ch = rng.integers(0, P, size=(3, 2), dtype=np.uint64)
    zs = orc.partial_products(routed, sigmas, ch[0], ch[1], 8)
    g = [qpzk.PolynomialBatch.from_values(ctx, a, rate_bits, None, 2) for a in (cs, wires, zs)]
    o = [orc.PolynomialBatch.from_values(a, rate_bits, 2) for a in (cs, wires, zs)]
    return g, o, ch


@pytest.mark.parametrize("log_n,R,num_wires,num_consts,G", [(5, 19, 22, 2, 0), (9, 80, 135, 5, 0), (6, 8, 8, 0, 0),
                                                           (7, 60, 135, 4, 5), (3, 1, 2, 1, 2)])
def test_quotient_polys_match_oracle(ctx, qpzk, orc, log_n, R, num_wires, num_consts, G):
    import torch
    g, o, ch = _gateless_batches(ctx, qpzk, orc, log_n, R, num_wires, num_consts, seed=300 + log_n)
    gate = gate_dev = None
    if G:
        gate = np.random.default_rng(log_n).integers(0, P, size=(8 << log_n, G), dtype=np.uint64)
        gate_t = torch.from_numpy(gate.view(np.int64)).cuda()
        gate_dev = gate_t.data_ptr()
    want = orc.quotient_polys(o[0], o[1], o[2], num_consts, R, ch[0], ch[1], ch[2], gate_terms=gate)
    got = qpzk.PolynomialBatch.quotient_polys(g[0], g[1], g[2], num_consts, R, ch[0], ch[1], ch[2],
                                              gate_terms_dev=gate_dev, num_gate_terms=G)
    assert got.shape == want.shape == (16, 1 << log_n)
    assert (got == want).all()
    qb = qpzk.PolynomialBatch.commit_quotient(g[0], g[1], g[2], num_consts, R, ch[0], ch[1], ch[2], 2,
                                              gate_terms_dev=gate_dev, num_gate_terms=G)
    assert (qb.cap == orc.PolynomialBatch.from_coeffs(want, 3, 2).cap()).all()
    assert (qb.polynomials == want).all()


def test_quotient_with_smaller_degree_factor_than_blowup(ctx, qpzk, orc):
    """quotient_degree_factor 4 on rate_bits 3 batches: the quotient domain is every second LDE point = the first
    half of the leaf rows."""
    from perm_util import satisfied_permutation
    log_n, R = 6, 11
    routed, sigmas, _ = satisfied_permutation(log_n, R, seed=8)
    rng = np.random.default_rng(8)
    ch = rng.integers(0, P, size=(3, 1), dtype=np.uint64)
    zs = orc.partial_products(routed, sigmas, ch[0], ch[1], 4)
    g = [qpzk.PolynomialBatch.from_values(ctx, a, 3, None, 2) for a in (sigmas, routed, zs)]
    o = [orc.PolynomialBatch.from_values(a, 3, 2) for a in (sigmas, routed, zs)]
    want = orc.quotient_polys(o[0], o[1], o[2], 0, R, ch[0], ch[1], ch[2], quotient_degree_factor=4)
    got = qpzk.PolynomialBatch.quotient_polys(g[0], g[1], g[2], 0, R, ch[0], ch[1], ch[2], quotient_degree_factor=4)
    assert got.shape == (4, 1 << log_n) and (got == want).all()


def test_quotient_rejects_bad_arguments(ctx, qpzk):
    n = 16
    z = np.zeros((4, n), np.uint64)
    b = qpzk.PolynomialBatch.from_values(ctx, z, 3, None, 2)
    small = qpzk.PolynomialBatch.from_values(ctx, z[:, :8], 3, None, 2)
    Q = qpzk.PolynomialBatch.quotient_polys
    with pytest.raises(qpzk.QpzkError):
        Q(b, b, small, 0, 4, [1], [1], [1])                    # different degrees
    with pytest.raises(qpzk.QpzkError):
        Q(b, b, b, 2, 4, [1], [1], [1])                        # sigma columns out of range
    with pytest.raises(qpzk.QpzkError):
        Q(b, b, b, 0, 4, [1], [1], [1], quotient_degree_factor=6)
    with pytest.raises(qpzk.QpzkError):
        Q(b, b, b, 0, 4, [1], [1], [1], quotient_degree_factor=16)   # exceeds the blowup
    with pytest.raises(qpzk.QpzkError):
        Q(b, b, b, 0, 4, [P], [1], [1])                        # non-canonical challenge
    with pytest.raises(qpzk.QpzkError):
        Q(b, b, b, 0, 4, [1] * 5, [1] * 5, [1] * 5)            # more than 4 challenges


def test_gateless_plonk_proof_end_to_end(ctx, qpzk, orc):
    """wires -> partial products -> quotient -> openings -> FRI, everything on the device from the routed wires and
    the circuit's sigmas: bit-exact with the oracle, FRI accepted by the oracle verifier, and the vanishing identity
    holds on the OPENED values (what upstream's verifier checks) -- a complete proof of a circuit without gates."""
    from qpzk import prover
    from perm_util import satisfied_permutation, verify_gateless_openings
    log_n, R, num_consts = 8, 80, 5
    n = 1 << log_n
    shape = dict(prover.LEAF_SHAPE, log_n=log_n)
    rng = np.random.default_rng(123)
    routed, sigmas, _ = satisfied_permutation(log_n, R, seed=17)
    wires = np.concatenate([routed, rng.integers(0, P, size=(shape["wires"] - R, n), dtype=np.uint64)])
    cs = np.concatenate([rng.integers(0, P, size=(num_consts, n), dtype=np.uint64), sigmas])
    circuit = prover.SyntheticCircuit(ctx, cs, shape, sigma_col0=num_consts)
    pih = rng.integers(0, P, size=4, dtype=np.uint64)
    got = prover.prove_synthetic(circuit, wires, None, None, pih, sigmas=sigmas)
    ocs = orc.PolynomialBatch.from_values(cs, 3, 4)
    want = orc.prove_synthetic(ocs, wires, None, None, pih, log_n, sigmas=sigmas, sigma_col0=num_consts)
    for a, b in zip(got["caps"], want["caps"]):
        assert (a == b).all()
    assert (got["openings"] == want["openings"]).all()
    assert (got["fri_proof"] == want["fri_proof"]).all()
    assert orc.fri_verify(log_n, want["leaf_lens"], want["all_caps"], want["instance"], got["openings"], want["params"],
                          want["verifier_challenger"], got["fri_proof"]) == 0
    bg, alphas, zeta = got["challenges"]
    assert verify_gateless_openings(log_n, got["openings"], num_consts, R, shape["wires"], 2, 8, bg[:2], bg[2:], alphas, zeta)
    # a wire that breaks a copy constraint: the proof is still produced, and the vanishing check rejects it
    bad = wires.copy()
    bad[2, 5] = (int(bad[2, 5]) + 1) % P
    got_bad = prover.prove_synthetic(circuit, bad, None, None, pih, sigmas=sigmas)
    bg, alphas, zeta = got_bad["challenges"]
    assert not verify_gateless_openings(log_n, got_bad["openings"], num_consts, R, shape["wires"], 2, 8, bg[:2], bg[2:],
                                        alphas, zeta)


# ---- secret hygiene: zeroize modes ----------------------------------------------------------------------------------
def _peek(dev_ptr, words):
    from cuda import cudart
    out = np.zeros(words, np.uint64)
    err, = cudart.cudaMemcpy(out.ctypes.data, dev_ptr, 8 * words, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    assert int(err) == 0
    return out


def test_zeroize_modes(qpzk):
    """Mode 1 zeroes the LDE / coefficient buffers when the batch is freed; mode 2 leaves them in the context's private
    block cache and zeroes them before they go back to the driver (here: when the mode is switched back)."""
    c = qpzk.Context(0)
    cols = np.random.default_rng(3).integers(1, P, size=(5, 64), dtype=np.uint64)
    b = qpzk.PolynomialBatch.from_values(c, cols, 3, None, 2)
    leaves, coeffs = c.L.qpzk_batch_leaves_dev(b.h), c.L.qpzk_batch_coeffs_dev(b.h)
    assert _peek(leaves, 5 * 512).any() and _peek(coeffs, 5 * 64).any()
    b.free()
    c.synchronize()
    assert not _peek(leaves, 5 * 512).any() and not _peek(coeffs, 5 * 64).any()      # mode 1 (default): zero at release
    c.set_zeroize(2)
    b = qpzk.PolynomialBatch.from_values(c, cols, 3, None, 2)
    cap = b.cap.copy()
    leaves = c.L.qpzk_batch_leaves_dev(b.h)
    b.free()
    c.synchronize()
    assert _peek(leaves, 5 * 512).any()                                               # still in the private cache
    b2 = qpzk.PolynomialBatch.from_values(c, cols, 3, None, 2)                        # steady state: same block, same result
    assert (b2.cap == cap).all()
    b2_coeffs = b2.polynomials
    b2.free()
    c.set_zeroize(1)                                                                  # nothing stays deferred
    c.synchronize()
    assert not _peek(leaves, 5 * 512).any()
    # switching the mode while a batch lives in a block that carries deferred contents must not touch the live batch
    c.set_zeroize(2)
    qpzk.PolynomialBatch.from_values(c, cols, 3, None, 2).free()
    live = qpzk.PolynomialBatch.from_values(c, cols, 3, None, 2)                      # reuses the dirty blocks
    c.set_zeroize(1)
    c.synchronize()
    assert (live.cap == cap).all() and (live.polynomials == b2_coeffs).all()
    lp = c.L.qpzk_batch_leaves_dev(live.h)
    live.free()
    c.synchronize()
    assert not _peek(lp, 5 * 512).any()
    with pytest.raises(qpzk.QpzkError):
        c.set_zeroize(3)
    c.set_zeroize(2)
    qpzk.PolynomialBatch.from_values(c, cols, 3, None, 2).free()
    c.trim()                                                                          # scrubbed, then returned to the driver
    assert c.cached_bytes() == 0


def test_zs_partial_products_from_device_resident_batches(ctx, qpzk, orc):
    """keep_values with DEVICE inputs (from_values_dev copies the values next to the batch): the partial products
    computed from two such batches give the same commitment as the host path and as the oracle."""
    import torch
    from perm_util import satisfied_permutation
    log_n, R, num_consts = 7, 12, 3
    rng = np.random.default_rng(91)
    routed, sigmas, _ = satisfied_permutation(log_n, R, seed=91)
    wires = np.concatenate([routed, rng.integers(0, P, size=(4, 1 << log_n), dtype=np.uint64)])
    cs = np.concatenate([rng.integers(0, P, size=(num_consts, 1 << log_n), dtype=np.uint64), sigmas])
    betas, gammas = rng.integers(0, P, size=2, dtype=np.uint64), rng.integers(0, P, size=2, dtype=np.uint64)
    dw = torch.from_numpy(wires.view(np.int64)).cuda()
    dc = torch.from_numpy(cs.view(np.int64)).cuda()
    torch.cuda.synchronize()
    ctx.set_keep_values(True)
    try:
        wb = qpzk.PolynomialBatch.from_device(ctx, dw.data_ptr(), log_n, wires.shape[0], 3, 2)
        cb = qpzk.PolynomialBatch.from_device(ctx, dc.data_ptr(), log_n, cs.shape[0], 3, 2)
    finally:
        ctx.set_keep_values(False)
    assert ctx.L.qpzk_batch_values_dev(wb.h) and ctx.L.qpzk_batch_values_dev(wb.h) != dw.data_ptr()
    zb = qpzk.PolynomialBatch.commit_zs_partial_products(cb, wb, num_consts, R, betas, gammas, 2)
    want = orc.partial_products(routed, sigmas, betas, gammas, 8)
    assert (zb.cap == orc.PolynomialBatch.from_values(want, 3, 2).cap()).all()
    assert (zb.polynomials == orc.PolynomialBatch.from_values(want, 3, 2).coeffs()).all()
    plain = qpzk.PolynomialBatch.from_values(ctx, wires, 3, None, 2)       # without the flag a batch holds no values ...
    assert not ctx.L.qpzk_batch_values_dev(plain.h)
    with pytest.raises(qpzk.QpzkError):                                    # ... and cannot feed the permutation argument
        qpzk.PolynomialBatch.commit_zs_partial_products(cb, plain, num_consts, R, betas, gammas, 2)


# ---- gate constraints handed over as a program (qpzk_gate_program) ---------------------------------------------------
def test_gate_program_quotient_matches_oracle(ctx, qpzk, orc):
    from perm_util import arithmetic_circuit, arithmetic_program
    log_n = 7
    wires, consts, sigmas = arithmetic_circuit(log_n, 5)
    b = arithmetic_program()
    code, imm = b.arrays()
    prog = qpzk.GateProgram.make(code, imm, b.next)
    ch = np.random.default_rng(6).integers(0, P, size=(3, 2), dtype=np.uint64)
    zs = orc.partial_products(wires, sigmas, ch[0], ch[1], 8)
    cs = np.concatenate([consts, sigmas])
    g = [qpzk.PolynomialBatch.from_values(ctx, a, 3, None, 2) for a in (cs, wires, zs)]
    o = [orc.PolynomialBatch.from_values(a, 3, 2) for a in (cs, wires, zs)]
    gt = orc.gate_terms(o[0], o[1], code, imm)
    want = orc.quotient_polys(o[0], o[1], o[2], 3, 16, ch[0], ch[1], ch[2], gate_terms=gt)
    got = qpzk.PolynomialBatch.quotient_polys_prog(g[0], g[1], g[2], 3, 16, ch[0], ch[1], ch[2], prog)
    assert (got == want).all()
    none = qpzk.PolynomialBatch.quotient_polys_prog(g[0], g[1], g[2], 3, 16, ch[0], ch[1], ch[2], None)
    assert (none == orc.quotient_polys(o[0], o[1], o[2], 3, 16, ch[0], ch[1], ch[2])).all() and not (none == want).all()
    qb = qpzk.PolynomialBatch.commit_quotient_prog(g[0], g[1], g[2], 3, 16, ch[0], ch[1], ch[2], prog, 2)
    assert (qb.polynomials == want).all()
    # malformed programs are rejected at the boundary, with the instruction number in the message
    for bad, what in ((np.array([[11, 0, 0, 0]], np.uint32), "opcode"), (np.array([[0x10003, 0, 0, 0]], np.uint32), "opcode"),
                      (np.array([[0, 0, 0, 0], [8, 1, 0, 900]], np.uint32), "immediate"), (np.array([[0, 0, 0, 0], [0x0507, 1, 0, 0]], np.uint32), "not yet written"), (np.array([[0, 70, 0, 0]], np.uint32), "destination"),
                      (np.array([[0, 300, 0, 0]], np.uint32), "destination"),
                      (np.array([[0, 0, 99, 0]], np.uint32), "wire column"), (np.array([[1, 0, 99, 0]], np.uint32), "column"),
                      (np.array([[2, 0, 3, 0]], np.uint32), "immediate"), (np.array([[3, 0, 1, 2]], np.uint32), "not yet written"),
                      (np.array([[6, 0, 5, 0]], np.uint32), "not yet written")):
        with pytest.raises(qpzk.QpzkError) as e:
            qpzk.PolynomialBatch.quotient_polys_prog(g[0], g[1], g[2], 3, 16, ch[0], ch[1], ch[2], qpzk.GateProgram.make(bad, imm, 8))
        assert what in str(e.value)


def test_plonk_proof_with_gates_end_to_end(ctx, qpzk, orc):
    """Arithmetic gates chained by copy constraints, the gates handed to the device as a program: wires -> partial
    products -> quotient (permutation AND gate terms) -> openings -> FRI.  Bit-exact with the oracle, FRI accepted, and
    the plonk verifier's vanishing identity -- gate constraints interpreted on the OPENED values -- holds; it fails when a
    gate output is wrong."""
    from qpzk import prover
    from perm_util import arithmetic_circuit, arithmetic_program, verify_gateless_openings
    log_n = 8
    wires, consts, sigmas = arithmetic_circuit(log_n, 12)
    b = arithmetic_program()
    code, imm = b.arrays()
    prog = qpzk.GateProgram.make(code, imm, b.next)
    shape = dict(prover.LEAF_SHAPE, log_n=log_n, wires=16, constants_sigmas=19, zs_partial_products=4, quotient=16)
    cs = np.concatenate([consts, sigmas])
    circuit = prover.SyntheticCircuit(ctx, cs, shape, sigma_col0=3)
    pih = np.arange(4, dtype=np.uint64) + 100
    got = prover.prove_synthetic(circuit, wires, None, None, pih, sigmas=sigmas, gate_program=prog)
    ocs = orc.PolynomialBatch.from_values(cs, 3, 4)
    want = orc.prove_synthetic(ocs, wires, None, None, pih, log_n, sigmas=sigmas, sigma_col0=3, gate_program=(code, imm))
    for a, c in zip(got["caps"], want["caps"]):
        assert (a == c).all()
    assert (got["openings"] == want["openings"]).all() and (got["fri_proof"] == want["fri_proof"]).all()
    assert orc.fri_verify(log_n, want["leaf_lens"], want["all_caps"], want["instance"], got["openings"], want["params"],
                          want["verifier_challenger"], got["fri_proof"]) == 0
    bg, alphas, zeta = got["challenges"]
    assert verify_gateless_openings(log_n, got["openings"], 3, 16, 16, 2, 8, bg[:2], bg[2:], alphas, zeta, program=(code, imm))
    assert not verify_gateless_openings(log_n, got["openings"], 3, 16, 16, 2, 8, bg[:2], bg[2:], alphas, zeta)   # gates matter
    bad = wires.copy()
    bad[7, 9] = (int(bad[7, 9]) + 1) % P
    gb = prover.prove_synthetic(circuit, bad, None, None, pih, sigmas=sigmas, gate_program=prog)
    bg, alphas, zeta = gb["challenges"]
    assert not verify_gateless_openings(log_n, gb["openings"], 3, 16, 16, 2, 8, bg[:2], bg[2:], alphas, zeta, program=(code, imm))


def test_gate_program_with_a_large_register_file(ctx, qpzk, orc):
    """More than 64 registers selects the 256-register interpreter; same results as the oracle's interpreter."""
    from perm_util import arithmetic_circuit, ProgramBuilder
    log_n = 5
    wires, consts, sigmas = arithmetic_circuit(log_n, 2)
    b = ProgramBuilder()
    regs = [b.wire(j % 16) if j % 3 else b.const(j % 3) for j in range(120)]
    acc = b.immediate(12345)
    for j, r in enumerate(regs):
        acc = b.add(b.mul(acc, r), regs[(7 * j + 1) % 120]) if b.next < 250 else acc
        if j % 40 == 39:
            b.emit(acc)
    assert 64 < b.next <= 256
    code, imm = b.arrays()
    ch = np.random.default_rng(1).integers(0, P, size=(3, 1), dtype=np.uint64)
    zs = orc.partial_products(wires, sigmas, ch[0], ch[1], 8)
    cs = np.concatenate([consts, sigmas])
    g = [qpzk.PolynomialBatch.from_values(ctx, a, 3, None, 2) for a in (cs, wires, zs)]
    o = [orc.PolynomialBatch.from_values(a, 3, 2) for a in (cs, wires, zs)]
    want = orc.quotient_polys(o[0], o[1], o[2], 3, 16, ch[0], ch[1], ch[2], gate_terms=orc.gate_terms(o[0], o[1], code, imm))
    got = qpzk.PolynomialBatch.quotient_polys_prog(g[0], g[1], g[2], 3, 16, ch[0], ch[1], ch[2], qpzk.GateProgram.make(code, imm, b.next))
    assert (got == want).all()


def test_poseidon_hash_chain_proved_end_to_end(ctx, qpzk, orc):
    """The reference's shape in miniature: 135 wires of which 80 routed, one PoseidonGate per row (123 constraints of
    degree 7, recorded as a gate program of ~5300 instructions), rows chained by copy constraints on the digest words.
    Device path = oracle bit for bit, FRI accepted, and the vanishing identity -- the gate program interpreted over the
    opened extension-field values -- holds; a wrong S-box wire is rejected."""
    import gen_poseidon_constants as gen
    from qpzk import prover, gates
    from perm_util import poseidon_chain_circuit, verify_gateless_openings
    log_n = 6
    rc = gen.round_constants()
    rec = gates.Recorder()
    gates.poseidon_gate(rec, rc, gen.derive_fast_partial(rc), selector=rec.const(0))
    code, imm, regs = rec.finish()
    prog = qpzk.GateProgram.make(code, imm, regs)
    wires, consts, sigmas = poseidon_chain_circuit(log_n, 4, rc)
    cs = np.concatenate([consts, sigmas])
    shape = dict(prover.LEAF_SHAPE, log_n=log_n, constants_sigmas=81)
    circuit = prover.SyntheticCircuit(ctx, cs, shape, sigma_col0=1)
    pih = np.arange(4, dtype=np.uint64) + 7
    got = prover.prove_synthetic(circuit, wires, None, None, pih, sigmas=sigmas, gate_program=prog)
    ocs = orc.PolynomialBatch.from_values(cs, 3, 4)
    want = orc.prove_synthetic(ocs, wires, None, None, pih, log_n, sigmas=sigmas, sigma_col0=1, gate_program=(code, imm))
    for a, c in zip(got["caps"], want["caps"]):
        assert (a == c).all()
    assert (got["openings"] == want["openings"]).all() and (got["fri_proof"] == want["fri_proof"]).all()
    assert orc.fri_verify(log_n, want["leaf_lens"], want["all_caps"], want["instance"], got["openings"], want["params"],
                          want["verifier_challenger"], got["fri_proof"]) == 0
    bg, alphas, zeta = got["challenges"]
    assert verify_gateless_openings(log_n, got["openings"], 1, 80, 135, 2, 8, bg[:2], bg[2:], alphas, zeta, program=(code, imm))
    bad = wires.copy()
    bad[gates.wire_partial_sbox(7), 3] = (int(bad[gates.wire_partial_sbox(7), 3]) + 1) % P
    gb = prover.prove_synthetic(circuit, bad, None, None, pih, sigmas=sigmas, gate_program=prog)
    bg, alphas, zeta = gb["challenges"]
    assert not verify_gateless_openings(log_n, gb["openings"], 1, 80, 135, 2, 8, bg[:2], bg[2:], alphas, zeta, program=(code, imm))


def test_poseidon2_hash_chain_proved_end_to_end(ctx, qpzk, orc):
    """The gate of the reference's in-circuit hashing (Poseidon2, the permutation pinned to the reference's own
    vectors) as a gate program: a chain of Poseidon2 rows, 135 wires / 80 routed, proved on the device.  Device path =
    oracle bit for bit, FRI accepted, the vanishing identity holds on the opened values, a wrong S-box wire is rejected."""
    import gen_poseidon2_constants as gp
    from qpzk import prover, gates
    from perm_util import poseidon_chain_circuit, verify_gateless_openings
    log_n = 6
    ext, internal, diag = gp.parameters()
    flat = [c for row in ext for c in row]
    rec = gates.Recorder()
    gates.poseidon2_gate(rec, flat, internal, diag, selector=rec.const(0))
    code, imm, regs = rec.finish()
    prog = qpzk.GateProgram.make(code, imm, regs)
    wires, consts, sigmas = poseidon_chain_circuit(log_n, 14, None, witness=lambda i, s: gates.poseidon2_gate_witness(i, s, flat, internal, diag))
    cs = np.concatenate([consts, sigmas])
    shape = dict(prover.LEAF_SHAPE, log_n=log_n, constants_sigmas=81)
    circuit = prover.SyntheticCircuit(ctx, cs, shape, sigma_col0=1)
    pih = np.arange(4, dtype=np.uint64) + 11
    got = prover.prove_synthetic(circuit, wires, None, None, pih, sigmas=sigmas, gate_program=prog)
    ocs = orc.PolynomialBatch.from_values(cs, 3, 4)
    want = orc.prove_synthetic(ocs, wires, None, None, pih, log_n, sigmas=sigmas, sigma_col0=1, gate_program=(code, imm))
    for a, c in zip(got["caps"], want["caps"]):
        assert (a == c).all()
    assert (got["openings"] == want["openings"]).all() and (got["fri_proof"] == want["fri_proof"]).all()
    assert orc.fri_verify(log_n, want["leaf_lens"], want["all_caps"], want["instance"], got["openings"], want["params"],
                          want["verifier_challenger"], got["fri_proof"]) == 0
    bg, alphas, zeta = got["challenges"]
    assert verify_gateless_openings(log_n, got["openings"], 1, 80, 135, 2, 8, bg[:2], bg[2:], alphas, zeta, program=(code, imm))
    bad = wires.copy()
    bad[gates.wire_partial_sbox(9), 5] = (int(bad[gates.wire_partial_sbox(9), 5]) + 1) % P
    gb = prover.prove_synthetic(circuit, bad, None, None, pih, sigmas=sigmas, gate_program=prog)
    bg, alphas, zeta = gb["challenges"]
    assert not verify_gateless_openings(log_n, gb["openings"], 1, 80, 135, 2, 8, bg[:2], bg[2:], alphas, zeta, program=(code, imm))


def test_gate_program_compiled_equals_interpreted(ctx, qpzk, orc):
    """The PoseidonGate program (2904 steps) runs as run-time compiled straight-line code (gate_jit.cu: NVRTC, sm_100a);
    the same call with the compiler switched off runs the interpreter.  Both give the oracle's quotient, bit for bit, and
    the counter shows that a compiled program was really used."""
    import gen_poseidon_constants as gen
    from qpzk import gates
    from perm_util import poseidon_chain_circuit
    if not ctx.L.qpzk_gate_jit_available():
        pytest.skip("NVRTC not loadable on this box: every gate program is interpreted")
    log_n = 6
    rc = gen.round_constants()
    rec = gates.Recorder()
    gates.poseidon_gate(rec, rc, gen.derive_fast_partial(rc), selector=rec.const(0))
    code, imm, regs = rec.finish()
    prog = qpzk.GateProgram.make(code, imm, regs)
    wires, consts, sigmas = poseidon_chain_circuit(log_n, 4, rc)
    cs = np.concatenate([consts, sigmas])
    ch = np.random.default_rng(16).integers(0, P, size=(3, 2), dtype=np.uint64)
    zs = orc.partial_products(wires[:80], sigmas, ch[0], ch[1], 8)
    g = [qpzk.PolynomialBatch.from_values(ctx, a, 3, None, 4) for a in (cs, wires, zs)]
    o = [orc.PolynomialBatch.from_values(a, 3, 4) for a in (cs, wires, zs)]
    want = orc.quotient_polys(o[0], o[1], o[2], 1, 80, ch[0], ch[1], ch[2], gate_terms=orc.gate_terms(o[0], o[1], code, imm))
    try:
        ctx.L.qpzk_gate_jit_set_mode(1)
        compiled = qpzk.PolynomialBatch.quotient_polys_prog(g[0], g[1], g[2], 1, 80, ch[0], ch[1], ch[2], prog)
        assert ctx.L.qpzk_gate_jit_compiled() >= 1
        ctx.L.qpzk_gate_jit_set_mode(0)
        interpreted = qpzk.PolynomialBatch.quotient_polys_prog(g[0], g[1], g[2], 1, 80, ch[0], ch[1], ch[2], prog)
    finally:
        ctx.L.qpzk_gate_jit_set_mode(1)
    assert (compiled == want).all() and (interpreted == want).all()


def test_hash_no_pad_matches_oracle(ctx, qpzk, orc):
    rng = np.random.default_rng(17)
    for k in (0, 1, 4, 7, 8, 9, 16, 21, 135):
        x = rng.integers(0, P, size=k, dtype=np.uint64)
        assert (qpzk.hash_no_pad(ctx, x) == orc.hash_no_pad(x)).all(), k
    assert not qpzk.hash_no_pad(ctx, []).any()


def test_prove_to_bytes_and_verify_from_bytes(ctx, qpzk, orc):
    """prove() for a circuit given as data (Poseidon hash chain, gates as a program) returns serialized bytes; an
    independent verifier parses them, re-derives every challenge from the transcript, checks the vanishing identity on
    the opened values and the FRI proof.  Any flipped byte of the proof is rejected."""
    import gen_poseidon_constants as gen
    from qpzk import prover, gates
    from perm_util import poseidon_chain_circuit, verify_proof_bytes
    log_n = 5
    rc = gen.round_constants()
    rec = gates.Recorder()
    gates.poseidon_gate(rec, rc, gen.derive_fast_partial(rc), selector=rec.const(0))
    code, imm, regs = rec.finish()
    wires, consts, sigmas = poseidon_chain_circuit(log_n, 21, rc)
    cs = np.concatenate([consts, sigmas])
    circuit = prover.SyntheticCircuit(ctx, cs, dict(prover.LEAF_SHAPE, log_n=log_n, constants_sigmas=81), sigma_col0=1)
    public_inputs = wires[12:16, (1 << log_n) - 3].copy()           # the last digest of the chain
    data, shape = prover.prove(circuit, wires, public_inputs, 80, qpzk.GateProgram.make(code, imm, regs))
    assert len(data) == qpzk.load_library().qpzk_proof_bytes_len(circuit.fri_params, shape)
    po = orc.fri_params(log_n, 3, 4, 16, 28, 4, 5)
    args = (circuit.fri_params, po, shape, circuit.constants_sigmas.cap, circuit.circuit_digest, 80)
    assert verify_proof_bytes(orc, qpzk, data, *args, program=(code, imm))
    assert not verify_proof_bytes(orc, qpzk, data, *args)                       # a verifier without the gates rejects
    rng = np.random.default_rng(0)
    for pos in [0, 600, len(data) // 2, len(data) - 40, len(data) - 9] + [int(x) for x in rng.integers(0, len(data) - 48, 6)]:
        bad = bytearray(data)
        bad[pos] ^= 1
        try:
            ok = verify_proof_bytes(orc, qpzk, bytes(bad), *args, program=(code, imm))
        except qpzk.QpzkError:
            ok = False                                                           # not even well formed any more
        assert not ok, pos
