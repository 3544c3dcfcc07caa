#!/usr/bin/env python3
"""Derive the 360 Poseidon-12 round constants of plonky2's PoseidonGoldilocks hasher.

The constants are not vendored in the reference tree (the hasher lives in the
crates.io dependency qp-plonky2 =1.5.5, Cargo.toml:20).  Upstream generated them as
`ChaCha8Rng::seed_from_u64(0)` followed by 360 draws of `gen_range(0..p)`
(SURVEY.md Appendix A.3).  This script restates that recipe (rand_core's
seed_from_u64 PCG32 expander, ChaCha with 8 rounds, rand's widening-multiply
uniform sampler), checks the self-checks listed in SURVEY.md A.3, and writes the
table as a C header for the oracle and as a CUDA header for the product.

Usage: python tools/gen_poseidon_constants.py   (rewrites both headers in place)
"""
import os, struct, sys

P = 0xFFFFFFFF00000001
M64 = (1 << 64) - 1
M32 = (1 << 32) - 1


def seed_from_u64(state):
    """rand_core::SeedableRng::seed_from_u64 — PCG32 stream filling a 32-byte seed."""
    out = b""
    for _ in range(8):
        state = (state * 6364136223846793005 + 11634580027462260723) & M64
        xorshifted = (((state >> 18) ^ state) >> 27) & M32
        rot = state >> 59
        x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & M32
        out += struct.pack("<I", x)
    return out


def rotl(v, n):
    return ((v << n) | (v >> (32 - n))) & M32


def chacha_block(key_words, counter, rounds=8):
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [
        counter & M32, (counter >> 32) & M32, 0, 0]
    w = st[:]

    def qr(a, b, c, d):
        w[a] = (w[a] + w[b]) & M32; w[d] = rotl(w[d] ^ w[a], 16)
        w[c] = (w[c] + w[d]) & M32; w[b] = rotl(w[b] ^ w[c], 12)
        w[a] = (w[a] + w[b]) & M32; w[d] = rotl(w[d] ^ w[a], 8)
        w[c] = (w[c] + w[d]) & M32; w[b] = rotl(w[b] ^ w[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(w[i] + st[i]) & M32 for i in range(16)]


class ChaCha8Rng:
    def __init__(self, seed_u64):
        self.key = struct.unpack("<8I", seed_from_u64(seed_u64))
        self.counter = 0
        self.buf = []

    def next_u32(self):
        if not self.buf:
            self.buf = chacha_block(self.key, self.counter)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self):
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)

    def gen_range_u64(self, high):
        """rand 0.8 UniformInt::sample_single for u64, range [0, high)."""
        rng_range = high
        zone = (rng_range << (rng_range.bit_length() and (64 - rng_range.bit_length()))) - 1 & M64
        while True:
            v = self.next_u64()
            m = v * rng_range
            if (m & M64) <= zone:
                return m >> 64


def round_constants():
    rng = ChaCha8Rng(0)
    return [rng.gen_range_u64(P) for _ in range(360)]


def self_check(rc):
    assert rc[0:4] == [0xB585F766F2144405, 0x7746A55F43921AD7, 0xB2FB0D31CEE799B4, 0x0F6760A4803427D7], [hex(x) for x in rc[:4]]
    assert rc[359] == 0xBC8DFB627FE558FC
    x = 0
    s = 0
    for c in rc:
        x ^= c
        s = (s + c) & M64
    assert x == 0xD95D3C3BB2FE42E3, hex(x)
    assert s == 0x695B4F31AC3DC147, hex(s)
    assert max(rc) < 0xFFFEEAC900011537


CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
DIAG = [8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
N_FULL_HALF, N_PARTIAL, WIDTH = 4, 22, 12


def mds_matrix():
    """out[r] = sum_i CIRC[i]*s[(i+r)%12] + DIAG[r]*s[r]  =>  M[r][c] = CIRC[(c-r)%12] + [r==c]*DIAG[r]."""
    return [[(CIRC[(c - r) % 12] + (DIAG[r] if r == c else 0)) % P for c in range(12)] for r in range(12)]


def matmul(a, b):
    n, m, k = len(a), len(b[0]), len(b)
    return [[sum(a[i][t] * b[t][j] for t in range(k)) % P for j in range(m)] for i in range(n)]


def matvec(a, x):
    return [sum(a[i][t] * x[t] for t in range(len(x))) % P for i in range(len(a))]


def matinv(a):
    n = len(a)
    m = [row[:] + [1 if i == j else 0 for j in range(n)] for i, row in enumerate(a)]
    for c in range(n):
        piv = next(r for r in range(c, n) if m[r][c] % P)
        m[c], m[piv] = m[piv], m[c]
        inv = pow(m[c][c], P - 2, P)
        m[c] = [x * inv % P for x in m[c]]
        for r in range(n):
            if r != c and m[r][c]:
                f = m[r][c]
                m[r] = [(x - f * y) % P for x, y in zip(m[r], m[c])]
    return [row[n:] for row in m]


def poseidon_naive(state, rc):
    """SURVEY.md A.3 verbatim: add constants, S-box (all lanes / lane 0), MDS; 4 + 22 + 4 rounds."""
    m = mds_matrix()
    s = list(state)
    for r in range(30):
        s = [(x + c) % P for x, c in zip(s, rc[12 * r:12 * r + 12])]
        if r < 4 or r >= 26:
            s = [pow(x, 7, P) for x in s]
        else:
            s[0] = pow(s[0], 7, P)
        s = matvec(m, s)
    return s


def derive_fast_partial(rc, m_sparse=N_PARTIAL):
    """Sparse re-factoring of the 22 partial rounds (Poseidon paper, App. B), derived here from
    the MDS matrix and round constants alone; algebraically identical to the naive rounds.

    Computed form:   x[1..] <- INIT * x[1..]
                     for r in 0..21:  x0 <- (x0 + K[r])^7
                                      x0' = 25*x0 + sum_i VHAT[r][i]*x[1+i];  x[1+i] += W[r][i]*x0;  x0 <- x0'
                     x <- x + POST   (folded into the constants of full round 26)
    """
    m = mds_matrix()
    first = N_PARTIAL - m_sparse   # partial rounds [0, first) stay in the literal form
    # 1. push lane>=1 constants forward so that only lane 0 receives a scalar per round
    ks, delta = [], [0] * 12
    for r in range(first, N_PARTIAL):
        c = rc[12 * (4 + r):12 * (4 + r) + 12]
        ks.append((c[0] + delta[0]) % P)
        u = matvec(m, [0] + [(delta[i] + c[i]) % P for i in range(1, 12)])
        delta = u
    post = delta
    # 2. sparse factorisation, last round first:  T = [[t00, v],[w, B]] = [[t00, v*B^-1],[w, I]] * diag(1, B)
    w_l, vhat_l = [None] * m_sparse, [None] * m_sparse
    t = m
    init = [[1 if i == j else 0 for j in range(11)] for i in range(11)]
    for r in reversed(range(m_sparse)):
        b = [row[1:] for row in t[1:]]
        binv = matinv(b)
        v = [t[0][1:]]
        vhat_l[r] = matmul(v, binv)[0]
        w_l[r] = [t[i][0] for i in range(1, 12)]
        assert t[0][0] == 25
        nprime = [[1] + [0] * 11] + [[0] + row for row in b]
        t = matmul(nprime, m)
        init = b
    return ks, init, w_l, vhat_l, post


def poseidon_fast(state, rc, fast):
    ks, init, w_l, vhat_l, post = fast
    m = mds_matrix()
    s = list(state)
    for r in range(4):
        s = [(x + c) % P for x, c in zip(s, rc[12 * r:12 * r + 12])]
        s = [pow(x, 7, P) for x in s]
        s = matvec(m, s)
    for r in range(4, 4 + N_PARTIAL - len(ks)):   # literal partial rounds
        s = [(x + c) % P for x, c in zip(s, rc[12 * r:12 * r + 12])]
        s[0] = pow(s[0], 7, P)
        s = matvec(m, s)
    s = [s[0]] + matvec(init, s[1:])
    for r in range(len(ks)):
        x0 = pow((s[0] + ks[r]) % P, 7, P)
        d = (25 * x0 + sum(a * b for a, b in zip(vhat_l[r], s[1:]))) % P
        s = [d] + [(s[1 + i] + w_l[r][i] * x0) % P for i in range(11)]
    s = [(x + c) % P for x, c in zip(s, post)]
    for r in range(26, 30):
        s = [(x + c) % P for x, c in zip(s, rc[12 * r:12 * r + 12])]
        s = [pow(x, 7, P) for x in s]
        s = matvec(m, s)
    return s


KATS = [
    ([0] * 12, "3c18a9786cb0b359 c4055e3364a246c3 7953db0ab48808f4 c71603f33a1144ca d7709673896996dc 46a84e87642f44ed d032648251ee0b3c 1c687363b207df62 df8565563e8045fe 40f5b37ff4254dae d070f637b431067c 1792b1c4342109d7"),
    (list(range(12)), "d64e1e3efc5b8e9e 53666633020aaa47 d40285597c6a8825 613a4f81e81231d2 414754bfebd051f0 cb1f8980294a023f 6eb2a9e4d54a9d0f 1902bc3af467e056 f045d5eafdc6021f e4150f77caaa3be5 c9bfd01d39b50cce 5c0a27fcb0e1459b"),
    ([P - 1] * 12, "be0085cfc57a8357 d95af71847d05c09 cf55a13d33c1c953 95803a74f4530e82 fcd99eb30a135df1 e095905e913a3029 de0392461b42919b 7d3260e24e81d031 10d3d0465d9deaa0 a87571083dfc2a47 e18263681e9958f8 e28e96f1ae5e60d3"),
]


def check_permutation(rc, fast):
    import random
    for inp, want in KATS:
        want = [int(w, 16) for w in want.split()]
        assert poseidon_naive(inp, rc) == want, "naive KAT mismatch"
        assert poseidon_fast(inp, rc, fast) == want, "fast KAT mismatch"
        for m_sparse in (0, 1, 7, 11, 21):
            assert poseidon_fast(inp, rc, derive_fast_partial(rc, m_sparse)) == want, "split KAT mismatch"
    rnd = random.Random(7)
    for _ in range(8):
        x = [rnd.randrange(P) for _ in range(12)]
        assert poseidon_naive(x, rc) == poseidon_fast(x, rc, fast)


def emit_fast(f, prefix, fast, qual):
    ks, init, w_l, vhat_l, post = fast
    def arr(name, vals, dims):
        f.write("%s %s%s%s = {\n" % (qual, prefix, name, dims))
        for i in range(0, len(vals), 4):
            f.write("  " + ", ".join("0x%016xULL" % v for v in vals[i:i + 4]) + ",\n")
        f.write("};\n")
    arr("PARTIAL_K", ks, "[22]")
    arr("PARTIAL_INIT", [x for row in init for x in row], "[121]")
    arr("PARTIAL_W", [x for row in w_l for x in row], "[242]")
    arr("PARTIAL_VHAT", [x for row in vhat_l for x in row], "[242]")
    arr("PARTIAL_POST", post, "[12]")


def emit_device_split(f, rc, m_sparse):
    """Tables for the CUDA permutation: the first 22-m partial rounds literal, the last m sparse."""
    ks, init, w_l, vhat_l, post = derive_fast_partial(rc, m_sparse)
    f.write("// Device permutation schedule: %d literal + %d sparse partial rounds.\n" % (N_PARTIAL - m_sparse, m_sparse))
    f.write("#define QPZK_SPARSE_ROUNDS %d\n" % m_sparse)
    def arr(name, vals, n):
        f.write("static const uint64_t QPZK_POSEIDON_DEV_%s[%d] = {\n" % (name, max(n, 1)))
        vals = list(vals) or [0]
        for i in range(0, len(vals), 4):
            f.write("  " + ", ".join("0x%016xULL" % v for v in vals[i:i + 4]) + ",\n")
        f.write("};\n")
    arr("K", ks, m_sparse)
    arr("INIT", [x for row in init for x in row], 121)
    arr("W", [x for row in w_l for x in row], 11 * m_sparse)
    arr("VHAT", [x for row in vhat_l for x in row], 11 * m_sparse)
    # constants added to the state after the sparse segment: POST + RC[round 26]
    arr("V26", [(a + b) % P for a, b in zip(post, rc[12 * 26:12 * 26 + 12])], 12)


def emit(path, rc, guard, decl, fast=None, prefix="", qual="static const uint64_t", device_split=None):
    lines = [
        "// GENERATED by tools/gen_poseidon_constants.py -- do not edit.",
        "// Poseidon-12 (Goldilocks) round constants: ChaCha8Rng::seed_from_u64(0), 360 x gen_range(0..p)",
        "// (SURVEY.md Appendix A.3; hasher of qp-plonky2 =1.5.5, reference Cargo.toml:20).",
        "#ifndef %s" % guard, "#define %s" % guard,
        "%s[360] = {" % decl]
    for r in range(30):
        row = ", ".join("0x%016xULL" % c for c in rc[12 * r:12 * r + 12])
        lines.append("  " + row + ",")
    lines += ["};", ""]
    with open(path, "w") as f:
        f.write("\n".join(lines))
        if fast is not None:
            f.write("// Sparse partial-round factorisation (derive_fast_partial in the generator).\n")
            emit_fast(f, prefix, fast, qual)
        if device_split is not None:
            emit_device_split(f, rc, device_split)
        f.write("#endif\n")


if __name__ == "__main__":
    rc = round_constants()
    self_check(rc)
    fast = derive_fast_partial(rc)
    check_permutation(rc, fast)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    emit(os.path.join(root, "oracle", "poseidon_rc.h"), rc, "ORACLE_POSEIDON_RC_H",
         "static const uint64_t POSEIDON_RC", fast, "POSEIDON_")
    emit(os.path.join(root, "qp-zk-circuits_b200", "csrc", "poseidon_rc.cuh"), rc, "QPZK_POSEIDON_RC_CUH",
         "static const uint64_t QPZK_POSEIDON_RC_HOST", None, "QPZK_POSEIDON_",
         device_split=int(os.environ.get("QPZK_SPARSE_ROUNDS", "22")))
    print("ok: 360 constants, self-checks pass; naive and sparse forms match the 3 permutation KATs")
