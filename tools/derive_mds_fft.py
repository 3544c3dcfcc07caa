#!/usr/bin/env python3
"""Derivation + self-check of the add/shift-only form of the Poseidon-12 MDS layer used by the CUDA kernels
(qp-zk-circuits_b200/csrc/poseidon.cuh, `mds_limb`).

The layer is  y[r] = sum_i CIRC[i] * s[(i + r) mod 12] + DIAG[r] * s[r]   (SURVEY.md A.3).  The circulant part is a
product in Z[x]/(x^12 - 1):  Y = S * K  with  K(x) = sum_i CIRC[i] x^(-i).  With z = x^3 (z^4 = 1) the ring splits,
by a 4-point DFT in z over the three stride-3 subsequences, into
    Z[x]/(x^3 - 1)  x  Z[i][x]/(x^3 - i)  x  Z[x]/(x^3 + 1)      (the 4th component is the conjugate of the 2nd),
and the transform of K turns out to be made of tiny integers:
    k^[0]/4 = [16, 32, 16]     k^[2]/4 = [-1, -8, 2]     k^[1]/2 = [(2,1), (-4,-1), (16,-1)]
so the whole layer is additions and shifts on integers.  It is applied to 22/22/20-bit limbs of the state lanes
in wrapping 32-bit arithmetic: every operation is a ring operation mod 2^32 and the true outputs are < 2^31, so
the wrapped results are exact.

This script (1) derives the constants from CIRC, (2) checks the limb formulas (exactly the expression order of the
CUDA code, mod 2^32) against the literal matrix-vector product on random states.  Run: python tools/derive_mds_fft.py
"""
import random

CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
DIAG = [8] + [0] * 11
P = 0xFFFFFFFF00000001
M32 = 0xFFFFFFFF


def derive():
    K = [0] * 12
    for i in range(12):
        K[(12 - i) % 12] += CIRC[i]
    g, h, q = [], [], []
    for j in range(3):
        k = [K[j + 3 * a] for a in range(4)]
        k0, k2 = sum(k), k[0] - k[1] + k[2] - k[3]
        k1 = (k[0] - k[2], k[1] - k[3])
        assert k0 % 4 == 0 and k2 % 4 == 0 and k1[0] % 2 == 0 and k1[1] % 2 == 0
        g.append(k0 // 4); h.append(k2 // 4); q.append((k1[0] // 2, k1[1] // 2))
    return g, h, q


def mds_literal(s):
    return [sum(CIRC[i] * s[(i + r) % 12] for i in range(12)) + DIAG[r] * s[r] for r in range(12)]


def mds_limb(e, rc):
    """12 limbs in, 12 limbs out, wrapping 32-bit arithmetic; mirrors `mds_limb` in poseidon.cuh."""
    w = lambda x: x & M32
    A, B, Cr, Ci = [0] * 3, [0] * 3, [0] * 3, [0] * 3
    for j in range(3):
        t, u = w(e[j] + e[j + 6]), w(e[j + 3] + e[j + 9])
        A[j], B[j] = w(t + u), w(t - u)
        Cr[j], Ci[j] = w(e[j] - e[j + 6]), w(e[j + 3] - e[j + 9])
    T = w(A[0] + A[1] + A[2])
    TA = [w(T + A[2]), w(T + A[0]), w(T + A[1])]
    P2 = [w(8 * B[2] - B[0] - 2 * B[1]), w(-8 * B[0] - B[1] - 2 * B[2]), w(2 * B[0] - 8 * B[1] - B[2])]
    (r0, r1, r2), (i0, i1, i2) = Cr, Ci
    P1r = [w(2 * r0 - i0 - 16 * i1 + r1 + 4 * i2 + r2),
           w(-4 * r0 + i0 + 2 * r1 - i1 + r2 - 16 * i2),
           w(16 * r0 + i0 - 4 * r1 + i1 + 2 * r2 - i2)]
    P1i = [w(2 * i0 + r0 + 16 * r1 + i1 - 4 * r2 + i2),
           w(-4 * i0 - r0 + 2 * i1 + r1 + 16 * r2 + i2),
           w(16 * i0 - r0 - 4 * i1 - r1 + 2 * i2 + r2)]
    y = [0] * 12
    for j in range(3):
        U, V = w(16 * TA[j] + P2[j]), w(16 * TA[j] - P2[j])
        y[j] = w(U + P1r[j] + rc[j])
        y[j + 3] = w(V + P1i[j] + rc[j + 3])
        y[j + 6] = w(U - P1r[j] + rc[j + 6])
        y[j + 9] = w(V - P1i[j] + rc[j + 9])
    y[0] = w(y[0] + 8 * e[0])
    return y


def limb_combine(y0, y1, y2):
    """mirrors `limb_combine` in poseidon.cuh: y0 + 2^22 y1 + 2^44 y2 - 4096 as a u64 representative mod p"""
    h = y2 >> 20
    t0 = y0 + 4096 - h
    t1 = y1 + (h << 10)
    assert 0 <= t0 < 2**32 and t1 < 2**32
    a = t0 + (t1 << 22)
    a0, a1, l12 = a & M32, a >> 32, (y2 << 12) & M32
    s1 = a1 + l12
    c = s1 >> 32
    s1 &= M32
    s0 = a0 - c
    b = 1 if s0 < 0 else 0
    s0 &= M32
    s1 = (s1 + c - b) & M32
    return (s1 << 32) | s0


def split(x):
    return [x & 0x3FFFFF, (x >> 22) & 0x3FFFFF, x >> 44]


def main():
    g, h, q = derive()
    print("k^[0]/4 =", g, " k^[2]/4 =", h, " k^[1]/2 =", q)
    assert g == [16, 32, 16] and h == [-1, -8, 2] and q == [(2, 1), (-4, -1), (16, -1)]
    rnd = random.Random(1)
    for it in range(2000):
        if it == 0:
            s = [2**64 - 1] * 12
            rc = [P - 1] * 12
        else:
            s = [rnd.getrandbits(64) for _ in range(12)]
            rc = [rnd.randrange(P) for _ in range(12)]
        want = [(v + c) for v, c in zip(mds_literal(s), rc)]
        limbs = [split(x) for x in s]
        rcl = [split(c) for c in rc]
        rcl4 = [split((c - 4096) % P) for c in rc]     # what the device table holds
        ys4 = [mds_limb([limbs[i][l] for i in range(12)], [rcl4[i][l] for i in range(12)]) for l in range(3)]
        for i in range(12):
            got = limb_combine(ys4[0][i], ys4[1][i], ys4[2][i])
            assert got < 2**64 and got % P == want[i] % P, (it, i)
        ys = [mds_limb([limbs[i][l] for i in range(12)], [rcl[i][l] for i in range(12)]) for l in range(3)]
        for i in range(12):
            assert all(ys[l][i] < 2**31 for l in range(3))
            got = ys[0][i] + (ys[1][i] << 22) + (ys[2][i] << 44)
            assert got == want[i], (it, i)
    print("limb MDS == literal MDS on 2000 random states (plus the all-ones state): ok; limb_combine ok")


if __name__ == "__main__":
    main()
