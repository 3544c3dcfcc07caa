#!/usr/bin/env python3
"""Try to reproduce the reference's Poseidon2 known-answer tests with publicly documented parameter sets.

The application hash of the reference (`Poseidon2Hash::hash_no_pad`, qp-plonky2 / qp-poseidon-core =3.1.0 -- neither
crate is vendored under /root/reference, Cargo.toml:20-21) is pinned by vectors the reference's own tests hold:
  * 5 secret -> address pairs: address = H(H(felts("wormhole") || secret_felts))
    (wormhole/tests/src/circuit/unspendable_account_tests.rs:9-24; construction wormhole/circuit/src/unspendable_account.rs:63-95;
     padding `x || 1 || 0*` to the rate 8: wormhole/circuit/tests/heap_zeroization.rs:132-163;
     felt encodings: common/src/serialization.rs:126-141 (4 bytes/felt + 0x01 terminator), :236-243 (8 bytes/felt LE))
  * 2 block-hash vectors (wormhole/tests/test-helpers/src/lib.rs:208-219) -- their preimage needs the header encoding,
    so the address vectors are the ones searched here.
What is NOT in the tree: the permutation's round constants and matrices.  This script restates the Poseidon2 permutation
(width 12, x^7, 4 + 22 + 4 rounds, external matrix circ(2 M4, M4, M4), internal matrix 1 + diag) generically and tries
every candidate constant derivation listed in CANDIDATES; it prints which (if any) reproduces all five vectors.
Run:  python tools/poseidon2_search.py        (pure Python, a few seconds per candidate)
"""
import itertools
import struct
import sys

sys.path.insert(0, __import__("os").path.dirname(__file__))
from gen_poseidon_constants import chacha_block, seed_from_u64  # noqa: E402

P = 0xFFFFFFFF00000001
M64 = (1 << 64) - 1

SECRETS = ["4c8587bd422e01d961acdc75e7d66f6761b7af7c9b1864a492f369c9d6724f05",
           "c6034553e5556630d24a593d2c92de9f1ede81d48f0fb3371764462cc3594b3f",
           "87f5fc11df0d12f332ccfeb92ddd8995e6c11709501a8b59c2aaf9eefee63ec1",
           "ef69da4e3aa2a6f15b3a9eec5e481f17260ac812faf1e685e450713327c3ab1c",
           "9aa84f99ef2de22e3070394176868df41d6a148117a36132d010529e19b018b7"]
ADDRESSES = ["de68c6fcb3e38d6736b79a010e4504b98c6321f1e4d11cd8484f67c187ca090e",
             "2dcf4b944c13da31a748ed04a251557d9bfb8eed7c4a8af5593c59d6142642b7",
             "e84689fa523459215ac1c5a930a7898ace511c7e201bdb7295b04fc87037988f",
             "2a96bf97fb63467396ddd298ac931eb22add9a527d02764520040e5e32a31ad6",
             "c6a587f78e270025b7e9517da47e07c7f614c21be9a1460b5214eee7ef33ac68"]

# internal diagonal of the Goldilocks width-12 instance as published with the Poseidon2 reference implementation
# (Horizen Labs `poseidon2_instance_goldilocks.rs`, MAT_DIAG12_M_1; Plonky3 `MATRIX_DIAG_12_GOLDILOCKS`) -- as recalled
DIAG12_HL = [0xc3b6c08e23ba9300, 0xd84b5de94a324fb6, 0x0d0c371c5b35b84f, 0x7964f570e7188037,
             0x5daf18bbd996604b, 0x6743bc47b9595257, 0x5528b9362c59bb70, 0xac45e25b7127b68b,
             0xa2077d7dfbb606b5, 0xf3faac6faee378ae, 0x0c6388b51545e883, 0xd27dbb6944917b60]
M4_P3 = [[2, 3, 1, 1], [1, 2, 3, 1], [1, 1, 2, 3], [3, 1, 1, 2]]          # Plonky3 MDSMat4
M4_HL = [[5, 7, 1, 3], [4, 6, 1, 1], [1, 3, 5, 7], [1, 1, 4, 6]]          # Poseidon2 paper / Horizen Labs


class ChaChaRng:
    """rand_chacha ChaCha{8,12,20}Rng::seed_from_u64 (PCG32 seed expansion, 64-bit block counter, words in order)."""

    def __init__(self, seed_u64, rounds):
        self.key = struct.unpack("<8I", seed_from_u64(seed_u64))
        self.counter, self.buf, self.rounds = 0, [], rounds

    def next_u32(self):
        if not self.buf:
            self.buf = chacha_block(self.key, self.counter, self.rounds)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self):
        lo = self.next_u32()
        return lo | (self.next_u32() << 32)


class Xoroshiro128Plus:
    """rand_xoshiro::Xoroshiro128Plus::seed_from_u64 (SplitMix64 seeding), as Plonky3's tests use."""

    def __init__(self, seed):
        def sm():
            nonlocal seed
            seed = (seed + 0x9E3779B97F4A7C15) & M64
            z = seed
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
            return z ^ (z >> 31)
        self.s0, self.s1 = sm(), sm()

    def next_u64(self):
        s0, s1 = self.s0, self.s1
        r = (s0 + s1) & M64
        s1 ^= s0
        self.s0 = (((s0 << 24) | (s0 >> 40)) & M64) ^ s1 ^ ((s1 << 16) & M64)
        self.s1 = ((s1 << 37) | (s1 >> 27)) & M64
        return r


def draw_felts(rng, count, reject=True):
    out = []
    while len(out) < count:
        v = rng.next_u64()
        if reject and v >= P:          # Plonky3: Distribution<Goldilocks> for Standard loops until the draw is canonical
            continue
        out.append(v % P)
    return out


def permute(state, ext_rc, int_rc, diag, m4):
    s = list(state)

    def ext():
        nonlocal s
        t = []
        for c in range(3):
            x = s[4 * c:4 * c + 4]
            t += [sum(m4[r][k] * x[k] for k in range(4)) % P for r in range(4)]
        sums = [(t[k] + t[4 + k] + t[8 + k]) % P for k in range(4)]
        s = [(t[i] + sums[i % 4]) % P for i in range(12)]

    ext()
    for r in range(4):
        s = [pow((s[i] + ext_rc[r][i]) % P, 7, P) for i in range(12)]
        ext()
    for r in range(22):
        s[0] = pow((s[0] + int_rc[r]) % P, 7, P)
        tot = sum(s) % P
        s = [(s[i] * diag[i] + tot) % P for i in range(12)]
    for r in range(4, 8):
        s = [pow((s[i] + ext_rc[r][i]) % P, 7, P) for i in range(12)]
        ext()
    return s


def hash_no_pad_p2(felts, perm):
    """x || 1 || 0* to a multiple of 8, absorbed 8 at a time (overwrite == add for a single block from the zero state;
    both are tried for longer inputs by the caller), first 4 words out."""
    x = list(felts) + [1]
    x += [0] * (-len(x) % 8)
    st = [0] * 12
    for off in range(0, len(x), 8):
        st[:8] = x[off:off + 8]
        st = perm(st)
    return st[:4]


def bytes_to_felts(b, order):
    b = bytes(b) + b"\x01"
    b += b"\x00" * (-len(b) % 4)
    return [int.from_bytes(b[i:i + 4], order) for i in range(0, len(b), 4)]


def digest_felts(b):
    return [int.from_bytes(b[i:i + 8], "little") % P for i in range(0, 32, 8)]


def candidates():
    rngs = []
    for seed in (0x3141592653589793, 0, 1, 42, 0x0123456789ABCDEF):
        for rounds in (20, 8, 12):
            rngs.append(("ChaCha%dRng::seed_from_u64(%#x)" % (rounds, seed), lambda s=seed, r=rounds: ChaChaRng(s, r)))
        rngs.append(("Xoroshiro128Plus::seed_from_u64(%#x)" % seed, lambda s=seed: Xoroshiro128Plus(s)))
    for (name, mk), reject, (m4n, m4) in itertools.product(rngs, (True, False), (("p3 MDSMat4", M4_P3), ("HL M4", M4_HL))):
        rng = mk()
        flat = draw_felts(rng, 96, reject)                       # Plonky3 new_from_rng_128: 4 + 4 external rounds x 12 ...
        ext_rc = [flat[12 * r:12 * r + 12] for r in range(8)]
        int_rc = draw_felts(rng, 22, reject)                      # ... then the 22 internal constants
        yield "%s, %s, %s" % (name, "rejection sampling" if reject else "mod p", m4n), ext_rc, int_rc, DIAG12_HL, m4


def main():
    found = []
    tried = 0
    for desc, ext_rc, int_rc, diag, m4 in candidates():
        perm = lambda st: permute(st, ext_rc, int_rc, diag, m4)
        for order in ("little", "big"):
            tried += 1
            ok = 0
            for sec, addr in zip(SECRETS, ADDRESSES):
                pre = bytes_to_felts(b"wormhole", order) + digest_felts(bytes.fromhex(sec))
                h = hash_no_pad_p2(hash_no_pad_p2(pre, perm), perm)
                if h == digest_felts(bytes.fromhex(addr)):
                    ok += 1
                else:
                    break
            if ok:
                found.append((desc, order, ok))
                print("MATCH %d/5: %s, salt bytes %s-endian" % (ok, desc, order))
    print("%d parameter candidates tried, %d reproduce the reference's address vectors" % (tried, len(found)))
    return 0 if found else 1


if __name__ == "__main__":
    sys.exit(main())
