#!/usr/bin/env python3
"""Derive the Poseidon2 (Goldilocks, width 12) parameters of the reference's application hash and write them as headers.

`Poseidon2Hash` (qp-plonky2 =1.5.5 / qp-poseidon-core =3.1.0, reference Cargo.toml:20-21) is not vendored under
/root/reference, but the reference's tests hold seven vectors that pin it (tools/poseidon2_search.py found the one
public parameter derivation that reproduces all of them):
  round constants  Plonky3 `Poseidon2::new_from_rng_128` driven by `ChaCha20Rng::seed_from_u64(0x3141592653589793)`:
                   4 + 4 external rounds x 12 constants, then 22 internal constants, each the next 64-bit draw that
                   is < p (rand_core PCG32 seed expansion; ChaCha with 20 rounds, 64-bit counter, words in order)
  external layer   circ(2 M4, M4, M4) with M4 = [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]], also applied once before round 0
  internal layer   s <- diag * s + sum(s), diag = the published Goldilocks width-12 diagonal
  S-box x^7; hash: additive sponge, rate 8, capacity 4, padding x || 1 || 0*, output 4 words
Checks (before anything is written):
  5 secret -> address pairs   wormhole/tests/src/circuit/unspendable_account_tests.rs:9-24
                              (H(H(felts("wormhole") || secret)), wormhole/circuit/src/unspendable_account.rs:63-95)
  2 block hashes              wormhole/tests/test-helpers/src/lib.rs:208-219 (preimage: wormhole/circuit/src/block_header/header.rs:132-141)
Writes oracle/poseidon2_rc.h, qp-zk-circuits_b200/csrc/poseidon2_rc.cuh and tests/golden/poseidon2_kats.json.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import poseidon2_search as ps  # noqa: E402

P = ps.P
SEED = 0x3141592653589793


def parameters():
    rng = ps.ChaChaRng(SEED, 20)
    flat = ps.draw_felts(rng, 96)
    return [flat[12 * r:12 * r + 12] for r in range(8)], ps.draw_felts(rng, 22), list(ps.DIAG12_HL)


def permute(state, ext, internal, diag):
    return ps.permute(state, ext, internal, diag, ps.M4_P3)


def hash_no_pad(felts, ext, internal, diag):
    x = list(felts) + [1]
    x += [0] * (-len(x) % 8)
    st = [0] * 12
    for off in range(0, len(x), 8):
        st = permute([(st[i] + x[off + i]) % P for i in range(8)] + st[8:], ext, internal, diag)
    return st[:4]


BLOCK_HEADERS = [   # wormhole/tests/test-helpers/src/lib.rs:208-273: (parent hash, number, state root, expected block hash)
    (bytes(32), 1,
     bytes.fromhex("7d5f043e068be9691efbc3c1d498788b5dc5c7d65f41c0e24e2211c3997c0811"),
     bytes([41, 121, 200, 114, 64, 33, 135, 21, 222, 235, 173, 167, 111, 217, 87, 68, 46, 165, 205, 243, 3, 7, 81, 12, 41, 10,
            36, 214, 38, 141, 199, 160])),
    (bytes([160, 247, 232, 22, 150, 117, 245, 140, 3, 70, 175, 175, 22, 247, 90, 37, 231, 80, 170, 11, 27, 183, 40, 51, 5, 19,
            164, 19, 188, 192, 229, 212]), 2,
     bytes.fromhex("d897231aa700e15b496af8a43da156cfafed9c644982e0de72a19bdbe8f873f5"),
     bytes([38, 25, 3, 20, 38, 119, 31, 217, 10, 35, 145, 19, 235, 69, 79, 98, 99, 92, 90, 146, 245, 234, 228, 68, 125, 205, 20,
            92, 178, 28, 153, 61])),
]
DIGEST_LOGS = bytes([8, 6, 112, 111, 119, 95, 128, 233, 182, 183, 107, 158, 1, 115, 19, 219, 126, 253, 86, 30, 208, 176, 70, 21,
                     45, 180, 229, 9, 62, 91, 4, 6, 53, 245, 52, 48, 38, 123, 225, 5, 112, 111, 119, 95, 1, 1] + [0] * 61 + [18, 79, 226])


def kats(ext, internal, diag):
    """[(name, preimage felts, expected 4 felts)] -- every reference-held Poseidon2 vector."""
    out = []
    for k, (sec, addr) in enumerate(zip(ps.SECRETS, ps.ADDRESSES)):
        pre = ps.bytes_to_felts(b"wormhole", "little") + ps.digest_felts(bytes.fromhex(sec))
        inner = hash_no_pad(pre, ext, internal, diag)
        out.append(("unspendable_account_%d_inner" % k, pre, inner))
        out.append(("unspendable_account_%d" % k, inner, ps.digest_felts(bytes.fromhex(addr))))
    for k, (parent, number, state_root, want) in enumerate(BLOCK_HEADERS):
        pre = (ps.digest_felts(parent) + [number] + ps.digest_felts(state_root) + ps.digest_felts(bytes(32)) +
               ps.digest_felts(bytes(32)) + ps.bytes_to_felts(DIGEST_LOGS, "little"))
        out.append(("block_hash_%d" % k, pre, ps.digest_felts(want)))
    return out


def c_array(name, vals, per_line=4, ctype="uint64_t", qual="static const"):
    lines = ["%s %s %s[%d] = {" % (qual, ctype, name, len(vals))]
    for i in range(0, len(vals), per_line):
        lines.append("    " + ", ".join("0x%016xULL" % v for v in vals[i:i + per_line]) + ",")
    lines.append("};")
    return "\n".join(lines)


def main():
    ext, internal, diag = parameters()
    vectors = kats(ext, internal, diag)
    for name, pre, want in vectors:
        got = hash_no_pad(pre, ext, internal, diag)
        assert got == want, "Poseidon2 vector %s not reproduced" % name
    head = ("Poseidon2 (Goldilocks, width 12) parameters of the reference's `Poseidon2Hash` -- generated by\n"
            " * tools/gen_poseidon2_constants.py (Plonky3 new_from_rng_128 on ChaCha20Rng::seed_from_u64(0x3141592653589793);\n"
            " * pinned by the reference's 5 address vectors and 2 block-hash vectors, checked before this file was written).\n"
            " * EXT: 8 external rounds x 12;  INT: 22 internal rounds;  DIAG: internal matrix = diag + all-ones.")
    flat = [v for row in ext for v in row]
    with open(os.path.join(HERE, "..", "oracle", "poseidon2_rc.h"), "w") as f:
        f.write("/* ORACLE (test infrastructure): %s */\n#ifndef ORACLE_POSEIDON2_RC_H\n#define ORACLE_POSEIDON2_RC_H\n#include <stdint.h>\n" % head)
        f.write(c_array("POSEIDON2_EXT_RC", flat) + "\n" + c_array("POSEIDON2_INT_RC", internal) + "\n" + c_array("POSEIDON2_DIAG", diag) + "\n#endif\n")
    with open(os.path.join(HERE, "..", "qp-zk-circuits_b200", "csrc", "poseidon2_rc.cuh"), "w") as f:
        f.write("// %s\n#pragma once\n#include <cstdint>\n" % head.replace("\n *", "\n//"))
        f.write(c_array("QPZK_POSEIDON2_EXT_RC_HOST", flat, qual="static const") + "\n" +
                c_array("QPZK_POSEIDON2_INT_RC_HOST", internal) + "\n" + c_array("QPZK_POSEIDON2_DIAG_HOST", diag) + "\n")
    golden = {"what": "reference-held Poseidon2 vectors (see tools/gen_poseidon2_constants.py for the sources, file:line) and the "
                      "intermediate inner hashes; `in` = preimage felts, `out` = Poseidon2Hash::hash_no_pad(in)",
              "seed": hex(SEED), "vectors": [{"name": n, "in": [str(v) for v in pre], "out": [str(v) for v in want]} for n, pre, want in vectors],
              "permutation_of_0_to_11": [str(v) for v in permute(list(range(12)), ext, internal, diag)]}
    with open(os.path.join(HERE, "..", "tests", "golden", "poseidon2_kats.json"), "w") as f:
        json.dump(golden, f, indent=1)
    print("Poseidon2: %d vectors reproduced; headers and tests/golden/poseidon2_kats.json written" % len(vectors))


if __name__ == "__main__":
    main()
