#!/usr/bin/env python3
"""Writes upstream's PoseidonGate, recorded as a qpzk_gate_program (qpzk/gates.py), to a binary file the native driver
(host/prove_bench.cpp --gates FILE) loads:  u32 magic 'QPGP', u32 num_ops, u32 num_imm, u32 num_regs, code, immediates.

    python tools/export_gate_program.py [out [poseidon|poseidon2]]      (default: both, into qp-zk-circuits_b200/host/)
"""
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main(out, which="poseidon"):
    from qpzk import gates
    rec = gates.Recorder()
    if which == "poseidon2":
        import gen_poseidon2_constants as gp
        ext, internal, diag = gp.parameters()
        gates.poseidon2_gate(rec, [c for row in ext for c in row], internal, diag, selector=rec.const(0))
    else:
        import gen_poseidon_constants as gen
        rc = gen.round_constants()
        gates.poseidon_gate(rec, rc, gen.derive_fast_partial(rc), selector=rec.const(0))
    code, imm, regs = rec.finish()
    with open(out, "wb") as f:
        f.write(struct.pack("<4sIII", b"QPGP", code.shape[0], imm.size, regs))
        f.write(code.tobytes())
        f.write(imm.tobytes())
    return code.shape[0], imm.size, regs


if __name__ == "__main__":
    if len(sys.argv) > 1:
        jobs = [(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "poseidon")]
    else:   # both gates of the reference: PoseidonGate (recursion / Merkle hashing) and Poseidon2Gate (application hashing)
        jobs = [(os.path.join(ROOT, "qp-zk-circuits_b200", "host", "poseidon_gate.qpgp"), "poseidon"),
                (os.path.join(ROOT, "qp-zk-circuits_b200", "host", "poseidon2_gate.qpgp"), "poseidon2")]
    for path, which in jobs:
        print("wrote %s: %d instructions, %d immediates, %d registers" % ((path,) + main(path, which)))
