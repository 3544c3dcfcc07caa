#!/usr/bin/env python3
"""Writes upstream's PoseidonGate, recorded as a qpzk_gate_program (qpzk/gates.py), to a binary file the native driver
(host/prove_bench.cpp --gates FILE) loads:  u32 magic 'QPGP', u32 num_ops, u32 num_imm, u32 num_regs, code, immediates.

    python tools/export_gate_program.py [out]      (default: qp-zk-circuits_b200/host/poseidon_gate.qpgp)
"""
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main(out):
    import gen_poseidon_constants as gen
    from qpzk import gates
    rc = gen.round_constants()
    rec = gates.Recorder()
    gates.poseidon_gate(rec, rc, gen.derive_fast_partial(rc), selector=rec.const(0))
    code, imm, regs = rec.finish()
    with open(out, "wb") as f:
        f.write(struct.pack("<4sIII", b"QPGP", code.shape[0], imm.size, regs))
        f.write(code.tobytes())
        f.write(imm.tobytes())
    return code.shape[0], imm.size, regs


if __name__ == "__main__":
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "qp-zk-circuits_b200", "host", "poseidon_gate.qpgp")
    print("wrote %s: %d instructions, %d immediates, %d registers" % ((path,) + main(path)))
