"""Gate constraints recorded as `qpzk_gate_program`s (include/qpzk.h): the host-side half of SURVEY.md 8a row a17.

The reference's gate set lives in the un-vendored crate qp-plonky2 =1.5.5 (`gates/*.rs`); a fork records it by running
its constraint code over a recording field type.  This module is that recorder in Python, plus two
gates written from upstream plonky2 as recalled [upstream-memory] -- `ArithmeticGate` and `PoseidonGate`, the gate that
fixes the reference's 135-wire configuration (reference common/src/circuit.rs:428-431) -- so that the device interpreter
is exercised and measured on a program of the real shape: 135 wires, 123 constraints of degree 7.

    rec = Recorder();  poseidon_gate(rec, round_constants, fast_partial, selector=rec.const(0));  prog = rec.finish()
"""
import numpy as np

from . import GateProgram, P

WIRE, CONST, IMM, ADD, SUB, MUL, EMIT, MULADD, MULI, ADDI, MULIADD = range(11)


class Recorder:
    """Records field operations on virtual registers; finish() maps them onto at most `max_regs` physical registers
    (linear scan over the straight-line program: a register is free again after the last read of its value)."""

    def __init__(self):
        self.ops, self.imms, self.n = [], [], 0

    def _new(self):
        self.n += 1
        return self.n - 1

    def wire(self, j):
        d = self._new(); self.ops.append((WIRE, d, j, 0)); return d

    def const(self, j):
        d = self._new(); self.ops.append((CONST, d, j, 0)); return d

    def imm(self, v):
        v = int(v) % P
        d = self._new(); self.imms.append(v); self.ops.append((IMM, d, len(self.imms) - 1, 0)); return d

    def add(self, a, b):
        d = self._new(); self.ops.append((ADD, d, a, b)); return d

    def sub(self, a, b):
        d = self._new(); self.ops.append((SUB, d, a, b)); return d

    def mul(self, a, b):
        d = self._new(); self.ops.append((MUL, d, a, b)); return d

    def emit(self, a):
        self.ops.append((EMIT, 0, a, 0))

    @staticmethod
    def _sources(o):
        """register operands of an internal op tuple (op, dst, a, b, c)"""
        op, _, a, b, c = o
        if op in (ADD, SUB, MUL): return (a, b)
        if op == MULADD: return (a, b, c)
        if op in (MULI, ADDI): return (a,)
        if op == MULIADD: return (a, c)
        if op == EMIT: return (a,)
        return ()

    def _peephole(self):
        """Immediate operands and multiply-adds: MUL/ADD with an immediate operand become MULI/ADDI, an ADD whose
        operand is a single-use MUL/MULI becomes MULADD/MULIADD; loads and products that lose their last use are dropped.
        The ops are in SSA form (every value has its own virtual register), so moving a product to the place of its sum
        is always valid."""
        ops = [(op, d, a, b, 0) for op, d, a, b in self.ops]
        for _ in range(2):
            defs = {o[1]: o for o in ops if o[0] != EMIT}
            for i, o in enumerate(ops):
                op, d, a, b, c = o
                if op in (MUL, ADD):
                    if defs[b][0] == IMM:
                        ops[i] = (MULI if op == MUL else ADDI, d, a, defs[b][2], 0)
                    elif defs[a][0] == IMM:
                        ops[i] = (MULI if op == MUL else ADDI, d, b, defs[a][2], 0)
            uses = {}
            for o in ops:
                for r in self._sources(o):
                    uses[r] = uses.get(r, 0) + 1
            defs = {o[1]: o for o in ops if o[0] != EMIT}
            for i, o in enumerate(ops):
                op, d, a, b, c = o
                if op != ADD:
                    continue
                for x, y in ((a, b), (b, a)):
                    m = defs[x]
                    if m[0] in (MUL, MULI) and uses.get(x, 0) == 1:
                        ops[i] = (MULADD, d, m[2], m[3], y) if m[0] == MUL else (MULIADD, d, m[2], m[3], y)
                        uses[x] = 0
                        break
            while True:                                  # drop values nobody reads any more
                uses = {}
                for o in ops:
                    for r in self._sources(o):
                        uses[r] = uses.get(r, 0) + 1
                keep = [o for o in ops if o[0] == EMIT or uses.get(o[1], 0) > 0]
                if len(keep) == len(ops):
                    break
                ops = keep
        return ops

    def finish(self, max_regs=256, fuse=True):
        """(code (num_ops, 4) uint32, imm uint64, num_regs) with physical registers."""
        ops = self._peephole() if fuse else [(op, d, a, b, 0) for op, d, a, b in self.ops]
        last = {}
        for i, o in enumerate(ops):
            for r in self._sources(o):
                last[r] = i
        free, phys, used, code = [], {}, 0, []
        for i, o in enumerate(ops):
            op, d, a, b, c = o
            src = self._sources(o)
            pa = phys[a] if a in src else a
            pb = phys[b] if b in src and op in (ADD, SUB, MUL, MULADD) else b
            pc = phys[c] if op in (MULADD, MULIADD) else 0
            for v in set(src):                         # released BEFORE the destination is chosen (dst may reuse them)
                if last.get(v) == i:
                    free.append(phys.pop(v))
            if op == EMIT:
                code.append((EMIT, 0, pa, 0))
                continue
            if free:
                pd = free.pop()
            else:
                pd = used; used += 1
                if used > max_regs:
                    raise ValueError("gate program needs more than %d registers" % max_regs)
            if d in last:
                phys[d] = pd
            else:
                free.append(pd)                        # value never read: the slot is free again at once
            code.append((op | (pc << 8), pd, pa, pb))
        return np.array(code, dtype=np.uint32).reshape(-1, 4), np.array(self.imms, dtype=np.uint64), max(used, 1)

    def program(self, max_regs=256):
        code, imm, regs = self.finish(max_regs)
        return GateProgram.make(code, imm, regs)


def arithmetic_gate(rec, num_ops, selector=None, const_base=1):
    """upstream ArithmeticGate: wires [multiplicand_0, multiplicand_1, addend, output] x num_ops, constants c0, c1:
    constraint_k = output_k - (c0 * m0_k * m1_k + c1 * addend_k)   (times the selector filter)."""
    c0, c1 = rec.const(const_base), rec.const(const_base + 1)
    for k in range(num_ops):
        m0, m1, ad, out = (rec.wire(4 * k + t) for t in range(4))
        c = rec.sub(out, rec.add(rec.mul(rec.mul(m0, m1), c0), rec.mul(ad, c1)))
        rec.emit(rec.mul(c, selector) if selector is not None else c)


# ---- selectors and the sum over gates [upstream-memory: plonky2/src/gates/gate.rs, plonk/vanishing_poly.rs] ---------------
UNUSED_SELECTOR = 0xFFFFFFFF


def selector_filter(rec, s, row, group, many_selectors=False):
    """upstream `compute_filter(row, group_range, s, many_selector)`: the product over the other gate indices i of the
    selector group of (i - s), times (UNUSED_SELECTOR - s) when the circuit has several selector polynomials.  It is
    non-zero exactly on the rows whose selector value is `row`."""
    f = None
    for i in list(group) + ([UNUSED_SELECTOR] if many_selectors else []):
        if i == row:
            continue
        t = rec.sub(rec.imm(i), s)
        f = t if f is None else rec.mul(f, t)
    return f if f is not None else rec.imm(1)


class _Collector:
    """Stands in for the recorder while ONE gate is recorded: its constraints are kept, not emitted."""

    def __init__(self, rec):
        self._rec, self.constraints = rec, []

    def __getattr__(self, name):
        return getattr(self._rec, name)

    def emit(self, a):
        self.constraints.append(a)


def evaluate_gate_constraints(rec, gate_list):
    """upstream `evaluate_gate_constraints`: constraint slot i of the circuit is the sum over the gates of
    filter_g * constraint_{g,i}.  gate_list: [(record_fn(recorder), filter register), ...]; emits max_g(#constraints)
    terms."""
    slots = []
    for record, filt in gate_list:
        col = _Collector(rec)
        record(col)
        for i, c in enumerate(col.constraints):
            t = rec.mul(c, filt)
            if i < len(slots):
                slots[i] = rec.add(slots[i], t)
            else:
                slots.append(t)
    for t in slots:
        rec.emit(t)
    return len(slots)


def constant_gate(rec, num_consts, const_base):
    """upstream ConstantGate: wire_i - constant_i."""
    for i in range(num_consts):
        rec.emit(rec.sub(rec.const(const_base + i), rec.wire(i)))


def public_input_gate(rec, public_inputs_hash):
    """upstream PublicInputGate: wires 0..3 carry the hash of the public inputs (immediates of THIS proof's program)."""
    for i in range(4):
        rec.emit(rec.sub(rec.wire(i), rec.imm(public_inputs_hash[i])))


# ---- PoseidonGate [upstream-memory: plonky2/src/gates/poseidon.rs] ------------------------------------------------------
WIRE_SWAP, START_DELTA, START_FULL_0, START_PARTIAL, START_FULL_1, POSEIDON_GATE_WIRES = 24, 25, 29, 65, 87, 135
MDS_CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]


def wire_input(i): return i
def wire_output(i): return 12 + i
def wire_delta(i): return START_DELTA + i
def wire_full_sbox_0(r, i): return START_FULL_0 + 12 * (r - 1) + i     # rounds 1..3 (round 0 reads the state itself)
def wire_partial_sbox(r): return START_PARTIAL + r
def wire_full_sbox_1(r, i): return START_FULL_1 + 12 * r + i


def poseidon_gate(rec, rc, fast, selector=None):
    """123 constraints: swap is a bit (1), delta_i = swap * (rhs_i - lhs_i) (4), the S-box inputs of full rounds 1-3
    (36), of the 22 partial rounds (22) and of the last 4 full rounds (48) equal the running state, and the outputs (12).
    rc: the 360 round constants; fast = (K, INIT, W, VHAT, POST), the sparse form of the partial rounds
    (tools/gen_poseidon_constants.py: derive_fast_partial) -- the S-box inputs are the same field elements as in the
    literal rounds, so the witness does not depend on this choice."""
    ks, init, w_l, vhat_l, post = fast
    out = (lambda c: rec.emit(rec.mul(c, selector))) if selector is not None else rec.emit
    consts = {}

    def k(v):
        # the small MDS entries are used in every round and stay in registers; a round constant or a coefficient of
        # the sparse matrices is used once and is loaded where it is needed (a cached one would pin a register)
        v = int(v) % P
        if v >= 64:
            return rec.imm(v)
        if v not in consts:
            consts[v] = rec.imm(v)
        return consts[v]

    def sbox(x):
        x2 = rec.mul(x, x); x4 = rec.mul(x2, x2); x3 = rec.mul(x, x2)
        return rec.mul(x3, x4)

    def mds(s):
        res = []
        for r in range(12):
            acc = None
            for i in range(12):
                c = MDS_CIRC[i] + (8 if (r == 0 and i == 0) else 0)
                t = rec.mul(s[(i + r) % 12], k(c))
                acc = t if acc is None else rec.add(acc, t)
            res.append(acc)
        return res

    swap = rec.wire(WIRE_SWAP)
    out(rec.mul(swap, rec.sub(swap, k(1))))
    state = [None] * 12
    for i in range(4):
        lhs, rhs, delta = rec.wire(wire_input(i)), rec.wire(wire_input(i + 4)), rec.wire(wire_delta(i))
        out(rec.sub(rec.mul(swap, rec.sub(rhs, lhs)), delta))
        state[i], state[i + 4] = rec.add(lhs, delta), rec.sub(rhs, delta)
    for i in range(8, 12):
        state[i] = rec.wire(wire_input(i))
    for r in range(4):                                                   # first full rounds
        state = [rec.add(state[i], k(rc[12 * r + i])) for i in range(12)]
        if r:
            for i in range(12):
                sin = rec.wire(wire_full_sbox_0(r, i))
                out(rec.sub(state[i], sin))
                state[i] = sin
        state = mds([sbox(x) for x in state])
    rest = []                                                            # x[1..] <- INIT * x[1..]
    for i in range(11):
        acc = None
        for j in range(11):
            if init[i][j]:
                t = rec.mul(state[1 + j], k(init[i][j]))
                acc = t if acc is None else rec.add(acc, t)
        rest.append(acc)
    state = [state[0]] + rest
    for r in range(22):                                                  # partial rounds, sparse form
        sin = rec.wire(wire_partial_sbox(r))
        out(rec.sub(rec.add(state[0], k(ks[r])), sin))
        x0 = sbox(sin)
        d = rec.mul(x0, k(25))
        for i in range(11):
            d = rec.add(d, rec.mul(state[1 + i], k(vhat_l[r][i])))
        state = [d] + [rec.add(state[1 + i], rec.mul(x0, k(w_l[r][i]))) for i in range(11)]
    state = [rec.add(state[i], k(post[i])) for i in range(12)]
    for r in range(4):                                                   # last full rounds
        state = [rec.add(state[i], k(rc[12 * (26 + r) + i])) for i in range(12)]
        for i in range(12):
            sin = rec.wire(wire_full_sbox_1(r, i))
            out(rec.sub(state[i], sin))
            state[i] = sin
        state = mds([sbox(x) for x in state])
    for i in range(12):
        out(rec.sub(state[i], rec.wire(wire_output(i))))


def poseidon_gate_witness(inputs, swap, rc):
    """The 135 wire values of one PoseidonGate row for `inputs` (12 field elements) and a swap bit: literal rounds."""
    w = [0] * POSEIDON_GATE_WIRES
    inputs = [int(x) % P for x in inputs]
    for i in range(12):
        w[wire_input(i)] = inputs[i]
    w[WIRE_SWAP] = int(swap)
    s = list(inputs)
    for i in range(4):
        d = int(swap) * (inputs[i + 4] - inputs[i]) % P
        w[wire_delta(i)] = d
        s[i], s[i + 4] = (inputs[i] + d) % P, (inputs[i + 4] - d) % P

    def mds(v):
        return [sum((MDS_CIRC[i] + (8 if (r == 0 and i == 0) else 0)) * v[(i + r) % 12] for i in range(12)) % P for r in range(12)]

    for r in range(30):
        s = [(x + c) % P for x, c in zip(s, rc[12 * r:12 * r + 12])]
        if r < 4 or r >= 26:
            if 1 <= r < 4:
                for i in range(12):
                    w[wire_full_sbox_0(r, i)] = s[i]
            if r >= 26:
                for i in range(12):
                    w[wire_full_sbox_1(r - 26, i)] = s[i]
            s = [pow(x, 7, P) for x in s]
        else:
            w[wire_partial_sbox(r - 4)] = s[0]
            s[0] = pow(s[0], 7, P)
        s = mds(s)
    for i in range(12):
        w[wire_output(i)] = s[i]
    return w


# ---- Poseidon2Gate ------------------------------------------------------------------------------------------------------
# The reference hashes inside its circuits with `builder.hash_n_to_hash_no_pad_p2::<Poseidon2Hash>` (wormhole/circuit/src/
# nullifier.rs:298-299, unspendable_account.rs:229-231, zk_merkle_proof.rs:504,606, block_header/mod.rs:66): the gate behind
# it lives in the un-vendored qp-plonky2 fork, so its wire layout is not in the reference tree.  What IS pinned is the
# permutation (tests/golden/poseidon2_kats.json: the reference's 5 address and 2 block-hash vectors).  This recorder lays
# the gate out the way upstream lays out PoseidonGate -- inputs, outputs, swap flag, deltas, one wire per S-box input so
# that every constraint has degree <= 7 -- over the pinned permutation: same 135 wires, same 123 constraints.  A fork
# records its own gate with its own layout; the device only ever sees the program.
def _p2_external_layer(rec, s):
    """circ(2 M4, M4, M4) with M4 = [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]]: additions only."""
    t = [None] * 12
    for c in range(3):
        a, b, d, e = s[4 * c:4 * c + 4]
        tot = rec.add(rec.add(a, b), rec.add(d, e))
        t[4 * c + 0] = rec.add(rec.add(tot, a), rec.add(b, b))
        t[4 * c + 1] = rec.add(rec.add(tot, b), rec.add(d, d))
        t[4 * c + 2] = rec.add(rec.add(tot, d), rec.add(e, e))
        t[4 * c + 3] = rec.add(rec.add(tot, e), rec.add(a, a))
    out = [None] * 12
    for k in range(4):
        col = rec.add(rec.add(t[k], t[4 + k]), t[8 + k])
        for c in range(3):
            out[4 * c + k] = rec.add(t[4 * c + k], col)
    return out


def poseidon2_gate(rec, ext_rc, int_rc, diag, selector=None):
    """123 constraints over the PoseidonGate wire layout: swap is a bit (1), the four deltas (4), the S-box inputs of
    external rounds 1-3 (36), of the 22 internal rounds (22) and of the last 4 external rounds (48), the outputs (12).
    ext_rc: 96 external round constants, int_rc: 22 internal ones, diag: the 12 diagonal entries of the internal matrix
    (tools/gen_poseidon2_constants.py: parameters())."""
    out = (lambda c: rec.emit(rec.mul(c, selector))) if selector is not None else rec.emit

    def sbox(x):
        x2 = rec.mul(x, x); x4 = rec.mul(x2, x2); x3 = rec.mul(x, x2)
        return rec.mul(x3, x4)

    swap = rec.wire(WIRE_SWAP)
    out(rec.mul(swap, rec.sub(swap, rec.imm(1))))
    state = [None] * 12
    for i in range(4):
        lhs, rhs, delta = rec.wire(wire_input(i)), rec.wire(wire_input(i + 4)), rec.wire(wire_delta(i))
        out(rec.sub(rec.mul(swap, rec.sub(rhs, lhs)), delta))
        state[i], state[i + 4] = rec.add(lhs, delta), rec.sub(rhs, delta)
    for i in range(8, 12):
        state[i] = rec.wire(wire_input(i))
    state = _p2_external_layer(rec, state)
    for r in range(4):
        state = [rec.add(state[i], rec.imm(int(ext_rc[12 * r + i]) % P)) for i in range(12)]
        if r:
            for i in range(12):
                sin = rec.wire(wire_full_sbox_0(r, i))
                out(rec.sub(state[i], sin))
                state[i] = sin
        state = _p2_external_layer(rec, [sbox(x) for x in state])
    for r in range(22):
        sin = rec.wire(wire_partial_sbox(r))
        out(rec.sub(rec.add(state[0], rec.imm(int(int_rc[r]) % P)), sin))
        state[0] = sbox(sin)
        tot = state[0]
        for i in range(1, 12):
            tot = rec.add(tot, state[i])
        state = [rec.add(rec.mul(state[i], rec.imm(int(diag[i]) % P)), tot) for i in range(12)]
    for r in range(4):
        state = [rec.add(state[i], rec.imm(int(ext_rc[12 * (4 + r) + i]) % P)) for i in range(12)]
        for i in range(12):
            sin = rec.wire(wire_full_sbox_1(r, i))
            out(rec.sub(state[i], sin))
            state[i] = sin
        state = _p2_external_layer(rec, [sbox(x) for x in state])
    for i in range(12):
        out(rec.sub(state[i], rec.wire(wire_output(i))))


def poseidon2_gate_witness(inputs, swap, ext_rc, int_rc, diag):
    """The 135 wire values of one Poseidon2Gate row; the output wires are the permutation of the (swapped) inputs."""
    w = [0] * POSEIDON_GATE_WIRES
    inputs = [int(x) % P for x in inputs]
    for i in range(12):
        w[wire_input(i)] = inputs[i]
    w[WIRE_SWAP] = int(swap)
    s = list(inputs)
    for i in range(4):
        d = int(swap) * (inputs[i + 4] - inputs[i]) % P
        w[wire_delta(i)] = d
        s[i], s[i + 4] = (inputs[i] + d) % P, (inputs[i + 4] - d) % P

    def ext(v):
        t = [0] * 12
        for c in range(3):
            a, b, d, e = v[4 * c:4 * c + 4]
            tot = a + b + d + e
            t[4 * c:4 * c + 4] = [tot + a + 2 * b, tot + b + 2 * d, tot + d + 2 * e, tot + e + 2 * a]
        return [(t[4 * c + k] + t[k] + t[4 + k] + t[8 + k]) % P for c in range(3) for k in range(4)]

    s = ext(s)
    for r in range(4):
        s = [(x + int(c)) % P for x, c in zip(s, ext_rc[12 * r:12 * r + 12])]
        if r:
            for i in range(12):
                w[wire_full_sbox_0(r, i)] = s[i]
        s = ext([pow(x, 7, P) for x in s])
    for r in range(22):
        s[0] = (s[0] + int(int_rc[r])) % P
        w[wire_partial_sbox(r)] = s[0]
        s[0] = pow(s[0], 7, P)
        tot = sum(s) % P
        s = [(x * int(dg) + tot) % P for x, dg in zip(s, diag)]
    for r in range(4):
        s = [(x + int(c)) % P for x, c in zip(s, ext_rc[12 * (4 + r):12 * (4 + r) + 12])]
        for i in range(12):
            w[wire_full_sbox_1(r, i)] = s[i]
        s = ext([pow(x, 7, P) for x in s])
    for i in range(12):
        w[wire_output(i)] = s[i]
    return w
