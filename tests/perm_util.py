"""Helpers for the permutation-argument tests: a random copy-constraint system whose wires satisfy it."""
import numpy as np

P = 0xFFFFFFFF00000001
G = 14293326489335486720
W32 = 7277203076849721926


def root_of_unity(log_n):
    w = W32
    for _ in range(32 - log_n):
        w = w * w % P
    return w


def satisfied_permutation(log_n, num_routed, seed):
    """(wires, sigmas, ids): ids[j][i] = k_j * w^i with k_j = g^j; sigmas = ids permuted by a random permutation of all
    positions; wires constant on the cycles of that permutation (so the grand product of the argument is 1)."""
    rng = np.random.default_rng(seed)
    n = 1 << log_n
    w = root_of_unity(log_n)
    xs = [pow(w, i, P) for i in range(n)]
    ids = np.array([[pow(G, j, P) * x % P for x in xs] for j in range(num_routed)], dtype=np.uint64)
    perm = rng.permutation(num_routed * n)
    sigmas = ids.reshape(-1)[perm].reshape(num_routed, n)
    vals = np.zeros(num_routed * n, dtype=np.uint64)
    seen = np.zeros(num_routed * n, bool)
    for s0 in range(num_routed * n):
        if not seen[s0]:
            v = int(rng.integers(0, P, dtype=np.uint64))
            k = s0
            while not seen[k]:
                seen[k] = True
                vals[k] = v
                k = perm[k]
    return vals.reshape(num_routed, n), sigmas, ids


def horner(coeffs, z):
    acc = 0
    for c in coeffs[::-1]:
        acc = (acc * z + int(c)) % P
    return acc


def quotient_identity_holds(log_n, wires_coeffs, sigma_coeffs, zs_pp_coeffs, quotient_chunks, betas, gammas, alphas,
                            chunk, zeta, gate_term_values=()):
    """The verifier's check at a base-field point zeta (upstream plonk/verifier.rs + eval_vanishing_poly):
        sum_t terms[t](zeta) alpha_a^t  ==  Z_H(zeta) * sum_k zeta^(k n) chunk_{a,k}(zeta)      for every challenge a,
    terms = [L_0 (Z_c - 1)]_c ++ [partial-product checks]_c ++ gate terms.  All polynomials are given by coefficients
    and evaluated with big-integer Horner, independently of any NTT."""
    n = 1 << log_n
    R, nch = len(wires_coeffs), len(betas)
    nchunks = (R + chunk - 1) // chunk
    np_ = nchunks - 1
    w = root_of_unity(log_n)
    wv = [horner(c, zeta) for c in wires_coeffs]
    sv = [horner(c, zeta) for c in sigma_coeffs]
    zv = [horner(c, zeta) for c in zs_pp_coeffs]
    z_next = [horner(zs_pp_coeffs[c], zeta * w % P) for c in range(nch)]
    zh = (pow(zeta, n, P) - 1) % P
    l0 = zh * pow(n * (zeta - 1) % P, P - 2, P) % P
    terms = [l0 * (zv[c] - 1) % P for c in range(nch)]
    for c in range(nch):
        accs = [zv[c]] + zv[nch + c * np_: nch + (c + 1) * np_] + [z_next[c]]
        for k in range(nchunks):
            pn = pd = 1
            for j in range(k * chunk, min(R, (k + 1) * chunk)):
                pn = pn * ((wv[j] + int(betas[c]) * pow(G, j, P) * zeta + int(gammas[c])) % P) % P
                pd = pd * ((wv[j] + int(betas[c]) * sv[j] + int(gammas[c])) % P) % P
            terms.append((accs[k] * pn - accs[k + 1] * pd) % P)
    terms += [int(v) % P for v in gate_term_values]
    ok = True
    for a in range(nch):
        lhs = sum(t * pow(int(alphas[a]), i, P) for i, t in enumerate(terms)) % P
        q = sum(pow(zeta, k * n, P) * horner(quotient_chunks[a * chunk + k], zeta) for k in range(chunk)) % P
        ok = ok and lhs == zh * q % P
    return ok


# ---- quadratic extension F[X]/(X^2 - 7), elements as (a, b) tuples of Python ints ----------------------------------
def e_add(x, y): return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
def e_sub(x, y): return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)
def e_mul(x, y): return ((x[0] * y[0] + 7 * x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
def e_scale(x, s): return (x[0] * s % P, x[1] * s % P)


def e_inv(x):
    norm = (x[0] * x[0] - 7 * x[1] * x[1]) % P
    ni = pow(norm, P - 2, P)
    return (x[0] * ni % P, (-x[1]) * ni % P)


def e_pow(x, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = e_mul(r, x)
        x = e_mul(x, x)
        e >>= 1
    return r


def interpret_program_ext(code, imm, wires, consts):
    """The gate program over extension-field values (the opened values of the wires / constants)."""
    r, out = {}, []
    for w0, d, a, b in [tuple(int(x) for x in row) for row in code]:
        op, c = w0 & 0xff, (w0 >> 8) & 0xff
        if op == 0: r[d] = wires[a]
        elif op == 1: r[d] = consts[a]
        elif op == 2: r[d] = (int(imm[a]) % P, 0)
        elif op == 3: r[d] = e_add(r[a], r[b])
        elif op == 4: r[d] = e_sub(r[a], r[b])
        elif op == 5: r[d] = e_mul(r[a], r[b])
        elif op == 7: r[d] = e_add(e_mul(r[a], r[b]), r[c])
        elif op == 8: r[d] = e_scale(r[a], int(imm[b]) % P)
        elif op == 9: r[d] = e_add(r[a], (int(imm[b]) % P, 0))
        elif op == 10: r[d] = e_add(e_scale(r[a], int(imm[b]) % P), r[c])
        else: out.append(r[a])
    return out


def verify_gateless_openings(log_n, openings, num_consts, num_routed, num_wires, nch, chunk, betas, gammas, alphas, zeta,
                             program=None):
    """The plonk verifier's vanishing check on OPENED values (upstream plonk/verifier.rs: eval_vanishing_poly at zeta
    against Z_H(zeta) * reduce_with_powers(quotient chunk openings, zeta^n)) for a circuit without gates.
    openings: (k, 2) array in the order constants | sigmas | wires | zs | partial products | quotient chunks | zs at g zeta."""
    n = 1 << log_n
    op = [(int(a), int(b)) for a, b in openings]
    nchunks = (num_routed + chunk - 1) // chunk
    np_ = nchunks - 1
    o = num_consts
    sig = op[o:o + num_routed]; o += num_routed
    wires = op[o:o + num_wires]; o += num_wires
    zs = op[o:o + nch]; o += nch
    pps = op[o:o + nch * np_]; o += nch * np_
    quot = op[o:o + nch * chunk]; o += nch * chunk
    zs_next = op[o:o + nch]; o += nch
    assert o == len(op)
    z = (int(zeta[0]), int(zeta[1]))
    zh = e_sub(e_pow(z, n), (1, 0))
    l0 = e_mul(zh, e_inv(e_scale(e_sub(z, (1, 0)), n % P)))
    terms = [e_mul(l0, e_sub(zs[c], (1, 0))) for c in range(nch)]
    for c in range(nch):
        accs = [zs[c]] + pps[c * np_:(c + 1) * np_] + [zs_next[c]]
        b, g = int(betas[c]), int(gammas[c])
        for k in range(nchunks):
            pn = pd = (1, 0)
            for j in range(k * chunk, min(num_routed, (k + 1) * chunk)):
                pn = e_mul(pn, e_add(e_add(wires[j], e_scale(z, b * pow(G, j, P) % P)), (g, 0)))
                pd = e_mul(pd, e_add(e_add(wires[j], e_scale(sig[j], b)), (g, 0)))
            terms.append(e_sub(e_mul(accs[k], pn), e_mul(accs[k + 1], pd)))
    if program is not None:      # gate constraints, evaluated on the opened wires and constants
        terms += interpret_program_ext(program[0], program[1], wires, op[:num_consts])
    zn = e_pow(z, n)
    for a in range(nch):
        lhs = (0, 0)
        for t in terms[::-1]:
            lhs = e_add(e_scale(lhs, int(alphas[a])), t)
        q = (0, 0)
        for k in range(chunk - 1, -1, -1):
            q = e_add(e_mul(q, zn), quot[a * chunk + k])
        if lhs != e_mul(zh, q):
            return False
    return True


# ---- a small circuit WITH gates: arithmetic gates chained by copy constraints ---------------------------------------
class ProgramBuilder:
    """Assembles a gate program in the format of qpzk_gate_program: (num_ops, 4) uint32 code + uint64 immediates."""
    WIRE, CONST, IMM, ADD, SUB, MUL, EMIT = range(7)

    def __init__(self):
        self.code, self.imm, self.next = [], [], 0

    def _new(self):
        self.next += 1
        return self.next - 1

    def load(self, kind, idx):
        d = self._new()
        self.code.append([kind, d, idx, 0])
        return d

    def wire(self, j): return self.load(self.WIRE, j)
    def const(self, j): return self.load(self.CONST, j)

    def immediate(self, v):
        self.imm.append(int(v) % P)
        return self.load(self.IMM, len(self.imm) - 1)

    def op(self, kind, a, b):
        d = self._new()
        self.code.append([kind, d, a, b])
        return d

    def add(self, a, b): return self.op(self.ADD, a, b)
    def sub(self, a, b): return self.op(self.SUB, a, b)
    def mul(self, a, b): return self.op(self.MUL, a, b)
    def emit(self, a): self.code.append([self.EMIT, 0, a, 0])

    def arrays(self):
        return np.array(self.code, dtype=np.uint32).reshape(-1, 4), np.array(self.imm, dtype=np.uint64)


def interpret_program(code, imm, wires, consts):
    """Big-integer interpreter of a gate program at one point: wires / consts are lists of field elements."""
    r, out = {}, []
    for w0, d, a, b in [tuple(int(x) for x in row) for row in code]:
        op, c = w0 & 0xff, (w0 >> 8) & 0xff
        if op == 0: r[d] = int(wires[a]) % P
        elif op == 1: r[d] = int(consts[a]) % P
        elif op == 2: r[d] = int(imm[a]) % P
        elif op == 3: r[d] = (r[a] + r[b]) % P
        elif op == 4: r[d] = (r[a] - r[b]) % P
        elif op == 5: r[d] = r[a] * r[b] % P
        elif op == 7: r[d] = (r[a] * r[b] + r[c]) % P
        elif op == 8: r[d] = r[a] * int(imm[b]) % P
        elif op == 9: r[d] = (r[a] + int(imm[b])) % P
        elif op == 10: r[d] = (r[a] * int(imm[b]) + r[c]) % P
        else: out.append(r[a])
    return out


ARITH_OPS = 4   # operations per row: wires [m0, m1, addend, out] x 4; constants columns [selector, c0, c1]


def arithmetic_program():
    """constraint_k = selector * (out_k - (c0 * m0_k * m1_k + c1 * addend_k)),  k < 4   (upstream ArithmeticGate with
    a selector filter)."""
    b = ProgramBuilder()
    s, c0, c1 = b.const(0), b.const(1), b.const(2)
    for k in range(ARITH_OPS):
        m0, m1, ad, out = (b.wire(4 * k + t) for t in range(4))
        t = b.add(b.mul(b.mul(m0, m1), c0), b.mul(ad, c1))
        b.emit(b.mul(b.sub(out, t), s))
    return b


def arithmetic_circuit(log_n, seed, inactive_rows=3):
    """A satisfying instance: (wires (16, n), constants (3, n), sigmas (16, n)).  Copy constraints: out_0 of row i feeds
    m0_0 of row i + 1 (a chain through the active rows), and m1_2 = m0_3 inside every active row."""
    rng = np.random.default_rng(seed)
    n, R = 1 << log_n, 4 * ARITH_OPS
    act = n - inactive_rows
    w = root_of_unity(log_n)
    xs = [pow(w, i, P) for i in range(n)]
    ids = [[pow(G, j, P) * x % P for x in xs] for j in range(R)]
    rnd = lambda: int(rng.integers(0, P, dtype=np.uint64))
    consts = [[1 if i < act else 0 for i in range(n)], [rnd() for _ in range(n)], [rnd() for _ in range(n)]]
    wires = [[rnd() for _ in range(n)] for _ in range(R)]
    nxt = {}                                            # position -> next position of its cycle
    link = lambda a, b: (nxt.__setitem__(a, b), nxt.__setitem__(b, a))
    for i in range(act):
        if i > 0:
            wires[0][i] = wires[3][i - 1]               # m0_0(i) = out_0(i - 1)
            link((3, i - 1), (0, i))
        wires[12][i] = wires[9][i]                      # m0_3(i) = m1_2(i)
        link((9, i), (12, i))
        for k in range(ARITH_OPS):
            m0, m1, ad = wires[4 * k][i], wires[4 * k + 1][i], wires[4 * k + 2][i]
            wires[4 * k + 3][i] = (consts[1][i] * m0 % P * m1 + consts[2][i] * ad) % P
    sig = [[ids[j][i] for i in range(n)] for j in range(R)]
    for (j, i), (j2, i2) in nxt.items():
        sig[j][i] = ids[j2][i2]
    u = lambda a: np.array(a, dtype=np.uint64)
    return u(wires), u(consts), u(sig)


def poseidon_chain_circuit(log_n, seed, rc, routed=80, inactive_rows=2, witness=None):
    """A hash chain in the reference's shape: every active row is one PoseidonGate (135 wires, the first 80 routed),
    output words 0..3 of row i are copied to input words 0..3 of row i + 1, the other input words are free.
    Returns (wires (135, n), constants (1, n) = the gate's selector, sigmas (routed, n)).
    witness(inputs, swap) -> the 135 wires of a row: another gate over the same layout (Poseidon2Gate); default PoseidonGate."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qp-zk-circuits_b200"))
    from qpzk import gates
    rng = np.random.default_rng(seed)
    n = 1 << log_n
    act = n - inactive_rows
    rnd = lambda: int(rng.integers(0, P, dtype=np.uint64))
    w = root_of_unity(log_n)
    xs = [pow(w, i, P) for i in range(n)]
    ids = [[pow(G, j, P) * x % P for x in xs] for j in range(routed)]
    sig = [list(r) for r in ids]
    wires = [[rnd() for _ in range(n)] for _ in range(gates.POSEIDON_GATE_WIRES)]
    prev = None
    for i in range(act):
        inp = [rnd() for _ in range(12)]
        if prev is not None:
            inp[:4] = prev[:4]
            for t in range(4):                        # out_t(i - 1) <-> in_t(i): a 2-cycle of the permutation
                a, b = (gates.wire_output(t), i - 1), (gates.wire_input(t), i)
                sig[a[0]][a[1]], sig[b[0]][b[1]] = ids[b[0]][b[1]], ids[a[0]][a[1]]
        row = witness(inp, i & 1) if witness else gates.poseidon_gate_witness(inp, i & 1, rc)
        for j in range(gates.POSEIDON_GATE_WIRES):
            wires[j][i] = row[j]
        prev = row[12:24] if not (i & 1) else row[12:24]
    u = lambda a: np.array(a, dtype=np.uint64)
    return u(wires), u([[1 if i < act else 0 for i in range(n)]]), u(sig)


# ---- a verifier over the serialized proof ---------------------------------------------------------------------------
def verify_proof_bytes(orc, qpzk, data, params_q, params_o, shape, cs_cap, circuit_digest, num_routed, chunk=8, program=None):
    """What upstream's `verify()` does with `ProofWithPublicInputs::from_bytes` output, for a circuit given as data:
    re-derive every challenge from the transcript (circuit digest, public-inputs hash, the three caps, the openings),
    check the vanishing identity on the opened values (gate program included) and verify the FRI proof.
    params_q / params_o: the same FRI parameters as qpzk / oracle structures.  Returns True iff the proof is accepted."""
    pr = qpzk.proof_from_bytes(params_q, shape, data)
    log_n, nch = shape.log_n, shape.num_zs
    ch = orc.Challenger()
    ch.observe(np.asarray(circuit_digest, np.uint64))
    ch.observe(orc.hash_no_pad(pr["public_inputs"]))
    ch.observe(pr["wires_cap"].reshape(-1))
    bg = ch.get(2 * nch)
    ch.observe(pr["zs_cap"].reshape(-1))
    alphas = ch.get(nch)
    ch.observe(pr["quotient_cap"].reshape(-1))
    zeta = ch.get(2)
    if not verify_gateless_openings(log_n, pr["openings"], shape.num_constants, num_routed, shape.num_wires, nch, chunk,
                                    bg[:nch], bg[nch:], alphas, zeta, program=program):
        return False
    ch.observe(pr["openings"].reshape(-1))
    g = root_of_unity(log_n)
    g_zeta = [int(zeta[0]) * g % P, int(zeta[1]) * g % P]
    lens = [shape.num_constants + shape.num_sigmas, shape.num_wires, shape.num_zs + shape.num_partial_products,
            shape.num_quotient_polys]
    all_polys = [(o, i) for o in range(4) for i in range(lens[o])]
    inst = orc.Instance([([int(zeta[0]), int(zeta[1])], all_polys), (g_zeta, [(2, i) for i in range(nch)])])
    caps = [np.asarray(cs_cap, np.uint64).reshape(-1, 4), pr["wires_cap"], pr["zs_cap"], pr["quotient_cap"]]
    return orc.fri_verify(log_n, list(shape.leaf_lens), caps, inst, pr["openings"], params_o, ch, pr["fri_proof"]) == 0
