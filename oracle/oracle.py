"""ORACLE (test infrastructure, NOT product code): ctypes wrapper over oracle/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (qp-zk-circuits_b200/qpzk) never does.

The algorithms restated are listed in oracle/oracle.c (the reference's hot
path lives in the un-vendored crate qp-plonky2 =1.5.5, reference Cargo.toml:20).
PARITY STATUS: parity unpinned beyond the Poseidon permutation KATs (see oracle/oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
P = 0xFFFFFFFF00000001
GENERATOR = 14293326489335486720
U64P = C.POINTER(C.c_uint64)
U32P = C.POINTER(C.c_uint32)
MAX_ROUNDS = 16


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class FriParams(C.Structure):
    _fields_ = [("rate_bits", C.c_uint32), ("cap_height", C.c_uint32), ("pow_bits", C.c_uint32),
                ("num_queries", C.c_uint32), ("num_rounds", C.c_uint32), ("arity_bits", C.c_uint32 * MAX_ROUNDS)]


class FriInstance(C.Structure):
    _fields_ = [("num_batches", C.c_uint32), ("points", U64P), ("batch_sizes", U32P),
                ("poly_oracle", U32P), ("poly_index", U32P)]


class ChallengerState(C.Structure):
    _fields_ = [("state", C.c_uint64 * 12), ("inp", C.c_uint64 * 8), ("out", C.c_uint64 * 8),
                ("in_len", C.c_uint32), ("out_len", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_gl_mul.restype = C.c_uint64; L.orc_gl_mul.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_gl_inv.restype = C.c_uint64; L.orc_gl_inv.argtypes = [C.c_uint64]
        L.orc_gl_exp.restype = C.c_uint64; L.orc_gl_exp.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_gl_root_of_unity.restype = C.c_uint64; L.orc_gl_root_of_unity.argtypes = [C.c_uint]
        for f in (L.orc_poseidon, L.orc_poseidon_naive, L.orc_poseidon_rc):
            f.restype = None; f.argtypes = [U64P]
        L.orc_partial_products.restype = C.c_int
        L.orc_partial_products.argtypes = [U64P, U64P, C.c_uint, C.c_uint, C.c_uint, C.c_uint, U64P, U64P, U64P]
        L.orc_quotient_polys.restype = C.c_int
        L.orc_quotient_polys.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint,
                                         U64P, U64P, U64P, U64P, C.c_uint, U64P]
        L.orc_gate_terms.restype = C.c_int
        L.orc_gate_terms.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, U32P, C.c_uint, U64P, U64P]
        for f in (L.orc_hash_no_pad, L.orc_hash_or_noop, L.orc_hash_no_pad_p2):
            f.restype = None; f.argtypes = [U64P, C.c_size_t, U64P]
        L.orc_poseidon2.restype = None; L.orc_poseidon2.argtypes = [U64P]
        L.orc_two_to_one.restype = None; L.orc_two_to_one.argtypes = [U64P, U64P, U64P]
        for f in (L.orc_fft, L.orc_ifft):
            f.restype = None; f.argtypes = [U64P, C.c_uint]
        L.orc_coset_fft.restype = None; L.orc_coset_fft.argtypes = [U64P, C.c_uint, C.c_uint64]
        L.orc_merkle_new.restype = C.c_void_p; L.orc_merkle_new.argtypes = [U64P, C.c_size_t, C.c_size_t, C.c_uint]
        L.orc_merkle_free.restype = None; L.orc_merkle_free.argtypes = [C.c_void_p]
        L.orc_merkle_cap.restype = C.c_size_t; L.orc_merkle_cap.argtypes = [C.c_void_p, U64P]
        L.orc_merkle_prove.restype = C.c_size_t; L.orc_merkle_prove.argtypes = [C.c_void_p, C.c_size_t, U64P]
        L.orc_merkle_leaf.restype = U64P; L.orc_merkle_leaf.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_merkle_verify.restype = C.c_int
        L.orc_merkle_verify.argtypes = [U64P, C.c_size_t, C.c_size_t, U64P, C.c_uint, U64P, C.c_size_t]
        for f in (L.orc_batch_from_values, L.orc_batch_from_coeffs):
            f.restype = C.c_void_p; f.argtypes = [U64P, C.c_uint, C.c_size_t, C.c_uint, C.c_uint, U64P]
        L.orc_batch_free.restype = None; L.orc_batch_free.argtypes = [C.c_void_p]
        L.orc_batch_tree.restype = C.c_void_p; L.orc_batch_tree.argtypes = [C.c_void_p]
        L.orc_batch_coeffs.restype = U64P; L.orc_batch_coeffs.argtypes = [C.c_void_p]
        L.orc_batch_leaf_len.restype = C.c_size_t; L.orc_batch_leaf_len.argtypes = [C.c_void_p]
        L.orc_batch_eval_ext.restype = None; L.orc_batch_eval_ext.argtypes = [C.c_void_p, U64P, U64P]
        L.orc_challenger_init.restype = None; L.orc_challenger_init.argtypes = [C.POINTER(ChallengerState)]
        L.orc_challenger_observe.restype = None
        L.orc_challenger_observe.argtypes = [C.POINTER(ChallengerState), U64P, C.c_size_t]
        L.orc_challenger_get.restype = None
        L.orc_challenger_get.argtypes = [C.POINTER(ChallengerState), U64P, C.c_size_t]
        L.orc_fri_params_constant_arity.restype = None
        L.orc_fri_params_constant_arity.argtypes = [C.POINTER(FriParams)] + [C.c_uint] * 7
        L.orc_fri_proof_words.restype = C.c_size_t
        L.orc_fri_proof_words.argtypes = [C.POINTER(FriParams), C.c_uint, C.c_uint, U32P]
        L.orc_fri_prove.restype = C.c_int
        L.orc_fri_prove.argtypes = [C.POINTER(C.c_void_p), C.c_uint, C.POINTER(FriInstance), C.POINTER(FriParams),
                                    C.POINTER(ChallengerState), C.c_uint64, U64P]
        L.orc_fri_verify.restype = C.c_int
        L.orc_fri_verify.argtypes = [C.c_uint, C.c_uint, U32P, U64P, C.POINTER(FriInstance), U64P,
                                     C.POINTER(FriParams), C.POINTER(ChallengerState), U64P]
        for f in (L.orc_eval_from_values, L.orc_eval_from_coeffs):
            f.restype = C.c_int; f.argtypes = [U64P, C.c_size_t, C.c_size_t, C.c_uint, C.c_uint64, U64P]
        L.orc_lde_subgroup_rows.restype = C.c_int
        L.orc_lde_subgroup_rows.argtypes = [U64P, C.c_size_t, C.c_size_t, C.c_uint, C.c_uint, C.c_uint64, C.c_uint, U64P]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.restype = None; L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p64(a):
    return a.ctypes.data_as(U64P)


def _p32(a):
    return a.ctypes.data_as(U32P)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def partial_products(wires, sigmas, betas, gammas, chunk=8):
    """all_wires_permutation_partial_products: wires, sigmas (num_routed, n) -> (len(betas) * ceil(R/chunk), n) in the
    order of the committed batch: Z per challenge, then the partial products of challenge 0, 1, ..."""
    w, sg = _u64(wires), _u64(sigmas)
    R, n = w.shape
    assert sg.shape == w.shape and n & (n - 1) == 0
    b, g = _u64(betas), _u64(gammas)
    nch = (R + chunk - 1) // chunk
    out = np.zeros((b.size * nch, n), np.uint64)
    rc = lib().orc_partial_products(_p64(w), _p64(sg), n.bit_length() - 1, R, chunk, b.size, _p64(b), _p64(g), _p64(out))
    assert rc == 0
    return out


def quotient_polys(constants_sigmas, wires, zs_partial_products, sigma_col0, num_routed, betas, gammas, alphas,
                   quotient_degree_factor=8, gate_terms=None):
    """compute_quotient_polys for the permutation argument + optional gate terms ((n*Q, G) array, row = leaf index):
    (len(alphas) * Q, n) quotient chunk coefficients."""
    b, g, a = _u64(betas), _u64(gammas), _u64(alphas)
    gt, G = None, 0
    if gate_terms is not None:
        gt = _u64(gate_terms)
        G = gt.shape[1]
    out = np.zeros((a.size * quotient_degree_factor, 1 << wires.log_n), np.uint64)
    rc = lib().orc_quotient_polys(constants_sigmas.h, wires.h, zs_partial_products.h, sigma_col0, num_routed,
                                  quotient_degree_factor, a.size, _p64(b), _p64(g), _p64(a),
                                  _p64(gt) if gt is not None else None, G, _p64(out))
    assert rc == 0, rc
    return out


def gate_terms(constants_sigmas, wires, code, imm, quotient_degree_factor=8):
    """Interpret a gate program (qpzk_gate_program format: (num_ops, 4) uint32 code, uint64 immediates) at every point
    of the quotient domain: (n*Q, num_terms) matrix, row = leaf index."""
    code = np.ascontiguousarray(code, dtype=np.uint32).reshape(-1, 4)
    imm = _u64(imm if len(imm) else [0])
    cp = code.ctypes.data_as(U32P)
    nt = lib().orc_gate_terms(constants_sigmas.h, wires.h, quotient_degree_factor, cp, code.shape[0], _p64(imm), None)
    assert nt >= 0, "bad gate program"
    out = np.zeros(((quotient_degree_factor << wires.log_n), max(nt, 1)), np.uint64)
    if nt:
        lib().orc_gate_terms(constants_sigmas.h, wires.h, quotient_degree_factor, cp, code.shape[0], _p64(imm), _p64(out))
    return out[:, :nt]


def round_constants():
    out = np.zeros(360, np.uint64)
    lib().orc_poseidon_rc(_p64(out))
    return out


def poseidon(state, naive=False):
    s = _u64(state).copy()
    (lib().orc_poseidon_naive if naive else lib().orc_poseidon)(_p64(s))
    return s


def hash_no_pad(x):
    x = _u64(x); out = np.zeros(4, np.uint64)
    lib().orc_hash_no_pad(_p64(x), x.size, _p64(out))
    return out


def poseidon2(state):
    """The Poseidon2 permutation of the reference's application hash (12 words)."""
    s = _u64(state).copy()
    lib().orc_poseidon2(_p64(s))
    return s


def hash_no_pad_p2(x):
    """Poseidon2Hash::hash_no_pad: additive sponge, rate 8, padding x || 1 || 0*, four words out."""
    x = _u64(x).reshape(-1); out = np.zeros(4, np.uint64)
    lib().orc_hash_no_pad_p2(_p64(x) if x.size else None, x.size, _p64(out))
    return out


def hash_or_noop(x):
    x = _u64(x); out = np.zeros(4, np.uint64)
    lib().orc_hash_or_noop(_p64(x), x.size, _p64(out))
    return out


def two_to_one(l, r):
    l = _u64(l); r = _u64(r); out = np.zeros(4, np.uint64)
    lib().orc_two_to_one(_p64(l), _p64(r), _p64(out))
    return out


def fft(a, inverse=False):
    a = _u64(a).copy()
    log_n = int(a.size).bit_length() - 1
    (lib().orc_ifft if inverse else lib().orc_fft)(_p64(a), log_n)
    return a


def coset_fft(a, shift=GENERATOR):
    a = _u64(a).copy()
    lib().orc_coset_fft(_p64(a), int(a.size).bit_length() - 1, shift)
    return a


class MerkleTree:
    """MerkleTree::new(leaves, cap_height): leaves is a (num_leaves, leaf_len) uint64 array."""

    def __init__(self, leaves=None, cap_height=0, _handle=None, _owner=None):
        if _handle is not None:
            self.h, self._own, self._owner = _handle, False, _owner
            return
        leaves = _u64(leaves)
        self.num_leaves, self.leaf_len = leaves.shape
        self.cap_height = cap_height
        self.h = lib().orc_merkle_new(_p64(leaves), self.num_leaves, self.leaf_len, cap_height)
        if not self.h:
            raise ValueError("orc_merkle_new failed")
        self._own = True

    def cap(self):
        out = np.zeros(4 << self.cap_height, np.uint64)
        lib().orc_merkle_cap(self.h, _p64(out))
        return out.reshape(-1, 4)

    def prove(self, i):
        out = np.zeros(4 * 64, np.uint64)
        k = lib().orc_merkle_prove(self.h, i, _p64(out))
        return out[:4 * k].reshape(k, 4).copy()

    def leaf(self, i):
        p = lib().orc_merkle_leaf(self.h, i)
        return np.ctypeslib.as_array(p, shape=(self.leaf_len,)).copy()

    def __del__(self):
        if getattr(self, "_own", False) and self.h:
            lib().orc_merkle_free(self.h)
            self.h = None


def merkle_verify(leaf, index, cap, siblings):
    leaf = _u64(leaf); cap = _u64(cap).reshape(-1); siblings = _u64(siblings).reshape(-1)
    cap_height = (cap.size // 4).bit_length() - 1
    return lib().orc_merkle_verify(_p64(leaf), leaf.size, index, _p64(cap), cap_height, _p64(siblings),
                                   siblings.size // 4) == 0


class PolynomialBatch:
    """PolynomialBatch::from_values / from_coeffs. cols: (num_cols, n) uint64 (one row per polynomial)."""

    def __init__(self, handle, log_n, num_cols, rate_bits, cap_height, salted):
        self.h, self.log_n, self.num_cols = handle, log_n, num_cols
        self.rate_bits, self.cap_height, self.salted = rate_bits, cap_height, salted
        self.leaf_len = num_cols + (4 if salted else 0)
        t = MerkleTree(_handle=lib().orc_batch_tree(handle), _owner=self)
        t.num_leaves, t.leaf_len, t.cap_height = 1 << (log_n + rate_bits), self.leaf_len, cap_height
        self.merkle_tree = t

    @classmethod
    def _make(cls, fn, cols, rate_bits, cap_height, salt):
        cols = _u64(cols)
        num_cols, n = cols.shape
        log_n = n.bit_length() - 1
        sp = None
        if salt is not None:
            salt = _u64(salt)
            assert salt.shape == (4, n << rate_bits)
            sp = _p64(salt)
        h = fn(_p64(cols), log_n, num_cols, rate_bits, cap_height, sp)
        if not h:
            raise ValueError("oracle batch construction failed")
        return cls(h, log_n, num_cols, rate_bits, cap_height, salt is not None)

    @classmethod
    def from_values(cls, cols, rate_bits, cap_height, salt=None):
        return cls._make(lib().orc_batch_from_values, cols, rate_bits, cap_height, salt)

    @classmethod
    def from_coeffs(cls, cols, rate_bits, cap_height, salt=None):
        return cls._make(lib().orc_batch_from_coeffs, cols, rate_bits, cap_height, salt)

    def coeffs(self):
        p = lib().orc_batch_coeffs(self.h)
        return np.ctypeslib.as_array(p, shape=(self.num_cols, 1 << self.log_n)).copy()

    def cap(self):
        return self.merkle_tree.cap()

    def leaves(self):
        L = 1 << (self.log_n + self.rate_bits)
        p = lib().orc_merkle_leaf(self.merkle_tree.h, 0)
        return np.ctypeslib.as_array(p, shape=(L, self.leaf_len)).copy()

    def eval_ext(self, point):
        pt = _u64(point); out = np.zeros(2 * self.num_cols, np.uint64)
        lib().orc_batch_eval_ext(self.h, _p64(pt), _p64(out))
        return out.reshape(-1, 2)

    def __del__(self):
        if self.h:
            lib().orc_batch_free(self.h)
            self.h = None


class Challenger:
    def __init__(self):
        self.s = ChallengerState()
        lib().orc_challenger_init(C.byref(self.s))

    def observe(self, elems):
        e = _u64(elems).reshape(-1)
        lib().orc_challenger_observe(C.byref(self.s), _p64(e), e.size)

    def get(self, n=1):
        out = np.zeros(n, np.uint64)
        lib().orc_challenger_get(C.byref(self.s), _p64(out), n)
        return out

    def clone(self):
        c = Challenger()
        C.memmove(C.byref(c.s), C.byref(self.s), C.sizeof(ChallengerState))
        return c

    def export(self):
        """(state[12], in[8], out[8], in_len, out_len) as plain python for cross-library comparison."""
        return (list(self.s.state), list(self.s.inp)[:self.s.in_len], list(self.s.out)[:self.s.out_len])


def fri_params(log_n, rate_bits=3, cap_height=4, pow_bits=16, num_queries=28, arity_bits=4, final_poly_bits=5):
    p = FriParams()
    lib().orc_fri_params_constant_arity(C.byref(p), log_n, rate_bits, cap_height, pow_bits, num_queries,
                                        arity_bits, final_poly_bits)
    return p


class Instance:
    """FriInstanceInfo: batches = [(point(2 words), [(oracle_index, poly_index), ...]), ...]"""

    def __init__(self, batches):
        self.points = _u64([w for pt, _ in batches for w in pt])
        self.sizes = np.array([len(ps) for _, ps in batches], np.uint32)
        self.po = np.array([o for _, ps in batches for o, _ in ps], np.uint32)
        self.pi = np.array([i for _, ps in batches for _, i in ps], np.uint32)
        self.c = FriInstance(len(batches), _p64(self.points), _p32(self.sizes), _p32(self.po), _p32(self.pi))


def fri_proof_words(params, log_n, leaf_lens):
    ll = np.array(leaf_lens, np.uint32)
    return lib().orc_fri_proof_words(C.byref(params), log_n, ll.size, _p32(ll))


def fri_prove(oracles, instance, params, challenger, pow_witness=None):
    arr = (C.c_void_p * len(oracles))(*[o.h for o in oracles])
    words = fri_proof_words(params, oracles[0].log_n, [o.leaf_len for o in oracles])
    out = np.zeros(words, np.uint64)
    w = 0xFFFFFFFFFFFFFFFF if pow_witness is None else pow_witness
    rc = lib().orc_fri_prove(arr, len(oracles), C.byref(instance.c), C.byref(params), C.byref(challenger.s), w,
                             _p64(out))
    if rc:
        raise RuntimeError("orc_fri_prove rc=%d" % rc)
    return out


def fri_verify(log_n, leaf_lens, caps, instance, openings, params, challenger, proof):
    ll = np.array(leaf_lens, np.uint32)
    caps = _u64(np.concatenate([_u64(c).reshape(-1) for c in caps]))
    openings = _u64(openings).reshape(-1)
    proof = _u64(proof)
    return lib().orc_fri_verify(log_n, ll.size, _p32(ll), _p64(caps), C.byref(instance.c), _p64(openings),
                                C.byref(params), C.byref(challenger.s), _p64(proof))


def eval_from_values(values, r):
    """P_c(r) for every column c, from the column's VALUES on the subgroup (barycentric formula; r outside the subgroup)."""
    v = _u64(values)
    out = np.zeros(v.shape[0], np.uint64)
    rc = lib().orc_eval_from_values(_p64(v), v.shape[1], v.shape[0], v.shape[1].bit_length() - 1, int(r), _p64(out))
    assert rc == 0, rc
    return out


def eval_from_coeffs(coeffs, r):
    """P_c(r) for every column c by Horner's rule on the coefficients."""
    c = _u64(coeffs)
    out = np.zeros(c.shape[0], np.uint64)
    rc = lib().orc_eval_from_coeffs(_p64(c), c.shape[1], c.shape[0], c.shape[1].bit_length() - 1, int(r), _p64(out))
    assert rc == 0, rc
    return out


def lde_subgroup_rows(coeffs, rate_bits, i0, log_m):
    """LDE values of every column at the natural LDE indices i0 + k * (L >> log_m), k < 2^log_m: (2^log_m, C) array."""
    c = _u64(coeffs)
    out = np.zeros((1 << log_m, c.shape[0]), np.uint64)
    rc = lib().orc_lde_subgroup_rows(_p64(c), c.shape[1], c.shape[0], c.shape[1].bit_length() - 1, rate_bits, int(i0), log_m, _p64(out))
    assert rc == 0, rc
    return out


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n=None):
    """Use n OpenMP threads (default: every core this process may run on) whatever OMP_NUM_THREADS says --
    torch.distributed.run sets it to 1 for every rank."""
    if n is None:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
    lib().orc_set_num_threads(int(n))
    return num_threads()


# ---- the commitment / opening half of prove() on the CPU (mirror of qpzk/prover.py) ----------------
def prove_synthetic(constants_sigmas_batch, wires, zs_partial_products, quotient_coeffs, public_inputs_hash,
                    log_n, num_zs=2, rate_bits=3, cap_height=4, pow_bits=16, num_queries=28, arity_bits=4,
                    final_poly_bits=5, sigmas=None, quotient_degree_factor=8, sigma_col0=None, gate_program=None):
    """Same transcript and the same outputs as qpzk.prover.prove_synthetic, computed by the CPU restatement."""
    ch = Challenger()
    ch.observe(np.arange(4, dtype=np.uint64) + 11)
    ch.observe(public_inputs_hash)
    wires_b = PolynomialBatch.from_values(wires, rate_bits, cap_height)
    ch.observe(wires_b.cap())
    betas_gammas = ch.get(2 * num_zs)
    if zs_partial_products is None:
        sg = _u64(sigmas)
        zs_partial_products = partial_products(_u64(wires)[:sg.shape[0]], sg, betas_gammas[:num_zs],
                                               betas_gammas[num_zs:], quotient_degree_factor)
    zs_b = PolynomialBatch.from_values(zs_partial_products, rate_bits, cap_height)
    ch.observe(zs_b.cap())
    alphas = ch.get(num_zs)
    if quotient_coeffs is None:
        gt = None
        if gate_program is not None:      # (code, imm) in the format of qpzk_gate_program
            gt = gate_terms(constants_sigmas_batch, wires_b, gate_program[0], gate_program[1], quotient_degree_factor)
        quotient_coeffs = quotient_polys(constants_sigmas_batch, wires_b, zs_b, sigma_col0, _u64(sigmas).shape[0],
                                         betas_gammas[:num_zs], betas_gammas[num_zs:], alphas, quotient_degree_factor, gt)
    quot_b = PolynomialBatch.from_coeffs(quotient_coeffs, rate_bits, cap_height)
    ch.observe(quot_b.cap())
    zeta = ch.get(2)
    g = lib().orc_gl_root_of_unity(log_n)
    g_zeta = [lib().orc_gl_mul(int(zeta[0]), g), lib().orc_gl_mul(int(zeta[1]), g)]
    oracles = [constants_sigmas_batch, wires_b, zs_b, quot_b]
    all_polys = [(o, i) for o, b in enumerate(oracles) for i in range(b.num_cols)]
    zs_next = [(2, i) for i in range(num_zs)]
    inst = Instance([([int(zeta[0]), int(zeta[1])], all_polys), (g_zeta, zs_next)])
    openings = np.concatenate([np.concatenate([b.eval_ext(zeta) for b in oracles]), zs_b.eval_ext(g_zeta)[:num_zs]])
    ch.observe(openings)
    params = fri_params(log_n, rate_bits, cap_height, pow_bits, num_queries, arity_bits, final_poly_bits)
    verifier_ch = ch.clone()
    proof = fri_prove(oracles, inst, params, ch)
    return dict(caps=[wires_b.cap(), zs_b.cap(), quot_b.cap()], openings=openings, fri_proof=proof,
                challenges=(betas_gammas, alphas, zeta), instance=inst, params=params, verifier_challenger=verifier_ch,
                leaf_lens=[b.leaf_len for b in oracles], all_caps=[b.cap() for b in oracles])


# ---- wire format (upstream util/serialization, recalled order -- parity unpinned, Appendix B A6) -----
def parse_fri_proof(words, params, log_n, leaf_lens):
    """Flat FriProof words -> the structured proof upstream holds: dict(commit_phase_merkle_caps, query_round_proofs
    [dict(initial_trees_proof=[(evals, siblings)], steps=[(evals, siblings)])], final_poly, pow_witness)."""
    w = [int(x) for x in _u64(words)]
    pos = [0]

    def take(n):
        v = w[pos[0]:pos[0] + n]
        assert len(v) == n
        pos[0] += n
        return v

    H, R = 1 << params.cap_height, params.num_rounds
    log_L = log_n + params.rate_bits
    caps = [[take(4) for _ in range(H)] for _ in range(R)]
    rounds = []
    for _ in range(params.num_queries):
        init = []
        for ll in leaf_lens:
            init.append((take(int(ll)), [take(4) for _ in range(log_L - params.cap_height)]))
        steps, lg = [], log_L
        for r in range(R):
            ab = params.arity_bits[r]
            evals = take(2 << ab)
            steps.append((evals, [take(4) for _ in range(max(lg - ab - params.cap_height, 0))]))
            lg -= ab
        rounds.append(dict(initial_trees_proof=init, steps=steps))
    final_poly = take(2 << (log_n - sum(params.arity_bits[r] for r in range(R))))
    pow_witness = take(1)[0]
    assert pos[0] == len(w)
    return dict(commit_phase_merkle_caps=caps, query_round_proofs=rounds, final_poly=final_poly, pow_witness=pow_witness)


def proof_to_bytes(caps3, opening_set, fri, public_inputs):
    """write_proof_with_public_inputs: caps3 = (wires, zs_partial_products, quotient) caps as lists of 4-word hashes;
    opening_set = dict of the upstream OpeningSet fields (lists of (a, b) extension elements); fri = parse_fri_proof()."""
    import struct
    out = bytearray()
    felt = lambda v: out.extend(struct.pack("<Q", int(v)))

    def merkle_proof(sibs):
        out.append(len(sibs))
        for h in sibs:
            for v in h:
                felt(v)

    for cap in caps3:
        for h in cap:
            for v in h:
                felt(v)
    for name in ("constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products", "quotient_polys",
                 "lookup_zs", "lookup_zs_next"):
        for e in opening_set.get(name, []):
            felt(e[0]); felt(e[1])
    for cap in fri["commit_phase_merkle_caps"]:
        for h in cap:
            for v in h:
                felt(v)
    for rd in fri["query_round_proofs"]:
        for evals, sibs in rd["initial_trees_proof"]:
            for v in evals:
                felt(v)
            merkle_proof(sibs)
        for evals, sibs in rd["steps"]:
            for v in evals:
                felt(v)
            merkle_proof(sibs)
    for v in fri["final_poly"]:
        felt(v)
    felt(fri["pow_witness"])
    out.extend(struct.pack("<Q", len(public_inputs)))
    for v in public_inputs:
        felt(v)
    return bytes(out)
