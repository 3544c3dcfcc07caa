"""Host-side mirror of the plonky2 commitment surface over libqpzk.so (ctypes on the C ABI of include/qpzk.h).

The reference's host language is Rust (no cargo/rustc in this image), so this module plays the role of the
`-sys` + safe wrapper crate that a `[patch.crates-io]` fork of qp-plonky2 =1.5.5 would contain (see
INTEGRATION.md).  Names, argument meaning and error behaviour follow the upstream API the reference reaches
through `prove()` (reference wormhole/prover/src/lib.rs:171-175):

    PolynomialBatch.from_values(values, rate_bits, blinding_salt, cap_height)
    PolynomialBatch.from_coeffs(polynomials, rate_bits, blinding_salt, cap_height)
    PolynomialBatch.prove_openings(instance, oracles, challenger, fri_params)
    MerkleTree.new(leaves, cap_height) / .cap / .prove(leaf_index)
    Challenger.observe_elements / observe_cap / get_n_challenges

There is no CPU fallback: importing works anywhere (so symbol checks run without a GPU), but creating a
Context fails loudly when the library or a Blackwell device is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QPZK_LIB") or os.path.join(os.path.dirname(_HERE), "libqpzk.so")
HEADER_PATH = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "qpzk.h")

P = 0xFFFFFFFF00000001
GENERATOR = 14293326489335486720
POW_SEARCH = 0xFFFFFFFFFFFFFFFF
SALT_SIZE = 4
MAX_FRI_ROUNDS = 16
U64P = C.POINTER(C.c_uint64)
U32P = C.POINTER(C.c_uint32)


class QpzkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("qpzk error %d: %s" % (code, msg))
        self.code = code


class FriParams(C.Structure):
    _fields_ = [("rate_bits", C.c_uint32), ("cap_height", C.c_uint32), ("proof_of_work_bits", C.c_uint32),
                ("num_query_rounds", C.c_uint32), ("num_rounds", C.c_uint32),
                ("arity_bits", C.c_uint32 * MAX_FRI_ROUNDS)]

    @property
    def reduction_arity_bits(self):
        return list(self.arity_bits)[:self.num_rounds]


class _FriInstance(C.Structure):
    _fields_ = [("num_batches", C.c_uint32), ("points", U64P), ("batch_sizes", U32P),
                ("poly_oracle", U32P), ("poly_index", U32P)]


class ProofShape(C.Structure):
    """qpzk_proof_shape: how many polynomials of each kind a proof of this circuit opens (wire format)."""
    _fields_ = [("log_n", C.c_uint32), ("num_constants", C.c_uint32), ("num_sigmas", C.c_uint32),
                ("num_wires", C.c_uint32), ("num_zs", C.c_uint32), ("num_partial_products", C.c_uint32),
                ("num_quotient_polys", C.c_uint32), ("leaf_lens", C.c_uint32 * 4), ("num_public_inputs", C.c_uint32)]


class GateProgram(C.Structure):
    """qpzk_gate_program: gate constraints as a straight-line register program (see include/qpzk.h)."""
    _fields_ = [("code", U32P), ("num_ops", C.c_uint32), ("imm", U64P), ("num_imm", C.c_uint32), ("num_regs", C.c_uint32)]

    @classmethod
    def make(cls, code, imm, num_regs):
        code = np.ascontiguousarray(code, dtype=np.uint32).reshape(-1, 4)
        imm = np.ascontiguousarray(imm, dtype=np.uint64).reshape(-1)
        g = cls(code.ctypes.data_as(U32P), code.shape[0], imm.ctypes.data_as(U64P) if imm.size else None, imm.size, num_regs)
        g._keep = (code, imm)
        return g


class ProverConfig(C.Structure):
    """qpzk_prover_config (see include/qpzk.h)."""
    _fields_ = [("num_wires", C.c_uint32), ("num_routed", C.c_uint32), ("sigma_col0", C.c_uint32),
                ("quotient_degree_factor", C.c_uint32), ("num_challenges", C.c_uint32), ("num_public_inputs", C.c_uint32),
                ("blinding", C.c_uint32), ("use_graph", C.c_uint32), ("circuit_digest", C.c_uint64 * 4),
                ("fri", FriParams), ("gates", C.POINTER(GateProgram))]


class ProverLayout(C.Structure):
    _fields_ = [("cap_words", C.c_uint64), ("cap_offset", C.c_uint64 * 3), ("openings_offset", C.c_uint64),
                ("num_openings", C.c_uint64), ("fri_offset", C.c_uint64), ("fri_words", C.c_uint64),
                ("total_words", C.c_uint64), ("leaf_lens", C.c_uint32 * 4)]


class ChallengerState(C.Structure):
    _fields_ = [("sponge_state", C.c_uint64 * 12), ("input_buffer", C.c_uint64 * 8),
                ("output_buffer", C.c_uint64 * 8), ("input_len", C.c_uint32), ("output_len", C.c_uint32)]


_lib = None


def load_library():
    """dlopen libqpzk.so and declare every prototype of include/qpzk.h.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QpzkError(-2, "libqpzk.so not built (run `make -C qp-zk-circuits_b200` or __graft_entry__.build()); "
                            "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    sig = {
        "qpzk_last_error": (C.c_char_p, []),
        "qpzk_version": (C.c_char_p, []),
        "qpzk_ctx_create": (i, [i, C.POINTER(vp)]),
        "qpzk_ctx_create_on_stream": (i, [i, vp, C.POINTER(vp)]),
        "qpzk_ctx_destroy": (None, [vp]),
        "qpzk_ctx_synchronize": (i, [vp]),
        "qpzk_ctx_set_zeroize": (i, [vp, i]),
        "qpzk_ctx_set_timing": (i, [vp, i]),
        "qpzk_ctx_set_keep_values": (i, [vp, i]),
        "qpzk_batch_values_dev": (vp, [vp]),
        "qpzk_commit_zs_partial_products": (i, [vp, vp, vp, u32, u32, u32, u32, vp, vp, u32, vp, C.POINTER(vp)]),
        "qpzk_ctx_last_commit_timing": (i, [vp, C.POINTER(C.c_float)]),
        "qpzk_ctx_launch_count": (u64, [vp]),
        "qpzk_ctx_trim": (i, [vp]),
        "qpzk_batches_eval_ext": (i, [vp, vp, u32, vp, vp]),
        "qpzk_pinned_alloc": (vp, [C.c_size_t]),
        "qpzk_pinned_free": (None, [vp]),
        "qpzk_poseidon_permute": (i, [vp, vp, u64, vp]),
        "qpzk_selftest_mds_layer": (i, [vp, vp, u64, u32, vp]),
        "qpzk_poseidon2_permute": (i, [vp, vp, u64, vp]),
        "qpzk_poseidon2_hash_no_pad": (i, [vp, vp, u64, u32, vp]),
        "qpzk_poseidon2_hash_no_pad_dev": (i, [vp, vp, u64, u32, vp]),
        "qpzk_ctx_cached_bytes": (u64, [vp]),
        "qpzk_ctx_device_allocs": (u64, [vp]),
        "qpzk_ctx_peak_bytes": (u64, [vp]),
        "qpzk_ctx_reset_peak": (None, [vp]),
        "qpzk_commit_from_values": (i, [vp, vp, u32, u32, u32, u32, vp, C.POINTER(vp)]),
        "qpzk_commit_from_coeffs": (i, [vp, vp, u32, u32, u32, u32, vp, C.POINTER(vp)]),
        "qpzk_commit_from_values_dev": (i, [vp, vp, u32, u32, u32, u32, vp, C.POINTER(vp)]),
        "qpzk_commit_from_coeffs_dev": (i, [vp, vp, u32, u32, u32, u32, vp, C.POINTER(vp)]),
        "qpzk_commit_cosets_from_coeffs_dev": (i, [vp, vp, u32, u32, u32, u32, u32, u32, C.POINTER(vp)]),
        "qpzk_intt_columns_dev": (i, [vp, vp, vp, u32, u32]),
        "qpzk_commit_cosets_from_coeff_slices_dev": (i, [vp, C.POINTER(vp), u32, u32, u32, u32, u32, u32, u32, u32, C.POINTER(vp)]),
        "qpzk_peer_buffer_create": (i, [vp, u64, C.POINTER(vp), vp]),
        "qpzk_peer_buffer_open": (i, [vp, vp, C.POINTER(vp)]),
        "qpzk_peer_buffer_close": (i, [vp, vp]),
        "qpzk_peer_buffer_free": (i, [vp, vp]),
        "qpzk_merkle_new": (i, [vp, vp, u64, u32, u32, C.POINTER(vp)]),
        "qpzk_merkle_new_dev": (i, [vp, vp, u64, u32, u32, C.POINTER(vp)]),
        "qpzk_batch_free": (None, [vp]),
        "qpzk_batch_num_leaves": (u64, [vp]),
        "qpzk_batch_leaf_len": (u32, [vp]),
        "qpzk_batch_cap_len": (u32, [vp]),
        "qpzk_batch_cap": (i, [vp, vp]),
        "qpzk_batch_coeffs": (i, [vp, vp]),
        "qpzk_batch_rows": (i, [vp, u64, u64, vp]),
        "qpzk_batch_digests": (i, [vp, u32, vp]),
        "qpzk_batch_open": (i, [vp, vp, u32, vp, vp]),
        "qpzk_batch_eval_ext": (i, [vp, vp, vp]),
        "qpzk_batch_leaves_dev": (vp, [vp]),
        "qpzk_batch_coeffs_dev": (vp, [vp]),
        "qpzk_challenger_create": (i, [vp, C.POINTER(vp)]),
        "qpzk_challenger_free": (None, [vp]),
        "qpzk_challenger_observe": (i, [vp, vp, u32]),
        "qpzk_challenger_observe_cap": (i, [vp, vp]),
        "qpzk_challenger_get": (i, [vp, vp, u32]),
        "qpzk_challenger_export": (i, [vp, C.POINTER(ChallengerState)]),
        "qpzk_challenger_import": (i, [vp, C.POINTER(ChallengerState)]),
        "qpzk_fri_params_constant_arity": (i, [C.POINTER(FriParams), u32, u32, u32, u32, u32, u32, u32]),
        "qpzk_fri_proof_words": (C.c_size_t, [C.POINTER(FriParams), u32, u32, U32P]),
        "qpzk_fri_prove": (i, [vp, C.POINTER(vp), u32, C.POINTER(_FriInstance), C.POINTER(FriParams), vp, u64, vp, vp]),
        "qpzk_partial_products": (i, [vp, vp, vp, u32, u32, u32, u32, vp, vp, vp]),
        "qpzk_partial_products_dev": (i, [vp, vp, vp, u32, u32, u32, u32, vp, vp, vp]),
        "qpzk_quotient_polys": (i, [vp, vp, vp, vp, u32, u32, u32, u32, vp, vp, vp, vp, u32, vp]),
        "qpzk_commit_quotient": (i, [vp, vp, vp, vp, u32, u32, u32, u32, vp, vp, vp, vp, u32, u32, vp, C.POINTER(vp)]),
        "qpzk_proof_bytes_len": (C.c_size_t, [C.POINTER(FriParams), C.POINTER(ProofShape)]),
        "qpzk_proof_to_bytes": (i, [C.POINTER(FriParams), C.POINTER(ProofShape), vp, vp, vp, vp, vp, vp, vp, C.c_size_t]),
        "qpzk_proof_from_bytes": (i, [C.POINTER(FriParams), C.POINTER(ProofShape), vp, C.c_size_t, vp, vp, vp, vp, vp, vp]),
        "qpzk_quotient_polys_prog": (i, [vp, vp, vp, vp, u32, u32, u32, u32, vp, vp, vp, C.POINTER(GateProgram), vp]),
        "qpzk_commit_quotient_prog": (i, [vp, vp, vp, vp, u32, u32, u32, u32, vp, vp, vp, C.POINTER(GateProgram), u32, vp, C.POINTER(vp)]),
        "qpzk_prover_create": (i, [vp, vp, C.POINTER(ProverConfig), C.POINTER(vp)]),
        "qpzk_prover_free": (None, [vp]),
        "qpzk_prover_get_layout": (i, [vp, C.POINTER(ProverLayout)]),
        "qpzk_prover_prove_async": (i, [vp, vp, vp, vp, u64]),
        "qpzk_prover_finish": (i, [vp, vp]),
        "qpzk_prover_prove": (i, [vp, vp, vp, vp, u64, vp]),
        "qpzk_prover_launches_per_proof": (u64, [vp]),
        "qpzk_prover_uses_graph": (i, [vp]),
        "qpzk_lde_pass_plan": (i, [u32, u32, i, vp, vp, vp]),
        "qpzk_selftest_field_ops": (i, [vp, u32, vp, vp, vp, u64, vp]),
        "qpzk_gate_jit_available": (i, []),
        "qpzk_gate_jit_set_mode": (None, [i]),
        "qpzk_gate_jit_compile_check": (i, [vp, vp]),
        "qpzk_gate_jit_compiled": (u64, []),
        "qpzk_bench_imad": (i, [vp, u32, C.POINTER(C.c_double)]),
        "qpzk_bench_poseidon": (i, [vp, u32, C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)  # AttributeError here = header/library mismatch
        f.restype, f.argtypes = res, args
    L._qpzk_symbols = sorted(sig)
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise QpzkError(rc, load_library().qpzk_last_error().decode())


def _host(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(C.c_void_p)


class Context:
    """One device + one CUDA stream (qpzk_ctx).  `stream` may be a raw cudaStream_t (int), e.g.
    torch.cuda.current_stream().cuda_stream, so that torch events on that stream time the kernels."""

    def __init__(self, device=0, stream=None):
        L = load_library()
        h = C.c_void_p()
        if stream is None:
            _check(L.qpzk_ctx_create(device, C.byref(h)))
        else:
            _check(L.qpzk_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h)))
        self.h, self.device, self.L = h, device, L

    def synchronize(self):
        _check(self.L.qpzk_ctx_synchronize(self.h))

    def set_zeroize(self, mode):
        """0 never, 1 (default) when a buffer is released, 2 deferred until the block leaves the library (trim / destroy)."""
        _check(self.L.qpzk_ctx_set_zeroize(self.h, int(mode)))

    def set_keep_values(self, on):
        """from_values batches also keep their input values on the device (needed by commit_zs_partial_products)."""
        _check(self.L.qpzk_ctx_set_keep_values(self.h, int(on)))

    def set_timing(self, on):
        _check(self.L.qpzk_ctx_set_timing(self.h, int(on)))

    def last_commit_timing(self):
        ms = (C.c_float * 4)()
        _check(self.L.qpzk_ctx_last_commit_timing(self.h, ms))
        return dict(zip(("intt_ms", "lde_ms", "leaf_hash_ms", "merkle_levels_ms"), [float(x) for x in ms]))

    def launch_count(self):
        return int(self.L.qpzk_ctx_launch_count(self.h))

    def trim(self):
        """Return the cached device blocks to the driver."""
        _check(self.L.qpzk_ctx_trim(self.h))

    def device_allocs(self):
        return int(self.L.qpzk_ctx_device_allocs(self.h))

    def cached_bytes(self):
        return int(self.L.qpzk_ctx_cached_bytes(self.h))

    def peak_bytes(self, reset=False):
        """High-water mark of the device memory held by this context (in use + cached), in bytes."""
        v = int(self.L.qpzk_ctx_peak_bytes(self.h))
        if reset:
            self.L.qpzk_ctx_reset_peak(self.h)
        return v

    def poseidon_permute(self, states, mds_only_round=None):
        """PoseidonPermutation::permute on an (N, 12) array; inputs may be non-canonical, outputs are canonical.
        mds_only_round = r applies only the linear layer + constants of round r (self-test hook)."""
        a = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1, 12)
        out = np.zeros_like(a)
        if mds_only_round is None:
            _check(self.L.qpzk_poseidon_permute(self.h, a.ctypes.data_as(C.c_void_p), a.shape[0], out.ctypes.data_as(C.c_void_p)))
        else:
            _check(self.L.qpzk_selftest_mds_layer(self.h, a.ctypes.data_as(C.c_void_p), a.shape[0], mds_only_round,
                                                  out.ctypes.data_as(C.c_void_p)))
        return out

    def poseidon2_permute(self, states):
        """The Poseidon2 permutation of the reference's application hash on an (N, 12) array."""
        a = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1, 12)
        out = np.zeros_like(a)
        _check(self.L.qpzk_poseidon2_permute(self.h, a.ctypes.data_as(C.c_void_p), a.shape[0], out.ctypes.data_as(C.c_void_p)))
        return out

    def poseidon2_hash_no_pad(self, msgs):
        """Poseidon2Hash::hash_no_pad of every row of an (N, len) array of canonical felts -> (N, 4)."""
        a = np.ascontiguousarray(msgs, dtype=np.uint64)
        if a.ndim == 1:
            a = a.reshape(1, -1)
        out = np.zeros((a.shape[0], 4), np.uint64)
        _check(self.L.qpzk_poseidon2_hash_no_pad(self.h, a.ctypes.data_as(C.c_void_p) if a.size else None, a.shape[0], a.shape[1],
                                                 out.ctypes.data_as(C.c_void_p)))
        return out

    def partial_products(self, wires, sigmas, betas, gammas, chunk=8):
        """all_wires_permutation_partial_products (qp-plonky2 plonk/prover.rs): wires, sigmas (num_routed, n) values on
        the subgroup -> (len(betas) * ceil(num_routed / chunk), n): Z of every challenge, then the partial products of
        challenge 0, 1, ... -- the columns of the zs_partial_products commitment."""
        w, wp = _host(wires)
        sg, sp = _host(sigmas)
        if w.ndim != 2 or sg.shape != w.shape or w.shape[1] & (w.shape[1] - 1) or not w.shape[1]:
            raise QpzkError(-1, "wires and sigmas must be (num_routed, 2^k) arrays of the same shape")
        b, bp = _host(betas)
        g, gp = _host(gammas)
        if b.ndim != 1 or b.shape != g.shape:
            raise QpzkError(-1, "betas and gammas must have the same length")
        R, n = w.shape
        out = np.zeros((b.size * ((R + chunk - 1) // chunk), n), np.uint64)
        _check(self.L.qpzk_partial_products(self.h, wp, sp, n.bit_length() - 1, R, chunk, b.size, bp, gp,
                                            out.ctypes.data_as(C.c_void_p)))
        return out

    def partial_products_device(self, wires_ptr, sigmas_ptr, log_n, num_routed, betas, gammas, out_ptr, chunk=8):
        """The same on raw device addresses (wires / sigmas: num_routed x 2^log_n column-major; out: Z and partial
        product columns, ready for PolynomialBatch.from_device)."""
        b, bp = _host(betas)
        g, gp = _host(gammas)
        _check(self.L.qpzk_partial_products_dev(self.h, C.c_void_p(wires_ptr), C.c_void_p(sigmas_ptr), log_n, num_routed,
                                                chunk, b.size, bp, gp, C.c_void_p(out_ptr)))

    def selftest_field_ops(self, op, a, b=None, c=None):
        """qpzk_selftest_field_ops: one device field operation per element (a test hook, see include/qpzk.h)."""
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1)
        b = np.zeros_like(a) if b is None else np.ascontiguousarray(b, dtype=np.uint64).reshape(-1)
        c = np.zeros_like(a) if c is None else np.ascontiguousarray(c, dtype=np.uint64).reshape(-1)
        out = np.empty_like(a)
        _check(self.L.qpzk_selftest_field_ops(self.h, op, a.ctypes.data, b.ctypes.data, c.ctypes.data, a.size, out.ctypes.data))
        return out

    def bench_imad(self, iters=4096):
        v = C.c_double()
        _check(self.L.qpzk_bench_imad(self.h, iters, C.byref(v)))
        return v.value

    def bench_poseidon(self, perms=64):
        v = C.c_double()
        _check(self.L.qpzk_bench_poseidon(self.h, perms, C.byref(v)))
        return v.value

    def close(self):
        if self.h:
            self.L.qpzk_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PeerBuffers:
    """The exchange buffers of a multi-GPU commitment (one process per GPU): this rank's buffer, exported as a CUDA IPC
    handle, and every peer's buffer mapped into this process.  `exchange(handles)` takes the list of all ranks' handle
    bytes in rank order (gathered by the caller, e.g. torch.distributed.all_gather_object)."""

    def __init__(self, ctx, words):
        self.ctx, self.words = ctx, words
        p = C.c_void_p()
        hb = (C.c_uint8 * 64)()
        _check(ctx.L.qpzk_peer_buffer_create(ctx.h, words, C.byref(p), hb))
        self.ptr, self.handle, self.peers, self._opened = p.value, bytes(hb), None, []

    def exchange(self, handles, rank):
        self.peers = []
        for r, hb in enumerate(handles):
            if r == rank:
                self.peers.append(self.ptr)
                continue
            p = C.c_void_p()
            buf = (C.c_uint8 * 64).from_buffer_copy(hb)
            _check(self.ctx.L.qpzk_peer_buffer_open(self.ctx.h, buf, C.byref(p)))
            self.peers.append(p.value)
            self._opened.append(p.value)
        return self.peers

    def close(self):
        for p in self._opened:
            self.ctx.L.qpzk_peer_buffer_close(self.ctx.h, C.c_void_p(p))
        self._opened = []
        if self.ptr:
            self.ctx.L.qpzk_peer_buffer_free(self.ctx.h, C.c_void_p(self.ptr))
            self.ptr = None


class MerkleTree:
    """MerkleTree<F, PoseidonHash>: `.cap`, `.prove(i)`, `.get(i)`."""

    def __init__(self, ctx, handle, owner=None):
        self.ctx, self.h, self._owner = ctx, handle, owner
        L = ctx.L
        self.num_leaves = int(L.qpzk_batch_num_leaves(handle))
        self.leaf_len = int(L.qpzk_batch_leaf_len(handle))
        self.cap_len = int(L.qpzk_batch_cap_len(handle))
        self.cap_height = self.cap_len.bit_length() - 1

    @classmethod
    def new(cls, ctx, leaves, cap_height):
        """leaves: (num_leaves, leaf_len) uint64, row-major."""
        a = np.ascontiguousarray(leaves, dtype=np.uint64)
        if a.ndim != 2:
            raise QpzkError(-1, "leaves must be a 2-D array")
        h = C.c_void_p()
        _check(ctx.L.qpzk_merkle_new(ctx.h, a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1], cap_height, C.byref(h)))
        t = cls(ctx, h)
        t._free = True
        return t

    @property
    def cap(self):
        out = np.zeros((self.cap_len, 4), np.uint64)
        _check(self.ctx.L.qpzk_batch_cap(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def digests(self, level=0):
        out = np.zeros((self.num_leaves >> level, 4), np.uint64)
        _check(self.ctx.L.qpzk_batch_digests(self.h, level, out.ctypes.data_as(C.c_void_p)))
        return out

    def get(self, i):
        return self.rows(i, 1)[0]

    def rows(self, first, count):
        out = np.zeros((count, self.leaf_len), np.uint64)
        _check(self.ctx.L.qpzk_batch_rows(self.h, first, count, out.ctypes.data_as(C.c_void_p)))
        return out

    def open(self, indices):
        """(rows, siblings) for several leaves: MerkleTree::prove + the leaf data."""
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        ns = (self.num_leaves.bit_length() - 1) - self.cap_height
        rows = np.zeros((idx.size, self.leaf_len), np.uint64)
        sib = np.zeros((idx.size, ns, 4), np.uint64)
        _check(self.ctx.L.qpzk_batch_open(self.h, idx.ctypes.data_as(C.c_void_p), idx.size,
                                          rows.ctypes.data_as(C.c_void_p), sib.ctypes.data_as(C.c_void_p)))
        return rows, sib

    def prove(self, leaf_index):
        return self.open([leaf_index])[1][0]

    def free(self):
        if getattr(self, "_free", False) and self.h:
            self.ctx.L.qpzk_batch_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PolynomialBatch:
    """PolynomialBatch<F, C, D>: coefficients + salted LDE leaves + Merkle tree, all device resident."""

    def __init__(self, ctx, handle, log_n, num_cols, rate_bits, cap_height, salted):
        self.ctx, self.h = ctx, handle
        self.degree_log, self.num_cols, self.rate_bits, self.cap_height = log_n, num_cols, rate_bits, cap_height
        self.blinding = salted
        self.merkle_tree = MerkleTree(ctx, handle, owner=self)
        self.leaf_len = self.merkle_tree.leaf_len

    @classmethod
    def _commit(cls, ctx, fn, cols, rate_bits, salt, cap_height):
        a = np.ascontiguousarray(cols, dtype=np.uint64)
        if a.ndim != 2:
            raise QpzkError(-1, "expected a (num_polys, n) array")
        num_cols, n = a.shape
        if n == 0 or n & (n - 1):
            raise QpzkError(-1, "polynomial length must be a power of two")
        log_n = n.bit_length() - 1
        sp = None
        if salt is not None:
            s = np.ascontiguousarray(salt, dtype=np.uint64)
            if s.shape != (SALT_SIZE, n << rate_bits):
                raise QpzkError(-1, "salt must have shape (4, n << rate_bits)")
            sp = s.ctypes.data_as(C.c_void_p)
        h = C.c_void_p()
        _check(fn(ctx.h, a.ctypes.data_as(C.c_void_p), log_n, num_cols, rate_bits, cap_height, sp, C.byref(h)))
        return cls(ctx, h, log_n, num_cols, rate_bits, cap_height, salt is not None)

    @classmethod
    def from_values(cls, ctx, values, rate_bits, blinding_salt, cap_height):
        return cls._commit(ctx, ctx.L.qpzk_commit_from_values, values, rate_bits, blinding_salt, cap_height)

    @classmethod
    def from_coeffs(cls, ctx, polynomials, rate_bits, blinding_salt, cap_height):
        return cls._commit(ctx, ctx.L.qpzk_commit_from_coeffs, polynomials, rate_bits, blinding_salt, cap_height)

    @classmethod
    def from_device(cls, ctx, dev_ptr, log_n, num_cols, rate_bits, cap_height, values=True, salt_ptr=None):
        """Inputs already resident in HBM (dev_ptr = raw device address, e.g. tensor.data_ptr())."""
        fn = ctx.L.qpzk_commit_from_values_dev if values else ctx.L.qpzk_commit_from_coeffs_dev
        h = C.c_void_p()
        _check(fn(ctx.h, C.c_void_p(dev_ptr), log_n, num_cols, rate_bits, cap_height,
                  C.c_void_p(salt_ptr) if salt_ptr else None, C.byref(h)))
        return cls(ctx, h, log_n, num_cols, rate_bits, cap_height, salt_ptr is not None)

    @classmethod
    def cosets_from_coeffs_device(cls, ctx, dev_ptr, log_n, num_cols, rate_bits, cap_height, coset_begin, coset_count):
        h = C.c_void_p()
        _check(ctx.L.qpzk_commit_cosets_from_coeffs_dev(ctx.h, C.c_void_p(dev_ptr), log_n, num_cols, rate_bits,
                                                        cap_height, coset_begin, coset_count, C.byref(h)))
        return cls(ctx, h, log_n, num_cols, rate_bits, cap_height, False)

    @classmethod
    def cosets_from_coeff_slices_device(cls, ctx, slice_ptrs, cols_per_slice, log_n, num_cols, rate_bits, cap_height,
                                        coset_begin, coset_count):
        """Coset-split commit whose coefficient columns live in several device buffers (own + peers mapped over NVLink,
        see PeerBuffers): the first LDE pass pulls them where they lie."""
        arr = (C.c_void_p * len(slice_ptrs))(*[C.c_void_p(p) for p in slice_ptrs])
        h = C.c_void_p()
        _check(ctx.L.qpzk_commit_cosets_from_coeff_slices_dev(ctx.h, arr, len(slice_ptrs), cols_per_slice, log_n, num_cols,
                                                              rate_bits, cap_height, coset_begin, coset_count, C.byref(h)))
        return cls(ctx, h, log_n, num_cols, rate_bits, cap_height, False)

    @staticmethod
    def _quotient_args(constants_sigmas, wires, zs_partial_products, sigma_col0, num_routed, betas, gammas, alphas,
                       gate_terms_dev, num_gate_terms):
        b, bp = _host(betas)
        g, gp = _host(gammas)
        al, ap = _host(alphas)
        if not (b.ndim == g.ndim == al.ndim == 1 and b.size == g.size == al.size):
            raise QpzkError(-1, "betas, gammas and alphas must have the same length")
        keep = (b, g, al)
        return keep, (wires.ctx.h, constants_sigmas.h, wires.h, zs_partial_products.h, sigma_col0, num_routed), \
            (b.size, bp, gp, ap, C.c_void_p(gate_terms_dev) if gate_terms_dev else None, num_gate_terms)

    @classmethod
    def commit_zs_partial_products(cls, constants_sigmas, wires, sigma_col0, num_routed, betas, gammas, cap_height,
                                   quotient_degree_factor=8, salt_ptr=None):
        """all_wires_permutation_partial_products + the zs/partial-products commitment on the device, from two batches
        committed with Context.set_keep_values(True).  salt_ptr: device address of the (4, n << rate_bits) blinding
        columns (zero-knowledge circuits blind this oracle), or None."""
        b, bp = _host(betas)
        g, gp = _host(gammas)
        h = C.c_void_p()
        ctx = wires.ctx
        _check(ctx.L.qpzk_commit_zs_partial_products(ctx.h, constants_sigmas.h, wires.h, sigma_col0, num_routed,
                                                     quotient_degree_factor, b.size, bp, gp, cap_height,
                                                     C.c_void_p(salt_ptr) if salt_ptr else None, C.byref(h)))
        cols = b.size * ((num_routed + quotient_degree_factor - 1) // quotient_degree_factor)
        return cls(ctx, h, wires.degree_log, cols, wires.rate_bits, cap_height, salt_ptr is not None)

    @classmethod
    def quotient_polys(cls, constants_sigmas, wires, zs_partial_products, sigma_col0, num_routed, betas, gammas, alphas,
                       quotient_degree_factor=8, gate_terms_dev=None, num_gate_terms=0):
        """compute_quotient_polys for the permutation argument (+ caller-supplied gate terms, a raw device address):
        returns the (len(alphas) * quotient_degree_factor, n) quotient chunk coefficients."""
        keep, head, tail = cls._quotient_args(constants_sigmas, wires, zs_partial_products, sigma_col0, num_routed,
                                              betas, gammas, alphas, gate_terms_dev, num_gate_terms)
        out = np.zeros((tail[0] * quotient_degree_factor, 1 << wires.degree_log), np.uint64)
        _check(wires.ctx.L.qpzk_quotient_polys(*head, quotient_degree_factor, *tail, out.ctypes.data_as(C.c_void_p)))
        return out

    @classmethod
    def commit_quotient(cls, constants_sigmas, wires, zs_partial_products, sigma_col0, num_routed, betas, gammas, alphas,
                        cap_height, quotient_degree_factor=8, gate_terms_dev=None, num_gate_terms=0, salt_ptr=None):
        """The same, committed (from_coeffs) without leaving the device."""
        keep, head, tail = cls._quotient_args(constants_sigmas, wires, zs_partial_products, sigma_col0, num_routed,
                                              betas, gammas, alphas, gate_terms_dev, num_gate_terms)
        h = C.c_void_p()
        _check(wires.ctx.L.qpzk_commit_quotient(*head, quotient_degree_factor, *tail, cap_height,
                                                C.c_void_p(salt_ptr) if salt_ptr else None, C.byref(h)))
        return cls(wires.ctx, h, wires.degree_log, tail[0] * quotient_degree_factor, wires.rate_bits, cap_height, salt_ptr is not None)

    @classmethod
    def quotient_polys_prog(cls, constants_sigmas, wires, zs_partial_products, sigma_col0, num_routed, betas, gammas, alphas,
                            program, quotient_degree_factor=8):
        """quotient_polys with the gate terms computed on the device from a GateProgram (None = no gates)."""
        keep, head, tail = cls._quotient_args(constants_sigmas, wires, zs_partial_products, sigma_col0, num_routed,
                                              betas, gammas, alphas, None, 0)
        out = np.zeros((tail[0] * quotient_degree_factor, 1 << wires.degree_log), np.uint64)
        _check(wires.ctx.L.qpzk_quotient_polys_prog(*head, quotient_degree_factor, *tail[:4],
                                                    C.byref(program) if program is not None else None,
                                                    out.ctypes.data_as(C.c_void_p)))
        return out

    @classmethod
    def commit_quotient_prog(cls, constants_sigmas, wires, zs_partial_products, sigma_col0, num_routed, betas, gammas, alphas,
                             program, cap_height, quotient_degree_factor=8, salt_ptr=None):
        keep, head, tail = cls._quotient_args(constants_sigmas, wires, zs_partial_products, sigma_col0, num_routed,
                                              betas, gammas, alphas, None, 0)
        h = C.c_void_p()
        _check(wires.ctx.L.qpzk_commit_quotient_prog(*head, quotient_degree_factor, *tail[:4],
                                                     C.byref(program) if program is not None else None, cap_height,
                                                     C.c_void_p(salt_ptr) if salt_ptr else None, C.byref(h)))
        return cls(wires.ctx, h, wires.degree_log, tail[0] * quotient_degree_factor, wires.rate_bits, cap_height, salt_ptr is not None)

    @property
    def polynomials(self):
        out = np.zeros((self.num_cols, 1 << self.degree_log), np.uint64)
        _check(self.ctx.L.qpzk_batch_coeffs(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    @property
    def cap(self):
        return self.merkle_tree.cap

    def get_lde_values(self, index, step=1):
        """Upstream get_lde_values(index, step): the row of LDE point index*step, without the salt."""
        log_L = self.degree_log + self.rate_bits
        i = index * step
        r = int(format(i, "0%db" % log_L)[::-1], 2) if log_L else 0
        return self.merkle_tree.get(r)[:self.num_cols]

    def eval_ext(self, point):
        pt, pp = _host(point)
        out = np.zeros((self.num_cols, 2), np.uint64)
        _check(self.ctx.L.qpzk_batch_eval_ext(self.h, pp, out.ctypes.data_as(C.c_void_p)))
        return out

    @staticmethod
    def eval_ext_many(batches, point):
        """eval_ext of several batches of equal degree at one point: one launch, one copy; returns (sum cols, 2)."""
        ctx = batches[0].ctx
        pt, pp = _host(point)
        arr = (C.c_void_p * len(batches))(*[b.h for b in batches])
        out = np.zeros((sum(b.num_cols for b in batches), 2), np.uint64)
        _check(ctx.L.qpzk_batches_eval_ext(ctx.h, arr, len(batches), pp, out.ctypes.data_as(C.c_void_p)))
        return out

    @staticmethod
    def prove_openings(instance, oracles, challenger, fri_params, pow_witness=None, return_openings=False):
        """PolynomialBatch::prove_openings -> flat FriProof (layout in include/qpzk.h)."""
        ctx = oracles[0].ctx
        L = ctx.L
        arr = (C.c_void_p * len(oracles))(*[o.h for o in oracles])
        lens = np.array([o.leaf_len for o in oracles], np.uint32)
        words = L.qpzk_fri_proof_words(C.byref(fri_params), oracles[0].degree_log, len(oracles),
                                       lens.ctypes.data_as(U32P))
        proof = np.zeros(words, np.uint64)
        openings = np.zeros((int(instance.sizes.sum()), 2), np.uint64)
        w = POW_SEARCH if pow_witness is None else pow_witness
        _check(L.qpzk_fri_prove(ctx.h, arr, len(oracles), C.byref(instance.c), C.byref(fri_params), challenger.h, w,
                                proof.ctypes.data_as(C.c_void_p), openings.ctypes.data_as(C.c_void_p)))
        return (proof, openings) if return_openings else proof

    def free(self):
        if self.h:
            self.ctx.L.qpzk_batch_free(self.h)
            self.h = None
            self.merkle_tree.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Prover:
    """qpzk_prover: `prove()` after witness generation as one device-resident call (CUDA graph from the third proof).

    constants_sigmas: the circuit's PolynomialBatch, committed with from_values under Context.set_keep_values(True).
    prove(wires, public_inputs) -> dict(caps=[wires, zs_pp, quotient], openings=(k, 2), fri_proof=flat words)."""

    def __init__(self, constants_sigmas, num_wires, num_routed, sigma_col0, fri_params, circuit_digest, num_public_inputs,
                 gate_program=None, quotient_degree_factor=8, num_challenges=2, blinding=False, use_graph=True):
        self.ctx = constants_sigmas.ctx
        self._keep = (constants_sigmas, gate_program)
        cfg = ProverConfig(num_wires, num_routed, sigma_col0, quotient_degree_factor, num_challenges, num_public_inputs,
                           int(blinding), int(use_graph))
        for k, v in enumerate(np.asarray(circuit_digest, dtype=np.uint64).reshape(4)):
            cfg.circuit_digest[k] = int(v)
        cfg.fri = fri_params
        cfg.gates = C.pointer(gate_program) if gate_program is not None else None
        h = C.c_void_p()
        _check(self.ctx.L.qpzk_prover_create(self.ctx.h, constants_sigmas.h, C.byref(cfg), C.byref(h)))
        self.h = h
        self.layout = ProverLayout()
        _check(self.ctx.L.qpzk_prover_get_layout(h, C.byref(self.layout)))
        self.num_wires, self.n = num_wires, 1 << constants_sigmas.degree_log
        self.num_public_inputs, self.blinding = num_public_inputs, blinding
        self._pending = None

    def prove_async(self, wires, public_inputs, salts=None, pow_witness=None):
        w = np.ascontiguousarray(wires, dtype=np.uint64)
        if w.shape != (self.num_wires, self.n):
            raise QpzkError(-1, "wires must have shape (%d, %d)" % (self.num_wires, self.n))
        pi = np.ascontiguousarray(public_inputs, dtype=np.uint64).reshape(-1)
        if pi.size != self.num_public_inputs:
            raise QpzkError(-1, "expected %d public inputs" % self.num_public_inputs)
        sp = None
        if salts is not None:
            sl = np.ascontiguousarray(salts, dtype=np.uint64)
            sp = sl.ctypes.data_as(C.c_void_p)
        _check(self.ctx.L.qpzk_prover_prove_async(self.h, w.ctypes.data_as(C.c_void_p), pi.ctypes.data_as(C.c_void_p) if pi.size else None,
                                                  sp, POW_SEARCH if pow_witness is None else pow_witness))

    def finish(self):
        lay = self.layout
        out = np.zeros(lay.total_words, np.uint64)
        _check(self.ctx.L.qpzk_prover_finish(self.h, out.ctypes.data_as(C.c_void_p)))
        H = lay.cap_words
        caps = [out[lay.cap_offset[k]:lay.cap_offset[k] + H].reshape(-1, 4).copy() for k in range(3)]
        op = out[lay.openings_offset:lay.openings_offset + 2 * lay.num_openings].reshape(-1, 2).copy()
        return dict(caps=caps, openings=op, fri_proof=out[lay.fri_offset:lay.fri_offset + lay.fri_words].copy())

    def prove(self, wires, public_inputs, salts=None, pow_witness=None):
        self.prove_async(wires, public_inputs, salts, pow_witness)
        return self.finish()

    @property
    def launches_per_proof(self):
        return int(self.ctx.L.qpzk_prover_launches_per_proof(self.h))

    @property
    def uses_graph(self):
        return bool(self.ctx.L.qpzk_prover_uses_graph(self.h))

    def free(self):
        if self.h:
            self.ctx.L.qpzk_prover_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class FriInstanceInfo:
    """batches = [(point (2 words), [(oracle_index, polynomial_index), ...]), ...]"""

    def __init__(self, batches):
        self.points = np.array([w for pt, _ in batches for w in pt], np.uint64)
        self.sizes = np.array([len(ps) for _, ps in batches], np.uint32)
        self.po = np.array([o for _, ps in batches for o, _ in ps], np.uint32)
        self.pi = np.array([i for _, ps in batches for _, i in ps], np.uint32)
        self.c = _FriInstance(len(batches), self.points.ctypes.data_as(U64P), self.sizes.ctypes.data_as(U32P),
                              self.po.ctypes.data_as(U32P), self.pi.ctypes.data_as(U32P))


def fri_params(log_n, rate_bits=3, cap_height=4, proof_of_work_bits=16, num_query_rounds=28, arity_bits=4,
               final_poly_bits=5):
    """FriConfig + ConstantArityBits(4, 5) as used by the reference (wormhole/circuit/src/profile.rs:168-174)."""
    p = FriParams()
    _check(load_library().qpzk_fri_params_constant_arity(C.byref(p), log_n, rate_bits, cap_height, proof_of_work_bits,
                                                         num_query_rounds, arity_bits, final_poly_bits))
    return p


def hash_no_pad(ctx, elements):
    """PoseidonHash::hash_no_pad on the device: overwrite-mode sponge of rate 8, no padding, first 4 state words
    (upstream hashing.rs; `public_inputs_hash = hash_no_pad(public_inputs)` in prove())."""
    x = np.ascontiguousarray(elements, dtype=np.uint64).reshape(-1)
    state = np.zeros(12, np.uint64)
    for off in range(0, x.size, 8):                  # no input, no permutation: the hash of nothing is four zeros
        chunk = x[off:off + 8]
        state[:chunk.size] = chunk
        state = ctx.poseidon_permute(state.reshape(1, 12))[0]
    return state[:4].copy()


def proof_shape(log_n, num_constants, num_sigmas, num_wires, num_zs, num_partial_products, num_quotient_polys,
                num_public_inputs, salted=(False, False, False, False)):
    s = ProofShape(log_n, num_constants, num_sigmas, num_wires, num_zs, num_partial_products, num_quotient_polys)
    lens = [num_constants + num_sigmas, num_wires, num_zs + num_partial_products, num_quotient_polys]
    for k in range(4):
        s.leaf_lens[k] = lens[k] + (SALT_SIZE if salted[k] else 0)
    s.num_public_inputs = num_public_inputs
    return s


def proof_to_bytes(params, shape, wires_cap, zs_cap, quotient_cap, openings, fri_proof, public_inputs):
    """ProofWithPublicInputs::to_bytes from the pieces the prover returns (openings in FRI-instance order)."""
    L = load_library()
    n = L.qpzk_proof_bytes_len(C.byref(params), C.byref(shape))
    if not n:
        raise QpzkError(-1, L.qpzk_last_error().decode())
    arrs = [np.ascontiguousarray(a, dtype=np.uint64).reshape(-1) for a in (wires_cap, zs_cap, quotient_cap, openings, fri_proof, public_inputs)]
    H = 4 << params.cap_height
    want = [H, H, H, 2 * (shape.num_constants + shape.num_sigmas + shape.num_wires + 2 * shape.num_zs +
                          shape.num_partial_products + shape.num_quotient_polys),
            L.qpzk_fri_proof_words(C.byref(params), shape.log_n, 4, shape.leaf_lens), shape.num_public_inputs]
    if [a.size for a in arrs] != want:
        raise QpzkError(-1, "proof pieces have sizes %s, the circuit needs %s" % ([a.size for a in arrs], want))
    out = np.zeros(n, np.uint8)
    _check(L.qpzk_proof_to_bytes(C.byref(params), C.byref(shape), *[a.ctypes.data_as(C.c_void_p) for a in arrs],
                                 out.ctypes.data_as(C.c_void_p), n))
    return out.tobytes()


def proof_from_bytes(params, shape, data):
    """The inverse: dict(wires_cap, zs_cap, quotient_cap, openings (k, 2), fri_proof, public_inputs)."""
    L = load_library()
    H = 1 << params.cap_height
    k = (shape.num_constants + shape.num_sigmas + shape.num_wires + 2 * shape.num_zs + shape.num_partial_products +
         shape.num_quotient_polys)
    out = dict(wires_cap=np.zeros((H, 4), np.uint64), zs_cap=np.zeros((H, 4), np.uint64), quotient_cap=np.zeros((H, 4), np.uint64),
               openings=np.zeros((k, 2), np.uint64),
               fri_proof=np.zeros(L.qpzk_fri_proof_words(C.byref(params), shape.log_n, 4, shape.leaf_lens), np.uint64),
               public_inputs=np.zeros(shape.num_public_inputs, np.uint64))
    buf = np.frombuffer(bytes(data), np.uint8)
    _check(L.qpzk_proof_from_bytes(C.byref(params), C.byref(shape), buf.ctypes.data_as(C.c_void_p), buf.size,
                                   *[out[f].ctypes.data_as(C.c_void_p) for f in ("wires_cap", "zs_cap", "quotient_cap", "openings",
                                                                                 "fri_proof", "public_inputs")]))
    return out


class Challenger:
    """Challenger<F, PoseidonHash>; the sponge lives on the device so that caps never travel to the host."""

    def __init__(self, ctx):
        self.ctx = ctx
        h = C.c_void_p()
        _check(ctx.L.qpzk_challenger_create(ctx.h, C.byref(h)))
        self.h = h

    def observe_elements(self, elems):
        a, p = _host(np.asarray(elems, dtype=np.uint64).reshape(-1))
        _check(self.ctx.L.qpzk_challenger_observe(self.h, p, a.size))

    observe_element = lambda self, x: self.observe_elements([x])
    observe_hash = observe_elements
    observe_extension_elements = observe_elements

    def observe_cap(self, cap_or_batch):
        if isinstance(cap_or_batch, (PolynomialBatch, MerkleTree)):
            _check(self.ctx.L.qpzk_challenger_observe_cap(self.h, cap_or_batch.h))
        else:
            self.observe_elements(cap_or_batch)

    def get_n_challenges(self, n):
        out = np.zeros(n, np.uint64)
        _check(self.ctx.L.qpzk_challenger_get(self.h, out.ctypes.data_as(C.c_void_p), n))
        return out

    def get_challenge(self):
        return int(self.get_n_challenges(1)[0])

    def get_extension_challenge(self):
        return self.get_n_challenges(2)

    def export_state(self):
        s = ChallengerState()
        _check(self.ctx.L.qpzk_challenger_export(self.h, C.byref(s)))
        return (list(s.sponge_state), list(s.input_buffer)[:s.input_len], list(s.output_buffer)[:s.output_len])

    def import_state(self, sponge_state, input_buffer=(), output_buffer=()):
        s = ChallengerState()
        for k, v in enumerate(sponge_state):
            s.sponge_state[k] = v
        for k, v in enumerate(input_buffer):
            s.input_buffer[k] = v
        for k, v in enumerate(output_buffer):
            s.output_buffer[k] = v
        s.input_len, s.output_len = len(input_buffer), len(output_buffer)
        _check(self.ctx.L.qpzk_challenger_import(self.h, C.byref(s)))

    def free(self):
        if self.h:
            self.ctx.L.qpzk_challenger_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
