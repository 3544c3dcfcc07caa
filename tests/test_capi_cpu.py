"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/qpzk.h declares, its pure
host functions agree with the oracle, and the product fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

P = 0xFFFFFFFF00000001


def _declared_symbols(header):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qpzk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(qpzk):
    lib = qpzk.load_library()
    declared = _declared_symbols(qpzk.HEADER_PATH)
    assert len(declared) >= 40
    raw = C.CDLL(qpzk.LIB_PATH)
    missing = [s for s in declared if not hasattr(raw, s)]
    assert not missing, missing
    # and the Python mirror binds exactly the declared set
    assert sorted(lib._qpzk_symbols) == declared


def test_rust_sys_crate_binds_every_declared_symbol(qpzk):
    """rust/qpzk-sys/src/lib.rs (the -sys crate a qp-plonky2 fork links) declares every entry point of
    include/qpzk.h.  Source-level check: there is no cargo/rustc in this image."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rs = open(os.path.join(root, "rust", "qpzk-sys", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (qpzk_[a-z0-9_]+)\s*\(", rs))
    declared = set(_declared_symbols(qpzk.HEADER_PATH))
    assert declared - bound == set(), sorted(declared - bound)
    assert bound - declared == set(), sorted(bound - declared)


def test_version_string(qpzk):
    assert b"sm_100a" in qpzk.load_library().qpzk_version()


def test_fri_params_and_proof_size_match_oracle(qpzk, orc):
    for log_n in (5, 6, 10, 13, 15, 16, 20):
        for fpb in (2, 5):
            g = qpzk.fri_params(log_n, 3, 4, 16, 28, 4, fpb)
            o = orc.fri_params(log_n, 3, 4, 16, 28, 4, fpb)
            assert g.reduction_arity_bits == list(o.arity_bits)[:o.num_rounds]
            lens = np.array([85, 135, 24, 20], np.uint32)
            words = qpzk.load_library().qpzk_fri_proof_words(C.byref(g), log_n, 4, lens.ctypes.data_as(qpzk.U32P))
            assert words == orc.fri_proof_words(o, log_n, lens)
    with pytest.raises(qpzk.QpzkError):
        qpzk.fri_params(10, arity_bits=5)


def test_challenger_state_layout_matches_header(qpzk, orc):
    assert C.sizeof(qpzk.ChallengerState) == C.sizeof(orc.ChallengerState) == 12 * 8 + 8 * 8 + 8 * 8 + 8


def test_gate_program_compiles_for_sm_100a_without_a_gpu(qpzk, tmp_path, monkeypatch):
    """The run-time compiler path of gate programs (csrc/gate_jit.cu): source generation + NVRTC for sm_100a need no GPU.
    A small arithmetic program (every opcode) compiles to a cubin, lands in the disk cache and is served from there."""
    import ctypes as C
    lib = qpzk.load_library()
    if not lib.qpzk_gate_jit_available():
        pytest.skip("libnvrtc.so.12 not found")
    monkeypatch.setenv("QPZK_JIT_CACHE_DIR", str(tmp_path))
    code = np.array([[0, 0, 0, 0], [1, 1, 0, 0], [2, 2, 0, 0], [3, 3, 0, 1], [4, 4, 3, 2], [5, 5, 4, 3], [0x0307, 6, 5, 4],
                     [8, 7, 6, 1], [9, 7, 7, 0], [0x060a, 7, 7, 1], [6, 0, 7, 0], [6, 0, 5, 0]], np.uint32)
    imm = np.array([5, P - 1], np.uint64)
    prog = qpzk.GateProgram.make(code, imm, 8)
    size = C.c_uint64(0)
    assert lib.qpzk_gate_jit_compile_check(C.byref(prog), C.byref(size)) == 0, lib.qpzk_last_error()
    assert size.value > 1000
    cached = list(tmp_path.glob("*_12.cubin"))
    assert len(cached) == 1 and cached[0].read_bytes()[:4] == b"\x7fELF"
    assert lib.qpzk_gate_jit_compile_check(C.byref(prog), C.byref(size)) == 0 and size.value == cached[0].stat().st_size
    bad = qpzk.GateProgram.make(np.array([[8, 0, 0, 7]], np.uint32), imm, 8)
    assert lib.qpzk_gate_jit_compile_check(C.byref(bad), C.byref(size)) != 0


def test_lde_pass_plan_for_every_size(qpzk):
    """The pass planner of the coset LDE (capi.cu: lde_pass_plan) is a pure host function: for every size of the sweep the
    passes cover exactly the bits above the last pass, every pass is one a kernel takes at its position (1..4 bits need
    b_lo >= 8, 5..8 bits need b_lo >= 12 - s), the pass next to the last one is wide, and the benchmarked shapes get the
    measured-best orders (2^22: 1+8+8 on one GPU, 8+1+8 when the first pass reads peer slices)."""
    import ctypes as C
    lib = qpzk.load_library()
    passes, n, last = (C.c_uint32 * 8)(), C.c_uint32(0), C.c_uint32(0)

    def plan(log_n, cols=135, peer=0):
        assert lib.qpzk_lde_pass_plan(log_n, cols, peer, passes, C.byref(n), C.byref(last)) == 0
        return list(passes[:n.value]), last.value

    for peer in (0, 1):
        for log_n in range(12, 27):
            p, s_last = plan(log_n, 135, peer)
            assert s_last == 5 and sum(p) == log_n - 5 and len(p) <= 4
            b_hi = log_n
            for s in p:
                b_lo = b_hi - s
                assert 1 <= s <= 8 and (b_lo >= 8 if s <= 4 else b_lo >= 12 - s), (log_n, p)
                b_hi = b_lo
            assert p[-1] >= 7
    assert plan(22)[0] == [1, 8, 8] and plan(22, peer=1)[0] == [8, 1, 8]
    assert plan(20)[0] == [7, 8] and plan(21)[0] == [8, 8] and plan(13)[0] == [8] and plan(16)[0] == [3, 8]
    assert plan(23, peer=1)[0] == [8, 2, 8] and plan(21, peer=1)[0] == [8, 8]
    p, s_last = plan(8, 6)                         # a short polynomial: everything in the last pass or one generic pass
    assert sum(p) + s_last == 8
    assert lib.qpzk_lde_pass_plan(30, 135, 0, passes, C.byref(n), C.byref(last)) != 0


def test_no_cpu_fallback_without_gpu(qpzk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present; the failure path is for GPU-less hosts")
    with pytest.raises(qpzk.QpzkError) as e:
        qpzk.Context(0)
    assert e.value.code in (-1, -2)


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "qp-zk-circuits_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "orc_" not in text and "import oracle" not in text, f


def test_native_driver_builds_and_fails_loudly_without_gpu(qpzk):
    """host/prove_bench.cpp (the C++ caller of the C ABI) is built by `make`; on a GPU-less host it must exit with
    the library's error, not fall back to anything."""
    import subprocess
    import torch
    exe = os.path.join(os.path.dirname(qpzk.LIB_PATH), "prove_bench")
    assert os.path.exists(exe), "run __graft_entry__.build() / make -C qp-zk-circuits_b200"
    if torch.cuda.is_available():
        pytest.skip("GPU present; the failure path is for GPU-less hosts")
    r = subprocess.run([exe, "--threads", "1", "--proofs", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "qpzk_ctx_create" in r.stderr
