"""GPU parity tests at the sizes and through the code paths the benchmark uses (run with -m gpu on a B200).

SURVEY.md 8d: "all k <= 18 full compare of LDE + digests; larger k: cap + 1024 sampled rows/paths".  The pass plans of
the iNTT / LDE depend on log_n (capi.cu LdePlan::prepare, run_intt), the last-pass tiling and the column loop on the
number of columns, the H2D overlap on the input size (>= 64 MB: chunked on the copy stream): every one of those is
exercised here against the oracle -- full compare where the restatement finishes in seconds, the O(n) definition
checks of oracle/scale_check.py above that (coefficients at a random point from VALUES, 1024 leaf rows x all
columns on subgroup cosets, their Merkle paths, the top of the tree)."""
import numpy as np
import pytest

from conftest import P, splitmix_columns

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sc():
    from oracle import scale_check
    return scale_check


def _device_columns(seed, C_, n):
    import torch
    from bench import splitmix_columns_device
    t = splitmix_columns_device(seed, C_, n, torch.device("cuda", 0))
    torch.cuda.synchronize()
    return t


@pytest.mark.parametrize("log_n", [16, 18])
def test_full_compare_at_sweep_sizes_135_columns(ctx, qpzk, orc, log_n):
    """BASELINE config 3 at its two smallest sizes, complete: coefficients, cap, every leaf digest, every top-level
    digest, 4096 contiguous + strided leaf rows and sampled paths against the full CPU restatement."""
    C_ = 135
    cols = splitmix_columns(0x51575A4B00000000 + log_n, C_, 1 << log_n)
    b = qpzk.PolynomialBatch.from_values(ctx, cols, 3, None, 4)          # host buffer: >= 64 MB, the chunked copy path
    o = orc.PolynomialBatch.from_values(cols, 3, 4)
    assert (b.polynomials == o.coeffs()).all()
    assert (b.cap == o.cap()).all()
    L = 1 << (log_n + 3)
    lv = o.leaves()
    want0 = np.stack([orc.hash_no_pad(lv[i]) for i in range(0, L, L // 2048)])
    got0 = b.merkle_tree.digests(0)
    assert (got0[:: L // 2048] == want0).all()
    for first in (0, L // 2 - 256, L - 512):
        assert (b.merkle_tree.rows(first, 512) == lv[first:first + 512]).all()
    stride = np.arange(0, L, L // 4096, dtype=np.uint64) + 3
    rows, sib = b.merkle_tree.open(stride)
    assert (rows == lv[stride.astype(np.int64)]).all()
    for k in range(0, 4096, 64):
        assert (sib[k] == o.merkle_tree.prove(int(stride[k]))).all()
    b.free()


@pytest.mark.parametrize("log_n,C_", [(19, 9), (20, 9), (21, 9), (22, 9), (19, 135), (20, 33)])
def test_pass_plans_of_the_large_sizes(ctx, qpzk, sc, log_n, C_):
    """One test per iNTT / LDE plan shape of log_n = 19..22 (the sizes of the sweep above the full restatement):
    device-resident input, the real plan, definition checks on all columns."""
    x = _device_columns(0x51575A4B00000000 + log_n, C_, 1 << log_n)
    b = qpzk.PolynomialBatch.from_device(ctx, x.data_ptr(), log_n, C_, 3, 4, values=True)
    r = sc.check_commit(b, x.cpu().numpy().view(np.uint64), 3, num_rows=1024, seed=log_n)
    assert r["ok"] and r["rows_checked"] == 1024 and r["columns"] == C_
    b.free()


def test_benchmarked_configuration_2p22_x_135(ctx, qpzk, sc):
    """The configuration bench.py times (2^22 x 135, device-resident input), checked the same way bench.py's
    parity_check does, plus the host-buffer (8-chunk H2D overlap) path on the same input: identical cap and rows."""
    log_n, C_ = 22, 135
    x = _device_columns(0x51575A4B00000000 + log_n, C_, 1 << log_n)
    host = x.cpu().numpy().view(np.uint64)
    b = qpzk.PolynomialBatch.from_device(ctx, x.data_ptr(), log_n, C_, 3, 4, values=True)
    del x
    r = sc.check_commit(b, host, 3, num_rows=1024, seed=22)
    assert r["ok"] and r["rows_checked"] == 1024
    cap, rows = b.cap, b.merkle_tree.rows(12345678, 64)
    b.free()
    h = qpzk.PolynomialBatch.from_values(ctx, host, 3, None, 4)          # pageable host memory, 8 chunks
    assert (h.cap == cap).all() and (h.merkle_tree.rows(12345678, 64) == rows).all()
    h.free()
    ctx.trim()


@pytest.mark.parametrize("log_n,C_,salted", [(16, 135, False), (16, 135, True), (18, 135, False), (20, 17, True), (13, 135, True)])
def test_host_buffer_commit_equals_device_commit(ctx, qpzk, sc, log_n, C_, salted):
    """qpzk_commit_from_values (host input; >= 64 MB is copied in 2..8 column ranges on the copy stream, overlapped
    with the per-column stages) against qpzk_commit_from_values_dev on the same data: same coefficients, same leaves,
    same cap -- with pinned and pageable input, with and without salt columns; then the definition checks."""
    import ctypes as C
    import torch
    n, L = 1 << log_n, 1 << (log_n + 3)
    x = _device_columns(77 + log_n + C_, C_, n)
    salt = _device_columns(99 + log_n, 4, L) if salted else None
    d = qpzk.PolynomialBatch.from_device(ctx, x.data_ptr(), log_n, C_, 3, 4, values=True,
                                         salt_ptr=salt.data_ptr() if salted else None)
    host = x.cpu().numpy().view(np.uint64)
    hsalt = salt.cpu().numpy().view(np.uint64) if salted else None
    pinned = torch.empty((C_, n), dtype=torch.int64).pin_memory()
    pinned.copy_(x.cpu())
    for src in (host, pinned.numpy().view(np.uint64)):
        h = qpzk.PolynomialBatch.from_values(ctx, src, 3, hsalt, 4)
        assert h.leaf_len == C_ + (4 if salted else 0)
        assert (h.cap == d.cap).all()
        for first in (0, L // 3, L - 256):
            assert (h.merkle_tree.rows(first, 256) == d.merkle_tree.rows(first, 256)).all()
        h.free()
    # the call may return only when the caller's buffer has been read -- overwrite it right after the call
    h = qpzk.PolynomialBatch.from_values(ctx, pinned.numpy().view(np.uint64), 3, hsalt, 4)
    pinned.zero_()
    assert (h.cap == d.cap).all()
    h.free()
    r = sc.check_commit(d, host, 3, num_rows=512, seed=5)
    assert r["ok"]
    if salted:                                                  # salt words sit behind the polynomial columns, leaf order
        i = np.array([0, 1, L // 2 + 5, L - 1], dtype=np.uint64)
        rows, _ = d.merkle_tree.open(i)
        for k, j in enumerate(i):
            nat = int(format(int(j), "0%db" % (log_n + 3))[::-1], 2)
            assert (rows[k][C_:] == hsalt[:, nat]).all()
    d.free()


def test_coefficient_slices_at_2p22_use_the_wide_first_pass(ctx, qpzk):
    """At 2^22 the slice path (first pass reads the ranks' buffers in place) orders its strided passes 8 + 1 + 8 instead of
    the 1 + 8 + 8 of a single-buffer commit; both must give the same leaves and cap."""
    import torch
    log_n, C_ = 22, 10
    n = 1 << log_n
    x = _device_columns(2222, C_, n)
    full = qpzk.PolynomialBatch.from_device(ctx, x.data_ptr(), log_n, C_, 3, 4, values=True)
    coeffs = torch.empty_like(x)
    qpzk._check(ctx.L.qpzk_intt_columns_dev(ctx.h, x.data_ptr(), coeffs.data_ptr(), log_n, C_))
    ctx.synchronize()                                          # the context has its own stream: torch must not clone before the iNTT has run
    slices = [coeffs[c0:min(c0 + 4, C_)].clone() for c0 in range(0, C_, 4)]
    torch.cuda.synchronize()
    cap = full.cap
    for begin, count in ((0, 8), (6, 2)):
        part = qpzk.PolynomialBatch.cosets_from_coeff_slices_device(ctx, [t.data_ptr() for t in slices], 4, log_n, C_, 3, 4, begin, count)
        assert (part.cap.reshape(-1) == cap.reshape(-1)[begin * 8:(begin + count) * 8]).all()
        for row in (0, 12345, n - 40):
            assert (part.merkle_tree.rows(row, 32) == full.merkle_tree.rows(begin * n + row, 32)).all()
        part.free()
    full.free()


def test_coset_split_at_scale_matches_full_commit(ctx, qpzk, sc):
    """The multi-GPU decomposition at a size with the multi-pass plans: iNTT on column slices, the
    coefficient matrix assembled as the all-gather would, every coset range committed separately -- the cap slices
    concatenate to the full cap and each slice passes the definition checks."""
    import torch
    log_n, C_ = 19, 21
    n = 1 << log_n
    x = _device_columns(4242, C_, n)
    host = x.cpu().numpy().view(np.uint64)
    full = qpzk.PolynomialBatch.from_device(ctx, x.data_ptr(), log_n, C_, 3, 4, values=True)
    coeffs = torch.empty_like(x)
    for c0 in range(0, C_, 6):                                  # ranks of 6 columns (the last one short)
        c1 = min(c0 + 6, C_)
        qpzk._check(ctx.L.qpzk_intt_columns_dev(ctx.h, x[c0:c1].data_ptr(), coeffs[c0:c1].data_ptr(), log_n, c1 - c0))
    ctx.synchronize()
    hco = coeffs.cpu().numpy().view(np.uint64)
    assert (hco == full.polynomials).all()
    cap = full.cap
    for count in (1, 2, 4):
        caps = []
        for begin in range(0, 8, count):
            part = qpzk.PolynomialBatch.cosets_from_coeffs_device(ctx, coeffs.data_ptr(), log_n, C_, 3, 4, begin, count)
            caps.append(part.cap)
            if begin in (0, 8 - count):
                r = sc.check_commit(part, host, 3, num_rows=64 * count, seed=begin, coeffs=hco, coset_begin=begin, coset_count=count)
                assert r["ok"]
            part.free()
        assert (np.concatenate(caps) == cap).all()
    # the same with the coefficient matrix left where the ranks' iNTTs wrote it (qpzk_commit_cosets_from_coeff_slices_dev:
    # the first LDE pass reads every slice in place -- peer memory over NVLink on a multi-GPU box; here three buffers of
    # 8 columns on one device, the last one short)
    slices = [coeffs[c0:min(c0 + 8, C_)].clone() for c0 in range(0, C_, 8)]
    torch.cuda.synchronize()
    for begin, count in ((0, 8), (4, 4), (6, 2)):
        part = qpzk.PolynomialBatch.cosets_from_coeff_slices_device(ctx, [t.data_ptr() for t in slices], 8, log_n, C_, 3, 4, begin, count)
        assert (part.cap.reshape(-1) == cap.reshape(-1)[begin * 8:(begin + count) * 8]).all()
        assert (part.merkle_tree.rows(5, 32) == full.merkle_tree.rows(begin * n + 5, 32)).all()
        part.free()
    with pytest.raises(qpzk.QpzkError):
        qpzk.PolynomialBatch.cosets_from_coeff_slices_device(ctx, [slices[0].data_ptr()], 8, log_n, C_, 3, 4, 0, 8)   # slices do not cover the columns
    pb = qpzk.PeerBuffers(ctx, 1 << 12)
    assert pb.ptr and len(pb.handle) == 64
    pb.close()
    full.free()
