"""ORACLE (test infrastructure, NOT product code): parity checks for commitments too large for the full CPU
restatement ("larger k: cap + 1024 sampled rows/paths").

Used by tests/ and by bench.py's `parity_check` (outside the timed region), never by the product.  What is checked,
for a `PolynomialBatch::from_values` result of ANY size, in O(n) CPU work per column:

  1. coefficients: P_c(r) from the committed VALUES by the barycentric formula == P_c(r) from the returned
     coefficients by Horner's rule, at a random r, for every column (a wrong coefficient anywhere is caught with
     probability 1 - n/p);
  2. LDE rows: `num_rows` leaf rows x ALL columns against the oracle's evaluation of those coefficients on a coset of
     a small subgroup inside every coset of the LDE domain the batch holds (fold + one small transform: definitions,
     not the FFT under test); leaf order = bitrev of the natural LDE index;
  3. Merkle: every sampled row's path is verified against the cap with the oracle's Poseidon, and the top of the tree
     (levels with <= 4096 nodes) is recomputed from the device's digests.
"""
import numpy as np

from . import oracle as orc

P = orc.P


def _brev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def check_commit(batch, values, rate_bits, num_rows=1024, seed=1, coeffs=None, coset_begin=0, coset_count=None,
                 check_tree_top=True):
    """batch: qpzk.PolynomialBatch (full, or a coset slice [coset_begin, coset_begin + coset_count) whose coefficients
    the caller passes in `coeffs`); values: the (C, n) host array that was committed.  Returns a dict; raises
    AssertionError on the first mismatch."""
    values = np.ascontiguousarray(values, dtype=np.uint64)
    C_, n = values.shape
    log_n = n.bit_length() - 1
    log_L = log_n + rate_bits
    cosets = 1 << rate_bits
    if coset_count is None:
        coset_count = cosets
    rng = np.random.default_rng(seed)
    if coeffs is None:
        coeffs = batch.polynomials
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64)
    assert coeffs.shape == values.shape
    # 1. coefficients vs values at a random point
    while True:
        r = int(rng.integers(2, P, dtype=np.uint64))
        if pow(r, n, P) != 1:
            break
    ev, ec = orc.eval_from_values(values, r), orc.eval_from_coeffs(coeffs, r)
    assert (ev == ec).all(), "coefficients do not interpolate the committed values (columns %s)" % np.nonzero(ev != ec)[0][:8]
    # 2. sampled rows, all columns: in every leaf block (= coset) of the batch, a coset of a subgroup of order mm
    per = max(1, num_rows // coset_count)
    log_mm = min(max(per.bit_length() - 1, 0), log_n)
    mm = 1 << log_mm
    cap = batch.cap
    tree = batch.merkle_tree
    ns = (tree.num_leaves.bit_length() - 1) - tree.cap_height
    rows_checked = 0
    for blk in range(coset_begin, coset_begin + coset_count):
        j = _brev(blk, rate_bits)                       # leaf block blk holds the natural indices i = j (mod 2^rate_bits)
        m0 = int(rng.integers(0, n >> log_mm)) if n >> log_mm else 0
        i0 = j + cosets * m0
        want = orc.lde_subgroup_rows(coeffs, rate_bits, i0, log_mm)
        step = (1 << log_L) >> log_mm
        idx = np.array([_brev(i0 + k * step, log_L) - coset_begin * n for k in range(mm)], dtype=np.uint64)
        assert ((idx >= (blk - coset_begin) * n) & (idx < (blk - coset_begin + 1) * n)).all()
        rows, sib = tree.open(idx)
        assert (rows[:, :C_] == want).all(), "LDE rows of leaf block %d differ from the oracle" % blk
        # 3a. Merkle paths of the sampled rows
        for k in range(mm):
            assert orc.merkle_verify(rows[k], int(idx[k]), cap, sib[k]), "Merkle path of leaf %d does not reach the cap" % idx[k]
        rows_checked += mm
    # 3b. top of the tree from the device's digests
    top_nodes = 0
    if check_tree_top and ns > 0:
        lv = max(0, ns - 8)                               # at most 2^8 nodes per cap entry
        d = tree.digests(lv)
        top_nodes = d.shape[0]
        while d.shape[0] > cap.shape[0]:
            d = np.stack([orc.two_to_one(d[2 * i], d[2 * i + 1]) for i in range(d.shape[0] // 2)])
        assert (d == cap).all(), "cap is not the root of the device's digests at level %d" % lv
    return {"ok": True, "columns": int(C_), "rows_checked": int(rows_checked), "paths_checked": int(rows_checked),
            "coeff_point_checks": int(C_), "tree_top_nodes": int(top_nodes), "log_n": int(log_n)}
