"""
BASELINE.json configurations at FULL size on one B200, checked through size-independent properties
(the CPU cannot recompute them in seconds): checksums of checksums (1ᵀ(AX) = (Aᵀ1)ᵀX, 1ᵀ(AB)1 = colsum(A)·rowsum(B)),
per-row linearity (C·1 = A·(B·1)), sampled rows / entries recomputed on the host in float64, sortedness,
bit-identical repeat runs, trace identities.  The timings and the checks themselves live in
scripts/run_configs.py (also used for DESIGN.md's tables); this module asserts on them.

configs[1] goes through the public API (`dot_product_mkl(csr, ndarray, out=, out_scalar=)`, host arrays);
the others keep their multi-GB results in HBM and read back what the checks need.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def test_config1_full_size_public_api():
    """configs[1]: CSR(1M x 1M, 50 nnz/row, fp32) x dense(1M x 128), out= accumulate (beta = 0.5)."""
    import bench
    import sparse_dot_b200 as sdb

    a, x, y0 = bench.make_workload(bench.M_ROWS, bench.K_COLS, bench.NNZ_PER_ROW, bench.N_DENSE, seed=0)
    assert a.nnz == 50_000_000 and a.has_sorted_indices
    beta = 0.5
    out = y0.copy()
    got = sdb.dot_product_mkl(a, x, out=out, out_scalar=beta)
    assert got is out
    # (1) checksum of checksums: column sums of Y = (A^T 1)^T X + beta * column sums of Y0, in float64
    w = np.bincount(a.indices, weights=a.data.astype(np.float64), minlength=a.shape[1])
    want_cols = w @ x.astype(np.float64) + beta * y0.sum(axis=0, dtype=np.float64)
    got_cols = got.sum(axis=0, dtype=np.float64)
    assert np.max(np.abs(got_cols - want_cols) / want_cols) < 1e-7
    # (2) sampled rows recomputed in float64 (the first, the last and 62 random ones)
    rows = np.unique(np.concatenate([[0, a.shape[0] - 1], np.random.default_rng(0).integers(0, a.shape[0], 62)]))
    want = a[rows].astype(np.float64) @ x.astype(np.float64) + beta * y0[rows].astype(np.float64)
    assert np.max(np.abs(got[rows] - want) / np.abs(want)) < 1e-5
    # (3) the kernel has no atomics: a repeat run is bit-identical
    again = sdb.dot_product_mkl(a, x, out=y0.copy(), out_scalar=beta)
    assert np.array_equal(again, got)
    # (4) beta = 0 path (no `out`): Y1 = A X, and the accumulate result is Y1 + beta * Y0 up to rounding
    y1 = sdb.dot_product_mkl(a, x)
    assert y1.dtype == np.float32 and y1.flags.c_contiguous
    assert np.max(np.abs((y1[rows].astype(np.float64) + beta * y0[rows]) - want) / np.abs(want)) < 1e-5


def test_config2_full_size_rmat_scale22():
    """configs[2]: two R-MAT scale-22 fp32 matrices, sparse output, reorder_output=True (edge factor 1: the
    one instance whose result, 7e8 entries, can be exported and checked on the host)."""
    import run_configs as rc

    r = rc.run_c3(22, 1, full_check=False)
    assert r["nnz_c"] > 6.9e8 and r["products"] >= r["nnz_c"]
    assert r["total_rel_err"] < 1e-7      # 1^T (A B) 1 = sum_k colsum_A(k) rowsum_B(k)
    assert r["rowsum_max_rel_err"] < 1e-5  # C 1 = A (B 1), every row
    assert r["sorted"]                     # every row ascending after sdb_order


def test_config3_full_size_gram_dense():
    """configs[3]: gram_matrix_mkl A^T A, CSR(2M x 100k, 100 nnz/row, fp32), dense upper triangle (40 GB in HBM)."""
    import run_configs as rc

    r = rc.run_c4(2_000_000, 100_000, 100)
    assert 199_990_000 <= r["nnz"] <= 200_000_000
    assert r["sampled_rows_max_rel_err"] < 1e-5   # C[i, i:] = A[:, i]^T A[:, i:] on 6 random rows
    assert r["diag_sample_rel_err"] < 1e-6        # sampled trace = sum of squared column norms


def test_config4_full_size_shard_and_bsr():
    """configs[4]: one GPU's shard of the row-sharded product (1M x 1M, 64 nnz/row, N = 256) and the BSR-16 variant."""
    import run_configs as rc

    r = rc.run_c5(1_000_000, 1_000_000, 64, 256)
    assert 63_999_000 <= r["nnz"] <= 64_000_000 and r["sampled_rows_max_rel_err"] < 1e-5
    r = rc.run_c5bsr(62_500, 4, 16, 256)
    assert r["nnz"] == 64_000_000 and r["sampled_rows_max_rel_err"] < 1e-5
