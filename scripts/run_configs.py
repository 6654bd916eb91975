#!/usr/bin/env python
"""
run_configs.py — the BASELINE.json parity-test configurations at (or near) full
size on one B200, each timed on the device and checked with size-independent
properties (linearity checksums, sampled entries, triangle rules) plus a full
oracle comparison where the CPU can finish in seconds.  Not a bench line: the
output (one JSON object per config) feeds DESIGN.md / profiles.

    python scripts/run_configs.py [c1] [c3] [c4] [c5bsr] [--scale 22 --ef 1]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import sparse_dot_b200 as sdb  # noqa: E402
from sparse_dot_b200 import _handles as H  # noqa: E402
from sparse_dot_b200 import _lib  # noqa: E402
from tests import _cases as cs  # noqa: E402

lib = _lib.SDB.lib


def sync():
    _lib.check(lib.sdb_device_synchronize(), "sync")


def timed(fn, reps=1):
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    sync()
    return (time.perf_counter() - t0) / reps * 1e3, out


def run_c1():
    """configs[0]: CSR 10k x 10k d=1e-3 fp64 x dense 10k x 64."""
    from oracle import oracle as orc

    a = sp.random(10_000, 10_000, density=1e-3, format="csr", dtype=np.float64, random_state=86)
    x = np.random.default_rng(88).random((10_000, 64))
    sdb.dot_product_mkl(a, x)  # first call pays CUDA context + pinned-ring creation
    ms, y = timed(lambda: sdb.dot_product_mkl(a, x), reps=5)
    want = orc.c_spmm(a, x)
    t0 = time.perf_counter()
    orc.c_spmm(a, x)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    return {"config": "c1", "e2e_ms": ms, "oracle_cpu_ms": cpu_ms, "max_rel_err": cs.rel_err(y, want),
            "kernel_ms": sdb.last_timing_ms()[1]}


def chunked_rowsums(vals, indptr, w=None, idx=None, chunk=200_000_000):
    """float64 row sums of a CSR value array (optionally weighted by w[column]) without a full float64 copy."""
    rows = indptr.shape[0] - 1
    out = np.zeros(rows)
    r0 = 0
    while r0 < rows:
        r1 = int(np.searchsorted(indptr, indptr[r0] + chunk, side="right")) - 1
        r1 = min(rows, max(r1, r0 + 1))
        p0, p1 = int(indptr[r0]), int(indptr[r1])
        if p1 > p0:
            v = vals[p0:p1].astype(np.float64)
            if w is not None:
                v *= w[idx[p0:p1]]
            local = indptr[r0:r1] - p0
            nonempty = np.flatnonzero(np.diff(indptr[r0:r1 + 1]) > 0)
            out[r0 + nonempty] = np.add.reduceat(v, local[nonempty])
        r0 = r1
    return out


def run_c3_resident(scale, ef):
    """configs[2] at an edge factor whose result does not fit a host export comfortably: the product stays
    in HBM; only the values come back (one array) for the checksum-of-checksums."""
    a = cs.rmat_csr(scale, ef, np.float32, seed=1)
    b = cs.rmat_csr(scale, ef, np.float32, seed=2)
    products = int(np.dot(np.bincount(a.indices, minlength=a.shape[1]).astype(np.int64),
                          np.diff(b.indptr).astype(np.int64)))
    res = {"config": "c3_resident", "scale": scale, "edge_factor": ef, "nnz_a": int(a.nnz), "nnz_b": int(b.nnz),
           "products": products}
    ha, _, _ = H.create(a)
    hb, _, _ = H.create(b)
    with ha, hb:
        def mult():
            ref = C.c_void_p()
            _lib.check(lib.sdb_spgemm(_lib.OP_N, ha.ref, hb.ref, C.byref(ref)), "sdb_spgemm")
            return H.Handle(ref, np.float32)

        ms0, hc = timed(mult)  # cold: grows the memory pool
        hc.destroy()

        def mult_ordered():
            ref = C.c_void_p()
            _lib.check(lib.sdb_spgemm_ordered(_lib.OP_N, ha.ref, hb.ref, C.byref(ref)), "sdb_spgemm_ordered")
            return H.Handle(ref, np.float32)

        ms_o, hco = timed(mult_ordered)
        hco.destroy()
        res["spgemm_ordered_ms"] = ms_o
        ms, hc = timed(mult)
        with hc:
            res["spgemm_cold_ms"] = ms0
            res["spgemm_ms"] = ms
            info = H.info(hc)
            nnz = int(info["nnz"])
            res["nnz_c"] = nnz
            res["result_gbytes"] = nnz * 8 / 1e9
            res["g_products_per_s"] = products / (ms * 1e-3) / 1e9
            try:
                ms2, _ = timed(lambda: H.order(hc))
                res["order_ms"] = ms2
            except ValueError as e:
                res["order_error"] = str(e)[:200]
            vals = np.empty(nnz, dtype=np.float32)
            t0 = time.perf_counter()
            _lib.check(lib.sdb_export(hc.ref, None, 64, None, 32, vals.ctypes.data_as(C.c_void_p)), "sdb_export")
            res["values_export_ms"] = (time.perf_counter() - t0) * 1e3
            colsum_a = np.bincount(a.indices, weights=a.data.astype(np.float64), minlength=a.shape[1])
            rowsum_b = np.asarray(b.astype(np.float64).sum(axis=1)).ravel()
            want_total = float(np.dot(colsum_a, rowsum_b))
            got_total = float(vals.sum(dtype=np.float64))
            res["total_rel_err"] = abs(got_total - want_total) / want_total
            # the checksum above does not see WHERE a product landed: per-row sums (C 1 = A (B 1)) catch a product in the
            # wrong row, and — when the result is small enough to bring its indices back — a random weighting of the
            # columns (C w = A (B w)) catches one in the wrong column of the right row
            indptr = np.empty(a.shape[0] + 1, dtype=np.int64)
            _lib.check(lib.sdb_export(hc.ref, indptr.ctypes.data_as(C.c_void_p), 64, None, 32, None), "sdb_export")
            rows_got = chunked_rowsums(vals, indptr)
            want_rows = a.astype(np.float64) @ rowsum_b
            res["rowsum_max_rel_err"] = float(np.max(np.abs(rows_got - want_rows) / np.maximum(want_rows, 1e-300)))
            if nnz < 1_500_000_000:
                idx = np.empty(nnz, dtype=np.int32)
                _lib.check(lib.sdb_export(hc.ref, None, 64, idx.ctypes.data_as(C.c_void_p), 32, None), "sdb_export")
                w = np.random.default_rng(7).random(b.shape[1]) + 0.5
                wrow = chunked_rowsums(vals, indptr, w, idx)
                want_w = a.astype(np.float64) @ (b.astype(np.float64) @ w)
                res["weighted_rowsum_max_rel_err"] = float(np.max(np.abs(wrow - want_w) / np.maximum(want_w, 1e-300)))
                dec = np.flatnonzero(np.diff(idx.astype(np.int64)) <= 0) + 1
                res["rows_strictly_sorted_after_order"] = bool(np.all(np.isin(dec, indptr)))
    return res


def run_c3(scale, ef, full_check):
    """configs[2]: CSR x CSR SpGEMM, two R-MAT fp32 matrices, reorder_output=True."""
    a = cs.rmat_csr(scale, ef, np.float32, seed=1)
    b = cs.rmat_csr(scale, ef, np.float32, seed=2)
    res = {"config": "c3", "scale": scale, "edge_factor": ef, "nnz_a": int(a.nnz), "nnz_b": int(b.nnz),
           "max_row_a": int(np.diff(a.indptr).max())}
    products = int(np.dot(np.bincount(a.indices, minlength=a.shape[1]).astype(np.int64),
                          np.diff(b.indptr).astype(np.int64)))
    res["products"] = products
    ha, _, _ = H.create(a)
    hb, _, _ = H.create(b)
    with ha, hb:
        def mult(ordered=False):
            ref = C.c_void_p()
            fn = lib.sdb_spgemm_ordered if ordered else lib.sdb_spgemm
            _lib.check(fn(_lib.OP_N, ha.ref, hb.ref, C.byref(ref)), "sdb_spgemm")
            return H.Handle(ref, np.float32)

        with mult(True) as warm:
            pass
        ms_o, hco = timed(lambda: mult(True))
        hco.destroy()
        res["spgemm_ordered_ms"] = ms_o
        with mult() as warm:  # first call pays for growing the device memory pool (GBs of cudaMalloc)
            H.order(warm)
        ms, hc = timed(mult)
        with hc:
            res["spgemm_ms"] = ms
            ms2, _ = timed(lambda: H.order(hc))
            res["order_ms"] = ms2
            info = H.info(hc)
            res["nnz_c"] = int(info["nnz"])
            res["compression"] = products / max(1, info["nnz"])
            res["gbytes_model"] = (a.nnz * 8 + products * (8 + 4) + info["nnz"] * 8) / 1e9
            # checksum of checksums: 1^T (A B) 1 = sum_k colsum_A(k) * rowsum_B(k)
            t0 = time.perf_counter()
            c = H.export(hc)
            res["export_ms"] = (time.perf_counter() - t0) * 1e3
            colsum_a = np.bincount(a.indices, weights=a.data.astype(np.float64), minlength=a.shape[1])
            rowsum_b = np.asarray(b.astype(np.float64).sum(axis=1)).ravel()
            want_total = float(np.dot(colsum_a, rowsum_b))
            got_total = float(c.data.astype(np.float64).sum())
            res["total_rel_err"] = abs(got_total - want_total) / want_total
            # per-row linearity: C 1 = A (B 1)
            rows_c = np.asarray(c.astype(np.float64).sum(axis=1)).ravel()
            want_rows = a.astype(np.float64) @ rowsum_b
            res["rowsum_max_rel_err"] = float(np.max(np.abs(rows_c - want_rows) / np.maximum(want_rows, 1e-300)))
            srt = True
            idx = c.indices
            ptr = c.indptr
            # sortedness: a decrease may only happen at a row boundary
            dec = np.flatnonzero(np.diff(idx.astype(np.int64)) < 0) + 1
            srt = bool(np.all(np.isin(dec, ptr)))
            res["sorted"] = srt
            if full_check:
                t0 = time.perf_counter()
                want = (a @ b).tocsr()
                want.sort_indices()
                res["scipy_ms"] = (time.perf_counter() - t0) * 1e3
                res["indptr_equal"] = bool(np.array_equal(c.indptr, want.indptr))
                res["indices_equal"] = bool(np.array_equal(c.indices, want.indices))
                res["values_max_rel_err"] = float(np.max(np.abs(c.data - want.data) / np.abs(want.data)))
    return res


def run_c4(m, n, per_row):
    """configs[3]: gram_matrix_mkl A^T A, CSR(m x n, per_row nnz/row, fp32), dense upper-triangular output,
    device-resident (the 40 GB result stays in HBM; sampled rows come back for the checks)."""
    a = cs.uniform_rows_csr(m, n, per_row, np.float32, seed=4)
    res = {"config": "c4", "m": m, "n": n, "nnz": int(a.nnz)}
    ha, _, _ = H.create(a)
    with ha:
        d_c = C.c_void_p()
        _lib.check(lib.sdb_dev_alloc(C.byref(d_c), n * n * 4), "sdb_dev_alloc")
        one, zero = _lib.scalar_pair(1.0), _lib.scalar_pair(0.0)
        try:
            # first call builds the transposed companion; time it separately
            ms0, _ = timed(lambda: _lib.check(
                lib.sdb_syrkd_dev(_lib.OP_T, ha.ref, one, zero, d_c, _lib.LAYOUT_C, n, None), "sdb_syrkd_dev"))
            ms, _ = timed(lambda: _lib.check(
                lib.sdb_syrkd_dev(_lib.OP_T, ha.ref, one, zero, d_c, _lib.LAYOUT_C, n, None), "sdb_syrkd_dev"))
            res["first_call_ms"] = ms0
            res["syrkd_ms"] = ms
            macs = float(per_row * (per_row + 1) / 2 * m)
            res["gflops"] = 2 * macs / (ms * 1e-3) / 1e9
            res["out_gbytes_upper"] = n * (n + 1) / 2 * 4 / 1e9
            # sampled rows: recompute C[i, i:] = A[:, i]^T A[:, i:] on the host from the rows that hold column i
            rng = np.random.default_rng(0)
            worst = 0.0
            for i in rng.choice(n, size=6, replace=False):
                row = np.empty(n, dtype=np.float32)
                _lib.check(lib.sdb_memcpy(row.ctypes.data_as(C.c_void_p), C.c_void_p(d_c.value + int(i) * n * 4),
                                          n * 4, 2), "sdb_memcpy")
                hits = np.flatnonzero(a.indices == i)
                src_rows = np.searchsorted(a.indptr, hits, side="right") - 1
                sub = a[src_rows].astype(np.float64)
                want = np.asarray(sub.T @ a.data[hits].astype(np.float64)).ravel()
                up = slice(int(i), n)
                worst = max(worst, float(np.max(np.abs(row[up] - want[up]) / np.maximum(want[up], 1e-30)
                                                * (want[up] > 0))))
                assert np.all(row[up][want[up] == 0] == 0)
                assert np.all(row[:int(i)] == 0), "strict lower triangle must be exactly zero"
            res["sampled_rows_max_rel_err"] = worst
            # trace = ||A||_F^2 (linearity): read the diagonal with a strided copy
            diag = np.empty(n, dtype=np.float32)
            rowbuf = np.empty(n, dtype=np.float32)
            step = max(1, n // 512)
            tr_got = tr_want = 0.0
            sq = np.bincount(a.indices, weights=a.data.astype(np.float64) ** 2, minlength=n)
            for i in range(0, n, step):
                _lib.check(lib.sdb_memcpy(rowbuf.ctypes.data_as(C.c_void_p),
                                          C.c_void_p(d_c.value + (i * n + i) * 4), 4, 2), "sdb_memcpy")
                tr_got += float(rowbuf[0])
                tr_want += float(sq[i])
            res["diag_sample_rel_err"] = abs(tr_got - tr_want) / tr_want
        finally:
            lib.sdb_dev_free(d_c)
    return res


def run_c5(rows, cols, per_row, n):
    """configs[4], one GPU's shard: CSR(rows x cols, per_row nnz/row, fp32) x dense(cols x n), resident operands."""
    a = cs.uniform_rows_csr(rows, cols, per_row, np.float32, seed=5)
    x = np.random.default_rng(6).random((cols, n), dtype=np.float32)
    res = {"config": "c5_shard", "rows": rows, "cols": cols, "nnz": int(a.nnz), "n": n}
    ha, _, _ = H.create(a)
    with ha:
        d_x, d_y = C.c_void_p(), C.c_void_p()
        _lib.check(lib.sdb_dev_alloc(C.byref(d_x), x.nbytes), "alloc")
        _lib.check(lib.sdb_dev_alloc(C.byref(d_y), rows * n * 4), "alloc")
        try:
            _lib.check(lib.sdb_memcpy(d_x, x.ctypes.data_as(C.c_void_p), x.nbytes, 1), "memcpy")
            one, zero = _lib.scalar_pair(1.0), _lib.scalar_pair(0.0)
            call = lambda: _lib.check(lib.sdb_spmm_dev(_lib.OP_N, one, ha.ref, _lib.LAYOUT_C, d_x, n, n, zero, d_y,
                                                       n, None), "sdb_spmm_dev")
            timed(call)
            ms, _ = timed(call, reps=5)
            res["spmm_ms"] = ms
            g = a.nnz * 8 + a.nnz * n * 4 + rows * n * 4 + (rows + 1) * 8
            res["gather_model_gbs"] = g / (ms * 1e-3) / 1e9
            res["gflops"] = 2.0 * a.nnz * n / (ms * 1e-3) / 1e9
            pick = np.sort(np.random.default_rng(0).choice(rows, size=16, replace=False))
            want = a[pick].astype(np.float64) @ x.astype(np.float64)
            worst = 0.0
            for j, r in enumerate(pick):
                got = np.empty(n, dtype=np.float32)
                _lib.check(lib.sdb_memcpy(got.ctypes.data_as(C.c_void_p), C.c_void_p(d_y.value + int(r) * n * 4),
                                          n * 4, 2), "memcpy")
                worst = max(worst, float(np.max(np.abs(got - want[j]) / np.abs(want[j]))))
            res["sampled_rows_max_rel_err"] = worst
        finally:
            lib.sdb_dev_free(d_x)
            lib.sdb_dev_free(d_y)
    return res


def run_c5bsr(block_rows, blocks_per_row, b, n):
    """configs[4] BSR variant on 1 GPU: natively blocked matrix, dense b x b fp32 blocks, x dense (k x n)."""
    rng = np.random.default_rng(5)
    cols = np.sort(np.stack([rng.choice(block_rows, size=blocks_per_row, replace=False)
                             for _ in range(min(block_rows, 4096))]), axis=1)
    reps = (block_rows + cols.shape[0] - 1) // cols.shape[0]
    cols = np.tile(cols, (reps, 1))[:block_rows]
    cols = (cols + np.arange(block_rows)[:, None] * 7919) % block_rows
    cols.sort(axis=1)
    indptr = np.arange(0, block_rows * blocks_per_row + 1, blocks_per_row, dtype=np.int32)
    data = (rng.random((block_rows * blocks_per_row, b, b), dtype=np.float32) + 0.5)
    a = sp.bsr_matrix((data, cols.ravel().astype(np.int32), indptr), shape=(block_rows * b, block_rows * b))
    x = rng.random((block_rows * b, n), dtype=np.float32)
    res = {"config": "c5bsr", "rows": block_rows * b, "nnz": int(data.size), "block": b, "n": n}
    ha, _, _ = H.create(a)
    with ha:
        d_x, d_y = C.c_void_p(), C.c_void_p()
        _lib.check(lib.sdb_dev_alloc(C.byref(d_x), x.nbytes), "alloc")
        _lib.check(lib.sdb_dev_alloc(C.byref(d_y), x.nbytes), "alloc")
        try:
            _lib.check(lib.sdb_memcpy(d_x, x.ctypes.data_as(C.c_void_p), x.nbytes, 1), "memcpy")
            one, zero = _lib.scalar_pair(1.0), _lib.scalar_pair(0.0)
            call = lambda: _lib.check(lib.sdb_spmm_dev(_lib.OP_N, one, ha.ref, _lib.LAYOUT_C, d_x, n, n, zero, d_y,
                                                       n, None), "sdb_spmm_dev")
            ms0, _ = timed(call)
            ms, _ = timed(call, reps=5)
            res["first_call_ms"] = ms0
            res["spmm_ms"] = ms
            res["kernel"] = sdb.last_spmm_kernel()
            g = data.size * 4 + cols.size * 4 + cols.size * b * n * 4 + block_rows * b * n * 4
            res["gather_model_gbs"] = g / (ms * 1e-3) / 1e9
            res["gflops"] = 2.0 * data.size * n / (ms * 1e-3) / 1e9
            rows = np.sort(rng.choice(block_rows * b, size=16, replace=False))
            worst = 0.0
            csr_rows = a.tocsr()[rows].astype(np.float64)
            want = csr_rows @ x.astype(np.float64)
            for j, r in enumerate(rows):
                got = np.empty(n, dtype=np.float32)
                _lib.check(lib.sdb_memcpy(got.ctypes.data_as(C.c_void_p), C.c_void_p(d_y.value + int(r) * n * 4),
                                          n * 4, 2), "memcpy")
                worst = max(worst, float(np.max(np.abs(got - want[j]) / np.abs(want[j]))))
            res["sampled_rows_max_rel_err"] = worst
        finally:
            lib.sdb_dev_free(d_x)
            lib.sdb_dev_free(d_y)
    return res


def run_spmv(rows, cols, per_row, dtype=np.float32, matrix=None):
    """y = beta*y + A x on the configs[1] matrix (one dense column), operands in HBM; the
    16-byte-load kernel and the scalar one ("spmv_wide" 0 / 1), every row checked against float64 numpy."""
    a = matrix if matrix is not None else cs.uniform_rows_csr(rows, cols, per_row, dtype, seed=2)
    rows, cols = a.shape
    rng = np.random.default_rng(7)
    x = rng.random(cols).astype(dtype)
    y0 = rng.random(rows).astype(dtype)
    es = np.dtype(dtype).itemsize
    res = {"config": "spmv", "rows": rows, "cols": cols, "nnz": int(a.nnz), "dtype": np.dtype(dtype).name,
           # A once (index + value per entry, 8 B of row offset per row), x once, y read and written
           "algorithmic_bytes": int(a.nnz * (4 + es) + rows * 8 + cols * es + 2 * rows * es)}
    want = 0.5 * y0.astype(np.float64) + a.astype(np.float64) @ x.astype(np.float64)
    ha, _, _ = H.create(a)
    with ha:
        d_x, d_y = C.c_void_p(), C.c_void_p()
        _lib.check(lib.sdb_dev_alloc(C.byref(d_x), x.nbytes), "alloc")
        _lib.check(lib.sdb_dev_alloc(C.byref(d_y), y0.nbytes), "alloc")
        try:
            _lib.check(lib.sdb_memcpy(d_x, x.ctypes.data_as(C.c_void_p), x.nbytes, 1), "memcpy")
            one, half = _lib.scalar_pair(1.0), _lib.scalar_pair(0.5)
            call = lambda: _lib.check(lib.sdb_spmm_dev(_lib.OP_N, one, ha.ref, _lib.LAYOUT_C, d_x, 1, 1, half, d_y,
                                                       1, None), "sdb_spmm_dev")
            for label, wide, tile in (("scalar", 1, 1), ("wide", 0, 1), ("tile", 0, 2), ("auto", 0, 0)):
                _lib.set_option("spmv_wide", wide)
                _lib.set_option("spmv_tile", tile)
                _lib.check(lib.sdb_invalidate(ha.ref), "sdb_invalidate")  # every variant starts from a fresh handle state
                _lib.check(lib.sdb_memcpy(d_y, y0.ctypes.data_as(C.c_void_p), y0.nbytes, 1), "memcpy")
                t0 = time.perf_counter()
                call()
                sync()
                res[f"{label}_first_call_ms"] = (time.perf_counter() - t0) * 1e3
                got = np.empty_like(y0)
                _lib.check(lib.sdb_memcpy(got.ctypes.data_as(C.c_void_p), d_y, y0.nbytes, 2), "memcpy")
                res[f"{label}_max_rel_err"] = float(np.max(np.abs(got - want) / np.abs(want)))
                call()
                ms, _ = timed(call, reps=50)
                res[f"{label}_ms"] = ms
                res[f"{label}_gbs"] = res["algorithmic_bytes"] / (ms * 1e-3) / 1e9
                res[f"{label}_kernel"] = sdb.last_spmm_kernel()
            res["spmv_ms"] = res["auto_ms"]  # what a caller repeating products with a resident matrix gets
            res["kernel"] = res["auto_kernel"]
        finally:
            _lib.set_option("spmv_wide", 0)
            _lib.set_option("spmv_tile", 0)
            lib.sdb_dev_free(d_x)
            lib.sdb_dev_free(d_y)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c1", "c3", "c4", "c5bsr"])
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--ef", type=int, default=1)
    ap.add_argument("--no-full-check", action="store_true")
    ap.add_argument("--gram-m", type=int, default=2_000_000)
    ap.add_argument("--gram-n", type=int, default=100_000)
    ap.add_argument("--bsr-block-rows", type=int, default=62_500)
    args = ap.parse_args()
    print(sdb.get_version_string(), flush=True)
    for w in args.which:
        t0 = time.perf_counter()
        if w == "c1":
            r = run_c1()
        elif w == "c3":
            r = run_c3(args.scale, args.ef, not args.no_full_check)
        elif w == "c3res":
            r = run_c3_resident(args.scale, args.ef)
        elif w == "c4":
            r = run_c4(args.gram_m, args.gram_n, 100)
        elif w == "c5":
            r = run_c5(1_000_000, 1_000_000, 64, 256)
        elif w == "c5bsr":
            r = run_c5bsr(args.bsr_block_rows, 4, 16, 256)
        elif w == "spmv":
            r = run_spmv(1_000_000, 1_000_000, 50)
        elif w == "spmv64":
            r = run_spmv(1_000_000, 1_000_000, 50, np.float64)
        elif w == "spmv_rmat":  # power-law rows and columns: BASELINE configs[2]'s operand times a vector
            r = run_spmv(0, 0, 0, matrix=cs.rmat_csr(args.scale, args.ef, np.float32, seed=1))
            r["config"] = f"spmv_rmat scale {args.scale} ef {args.ef}"
        else:
            raise SystemExit(f"unknown config {w}")
        r["wall_s"] = time.perf_counter() - t0
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
