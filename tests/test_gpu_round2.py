"""Round-2 additions, through the C-ABI on the GPU: the pageable-memory host pipeline, create-time
validation, cache invalidation for borrowed arrays, the bandwidth probes."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

import sparse_dot_b200 as sdb
from oracle import oracle as orc
from sparse_dot_b200 import _handles as H
from sparse_dot_b200 import _lib
from tests import _cases as cs

pytestmark = pytest.mark.gpu
lib = _lib.SDB.lib


def _pinned_like(arr):
    p = C.c_void_p()
    _lib.check(lib.sdb_host_alloc(C.byref(p), arr.nbytes), "sdb_host_alloc")
    buf = (C.c_char * arr.nbytes).from_address(p.value)
    out = np.frombuffer(buf, dtype=arr.dtype).reshape(arr.shape)
    out[...] = arr
    return out, p


# ===================================================== sdb_spmm_csr_host on ordinary (pageable) numpy arrays
@pytest.mark.parametrize("beta", [0.0, 0.5])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pageable_pipeline_matches_oracle(dtype, beta):
    """Large enough (> 32 MiB of operands) to take the row-chunk pipeline; every array is pageable, so uploads go
    through the staging ring and the result comes back through the download thread (_sparse_dense.py:34-132)."""
    a, x, y0 = cs.c2_workload(70_000, 60_000, 40, 128, seed=3, dtype=dtype)
    if beta == 0.0:
        got = sdb.dot_product_mkl(a, x)
        want = orc.c_spmm(a, x)
    else:
        got = sdb.dot_product_mkl(a, x, out=y0.copy(), out_scalar=beta)
        want = orc.c_spmm(a, x, beta=beta, y=y0.copy())
    assert cs.rel_err(got, want) <= cs.TOL[np.dtype(dtype)]
    t = sdb.last_timing_ms()
    assert t[2] > 0.0  # the pipelined call reports its spans


def test_mixed_pinned_and_pageable_operands():
    """X page-locked, A and Y pageable: each array takes its own route inside one call."""
    a, x, y0 = cs.c2_workload(70_000, 60_000, 40, 128, seed=4)
    xp, handle = _pinned_like(x)
    try:
        got = sdb.dot_product_mkl(a, xp, out=y0.copy(), out_scalar=2.0)
        want = orc.c_spmm(a, x, beta=2.0, y=y0.copy())
        assert cs.rel_err(got, want) <= 1e-5
    finally:
        lib.sdb_host_free(handle)


def test_pageable_pipeline_repeated_calls_reuse_the_rings():
    a, x, y0 = cs.c2_workload(70_000, 60_000, 40, 128, seed=5)
    y = y0.copy()
    for _ in range(3):
        sdb.dot_product_mkl(a, x, out=y, out_scalar=0.5)
    want = y0.copy()
    for _ in range(3):
        want = orc.c_spmm(a, x, beta=0.5, y=want)
    assert cs.rel_err(y, want) <= 1e-5


# ===================================================== create never trusts the host arrays
def test_malformed_matrix_is_a_value_error_not_a_crash():
    m1, _ = cs.fixture_pair(np.float64)
    bad = m1.copy()
    bad.indices[7] = bad.shape[1] + 5  # column out of range
    with pytest.raises(ValueError, match="sdb_create_csr returned 3"):
        H.create(bad)
    bad = m1.copy()
    bad.indices[3] = -1
    with pytest.raises(ValueError, match="sdb_create_csr returned 3"):
        H.create(bad)
    bad = m1.copy()
    bad.indptr[5], bad.indptr[6] = bad.indptr[6] + 3, bad.indptr[5]  # offsets decrease
    with pytest.raises(ValueError, match="sdb_create_csr returned 3"):
        H.create(bad)
    # the context is still healthy afterwards
    x = np.random.default_rng(0).random((m1.shape[1], 8))
    assert cs.rel_err(sdb.dot_product_mkl(m1, x), orc.c_spmm(m1, x), orc.value_bound(abs(m1), abs(x))) <= 1e-12


# ===================================================== borrowed device arrays + sdb_invalidate
def test_invalidate_drops_cached_copies_of_borrowed_arrays():
    import torch

    m1, _ = cs.fixture_pair(np.float32)
    ip = torch.from_numpy(m1.indptr.astype(np.int64)).cuda()
    ix = torch.from_numpy(m1.indices.astype(np.int32)).cuda()
    va = torch.from_numpy(m1.data).cuda()
    ref = C.c_void_p()
    _lib.check(lib.sdb_create_csr_dev(C.byref(ref), m1.shape[0], m1.shape[1], m1.nnz, C.c_void_p(ip.data_ptr()),
                                      C.c_void_p(ix.data_ptr()), C.c_void_p(va.data_ptr()), _lib.F32),
               "sdb_create_csr_dev")
    h = H.Handle(ref, np.float32)
    x = np.random.default_rng(1).random((m1.shape[0], 16), dtype=np.float32)
    xt = torch.from_numpy(x).cuda()
    yt = torch.empty((m1.shape[1], 16), dtype=torch.float32, device="cuda")
    one, zero = _lib.scalar_pair(1.0), _lib.scalar_pair(0.0)

    def at_times_x():
        _lib.check(lib.sdb_spmm_dev(_lib.OP_T, one, h.ref, _lib.LAYOUT_C, C.c_void_p(xt.data_ptr()), 16, 16, zero,
                                    C.c_void_p(yt.data_ptr()), 16, None), "sdb_spmm_dev")
        _lib.check(lib.sdb_device_synchronize(), "sync")
        return yt.cpu().numpy()

    with h:
        want = orc.c_spmm(m1.T.tocsr(), x)
        assert cs.rel_err(at_times_x(), want, orc.value_bound(abs(m1.T.tocsr()), abs(x))) <= 1e-5
        va.mul_(2.0)  # the owner rewrites the values in place: the cached transposed copy is now stale
        torch.cuda.synchronize()
        _lib.check(lib.sdb_invalidate(h.ref), "sdb_invalidate")
        assert cs.rel_err(at_times_x(), 2.0 * want, 2.0 * orc.value_bound(abs(m1.T.tocsr()), abs(x))) <= 1e-5


# ===================================================== bandwidth probes (bench.py's roofline denominators)
def test_bandwidth_probes_are_sane():
    hbm = _lib.probe_bandwidth(0, 1 << 30, 2)
    l2 = _lib.probe_bandwidth(1, 32 << 20, 2)
    gather = _lib.probe_bandwidth(2, 32 << 20, 2)
    assert 1000.0 < hbm < 9000.0, hbm
    assert l2 > hbm and gather > hbm, (hbm, l2, gather)
    with pytest.raises(ValueError):
        _lib.probe_bandwidth(7, 1 << 30, 1)


# ===================================================== BSR x dense on tensor cores (DMMA / 3xTF32) vs the oracle
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("b", [8, 16, 32])
@pytest.mark.parametrize("n", [8, 40, 128, 264])
def test_bsr_tensor_core_kernel(dtype, b, n):
    """_common.py:327-384 create_bsr -> _sparse_dense.py:111-123 mm, with the tensor-core kernel forced on:
    every block size it covers, chunk widths with dead n-tiles (n = 40, 264), beta != 0, both block layouts."""
    rng = np.random.default_rng(b * 1000 + n)
    dense = sp.random(6 * b, 5 * b, density=0.02, format="csr", dtype=np.float64, random_state=b + n)
    a = (dense + sp.eye(6 * b, 5 * b, format="csr")).tobsr(blocksize=(b, b)).astype(dtype)
    a.data[:] = (rng.random(a.data.shape) + 0.5).astype(dtype)
    x = rng.random((a.shape[1], n)).astype(dtype)
    y0 = rng.random((a.shape[0], n)).astype(dtype)
    want = orc.c_spmm(a.tocsr(), x, beta=0.5, y=y0.copy())
    want0 = orc.c_spmm(a.tocsr(), x)
    tol = cs.TOL[np.dtype(dtype)]
    _lib.set_option("bsr_mma", 1)
    try:
        for mat in (a, sp.bsr_matrix((np.asfortranarray(a.data.transpose(0, 2, 1)).transpose(0, 2, 1), a.indices, a.indptr),
                                     shape=a.shape)):
            got = sdb.dot_product_mkl(mat, x, out=y0.copy(), out_scalar=0.5)
            assert "spmm_bsr_mma_kernel" in sdb.last_spmm_kernel()
            assert cs.rel_err(got, want) <= tol
            assert cs.rel_err(sdb.dot_product_mkl(mat, x), want0) <= tol
    finally:
        _lib.set_option("bsr_mma", -1)


def test_bsr_tensor_core_falls_back_when_not_covered():
    a = sp.random(40, 40, density=0.2, format="csr", dtype=np.float32, random_state=1).tobsr(blocksize=(4, 4))
    x = np.random.default_rng(0).random((40, 12), dtype=np.float32)
    _lib.set_option("bsr_mma", 1)
    try:
        got = sdb.dot_product_mkl(a, x)  # block 4 and n % 8 != 0: the FMA kernel
        assert "mma" not in sdb.last_spmm_kernel()
        assert cs.rel_err(got, orc.c_spmm(a.tocsr(), x), orc.value_bound(abs(a.tocsr()), abs(x))) <= 1e-5
    finally:
        _lib.set_option("bsr_mma", -1)


# ===================================================== SpGEMM: both wide-row formulations, all four bins
@pytest.mark.parametrize("wide", [0, 1])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_spgemm_wide_formulations_agree_with_the_oracle(dtype, wide):
    """_sparse_sparse.py:35-40 (mkl_sparse_spmm) + reorder_output: power-law rows reach the warp, small, CTA and
    wide bins; `spgemm_wide` picks the summary formulation (0, default) or the full-sweep bitmap (1)."""
    a = cs.rmat_csr(13, 8, dtype, seed=1)
    b = cs.rmat_csr(13, 8, dtype, seed=2)
    _lib.set_option("spgemm_wide", wide)
    try:
        got = sdb.dot_product_mkl(a, b, reorder_output=True)
        want = orc.c_spgemm(a, b, sort=True)
        assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
        assert cs.rel_err(got.data, want.data) <= cs.TOL[np.dtype(dtype)]
        unsorted = sdb.dot_product_mkl(a, b)
        unsorted.sort_indices()
        assert np.array_equal(unsorted.indices, want.indices)
        assert cs.rel_err(unsorted.data, want.data) <= cs.TOL[np.dtype(dtype)]
        g = sdb.gram_matrix_mkl(a, reorder_output=True)
        wg = orc.c_syrk(a, sort=True)
        assert np.array_equal(g.indptr, wg.indptr) and np.array_equal(g.indices, wg.indices)
        assert cs.rel_err(g.data, wg.data) <= cs.TOL[np.dtype(dtype)]
    finally:
        _lib.set_option("spgemm_wide", 0)


def test_spgemm_duplicate_columns_in_b_rows_use_atomics():
    """B rows with repeated column indices (not canonical): the warp bin must not take the plain read-modify-write."""
    rng = np.random.default_rng(3)
    a = sp.random(300, 200, density=0.05, format="csr", dtype=np.float64, random_state=5)
    rows = np.repeat(np.arange(200), 6)
    cols = rng.integers(0, 20, size=rows.shape[0])  # many duplicates inside a row
    vals = rng.random(rows.shape[0]) + 0.5
    order = np.lexsort((cols, rows))
    indptr = np.arange(0, 200 * 6 + 1, 6)
    b = sp.csr_matrix((vals[order], cols[order], indptr), shape=(200, 20))  # duplicates kept as stored
    assert not b.has_canonical_format
    got = sdb.dot_product_mkl(a, b, dense=True)
    want = a.toarray() @ b.toarray()
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    c = sdb.dot_product_mkl(a, b, reorder_output=True)
    assert np.abs(c.toarray() - want).max() <= 1e-12 * np.abs(want).max()


def test_options_reject_unknown_names():
    with pytest.raises(ValueError, match="sdb_set_option returned 3"):
        _lib.set_option("no_such_switch", 1)


# ===================================================== dense x dense (cblas_?gemm / cblas_?syrk callers)
ALL4 = [np.float32, np.float64, np.complex64, np.complex128]


def _dense(shape, dtype, order, seed):
    rng = np.random.default_rng(seed)
    a = rng.random(shape)
    if np.dtype(dtype).kind == "c":
        a = a + 1j * rng.random(shape)
    return np.asarray(a.astype(dtype), order=order)


@pytest.mark.parametrize("dtype", ALL4)
@pytest.mark.parametrize("order_a", ["C", "F"])
@pytest.mark.parametrize("order_b", ["C", "F"])
def test_dense_dot_dense(dtype, order_a, order_b):
    """_dense_dense.py:14-88: the result takes A's memory order, B may have the other one; out / out_scalar."""
    a = _dense((70, 130), dtype, order_a, 1)
    b = _dense((130, 45), dtype, order_b, 2)
    want = a.astype(np.complex128) @ b.astype(np.complex128)
    tol = cs.TOL[np.dtype(dtype)] * 130
    got = sdb.dot_product_mkl(a, b)
    assert got.dtype == np.dtype(dtype) and got.shape == (70, 45)
    assert got.flags.c_contiguous if order_a == "C" else got.flags.f_contiguous
    assert np.abs(got - want).max() <= tol * np.abs(want).max()
    out = np.asarray(np.ones((70, 45), dtype=dtype), order=order_a)
    res = sdb.dot_product_mkl(a, b, out=out, out_scalar=3.0)
    assert res is out and np.abs(out - (want + 3.0)).max() <= tol * np.abs(want).max()
    # a 1-d right operand is a column and the result is flattened (_dense_dense.py:19-21,71)
    v = sdb.dot_product_mkl(a, b[:, 0].copy())
    assert v.shape == (70,) and np.abs(v - want[:, 0]).max() <= tol * np.abs(want).max()


def test_dense_vector_dot_vector_is_numpys():
    x, y = np.arange(5.0), np.arange(5.0) + 1
    assert sdb.dot_product_mkl(x, y) == np.dot(x, y)


def test_dense_dot_dense_errors_and_cast():
    a = np.ones((4, 5), dtype=np.float32)
    b = np.ones((5, 3), dtype=np.float64)
    with pytest.raises(ValueError):
        sdb.dot_product_mkl(a, b)
    assert sdb.dot_product_mkl(a, b, cast=True).dtype == np.float64
    with pytest.raises(ValueError):
        sdb.dot_product_mkl(a, np.ones((4, 3), dtype=np.float32))
    assert sdb.dot_product_mkl(np.ones((0, 5), dtype=np.float32), np.ones((5, 3), dtype=np.float32)).shape == (0, 3)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("transpose", [False, True])
def test_gram_of_a_dense_array(dtype, order, transpose):
    """_gram_matrix.py:196-249 (cblas_?syrk): upper triangle in A's order; a fresh result is zero below the diagonal,
    a given out keeps what it had there."""
    a = _dense((60, 35), dtype, order, 3)
    full = (a @ a.T if transpose else a.T @ a).astype(np.float64)
    n = full.shape[0]
    got = sdb.gram_matrix_mkl(a, transpose=transpose)
    tol = cs.TOL[np.dtype(dtype)] * 60 * np.abs(full).max()
    assert got.flags.c_contiguous if order == "C" else got.flags.f_contiguous
    assert np.abs(np.triu(got) - np.triu(full)).max() <= tol and np.all(np.tril(got, -1) == 0)
    out = np.asarray(np.full((n, n), 7.0, dtype=dtype), order=order)
    res = sdb.gram_matrix_mkl(a, transpose=transpose, out=out, out_scalar=2.0)
    assert res is out and np.all(out[np.tril_indices(n, -1)] == 7.0)
    assert np.abs(np.triu(out) - np.triu(full + 14.0)).max() <= tol


# ===================================================== optimize(): the public call reaches the streaming kernel
def test_optimize_makes_dot_product_mkl_reuse_the_handle_and_reach_the_streaming_kernel():
    a, x, y0 = cs.c2_workload(120_000, 500_000, 40, 128, seed=6)
    want = orc.c_spmm(a, x, beta=0.5, y=y0.copy())
    _lib_env = None
    with sdb.optimize(a) as ra:
        kernels = []
        for _ in range(3):
            got = sdb.dot_product_mkl(ra, x, out=y0.copy(), out_scalar=0.5)
            kernels.append(sdb.last_spmm_kernel())
            assert cs.rel_err(got, want) <= 1e-5
        assert kernels[0].startswith("spmm_rowmajor")  # first product: the row-gather kernel
        # X is 256 MB (> 192 MB) and one wave of rows reuses X rows >= 1.5 times: the inspector runs on the 2nd call
        assert kernels[2].startswith("spmm_stream"), kernels
        with pytest.raises(ValueError):
            sdb.dot_product_mkl(ra, a)


@pytest.mark.gpu
@pytest.mark.parametrize("rows,row_elems,hpitch_elems,dpitch_elems", [(37, 1000, 1500, 1200), (3000, 9000, 20000, 9000),
                                                                    (5, 3, 7, 3), (2000, 20000, 20000, 20000)])
def test_memcpy_2d_round_trip(rows, row_elems, hpitch_elems, dpitch_elems):
    """sdb_memcpy_2d: a column range of a pageable row-major array to the device and back (strided on either side,
    staged through the page-locked ring by the copy threads), untouched bytes left alone."""
    import ctypes as ct

    rng = np.random.default_rng(rows)
    src = rng.standard_normal((rows, hpitch_elems)).astype(np.float32)
    dev = ct.c_void_p()
    _lib.check(_lib.SDB.lib.sdb_dev_alloc(ct.byref(dev), rows * dpitch_elems * 4), "sdb_dev_alloc")
    try:
        off = (hpitch_elems - row_elems) // 2
        _lib.check(_lib.SDB.lib.sdb_memcpy_2d(dev, dpitch_elems * 4, ct.c_void_p(src.ctypes.data + off * 4), hpitch_elems * 4,
                                    row_elems * 4, rows, 1), "sdb_memcpy_2d")
        back = np.full((rows, hpitch_elems + 3), 7.0, dtype=np.float32)
        _lib.check(_lib.SDB.lib.sdb_memcpy_2d(ct.c_void_p(back.ctypes.data + 2 * 4), (hpitch_elems + 3) * 4, dev, dpitch_elems * 4,
                                    row_elems * 4, rows, 2), "sdb_memcpy_2d")
    finally:
        _lib.check(_lib.SDB.lib.sdb_dev_free(dev), "sdb_dev_free")
    assert np.array_equal(back[:, 2:2 + row_elems], src[:, off:off + row_elems])
    assert np.all(back[:, :2] == 7.0) and np.all(back[:, 2 + row_elems:] == 7.0)
    with pytest.raises(Exception):
        _lib.check(_lib.SDB.lib.sdb_memcpy_2d(ct.c_void_p(back.ctypes.data), 4, dev, 8, 16, 2, 2), "sdb_memcpy_2d")


# ===================================================== SpMV with 16-byte loads of A (spmv_wide_kernel)
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ALL4)
@pytest.mark.parametrize("mean", [8, 15, 30, 60, 130])
def test_spmv_wide_kernel_ragged_rows(dtype, mean):
    """Every lanes-per-row specialisation of the 16-byte-load SpMV on rows of ragged length (empty rows, rows
    shorter than a pack, rows starting at every position modulo 4), against float64 numpy and against the scalar
    kernel (option spmv_wide = 1); conjugate transpose through the cached companion as well."""
    rng = np.random.default_rng(mean)
    rows, cols = 3001, 2500
    lens = rng.integers(0, 2 * mean + 1, size=rows)
    lens[::17] = 0
    lens[5::31] = 1
    indptr = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = np.concatenate([np.sort(rng.choice(cols, size=k, replace=False)) for k in lens]).astype(np.int32)
    data = rng.random(indptr[-1]) + 0.5
    if np.dtype(dtype).kind == "c":
        data = data + 1j * (rng.random(indptr[-1]) - 0.5)
    a = sp.csr_matrix((data.astype(dtype), indices, indptr), shape=(rows, cols))
    v = rng.random(cols).astype(dtype)
    if np.dtype(dtype).kind == "c":
        v = (v + 1j * rng.random(cols)).astype(dtype)
    want = a.astype(np.complex128) @ v.astype(np.complex128)
    tol = cs.TOL[np.dtype(dtype)]
    try:
        got = {}
        for opt in (0, 1):
            _lib.set_option("spmv_wide", opt)
            got[opt] = sdb.dot_product_mkl(a, v)
            assert np.abs(got[opt] - want).max() <= tol * np.abs(want).max()
        assert np.abs(got[0] - got[1]).max() <= tol * np.abs(want).max()
        _lib.set_option("spmv_wide", 0)
        w = rng.random(rows).astype(dtype)
        got_t = sdb.dot_product_mkl(w, a)
        want_t = w.astype(np.complex128) @ a.astype(np.complex128)
        assert np.abs(got_t - want_t).max() <= tol * np.abs(want_t).max()
    finally:
        _lib.set_option("spmv_wide", 0)


# ===================================================== SpMV with x staged in shared memory (spmv_tile.cu)
def _ragged_csr(rows, cols, mean, dtype, seed, banded=False):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, 2 * mean + 1, size=rows)
    lens[::17] = 0
    indptr = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    if banded:  # columns next to the diagonal: whole rows fall into one column slab
        start = (np.arange(rows) * (cols - 2 * mean - 1) // max(rows - 1, 1)).astype(np.int64)
        indices = np.concatenate([start[i] + np.arange(k) for i, k in enumerate(lens)]).astype(np.int32)
    else:
        indices = np.concatenate([np.sort(rng.choice(cols, size=k, replace=False)) for k in lens]).astype(np.int32)
    data = (rng.random(indptr[-1]) + 0.5).astype(dtype)
    return sp.csr_matrix((data, indices, indptr), shape=(rows, cols))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(3001, 2500, 12, False), (40_000, 70_000, 9, False), (20_000, 50_000, 40, True),
                                   (5, 40_000, 3, False)])
def test_spmv_tile_kernel_forced(dtype, shape):
    """Option spmv_tile = 2: the shared-memory SpMV on one-shot calls, several column slabs and row blocks, ragged
    and empty rows, banded rows (every entry of a row in one slab), out= with beta 0 / 0.5 / 1, and the transposed
    product through the cached companion; against float64 numpy."""
    rows, cols, mean, banded = shape
    a = _ragged_csr(rows, cols, mean, dtype, seed=rows + mean, banded=banded)
    rng = np.random.default_rng(3)
    v = rng.random(cols).astype(dtype)
    want = a.astype(np.float64) @ v.astype(np.float64)
    tol = cs.TOL[np.dtype(dtype)]
    scale = max(np.abs(want).max(), 1e-30)
    try:
        _lib.set_option("spmv_tile", 2)
        got = sdb.dot_product_mkl(a, v)
        assert sdb.last_spmm_kernel().startswith("spmv_tile_kernel")
        assert np.abs(got - want).max() <= tol * scale
        for beta in (0.0, 0.5, 1.0):
            out = np.full((rows, 1), np.nan if beta == 0.0 else 2.0, dtype=dtype)
            got = sdb.dot_product_mkl(a, v.reshape(-1, 1), out=out, out_scalar=beta)
            assert got is out
            assert np.abs(got.ravel() - (want + 2.0 * beta)).max() <= tol * (scale + 2.0)
        w = rng.random(rows).astype(dtype)
        got_t = sdb.dot_product_mkl(w, a)
        want_t = w.astype(np.float64) @ a.astype(np.float64)
        assert np.abs(got_t - want_t).max() <= tol * max(np.abs(want_t).max(), 1e-30)
    finally:
        _lib.set_option("spmv_tile", 0)


@pytest.mark.gpu
def test_spmv_tile_kernel_is_picked_for_repeated_products():
    """Automatic policy: a resident matrix large enough (6 M entries, x of 1.2 MB) runs the gather kernel on
    its first product with a vector and the shared-memory kernel from the second on; sdb_invalidate (the values
    changed) drops the cached tiles; a skewed matrix is balanced by work; one with too few entries per tile is
    declined by the inspector and stays on the gather kernel."""
    a = cs.uniform_rows_csr(150_000, 300_000, 40, np.float32, seed=9)
    x = np.random.default_rng(1).random((300_000, 1)).astype(np.float32)
    want = a.astype(np.float64) @ x.astype(np.float64)
    with sdb.optimize(a) as h:
        names = []
        for _ in range(3):
            got = h.dot(x)
            names.append(sdb.last_spmm_kernel())
            assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
        assert not names[0].startswith("spmv_tile_kernel")
        assert names[1].startswith("spmv_tile_kernel") and names[2].startswith("spmv_tile_kernel")
        _lib.check(_lib.SDB.lib.sdb_invalidate(h.handle.ref), "sdb_invalidate")
        h.dot(x)
        assert not sdb.last_spmm_kernel().startswith("spmv_tile_kernel")
    # half of the entries in the first 2 % of the rows: row blocks of equal work (not equal height) still balance it
    rng = np.random.default_rng(2)
    lens = np.full(150_000, 25)
    lens[:3000] = 800
    indptr = np.zeros(150_001, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = rng.integers(0, 300_000, size=indptr[-1]).astype(np.int32)
    skew = sp.csr_matrix((rng.random(indptr[-1]).astype(np.float32) + 0.5, indices, indptr), shape=(150_000, 300_000))
    skew.sum_duplicates()
    want = skew.astype(np.float64) @ x.astype(np.float64)
    with sdb.optimize(skew) as h:
        for i in range(3):
            got = h.dot(x)
            assert sdb.last_spmm_kernel().startswith("spmv_tile_kernel") == (i >= 1)
            assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    # too few entries per tile (12 per row over 2 M columns: every row block would stream 8 MB of x through shared
    # memory for 130 k entries): declined by the inspector, the gather kernel keeps serving it
    thin = cs.uniform_rows_csr(400_000, 2_000_000, 12, np.float32, seed=11)
    xt = np.random.default_rng(5).random((2_000_000, 1)).astype(np.float32)
    want = thin.astype(np.float64) @ xt.astype(np.float64)
    with sdb.optimize(thin) as h:
        for _ in range(3):
            got = h.dot(xt)
            assert not sdb.last_spmm_kernel().startswith("spmv_tile_kernel")
            assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_spmv_tile_kernel_on_a_power_law_matrix(dtype):
    """R-MAT rows and columns (BASELINE configs[2]'s generator): a few rows and columns hold most of the entries;
    forced shared-memory SpMV against float64 numpy, both directions."""
    a = cs.rmat_csr(16, 8, dtype, seed=5)
    rng = np.random.default_rng(6)
    v = rng.random(a.shape[1]).astype(dtype)
    want = a.astype(np.float64) @ v.astype(np.float64)
    tol = cs.TOL[np.dtype(dtype)]
    try:
        _lib.set_option("spmv_tile", 2)
        got = sdb.dot_product_mkl(a, v)
        assert sdb.last_spmm_kernel().startswith("spmv_tile_kernel")
        assert np.abs(got - want).max() <= tol * np.abs(want).max()
        got_t = sdb.dot_product_mkl(v, a)
        want_t = v.astype(np.float64) @ a.astype(np.float64)
        assert np.abs(got_t - want_t).max() <= tol * np.abs(want_t).max()
    finally:
        _lib.set_option("spmv_tile", 0)


@pytest.mark.gpu
def test_resident_matrix_times_vector_in_a_loop():
    """What a CG / power-iteration caller does: y = A x with a resident A and 1-D vectors, again and again,
    out= / out_scalar= included; the iterates match scipy's to fp32 tolerance."""
    a = cs.uniform_rows_csr(4000, 4000, 12, np.float32, seed=3)
    x = np.random.default_rng(4).random(4000).astype(np.float32)
    with sdb.optimize(a) as h:
        v, ref = x.copy(), x.astype(np.float64)
        for _ in range(4):
            v = sdb.dot_product_mkl(h, v)
            v /= np.abs(v).max()
            ref = a.astype(np.float64) @ ref
            ref /= np.abs(ref).max()
            assert v.ndim == 1 and np.abs(v - ref).max() <= 1e-5
        out = np.ones(4000, dtype=np.float32)
        got = h.dot(x, out=out, out_scalar=2.0)
        assert got is out
        assert np.abs(out - (a.astype(np.float64) @ x.astype(np.float64) + 2.0)).max() <= 1e-4
        with pytest.raises(ValueError):
            h.dot(x[:-1])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ALL4)
def test_spmv_rows_far_longer_than_the_rest(dtype):
    """Power-law rows: a mean of ~6 entries per row selects 2-4 lanes per row, and the rows beyond 512 entries per
    lane (here 2 600 and 2 999 entries, and an R-MAT matrix's head rows) are reduced by one CTA each
    (spmv_long_rows_kernel); both directions, out= with a scalar, against float64 numpy."""
    rng = np.random.default_rng(12)
    rows, cols = 20_000, 3000
    lens = rng.integers(0, 13, size=rows)
    lens[5], lens[777], lens[rows - 1] = 2600, 2999, 2100
    indptr = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = np.concatenate([np.sort(rng.choice(cols, size=k, replace=False)) for k in lens]).astype(np.int32)
    data = rng.random(indptr[-1]) + 0.5
    v = rng.random(cols)
    if np.dtype(dtype).kind == "c":
        data = data + 1j * (rng.random(indptr[-1]) - 0.5)
        v = v + 1j * rng.random(cols)
    a = sp.csr_matrix((data.astype(dtype), indices, indptr), shape=(rows, cols))
    v = v.astype(dtype)
    tol = cs.TOL[np.dtype(dtype)]
    want = a.astype(np.complex128) @ v.astype(np.complex128)
    for wide in (0, 1):
        try:
            _lib.set_option("spmv_wide", wide)
            got = sdb.dot_product_mkl(a, v)
            assert np.abs(got - want).max() <= tol * np.abs(want).max()
            out = np.ones(rows, dtype=dtype)
            got = sdb.dot_product_mkl(a, v, out=out, out_scalar=0.5)
            assert got is out and np.abs(out - (want + 0.5)).max() <= tol * np.abs(want).max()
        finally:
            _lib.set_option("spmv_wide", 0)
    w = rng.random(rows).astype(dtype)
    got_t = sdb.dot_product_mkl(w, a)
    want_t = w.astype(np.complex128) @ a.astype(np.complex128)
    assert np.abs(got_t - want_t).max() <= tol * np.abs(want_t).max()
    if np.dtype(dtype).kind != "c":
        g = cs.rmat_csr(17, 8, dtype, seed=3)
        x = rng.random(g.shape[1]).astype(dtype)
        with sdb.optimize(g) as h:
            for _ in range(2):
                got = h.dot(x)
                want = g.astype(np.float64) @ x.astype(np.float64)
                assert np.abs(got - want).max() <= tol * np.abs(want).max()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ALL4)
@pytest.mark.parametrize("n", [3, 32, 100, 128, 200])
def test_spmm_rows_far_longer_than_the_rest(dtype, n):
    """Power-law rows in sparse x dense: rows beyond 1024 entries (here 2 600, 2 999 and 1 025, next to a mean of 6)
    go to spmm_long_rows_kernel, the others stay in the row-gather kernel's 'long' instantiation; C and F ordered
    panels, out= with a scalar, the transposed product; against complex128 numpy."""
    rng = np.random.default_rng(n)
    rows, cols = 6000, 3000
    lens = rng.integers(0, 13, size=rows)
    lens[5], lens[777], lens[rows - 1], lens[4000] = 2600, 2999, 1025, 1024
    indptr = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = np.concatenate([np.sort(rng.choice(cols, size=k, replace=False)) for k in lens]).astype(np.int32)
    data = rng.random(indptr[-1]) + 0.5
    x = rng.random((cols, n))
    if np.dtype(dtype).kind == "c":
        data = data + 1j * (rng.random(indptr[-1]) - 0.5)
        x = x + 1j * rng.random((cols, n))
    a = sp.csr_matrix((data.astype(dtype), indices, indptr), shape=(rows, cols))
    x = x.astype(dtype)
    tol = cs.TOL[np.dtype(dtype)]
    want = a.astype(np.complex128) @ x.astype(np.complex128)
    scale = np.abs(want).max()
    got = sdb.dot_product_mkl(a, x)
    assert np.abs(got - want).max() <= tol * scale
    with sdb.optimize(a) as h:  # (large panels go through the row-chunk pipeline, whose chunks carry no handle)
        got = h.dot(x)
        assert "long" in sdb.last_spmm_kernel()
        assert np.abs(got - want).max() <= tol * scale
    got_f = sdb.dot_product_mkl(a, np.asfortranarray(x))
    assert np.abs(got_f - want).max() <= tol * scale
    out = np.ones((rows, n), dtype=dtype)
    got = sdb.dot_product_mkl(a, x, out=out, out_scalar=0.5)
    assert got is out and np.abs(out - (want + 0.5)).max() <= tol * scale
    w = rng.random((n, rows)).astype(dtype)
    got_t = sdb.dot_product_mkl(w, a)  # dense x sparse: the transposed companion has long rows of its own or none
    want_t = w.astype(np.complex128) @ a.astype(np.complex128)
    assert np.abs(got_t - want_t).max() <= tol * np.abs(want_t).max()


@pytest.mark.gpu
def test_spmm_on_a_power_law_matrix_keeps_its_results_when_repeated():
    """R-MAT scale 17 x dense 64 with a resident handle: the long-row list is built once and reused; csc input too."""
    g = cs.rmat_csr(17, 16, np.float32, seed=4)
    x = np.random.default_rng(1).random((g.shape[1], 64)).astype(np.float32)
    want = g.astype(np.float64) @ x.astype(np.float64)
    with sdb.optimize(g) as h:
        for _ in range(3):
            got = h.dot(x)
            assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    got = sdb.dot_product_mkl(g.tocsc(), x)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
