"""
GPU parity tests (run with `-m gpu` on the B200 box): the CUDA path, reached
through the public API and therefore through the C-ABI, against the CPU oracle
on the same seeded inputs.  Bar: bit-exact indptr / (sorted) indices; values
within 1e-5 (fp32 / complex64) or 1e-12 (fp64 / complex128) of the oracle,
relative to |A|@|B| (BASELINE.json north_star).

Modelled on the reference's own suites: tests/test_sparse_dense.py,
test_sparse_sparse.py, test_gram_matrix.py, test_sparse_vector.py, test_mkl.py.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc
from tests import _cases as cs

pytestmark = pytest.mark.gpu

import sparse_dot_b200 as sdb  # noqa: E402
from sparse_dot_b200 import _handles as H  # noqa: E402

REAL = [np.float32, np.float64]
ALL = [np.float32, np.float64, np.complex64, np.complex128]


def _pair(dtype):
    dtype = np.dtype(dtype)
    m1, m2 = cs.fixture_pair(np.float64)
    if dtype.kind == "c":
        m1, m2 = cs.complexify(m1, 11), cs.complexify(m2, 12)
    return m1.astype(dtype), m2.astype(dtype)


def _close(got, a, b, dtype, want=None, offset=0.0):
    """got ~= a @ b (+ offset) within the dtype's tolerance, relative to |a| @ |b| (+ |offset|)."""
    a64 = a.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64)
    b64 = b.astype(a64.dtype)
    if want is None:
        want = a64 @ b64
    want = (want.toarray() if sp.issparse(want) else np.asarray(want)) + offset
    bound = abs(a64) @ abs(b64)
    bound = (bound.toarray() if sp.issparse(bound) else np.asarray(bound)) + abs(offset)
    got = got.toarray() if sp.issparse(got) else np.asarray(got)
    assert got.shape == want.shape
    err = cs.rel_err(got, want, bound)
    assert err <= cs.TOL[np.dtype(dtype)], f"relative error {err:.3e}"


# ===================================================================== handles
@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_create_export_roundtrip(dtype, fmt):
    """test_mkl.py:204-228: create -> export reproduces the three arrays exactly."""
    m1, _ = _pair(dtype)
    m = m1.asformat(fmt)
    h, dbl, cplx = H.create(m)
    with h:
        back = H.export(h, output_type=f"{fmt}_matrix")
    assert (dbl, cplx) == (np.dtype(dtype).itemsize // (2 if cplx else 1) == 8, np.dtype(dtype).kind == "c")
    assert np.array_equal(back.indptr, m.indptr)
    assert np.array_equal(back.indices, m.indices)
    assert np.array_equal(back.data, m.data)
    assert back.dtype == m.dtype and back.shape == m.shape


@pytest.mark.parametrize("index_dtype", [np.int32, np.int64])
def test_create_accepts_both_index_widths(index_dtype):
    m1, _ = _pair(np.float64)
    m = m1.copy()
    m.indptr, m.indices = m.indptr.astype(index_dtype), m.indices.astype(index_dtype)
    h, _, _ = H.create(m)
    with h:
        back = H.export(h)
    assert np.array_equal(back.indices, m1.indices) and np.array_equal(back.indptr, m1.indptr)
    assert m.indices.dtype == index_dtype  # caller's arrays untouched


@pytest.mark.parametrize("dtype", REAL)
def test_bsr_roundtrip_and_convert(dtype):
    """test_mkl.py:230-268: BSR create/export, and BSR -> CSR conversion."""
    m1, _ = _pair(dtype)
    bsr = m1.tobsr(blocksize=(10, 10))
    h, _, _ = H.create(bsr)
    with h:
        back = H.export(h, output_type="bsr_matrix")
        assert np.array_equal(back.indptr, bsr.indptr)
        assert np.array_equal(back.indices, bsr.indices)
        assert np.array_equal(back.data, bsr.data)
        with H.convert_to_csr(h) as hc:
            csr = H.export(hc, output_type="csr_matrix")
    want = orc.c_bsr_to_csr(bsr)
    assert np.array_equal(csr.indptr, want.indptr)
    assert np.array_equal(csr.indices, want.indices)
    assert np.array_equal(csr.data, want.data)
    assert np.array_equal(csr.toarray(), m1.toarray())


@pytest.mark.parametrize("dtype", REAL)
def test_csc_convert_to_csr(dtype):
    m1, _ = _pair(dtype)
    csc = m1.tocsc()
    h, _, _ = H.create(csc)
    with h, H.convert_to_csr(h) as hc:
        csr = H.export(hc)
    want = orc.canonical(m1)
    assert np.array_equal(csr.indptr, want.indptr)
    assert np.array_equal(csr.indices, want.indices)
    assert np.array_equal(csr.data, want.data)


def test_order_sorts_rows_of_every_length():
    """mkl_sparse_order: warp (<=32), CTA (<=4096) and global (>4096) row bins."""
    rng = np.random.default_rng(5)
    lens = np.concatenate([rng.integers(0, 33, 300), rng.integers(33, 4097, 20), [4096, 4097, 9000, 20000]])
    n_cols = 50000
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    indices = np.concatenate([rng.choice(n_cols, size=int(k), replace=False) for k in lens]).astype(np.int32)
    data = rng.random(indices.shape[0])
    m = sp.csr_matrix((data, indices, indptr), shape=(len(lens), n_cols))
    want = orc.c_order(m.copy())
    h, _, _ = H.create(m)
    with h:
        H.order(h)
        got = H.export(h)
    assert np.array_equal(got.indptr, want.indptr)
    assert np.array_equal(got.indices, want.indices)
    assert np.array_equal(got.data, want.data)


def test_null_handle_is_value_error():
    """test_mkl.py:128-141."""
    import ctypes

    with pytest.raises(ValueError):
        H.Handle(ctypes.c_void_p(), np.float64).destroy()
    with pytest.raises(ValueError):
        H.export(H.Handle(ctypes.c_void_p(), np.float64))


# ================================================================ sparse x dense
@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("fmt", ["csr", "csc", "bsr"])
@pytest.mark.parametrize("order", ["C", "F"])
def test_sparse_dense_formats(dtype, fmt, order):
    """test_sparse_dense.py:31-279 across format x order x dtype."""
    m1, m2 = _pair(dtype)
    a = m1.asformat(fmt) if fmt != "bsr" else m1.tobsr(blocksize=(10, 10))
    a_before = a.copy()
    b = np.asarray(m2.toarray(), order=order)
    got = sdb.dot_product_mkl(a, b)
    assert got.dtype == np.dtype(dtype)
    assert got.flags["C_CONTIGUOUS" if order == "C" else "F_CONTIGUOUS"]
    _close(got, m1, b, dtype)
    # inputs are left alone (test_sparse_dense.py:95,112)
    assert np.array_equal(a.indices, a_before.indices) and np.array_equal(a.data, a_before.data)
    # dense @ sparse
    d1 = np.asarray(m1.toarray(), order=order)
    s2 = m2.asformat(fmt) if fmt != "bsr" else m2.tobsr(blocksize=(10, 10))
    got = sdb.dot_product_mkl(d1, s2)
    _close(got, d1, m2, dtype)


@pytest.mark.parametrize("dtype", REAL)
@pytest.mark.parametrize("order", ["C", "F"])
def test_sparse_dense_out_and_scalar(dtype, order):
    """test_sparse_dense.py:42-52: out=ones, out_scalar=3 -> AB + 3; identity kept."""
    m1, m2 = _pair(dtype)
    b = np.asarray(m2.toarray(), order=order)
    want = (m1.astype(np.float64) @ b.astype(np.float64))
    out = np.ones((200, 100), dtype=dtype, order=order)
    got = sdb.dot_product_mkl(m1, b, out=out, out_scalar=3.0)
    assert got is out
    _close(got, m1, b, dtype, want=want, offset=3.0)
    out = np.ones((200, 100), dtype=dtype, order=order)
    got = sdb.dot_product_mkl(m1, b, out=out)  # beta defaults to 1
    _close(got, m1, b, dtype, want=want, offset=1.0)
    # dense @ sparse with out
    d1 = np.asarray(m1.toarray(), order=order)
    out = np.ones((200, 100), dtype=dtype, order=order)
    got = sdb.dot_product_mkl(d1, m2, out=out, out_scalar=0.5)
    assert got is out
    _close(got, d1, m2, dtype, want=want, offset=0.5)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 7, 8, 31, 33, 64, 100, 128, 130, 256, 300])
@pytest.mark.parametrize("dtype", REAL)
def test_spmm_widths_against_c_oracle(n, dtype):
    """Every lane-group / vector-width specialisation, vs the C restatement."""
    a = cs.uniform_rows_csr(3000, 2000, 17, dtype, seed=n)
    # a few empty and a few long rows
    a = a.tolil()
    a[5, :] = 0
    a[6, :] = 0
    a = a.tocsr()
    long_row = sp.random(1, 2000, density=0.6, format="csr", dtype=dtype, random_state=3)
    a = sp.vstack([a, long_row]).tocsr()
    x = np.random.default_rng(n).random((2000, n)).astype(dtype)
    if n in (1,):
        x2 = x.copy()
        got = sdb.dot_product_mkl(a, x2.reshape(-1, 1))
    else:
        got = sdb.dot_product_mkl(a, x)
    want = orc.c_spmm(a, x)
    bound = orc.value_bound(abs(a), abs(x))
    assert cs.rel_err(got.reshape(want.shape), want, bound) <= cs.TOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", REAL)
@pytest.mark.parametrize("b", [4, 8, 16, 32])
@pytest.mark.parametrize("n", [4, 36, 100, 128, 256, 260])
def test_bsr_native_kernel(dtype, b, n):
    """The TMA-staged BSR x dense kernel (block sizes 4/8/16/32), incl. ragged
    column chunks, empty block rows, beta != 0 and column-major blocks."""
    rng = np.random.default_rng(b * 1000 + n)
    mb, kb = 37, 29
    dense_mask = rng.random((mb, kb)) < 0.2
    dense_mask[3, :] = False  # an empty block row
    dense_mask[5, :] = True   # a full one (more blocks than pipeline stages)
    a = sp.bsr_matrix(sp.kron(sp.csr_matrix(dense_mask.astype(dtype)), np.ones((b, b), dtype=dtype)), blocksize=(b, b))
    a.data[:] = rng.random(a.data.shape) + 0.5
    x = rng.random((kb * b, n)).astype(dtype)
    want = orc.c_spmm(a.tocsr(), x)
    bound = orc.value_bound(abs(a.tocsr()), abs(x))
    got = sdb.dot_product_mkl(a, x)
    assert cs.rel_err(got, want, bound) <= cs.TOL[np.dtype(dtype)]
    y0 = rng.random((mb * b, n)).astype(dtype)
    got = sdb.dot_product_mkl(a, x, out=y0.copy(), out_scalar=0.5)
    want2 = orc.c_spmm(a.tocsr(), x, beta=0.5, y=y0.copy())
    assert cs.rel_err(got, want2, bound + 0.5 * y0) <= cs.TOL[np.dtype(dtype)]
    # column-major blocks: same matrix, data stored transposed per block
    af = a.copy()
    af.data = np.ascontiguousarray(a.data.transpose(0, 2, 1)).transpose(0, 2, 1)
    assert not af.data.flags.c_contiguous
    got = sdb.dot_product_mkl(af, x)
    assert cs.rel_err(got, want, bound) <= cs.TOL[np.dtype(dtype)]


def test_spmm_ignores_garbage_in_fresh_output_and_nan_free():
    a = cs.uniform_rows_csr(500, 400, 9, np.float32, seed=1)
    x = np.random.default_rng(0).random((400, 16)).astype(np.float32)
    x[0, :] = np.inf  # only rows that reference column 0 may become inf
    got = sdb.dot_product_mkl(a, x)
    touches0 = np.asarray((a[:, 0] != 0).todense()).ravel()
    assert np.all(np.isfinite(got[~touches0]))
    assert np.all(np.isinf(got[touches0]))


def test_config1_fp64_against_mkl_golden():
    """BASELINE configs[0]: CSR 10k x 10k d=1e-3 fp64 x dense 10k x 64, checked
    against the output of real oneMKL committed by oracle/gen_golden.py."""
    import os

    path = os.path.join(os.path.dirname(__file__), "golden", "c1_spmm_f64.npz")
    g = np.load(path)
    a = sp.random(10_000, 10_000, density=1e-3, format="csr", dtype=np.float64, random_state=86)
    b = np.random.default_rng(88).random((10_000, 64))
    got = sdb.dot_product_mkl(a, b)
    rows = g["rows"]
    assert cs.rel_err(got[rows], g["y_rows"]) <= 1e-12
    assert abs(got.sum() - float(g["y_sum"])) <= 1e-9 * abs(float(g["y_sum"]))
    want = orc.c_spmm(a, b)
    assert cs.rel_err(got, want, orc.value_bound(abs(a), abs(b))) <= 1e-12


def test_sparse_dense_errors():
    m1, m2 = _pair(np.float64)
    b = m2.toarray()
    with pytest.raises(ValueError):  # mixed dtypes without cast
        sdb.dot_product_mkl(m1.astype(np.float32), b)
    got = sdb.dot_product_mkl(m1.astype(np.float32), b, cast=True)
    assert got.dtype == np.float64
    with pytest.raises(ValueError):  # wrong out dtype / order / shape
        sdb.dot_product_mkl(m1, b, out=np.ones((200, 100), dtype=np.float32))
    with pytest.raises(ValueError):
        sdb.dot_product_mkl(m1, b, out=np.ones((200, 100), order="F"))
    with pytest.raises(ValueError):
        sdb.dot_product_mkl(m1, b, out=np.ones((100, 200)))
    with pytest.raises(ValueError):  # misaligned
        sdb.dot_product_mkl(m1, b.T)
    with pytest.raises(ValueError):  # not contiguous
        sdb.dot_product_mkl(m1, np.ones((600, 100))[::2])
    with pytest.raises(ValueError):  # COO
        sdb.dot_product_mkl(m1.tocoo(), b)


def test_empty_products():
    """test_mkl.py:70-103."""
    empty = sp.csr_matrix((200, 300), dtype=np.float64)
    _, m2 = _pair(np.float64)
    got = sdb.dot_product_mkl(empty, m2.toarray())
    assert got.shape == (200, 100) and not got.any()
    got = sdb.dot_product_mkl(empty, m2)
    assert sp.issparse(got) and got.shape == (200, 100) and got.nnz == 0
    got = sdb.dot_product_mkl(empty, m2, dense=True)
    assert got.shape == (200, 100) and not got.any()


# =============================================================== sparse x vector
@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("fmt", ["csr", "csc", "bsr"])
def test_sparse_vector(dtype, fmt):
    """test_sparse_vector.py: (N,) and (N,1) vectors on either side."""
    m1, m2 = _pair(dtype)
    a = m1.asformat(fmt) if fmt != "bsr" else m1.tobsr(blocksize=(10, 10))
    v = cs.make_vector(300, complex=np.dtype(dtype).kind == "c").astype(dtype)
    got = sdb.dot_product_mkl(a, v)
    assert got.shape == (200,)
    _close(got.reshape(-1, 1), m1, v.reshape(-1, 1), dtype)
    got = sdb.dot_product_mkl(a, v.reshape(-1, 1))
    assert got.shape == (200, 1)
    _close(got, m1, v.reshape(-1, 1), dtype)
    b = m2.asformat(fmt) if fmt != "bsr" else m2.tobsr(blocksize=(10, 10))
    got = sdb.dot_product_mkl(v, b)
    assert got.shape == (100,)
    _close(got.reshape(1, -1), v.reshape(1, -1), m2, dtype)
    got = sdb.dot_product_mkl(v.reshape(1, -1), b)
    assert got.shape == (1, 100)
    _close(got, v.reshape(1, -1), m2, dtype)
    out = np.ones(200, dtype=dtype)
    got = sdb.dot_product_mkl(a, v, out=out, out_scalar=2.0)
    assert got is out
    _close(got.reshape(-1, 1), m1, v.reshape(-1, 1), dtype, offset=2.0)


# =============================================================== sparse x sparse
def _check_sparse_product(got, a, b, dtype, sorted_required):
    want = orc.canonical(orc.np_spgemm(a, b))
    if sorted_required:
        assert got.has_canonical_format or np.all(
            [np.all(np.diff(got.indices[got.indptr[i]:got.indptr[i + 1]]) > 0) for i in range(got.shape[0])]
        ), "reorder_output=True must return ascending columns"
        canon = got
    else:
        canon = orc.canonical(got)
    # positive inputs: structural == numerical nonzeros, so scipy's pattern is the oracle's
    assert np.array_equal(canon.indptr, want.indptr)
    assert np.array_equal(canon.indices, want.indices)
    bound = orc.canonical(abs(a) @ abs(b))
    err = np.abs(canon.data - want.data) / np.abs(bound.data)
    assert err.max() <= cs.TOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("reorder", [False, True])
def test_spgemm_fixture(dtype, reorder):
    """test_sparse_sparse.py:87-118."""
    m1, m2 = _pair(dtype)
    got = sdb.dot_product_mkl(m1, m2, reorder_output=reorder)
    assert isinstance(got, sp.csr_matrix) and got.dtype == np.dtype(dtype)
    if np.dtype(dtype).kind == "c":
        _close(got, m1, m2, dtype)
    else:
        assert got.nnz == 10491  # SURVEY §4 probe
        _check_sparse_product(got, m1, m2, dtype, reorder)


@pytest.mark.parametrize("dtype", REAL)
def test_spgemm_against_c_oracle_and_mkl_golden(dtype):
    import os

    m1, m2 = _pair(dtype)
    got = sdb.dot_product_mkl(m1, m2, reorder_output=True)
    want = orc.c_spgemm(m1, m2, sort=True)
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert cs.rel_err(got.data, want.data) <= cs.TOL[np.dtype(dtype)]
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fixture_spgemm.npz"))
    assert np.array_equal(got.indptr, g["indptr"]) and np.array_equal(got.indices, g["indices"])
    assert cs.rel_err(got.data, g["data"].astype(dtype)) <= cs.TOL[np.dtype(dtype)]


@pytest.mark.parametrize("fa,fb", [("csc", "csc"), ("csr", "csc"), ("csc", "csr")])
def test_spgemm_csc_and_mixed(fa, fb):
    """test_sparse_sparse.py:156-172: result comes back in A's container."""
    m1, m2 = _pair(np.float64)
    got = sdb.dot_product_mkl(m1.asformat(fa), m2.asformat(fb), reorder_output=True)
    assert got.format == fa
    _close(got, m1, m2, np.float64)
    assert got.nnz == 10491


def test_spgemm_array_containers():
    m1, m2 = _pair(np.float64)
    got = sdb.dot_product_mkl(sp.csr_array(m1), sp.csr_array(m2))
    assert isinstance(got, sp.csr_array)
    _close(got, m1, m2, np.float64)


@pytest.mark.parametrize("dtype", REAL)
def test_spgemm_all_row_bins(dtype):
    """Power-law rows hit the warp, CTA and wide (dense-accumulator) bins."""
    a = cs.rmat_csr(13, 8, dtype, seed=1)
    b = cs.rmat_csr(13, 8, dtype, seed=2)
    got = sdb.dot_product_mkl(a, b, reorder_output=True)
    _check_sparse_product(got, a, b, dtype, True)
    got = sdb.dot_product_mkl(a, b)
    _check_sparse_product(got, a, b, dtype, False)


@pytest.mark.parametrize("dtype", REAL)
@pytest.mark.parametrize("upper", [False, True])
def test_spgemm_every_bin_by_construction(dtype, upper):
    """Rows built to land in the warp (<=256), CTA (<=4096), large (<=65536, global hash) and
    wide (>65536, bitmap + dense accumulator) bins; also the triangular (syrk) filter."""
    rng = np.random.default_rng(7)
    k, n = 2000, 30000
    b = sp.random(k, n, density=0.01, format="csr", dtype=dtype, random_state=3)
    b.data[:] = rng.random(b.nnz) + 0.5
    per_row = [0, 1, 2, 10, 12, 60, 100, 150, 290, 320, 1, 0, 7]
    rows = []
    for cnt in per_row:
        cols = np.sort(rng.choice(k, size=cnt, replace=False))
        rows.append(sp.csr_matrix((rng.random(cnt).astype(dtype) + 0.5, (np.zeros(cnt, dtype=int), cols)), shape=(1, k)))
    a = sp.vstack(rows).tocsr().astype(dtype)
    if not upper:
        got = sdb.dot_product_mkl(a, b, reorder_output=True)
        want = orc.c_spgemm(a, b, sort=True)
    else:
        # gram of M = [a; b-ish]: use syrk on a matrix whose A^T A rows span the bins
        m = sp.vstack([a, a[::-1]]).tocsr()
        got = sdb.gram_matrix_mkl(m, reorder_output=True)
        want = orc.c_syrk(m, sort=True)
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert cs.rel_err(got.data, want.data) <= cs.TOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("b", [10, 4])
def test_spgemm_bsr(dtype, b):
    """test_sparse_sparse.py:250-262 (TestMultiplicationBSR): BSR x BSR -> BSR with the same blocks."""
    m1, m2 = _pair(dtype)
    a1, a2 = m1.tobsr(blocksize=(b, b)), m2.tobsr(blocksize=(b, b))
    got = sdb.dot_product_mkl(a1, a2)
    assert isinstance(got, sp.bsr_matrix) and got.blocksize == (b, b) and got.dtype == np.dtype(dtype)
    _close(got, m1, m2, dtype)
    want = (a1 @ a2).tobsr(blocksize=(b, b))
    want.sort_indices()
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    got_arr = sdb.dot_product_mkl(sp.bsr_array(a1), sp.bsr_array(a2), reorder_output=True)
    assert isinstance(got_arr, sp.bsr_array)
    dense = sdb.dot_product_mkl(a1, a2, dense=True)
    _close(dense, m1, m2, dtype)
    with pytest.raises(ValueError):
        sdb.dot_product_mkl(m1, a2)  # CSR x BSR: not supported (the reference skips it too, :174-182)


@pytest.mark.parametrize("dtype", REAL)
@pytest.mark.parametrize("per_row", [1, 5, 20, 70])
def test_spmv_kernel_widths(dtype, per_row):
    """Dedicated SpMV kernel: every lanes-per-row specialisation, both ops, strided out."""
    a = cs.uniform_rows_csr(5000, 3000, per_row, dtype, seed=per_row)
    v = np.random.default_rng(1).random(3000).astype(dtype)
    got = sdb.dot_product_mkl(a, v)
    want = orc.c_spmm(a, v.reshape(-1, 1)).ravel()
    assert cs.rel_err(got, want) <= cs.TOL[np.dtype(dtype)]
    w = np.random.default_rng(2).random(5000).astype(dtype)
    got = sdb.dot_product_mkl(w, a)
    want = orc.c_spmm(a, w.reshape(-1, 1), op=orc.OP_T).ravel()
    assert cs.rel_err(got, want) <= cs.TOL[np.dtype(dtype)]
    out = np.ones((5000, 1), dtype=dtype)
    got = sdb.dot_product_mkl(a, v.reshape(-1, 1), out=out, out_scalar=0.5)
    assert got is out
    assert cs.rel_err(got.ravel(), orc.c_spmm(a, v.reshape(-1, 1)).ravel() + 0.5) <= cs.TOL[np.dtype(dtype)]


def test_spgemm_structural_zeros_are_kept():
    """MKL convention (SURVEY §8c hazard 2): cancellation keeps the entry."""
    a = sp.csr_matrix(np.array([[1.0, -1.0], [2.0, 0.0]]))
    b = sp.csr_matrix(np.array([[1.0, 3.0], [1.0, 0.0]]))
    got = sdb.dot_product_mkl(a, b, reorder_output=True)
    assert got.nnz == 4
    assert np.array_equal(got.toarray(), np.array([[0.0, 3.0], [2.0, 6.0]]))


@pytest.mark.parametrize("dtype", ALL)
def test_spgemm_dense_output(dtype):
    """test_sparse_sparse.py:264-297: dense=True, and out= is OVERWRITTEN."""
    m1, m2 = _pair(dtype)
    got = sdb.dot_product_mkl(m1, m2, dense=True)
    assert isinstance(got, np.ndarray)
    _close(got, m1, m2, dtype)
    out = np.full((200, 100), np.nan, dtype=dtype)
    got = sdb.dot_product_mkl(m1, m2, dense=True, out=out)
    assert got is out
    _close(got, m1, m2, dtype)
    if np.dtype(dtype).kind != "c":
        want = orc.c_spmmd(m1, m2)
        assert cs.rel_err(got, want, orc.value_bound(abs(m1), abs(m2))) <= cs.TOL[np.dtype(dtype)]
    with pytest.raises(ValueError):
        sdb.dot_product_mkl(m1, m2, out=out)  # out needs dense=True
    with pytest.raises(ValueError):
        sdb.dot_product_mkl(m1, m2, dense=True, out=np.ones((200, 100), dtype=dtype, order="F"))


# ===================================================================== gram
@pytest.mark.parametrize("dtype", REAL)
@pytest.mark.parametrize("aat", [False, True])
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_gram_sparse(dtype, aat, fmt):
    """test_gram_matrix.py:35-64."""
    m1, _ = _pair(dtype)
    a = m1.asformat(fmt)
    got = sdb.gram_matrix_mkl(a, transpose=aat, cast=True, reorder_output=True)
    assert isinstance(got, sp.csr_matrix)
    want = orc.c_syrk(m1, aat=aat, sort=True)
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert cs.rel_err(got.data, want.data) <= cs.TOL[np.dtype(dtype)]
    dense_want = orc.np_gram_upper(m1.astype(np.float64), aat=aat)
    assert np.abs(got.toarray() - dense_want).max() <= (1.5e-5 if dtype == np.float32 else 1.5e-6)
    if fmt == "csc":
        with pytest.raises(ValueError):
            sdb.gram_matrix_mkl(a, transpose=aat)


@pytest.mark.parametrize("dtype", REAL)
@pytest.mark.parametrize("aat", [False, True])
def test_gram_dense(dtype, aat):
    """test_gram_matrix.py:66-114."""
    m1, _ = _pair(dtype)
    got = sdb.gram_matrix_mkl(m1, transpose=aat, dense=True)
    want = orc.c_syrkd(m1, aat=aat)
    n = want.shape[0]
    assert got.shape == (n, n)
    assert np.all(got[np.tril_indices(n, -1)] == 0)
    bound = orc.np_gram_upper(abs(m1).astype(np.float64), aat=aat)
    assert cs.rel_err(got, want, np.maximum(bound, 0)) <= cs.TOL[np.dtype(dtype)]
    # out + out_scalar: upper = gram + 2 * out, lower untouched
    out = np.full((n, n), 7.0, dtype=dtype)
    got2 = sdb.gram_matrix_mkl(m1, transpose=aat, dense=True, out=out, out_scalar=2.0)
    assert got2 is out
    want2 = orc.c_syrkd(m1, aat=aat, beta=2.0, out=np.full((n, n), 7.0, dtype=dtype))
    assert np.all(got2[np.tril_indices(n, -1)] == 7.0)
    iu = np.triu_indices(n)
    assert cs.rel_err(got2[iu], want2[iu], (bound + 14.0)[iu]) <= cs.TOL[np.dtype(dtype)]


@pytest.mark.parametrize("aat", [False, True])
def test_gram_unsorted_rows_and_duplicate_entries(aat):
    """Triangular products start each R-row walk at the diagonal only when every row is strictly
    ascending; shuffled rows and duplicate (unsummed) entries must take the plain filtered walk."""
    m1, _ = _pair(np.float64)
    rng = np.random.default_rng(3)
    shuffled = m1.copy()
    for i in range(shuffled.shape[0]):
        s0, e0 = shuffled.indptr[i], shuffled.indptr[i + 1]
        p = rng.permutation(e0 - s0)
        shuffled.indices[s0:e0] = shuffled.indices[s0:e0][p]
        shuffled.data[s0:e0] = shuffled.data[s0:e0][p]
    shuffled.has_sorted_indices = False
    dup = sp.csr_matrix((np.concatenate([m1.data, m1.data[:50]]), np.concatenate([m1.indices, m1.indices[:50]]),
                         np.concatenate([m1.indptr[:1], m1.indptr[1:] + 50])), shape=m1.shape)
    dup.indices[: m1.indptr[1] + 50] = np.concatenate([m1.indices[: m1.indptr[1]], m1.indices[:50]])
    for a in (shuffled, dup):
        ref = sp.csr_matrix(a.toarray())  # canonical copy: duplicates summed, sorted
        want = orc.np_gram_upper(ref, aat=aat)
        got = sdb.gram_matrix_mkl(a, transpose=aat, dense=True)
        assert np.abs(got - want).max() <= 1e-10
        got_s = sdb.gram_matrix_mkl(a, transpose=aat, reorder_output=True)
        assert np.abs(got_s.toarray() - want).max() <= 1e-10


def test_gram_errors():
    m1, _ = _pair(np.float64)
    with pytest.raises(ValueError):
        sdb.gram_matrix_mkl(cs.complexify(m1, 1))
    with pytest.raises(ValueError):
        sdb.gram_matrix_mkl(m1, out=np.zeros((300, 300)))
    with pytest.raises(ValueError):
        sdb.gram_matrix_mkl(m1, dense=True, out=np.zeros((300, 300), dtype=np.float32))
    with pytest.raises(ValueError):
        sdb.gram_matrix_mkl(m1.tobsr(blocksize=(10, 10)))


def test_gram_reduced_config4():
    """BASELINE configs[3] recipe at 20k x 2k, 100 nnz/row fp32, dense upper."""
    a = cs.uniform_rows_csr(20_000, 2_000, 100, np.float32, seed=4)
    got = sdb.gram_matrix_mkl(a, dense=True)
    want = orc.c_syrkd(a)
    assert np.all(got[np.tril_indices(2000, -1)] == 0)
    assert cs.rel_err(got, want, np.maximum(want, 0)) <= 1e-5


# ===================================================== real-MKL golden vectors (tests/golden)
def _gold(name):
    import os

    return np.load(os.path.join(os.path.dirname(__file__), "golden", name))


@pytest.mark.parametrize("dtype,tag", [(np.float32, "f32"), (np.float64, "f64")])
def test_spmm_matches_mkl_golden(dtype, tag):
    """The reference's own fixture through real oneMKL (oracle/gen_golden.py) vs the CUDA path."""
    g = _gold(f"fixture_spmm_{tag}.npz")
    m1, m2 = cs.fixture_pair(dtype)
    b = m2.toarray()
    tol = cs.TOL[np.dtype(dtype)]
    bound = orc.value_bound(abs(m1), abs(b))
    assert cs.rel_err(sdb.dot_product_mkl(m1, b), g["y"], bound) <= tol
    got = sdb.dot_product_mkl(m1, b, out=np.ones((200, 100), dtype=dtype), out_scalar=3.0)
    assert cs.rel_err(got, g["y_out3"], bound + 3.0) <= tol
    x = np.ascontiguousarray(m1.toarray()[:, :50])
    got = sdb.dot_product_mkl(x.T.copy(), m1).T  # (X^T A)^T = A^T X
    assert cs.rel_err(got, g["yt"], orc.value_bound(abs(m1.T), abs(x))) <= tol


def test_spmmd_gram_and_cancellation_match_mkl_golden():
    m1, m2 = cs.fixture_pair(np.float64)
    got = sdb.dot_product_mkl(m1, m2, dense=True)
    assert cs.rel_err(got, _gold("fixture_spmmd.npz")["c"], orc.value_bound(abs(m1), abs(m2))) <= 1e-12
    g = _gold("fixture_gram.npz")
    for aat, key in ((False, "ata"), (True, "aat")):
        c = sdb.gram_matrix_mkl(m1, transpose=aat, reorder_output=True)
        assert np.array_equal(c.indptr, g[f"{key}_indptr"]) and np.array_equal(c.indices, g[f"{key}_indices"])
        assert cs.rel_err(c.data, g[f"{key}_data"]) <= 1e-12
    z = _gold("cancel_spgemm.npz")
    a = sp.csr_matrix(np.array([[1.0, -1.0], [2.0, 0.0]]))
    b = sp.csr_matrix(np.array([[1.0, 3.0], [1.0, 0.0]]))
    c = sdb.dot_product_mkl(a, b, reorder_output=True)
    assert np.array_equal(c.indptr, z["indptr"]) and np.array_equal(c.indices, z["indices"])
    assert np.array_equal(c.data, z["data"])


def test_config2_small_matches_mkl_golden():
    g = _gold("c2_small_spmm_f32.npz")
    a = cs.uniform_rows_csr(20_000, 20_000, 50, np.float32, seed=0)
    x = np.random.default_rng(2).random((20_000, 128), dtype=np.float32)
    y0 = np.random.default_rng(3).random((20_000, 128), dtype=np.float32)
    y = sdb.dot_product_mkl(a, x, out=y0.copy(), out_scalar=0.5)
    assert cs.rel_err(y[g["rows"]], g["y_rows"]) <= 1e-5
    assert np.allclose(y.astype(np.float64).sum(axis=0), g["y_colsum"], rtol=1e-5)


# ===================================================== device-resident operands
def test_resident_operands():
    import torch

    m1, m2 = cs.fixture_pair(np.float32)
    x = np.random.default_rng(0).random((300, 64), dtype=np.float32)
    with sdb.ResidentCSR(m1) as a, sdb.ResidentCSR(m2) as b:
        assert a.shape == (200, 300) and a.nnz == m1.nnz
        want = orc.c_spmm(m1, x)
        bound = orc.value_bound(abs(m1), abs(x))
        for _ in range(3):  # repeated products, A uploaded once
            assert cs.rel_err(a.dot(x), want, bound) <= 1e-5
        got = a.dot(np.asfortranarray(x))
        assert got.flags.f_contiguous and cs.rel_err(got, want, bound) <= 1e-5
        xt = torch.from_numpy(x).cuda()
        yt = torch.ones((200, 64), dtype=torch.float32, device="cuda")
        a.dot_device(xt, yt, alpha=1.0, beta=3.0)
        torch.cuda.synchronize()
        assert cs.rel_err(yt.cpu().numpy(), want + 3.0, bound + 3.0) <= 1e-5
        with a.matmat(b, reorder_output=True) as c:
            got = c.to_scipy()
            w = orc.c_spgemm(m1, m2, sort=True)
            assert np.array_equal(got.indptr, w.indptr) and np.array_equal(got.indices, w.indices)
            assert cs.rel_err(got.data, w.data) <= 1e-5
        with a.gram(reorder_output=True) as g:
            w = orc.c_syrk(m1, sort=True)
            got = g.to_scipy()
            assert np.array_equal(got.indices, w.indices) and cs.rel_err(got.data, w.data) <= 1e-5
        with pytest.raises(ValueError):
            a.dot(x.astype(np.float64))
        with pytest.raises(ValueError):
            a.dot(x[:100])


def test_order_bsr_and_csc_handles():
    """mkl_sparse_order on BSR (whole blocks move with their column index) and CSC handles."""
    m1, _ = cs.fixture_pair(np.float64)
    bsr = m1.tobsr(blocksize=(10, 10))
    rng = np.random.default_rng(9)
    shuffled = bsr.copy()
    for i in range(shuffled.indptr.shape[0] - 1):
        s0, e0 = shuffled.indptr[i], shuffled.indptr[i + 1]
        p = rng.permutation(e0 - s0)
        shuffled.indices[s0:e0] = shuffled.indices[s0:e0][p]
        shuffled.data[s0:e0] = shuffled.data[s0:e0][p]
    shuffled.has_sorted_indices = False
    h, _, _ = H.create(shuffled)
    with h:
        H.order(h)
        got = H.export(h, output_type="bsr_matrix")
    want = bsr.copy()
    want.sort_indices()  # tobsr() does not order the block columns itself
    assert np.array_equal(got.indptr, want.indptr)
    assert np.array_equal(got.indices, want.indices) and np.array_equal(got.data, want.data)
    csc = m1.tocsc()
    sh = csc.copy()
    for j in range(sh.shape[1]):
        s0, e0 = sh.indptr[j], sh.indptr[j + 1]
        sh.indices[s0:e0] = sh.indices[s0:e0][::-1].copy()
        sh.data[s0:e0] = sh.data[s0:e0][::-1].copy()
    sh.has_sorted_indices = False
    h, _, _ = H.create(sh)
    with h:
        H.order(h)
        got = H.export(h, output_type="csc_matrix")
    assert np.array_equal(got.indices, csc.indices) and np.array_equal(got.data, csc.data)


@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("shape_case", ["row", "col"])
@pytest.mark.parametrize("dtype", REAL)
def test_one_row_and_one_column_operands(order, shape_case, dtype):
    """test_sparse_dense.py:293-334: the same method set with a 1-row left operand or a 1-column
    right operand (arrays that are both C- and F-contiguous; some calls dispatch to the vector path)."""
    M1, M2 = _pair(dtype)
    m1 = M1[[0], :] if shape_case == "row" else M1
    m2 = M2 if shape_case == "row" else M2[:, [0]]
    m1_d = np.asarray(M1.toarray(), order=order)[[0], :] if shape_case == "row" else np.asarray(M1.toarray(), order=order)
    m2_d = np.asarray(M2.toarray(), order=order) if shape_case == "row" else np.asarray(M2.toarray(), order=order)[:, [0]]
    want = m1_d.astype(np.float64) @ m2_d.astype(np.float64)
    tol = 1.5e-6 if dtype == np.float64 else 1.5e-5  # the reference's assert_array_almost_equal decimals
    for a, b in ((m1, m2_d), (m1_d, m2), (m1.tocsc(), m2_d), (m1_d, m2.tocsc())):
        got = sdb.dot_product_mkl(a, b)
        assert got.shape == want.shape and np.abs(got - want).max() < tol
        out = np.ones(want.shape, dtype=dtype, order=order)
        got = sdb.dot_product_mkl(a, b, out=out, out_scalar=3.0)
        assert got is out and np.abs(got - (want + 3.0)).max() < tol
    got = sdb.dot_product_mkl(m1, m2)
    assert sp.issparse(got) and np.abs(got.toarray() - want).max() < tol
    got = sdb.dot_product_mkl(m1, m2, dense=True)
    assert np.abs(got - want).max() < tol
