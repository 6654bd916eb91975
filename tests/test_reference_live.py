"""
The UNMODIFIED reference package as the checker.

`oracle/ref_pkg.py` imports sparse_dot_mkl 0.9.6 from baseline/_ref with $MKL_RT pointing at
oracle/_ref/libmkl_fwd.so, i.e. the reference's own Python (validation, dispatch, handle creation, export —
sparse_dot.py:18-253, _sparse_dense.py, _sparse_sparse.py, _sparse_vector.py) over Intel's real
mkl_sparse_?_mm / mkl_sparse_spmm / ?_spmmd / ?_mv inside torch's libtorch_cpu.so.

CPU (`not gpu`): the reference reproduces the committed golden vectors (so the goldens ARE the reference's
outputs) and the C oracle agrees with the reference on seeded cases (the pin of oracle/sdb_oracle.c).
GPU: `sparse_dot_b200.dot_product_mkl` against `sparse_dot_mkl.dot_product_mkl` on the same inputs, same
kwargs — bit-exact indptr / indices, values within 1e-5 (fp32, c64) / 1e-12 (fp64, c128) of |A|·|B|.

Gram is absent here: the embedded oneMKL has no syrk/syrkd (tests/test_gpu_parity.py checks gram against
the C oracle and numpy instead).
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc
from oracle import ref_pkg
from tests import _cases as cs

REF, REF_STATUS = ref_pkg.load()
pytestmark = pytest.mark.skipif(REF is None, reason=f"unmodified reference not loadable: {REF_STATUS}")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _bound(a, b):
    return orc.value_bound(abs(a), abs(b))


def _dense_pair(dtype, seed=11, m=300, k=400, n=24, density=0.03):
    rng = np.random.default_rng(seed)
    a = sp.random(m, k, density=density, format="csr", dtype=np.float64, random_state=seed)
    b = rng.random((k, n))
    if np.dtype(dtype).kind == "c":
        a = cs.complexify(a, seed + 1)
        b = b + 1j * rng.random((k, n))
    return a.astype(dtype), b.astype(dtype)


# ----------------------------------------------------------------------------- CPU: pin goldens and oracle
def test_reference_reproduces_the_golden_vectors():
    m1, m2 = cs.fixture_pair(np.float64)
    g = np.load(os.path.join(GOLD, "fixture_spgemm.npz"))
    c = REF.dot_product_mkl(m1, m2, reorder_output=True)
    assert np.array_equal(c.indptr, g["indptr"]) and np.array_equal(c.indices, g["indices"])
    assert cs.rel_err(c.data, g["data"]) <= 1e-12
    d = REF.dot_product_mkl(m1, m2, dense=True)
    assert cs.rel_err(d, np.load(os.path.join(GOLD, "fixture_spmmd.npz"))["c"], _bound(m1, m2)) <= 1e-12
    for dtype, tag in ((np.float32, "f32"), (np.float64, "f64")):
        a, b = cs.fixture_pair(dtype)
        gg = np.load(os.path.join(GOLD, f"fixture_spmm_{tag}.npz"))
        bd = b.toarray()
        tol = cs.TOL[np.dtype(dtype)]
        assert cs.rel_err(REF.dot_product_mkl(a, bd), gg["y"], _bound(a, bd)) <= tol
        y = REF.dot_product_mkl(a, bd, out=np.ones((200, 100), dtype=dtype), out_scalar=3.0)
        assert cs.rel_err(y, gg["y_out3"], _bound(a, bd) + 3.0) <= tol


def test_reference_reproduces_baseline_config0_golden():
    """BASELINE.json configs[0]: CSR(10k x 10k, density 1e-3, fp64) x dense(10k x 64) via dot_product_mkl."""
    g = np.load(os.path.join(GOLD, "c1_spmm_f64.npz"))
    a = sp.random(10_000, 10_000, density=1e-3, format="csr", dtype=np.float64, random_state=86)
    b = np.random.default_rng(88).random((10_000, 64))
    y = REF.dot_product_mkl(a, b)
    assert a.nnz == int(g["nnz"])
    assert cs.rel_err(y[g["rows"]], g["y_rows"]) <= 1e-12
    assert abs(y.sum() - float(g["y_sum"])) <= 1e-10 * abs(float(g["y_sum"]))
    assert cs.rel_err(y.sum(axis=0), g["y_colsum"]) <= 1e-11


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_c_oracle_agrees_with_the_reference(dtype):
    tol = cs.TOL[np.dtype(dtype)]
    a, b = _dense_pair(dtype)
    assert cs.rel_err(orc.c_spmm(a, b), REF.dot_product_mkl(a, b), _bound(a, b)) <= tol
    y0 = np.random.default_rng(5).random((a.shape[0], b.shape[1])).astype(dtype)
    want = REF.dot_product_mkl(a, b, out=y0.copy(), out_scalar=0.5)
    assert cs.rel_err(orc.c_spmm(a, b, beta=0.5, y=y0.copy()), want, _bound(a, b) + 0.5 * y0) <= tol
    m1, m2 = cs.fixture_pair(dtype)
    c, w = orc.c_spgemm(m1, m2, sort=True), REF.dot_product_mkl(m1, m2, reorder_output=True)
    assert np.array_equal(c.indptr, w.indptr) and np.array_equal(c.indices, w.indices)
    assert cs.rel_err(c.data, w.data) <= tol
    assert cs.rel_err(orc.c_spmmd(m1, m2), REF.dot_product_mkl(m1, m2, dense=True), _bound(m1, m2)) <= tol


def _error_cases():
    a = sp.random(30, 40, density=0.2, format="csr", random_state=1)
    b = np.random.default_rng(0).random((40, 5))
    v = np.random.default_rng(0).random(40)
    sq = sp.random(12, 12, density=0.3, format="csr", random_state=2)
    return {
        "coo x coo": lambda m: m.dot_product_mkl(a.tocoo(), a.T.tocoo()),
        "misaligned sparse x dense": lambda m: m.dot_product_mkl(a, b[:-1]),
        "misaligned sparse x sparse": lambda m: m.dot_product_mkl(a, a),
        "misaligned sparse x vector": lambda m: m.dot_product_mkl(a, v[:-1]),
        "3-d operand": lambda m: m.dot_product_mkl(a, np.zeros((40, 5, 2))),
        "mixed precision without cast": lambda m: m.dot_product_mkl(a.astype(np.float32), b),
        "integer data without cast": lambda m: m.dot_product_mkl(a.astype(np.int32), b),
        "out: wrong shape": lambda m: m.dot_product_mkl(a, b, out=np.zeros((3, 3))),
        "out: wrong dtype": lambda m: m.dot_product_mkl(a, b, out=np.zeros((30, 5), dtype=np.float32)),
        "out: wrong order": lambda m: m.dot_product_mkl(a, b, out=np.zeros((30, 5), order="F")),
        "out: not contiguous": lambda m: m.dot_product_mkl(a, b, out=np.zeros((30, 10))[:, ::2]),
        "out with a sparse result": lambda m: m.dot_product_mkl(a, a.T.tocsr(), out=np.zeros((30, 30))),
        "dense operand not contiguous": lambda m: m.dot_product_mkl(a, np.zeros((40, 10))[:, ::2]),
        "gram: complex": lambda m: m.gram_matrix_mkl(a.astype(np.complex128)),
        "gram: CSC without cast": lambda m: m.gram_matrix_mkl(a.tocsc()),
        "gram: BSR": lambda m: m.gram_matrix_mkl(sq.tobsr((3, 3))),
        "gram: out with a sparse result": lambda m: m.gram_matrix_mkl(a, out=np.zeros((40, 40))),
        "empty sparse x dense": lambda m: m.dot_product_mkl(sp.csr_matrix((30, 40)), b),
        "empty sparse x sparse": lambda m: m.dot_product_mkl(sp.csr_matrix((30, 40)), sp.csr_matrix((40, 7))),
    }


def _outcome(call, module):
    try:
        r = call(module)
        return ("returned", type(r).__name__, tuple(r.shape), str(r.dtype), float(abs(r).sum()))
    except Exception as e:  # noqa: BLE001
        return (type(e).__name__, str(e))


@pytest.mark.parametrize("name", sorted(_error_cases()))
def test_validation_paths_behave_like_the_reference(name):
    """Everything the reference decides BEFORE it touches MKL — refusals and the empty-product shortcut
    (_common.py:725-955, 1003-1024; sparse_dot.py; _gram_matrix.py:283-310) — needs no GPU in this package
    either: same exception type AND the same message, or the same kind of empty result."""
    import sparse_dot_b200 as sdb

    call = _error_cases()[name]
    assert _outcome(call, sdb) == _outcome(call, REF)


# ----------------------------------------------------------------------------- GPU: ours vs the reference, live
ALL = [np.float32, np.float64, np.complex64, np.complex128]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("fmt", ["csr", "csc"])
@pytest.mark.parametrize("order", ["C", "F"])
def test_gpu_sparse_dense_matches_reference(dtype, fmt, order):
    import sparse_dot_b200 as sdb

    tol = cs.TOL[np.dtype(dtype)]
    a, b = _dense_pair(dtype)
    a = a.asformat(fmt)
    b = np.asarray(b, order=order)
    want = REF.dot_product_mkl(a, b)
    got = sdb.dot_product_mkl(a, b)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert got.flags.c_contiguous == want.flags.c_contiguous and got.flags.f_contiguous == want.flags.f_contiguous
    assert cs.rel_err(got, want, _bound(a, b)) <= tol
    # out= accumulate (beta = out_scalar), the BASELINE configs[1] path
    y0 = np.asarray(np.random.default_rng(5).random(want.shape).astype(dtype), order=order)
    w = REF.dot_product_mkl(a, b, out=y0.copy(order=order), out_scalar=0.5)
    o = y0.copy(order=order)
    g = sdb.dot_product_mkl(a, b, out=o, out_scalar=0.5)
    assert g is o
    assert cs.rel_err(g, w, _bound(a, b) + 0.5 * np.abs(y0)) <= tol
    # dense x sparse (the reference runs it as (A^T B^T)^T, _sparse_dense.py:191-208)
    bt = np.asarray(b.T, order=order)
    w = REF.dot_product_mkl(bt, a.T.asformat(fmt))
    g = sdb.dot_product_mkl(bt, a.T.asformat(fmt))
    assert cs.rel_err(g, w, _bound(a, b).T) <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_gpu_sparse_sparse_matches_reference(dtype, fmt):
    import sparse_dot_b200 as sdb

    tol = cs.TOL[np.dtype(dtype)]
    m1, m2 = cs.make_matrixes(400, 250, 600, 0.02)
    if np.dtype(dtype).kind == "c":
        m1, m2 = cs.complexify(m1, 1), cs.complexify(m2, 2)
    m1, m2 = m1.astype(dtype).asformat(fmt), m2.astype(dtype).asformat(fmt)
    want = REF.dot_product_mkl(m1, m2, reorder_output=True)
    got = sdb.dot_product_mkl(m1, m2, reorder_output=True)
    assert type(got) is type(want) and got.shape == want.shape and got.dtype == want.dtype
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert cs.rel_err(got.data, want.data) <= 50 * tol  # relative to the entry itself: sums of ~12 products
    # unsorted output: same structure once both are put in canonical order
    gu, wu = sdb.dot_product_mkl(m1, m2), REF.dot_product_mkl(m1, m2)
    gu.sort_indices(), wu.sort_indices()
    assert np.array_equal(gu.indptr, wu.indptr) and np.array_equal(gu.indices, wu.indices)
    # dense output, overwrite semantics (garbage in `out` must not leak, test_sparse_sparse.py:286-297)
    wd = REF.dot_product_mkl(m1, m2, dense=True)
    out = np.full(wd.shape, 7.0, dtype=dtype)
    gd = sdb.dot_product_mkl(m1, m2, dense=True, out=out)
    assert gd is out and cs.rel_err(gd, wd, _bound(m1, m2)) <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ALL)
def test_gpu_sparse_vector_matches_reference(dtype):
    import sparse_dot_b200 as sdb

    tol = cs.TOL[np.dtype(dtype)]
    a, b = _dense_pair(dtype)
    v = np.ascontiguousarray(b[:, 0])
    want, got = REF.dot_product_mkl(a, v), sdb.dot_product_mkl(a, v)
    assert got.shape == want.shape and cs.rel_err(got, want, _bound(a, v.reshape(-1, 1)).ravel()) <= tol
    u = np.ascontiguousarray(b[: a.shape[0], 1]) if b.shape[0] >= a.shape[0] else np.ones(a.shape[0], dtype=dtype)
    want, got = REF.dot_product_mkl(u, a), sdb.dot_product_mkl(u, a)
    assert got.shape == want.shape and cs.rel_err(got, want, _bound(a.T, u.reshape(-1, 1)).ravel()) <= tol


@pytest.mark.gpu
def test_gpu_error_behaviour_matches_reference():
    """Same exception type from both packages on the error paths the reference's tests pin."""
    import sparse_dot_b200 as sdb

    a, b = _dense_pair(np.float64)
    cases = [
        lambda f: f(a.tocoo(), a.T.tocoo()),                                  # COO refused (test_sparse_sparse.py:184-188)
        lambda f: f(a, b[:-1]),                                               # misaligned (test_mkl.py:143-170)
        lambda f: f(a.astype(np.float32), b),                                 # mixed dtype without cast
        lambda f: f(a, b, out=np.zeros((3, 3))),                              # bad out shape
        lambda f: f(a, b, out=np.zeros((a.shape[0], b.shape[1]), dtype=np.float32)),  # bad out dtype
        lambda f: f(a, b, out=np.zeros((a.shape[0], b.shape[1]), order="F")),  # bad out order
        lambda f: f(a, a.T.tocsr(), out=np.zeros((a.shape[0], a.shape[0]))),   # out without dense=True
    ]
    for i, case in enumerate(cases):
        errs = []
        for f in (REF.dot_product_mkl, sdb.dot_product_mkl):
            try:
                case(f)
                errs.append(None)
            except Exception as e:  # noqa: BLE001
                errs.append(type(e))
        assert errs[0] is errs[1], (i, errs)
