"""
Pins the CPU oracle (oracle/sdb_oracle.c through oracle/oracle.py) BEFORE it is
trusted as the checker of the CUDA path:
  1. against tests/golden/*.npz — outputs of real oneMKL driven with the
     reference's call sequence on the reference's own fixtures
     (oracle/gen_golden.py, oracle/mkl_ref.py);
  2. against numpy/scipy on the same inputs — the comparator every reference
     test uses (sparse_dot_mkl/tests/test_mkl.py:53-67);
  3. live against the embedded oneMKL when it is reachable in this process.
CPU only.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import mkl_ref
from oracle import oracle as orc
from tests import _cases as cs

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REAL = [np.float32, np.float64]


def gold(name):
    return np.load(os.path.join(GOLD, name))


@pytest.mark.parametrize("dtype,tag", [(np.float32, "f32"), (np.float64, "f64")])
def test_spmm_matches_mkl_golden(dtype, tag):
    g = gold(f"fixture_spmm_{tag}.npz")
    m1, m2 = cs.fixture_pair(dtype)
    b = m2.toarray()
    tol = cs.TOL[np.dtype(dtype)]
    bound = orc.value_bound(abs(m1), abs(b))
    assert cs.rel_err(orc.c_spmm(m1, b), g["y"], bound) <= tol
    y = orc.c_spmm(m1, b, beta=3.0, y=np.ones((200, 100), dtype=dtype))
    assert cs.rel_err(y, g["y_out3"], bound + 3.0) <= tol
    x = np.ascontiguousarray(m1.toarray()[:, :50])
    yt = orc.c_spmm(m1, x, op=orc.OP_T)
    assert cs.rel_err(yt, g["yt"], orc.value_bound(abs(m1.T), abs(x))) <= tol
    # column-major panels give the same numbers
    yf = orc.c_spmm(m1, np.asfortranarray(b))
    assert yf.flags.f_contiguous and cs.rel_err(yf, g["y"], bound) <= tol


def test_spgemm_matches_mkl_golden_bit_exact_structure():
    g = gold("fixture_spgemm.npz")
    m1, m2 = cs.fixture_pair(np.float64)
    c = orc.c_spgemm(m1, m2, sort=True)
    assert c.nnz == 10491 == int(g["raw_nnz"])
    assert np.array_equal(c.indptr, g["indptr"])
    assert np.array_equal(c.indices, g["indices"])
    assert cs.rel_err(c.data, g["data"]) <= 1e-12
    unsorted = orc.canonical(orc.c_spgemm(m1, m2, sort=False))
    assert np.array_equal(unsorted.indices, g["indices"])
    d = orc.c_spmmd(m1, m2)
    assert cs.rel_err(d, gold("fixture_spmmd.npz")["c"], orc.value_bound(abs(m1), abs(m2))) <= 1e-12


def test_structural_zero_convention_matches_mkl():
    g = gold("cancel_spgemm.npz")
    a = sp.csr_matrix(np.array([[1.0, -1.0], [2.0, 0.0]]))
    b = sp.csr_matrix(np.array([[1.0, 3.0], [1.0, 0.0]]))
    c = orc.c_spgemm(a, b, sort=True)
    assert c.nnz == 4 == g["data"].shape[0]
    assert np.array_equal(c.indptr, g["indptr"]) and np.array_equal(c.indices, g["indices"])
    assert np.array_equal(c.data, g["data"])


@pytest.mark.parametrize("aat", [False, True])
def test_gram_matches_mkl_standin_and_numpy(aat):
    g = gold("fixture_gram.npz")
    key = "aat" if aat else "ata"
    m1, _ = cs.fixture_pair(np.float64)
    c = orc.c_syrk(m1, aat=aat, sort=True)
    assert np.array_equal(c.indptr, g[f"{key}_indptr"])
    assert np.array_equal(c.indices, g[f"{key}_indices"])
    assert cs.rel_err(c.data, g[f"{key}_data"]) <= 1e-12
    want = orc.np_gram_upper(m1, aat=aat)
    assert np.abs(c.toarray() - want).max() < 1.5e-6  # the reference's decimal=6
    d = orc.c_syrkd(m1, aat=aat)
    assert np.abs(d - want).max() < 1.5e-6
    assert not d[np.tril_indices(d.shape[0], -1)].any()
    out = np.full(want.shape, 7.0)
    d2 = orc.c_syrkd(m1, aat=aat, beta=2.0, out=out)
    assert np.all(d2[np.tril_indices(d.shape[0], -1)] == 7.0)
    iu = np.triu_indices(d.shape[0])
    assert np.abs(d2[iu] - (want[iu] + 14.0)).max() < 1e-9


def test_config1_matches_mkl_golden():
    g = gold("c1_spmm_f64.npz")
    a = sp.random(10_000, 10_000, density=1e-3, format="csr", dtype=np.float64, random_state=86)
    assert a.nnz == int(g["nnz"]) == 100_000
    x = np.random.default_rng(88).random((10_000, 64))
    y = orc.c_spmm(a, x)
    assert cs.rel_err(y[g["rows"]], g["y_rows"]) <= 1e-12
    assert np.allclose(y.sum(axis=0), g["y_colsum"], rtol=1e-11, atol=0)


def test_config2_small_matches_mkl_golden():
    g = gold("c2_small_spmm_f32.npz")
    a = cs.uniform_rows_csr(20_000, 20_000, 50, np.float32, seed=0)
    assert a.nnz == int(g["nnz"])
    x = np.random.default_rng(2).random((20_000, 128), dtype=np.float32)
    y0 = np.random.default_rng(3).random((20_000, 128), dtype=np.float32)
    y = orc.c_spmm(a, x, beta=0.5, y=y0.copy())
    assert cs.rel_err(y[g["rows"]], g["y_rows"]) <= 1e-5
    assert np.allclose(y.astype(np.float64).sum(axis=0), g["y_colsum"], rtol=1e-5)


@pytest.mark.parametrize("dtype", REAL)
def test_order_transpose_bsr_match_scipy(dtype):
    m1, _ = cs.fixture_pair(dtype)
    rng = np.random.default_rng(0)
    shuffled = m1.copy()
    for i in range(shuffled.shape[0]):
        s, e = shuffled.indptr[i], shuffled.indptr[i + 1]
        p = rng.permutation(e - s)
        shuffled.indices[s:e] = shuffled.indices[s:e][p]
        shuffled.data[s:e] = shuffled.data[s:e][p]
    shuffled.has_sorted_indices = False
    got = orc.c_order(shuffled)
    assert np.array_equal(got.indices, m1.indices) and np.array_equal(got.data, m1.data)
    t = orc.c_transpose(m1)
    want = m1.T.tocsr()
    want.sort_indices()
    assert np.array_equal(t.indptr, want.indptr) and np.array_equal(t.indices, want.indices)
    assert np.array_equal(t.data, want.data)
    bsr = m1.tobsr(blocksize=(10, 10))
    e = orc.c_bsr_to_csr(bsr)
    assert e.nnz == bsr.indices.shape[0] * 100
    assert np.array_equal(e.toarray(), m1.toarray())


@pytest.mark.skipif(not mkl_ref.available(), reason="embedded oneMKL not reachable")
@pytest.mark.parametrize("dtype", REAL)
def test_live_mkl_agrees_with_oracle(dtype):
    mkl_ref.set_threads(1)
    a = cs.uniform_rows_csr(4000, 3000, 20, dtype, seed=9)
    x = np.random.default_rng(1).random((3000, 48)).astype(dtype)
    tol = cs.TOL[np.dtype(dtype)]
    assert cs.rel_err(orc.c_spmm(a, x), mkl_ref.spmm(a, x)) <= tol
    b = cs.uniform_rows_csr(3000, 2500, 6, dtype, seed=10)
    cm = mkl_ref.spgemm(a, b)
    cm.has_sorted_indices = False
    cm.sort_indices()
    co = orc.c_spgemm(a, b, sort=True)
    assert np.array_equal(co.indptr, cm.indptr) and np.array_equal(co.indices, cm.indices)
    assert cs.rel_err(co.data, cm.data) <= tol
