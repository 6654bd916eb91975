"""Seeded inputs shared by the tests: the reference's own fixtures
(sparse_dot_mkl/tests/test_mkl.py:27-50) and scaled-down BASELINE configs."""
import numpy as np
import scipy.sparse as sp

SEED = 86  # test_mkl.py:27


def make_matrixes(a, b, n, density, dtype=np.float64):
    """test_mkl.py:30-37: (a x n) and (n x b) CSR matrices."""
    m1 = sp.random(a, n, density=density, format="csr", dtype=dtype, random_state=SEED)
    m2 = sp.random(n, b, density=density, format="csr", dtype=dtype, random_state=SEED + 1)
    return m1, m2


def make_vector(n, complex=False):
    rng = np.random.default_rng(SEED + 2)
    if not complex:
        return rng.random(n).astype(np.float64)
    return rng.random(n).astype(np.float64) + rng.random(n).astype(np.float64) * 1j


def fixture_pair(dtype=np.float64):
    """MATRIX_1, MATRIX_2 of test_mkl.py:48 (200x300 nnz 3000, 300x100 nnz 1500)."""
    m1, m2 = make_matrixes(200, 100, 300, 0.05)
    return m1.astype(dtype), m2.astype(dtype)


def complexify(m, seed):
    rng = np.random.default_rng(seed)
    out = m.astype(np.complex128)
    out.data = out.data + 1j * rng.random(out.data.shape[0])
    return out


def uniform_rows_csr(m, k, per_row, dtype, seed):
    """BASELINE C2/C5 recipe: exactly `per_row` distinct sorted columns per row,
    values in [0.5, 1.5) (no cancellation)."""
    rng = np.random.default_rng(seed)
    cols = np.empty((m, per_row), dtype=np.int64)
    # distinct columns per row: sample with a stride trick, then sort
    base = rng.integers(0, k, size=(m, per_row), dtype=np.int64)
    base.sort(axis=1)
    # push duplicates apart (wrap-around keeps them in range); cheap and deterministic
    for _ in range(4):
        dup = np.zeros_like(base, dtype=bool)
        dup[:, 1:] = base[:, 1:] == base[:, :-1]
        if not dup.any():
            break
        base[dup] = rng.integers(0, k, size=int(dup.sum()), dtype=np.int64)
        base.sort(axis=1)
    cols[:] = base
    indptr = np.arange(0, m * per_row + 1, per_row, dtype=np.int64)
    data = (rng.random(m * per_row) + 0.5).astype(dtype)
    a = sp.csr_matrix((data, cols.ravel().astype(np.int32), indptr.astype(np.int32 if m * per_row < 2**31 else np.int64)),
                      shape=(m, k))
    a.sum_duplicates()
    return a


def c2_workload(rows, cols, per_row, n_dense, seed, dtype=np.float32):
    """The BASELINE configs[1] / configs[4] SpMM workload: A = uniform_rows_csr, X and Y0 uniform
    random panels.  The ONE generator of this recipe: bench.py, scripts/run_configs.py and the tests use it."""
    a = uniform_rows_csr(rows, cols, per_row, dtype, seed)
    x = np.random.default_rng(seed + 2).random((cols, n_dense), dtype=np.float32).astype(dtype, copy=False)
    y = np.random.default_rng(seed + 3).random((rows, n_dense), dtype=np.float32).astype(dtype, copy=False)
    return a, x, y


def rmat_csr(scale, edge_factor, dtype, seed, abcd=(0.57, 0.19, 0.19, 0.05)):
    """Graph500 R-MAT (BASELINE C3 recipe): duplicates summed, values U[0.5,1.5)."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    ne = edge_factor * n
    a, b, c, _ = abcd
    rows = np.zeros(ne, dtype=np.int64)
    cols = np.zeros(ne, dtype=np.int64)
    for bit in range(scale):
        r = rng.random(ne)
        right = (r >= a) & (r < a + b) | (r >= a + b + c)
        down = r >= a + b
        rows |= down.astype(np.int64) << bit
        cols |= right.astype(np.int64) << bit
    vals = (rng.random(ne) + 0.5).astype(dtype)
    m = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    m.sum_duplicates()
    m.sort_indices()
    return m


def rel_err(got, want, bound=None):
    """max |got - want| / bound, bound defaulting to max(|want|, tiny)."""
    got = np.asarray(got)
    want = np.asarray(want)
    if bound is None:
        bound = np.maximum(np.abs(want), np.finfo(np.float64).tiny)
    bound = np.asarray(bound)
    ok = bound > 0
    if not ok.any():
        return float(np.abs(got - want).max()) if got.size else 0.0
    err = np.zeros(got.shape, dtype=np.float64)
    err[ok] = np.abs(got - want)[ok] / bound[ok]
    assert np.all(got[~ok] == want[~ok]), "entries with a zero error bound must match exactly"
    return float(err.max()) if err.size else 0.0


TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12,
       np.dtype(np.complex64): 1e-5, np.dtype(np.complex128): 1e-12}
