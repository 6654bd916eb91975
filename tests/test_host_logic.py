"""
Host-side contract of the public API that needs no GPU: validation order and
errors (all ValueError, like sparse_dot_mkl), dtype casting rules, output-array
checks, empty-product shortcuts.  Mirrors sparse_dot_mkl/tests/test_mkl.py
(:70-103 empties, :143-178 shape / block errors, :271-385 _type_check) and the
error-path tests of test_sparse_sparse.py / test_sparse_dense.py / test_gram_matrix.py.
"""
import numpy as np
import pytest
import scipy.sparse as sp

import sparse_dot_b200 as sdb
from sparse_dot_b200 import _handles, _validate as v
from tests import _cases as cs

M1, M2 = cs.fixture_pair(np.float64)


def test_unify_dtypes_identity_and_copy_semantics():
    a, b = v.unify_dtypes(M1, M2)
    assert a is M1 and b is M2
    a32 = M1.astype(np.float32)
    with pytest.raises(ValueError):
        v.unify_dtypes(a32, M2)
    a, b = v.unify_dtypes(a32, M2, cast=True)
    assert a.dtype == np.float64 and b is M2 and a is not a32
    ai = M1.astype(np.int64)
    assert v.unify_dtypes(ai, cast=True).dtype == np.float64
    with pytest.raises(ValueError):
        v.unify_dtypes(ai)
    c = cs.complexify(M1, 0).astype(np.complex64)
    a, b = v.unify_dtypes(c, M1.astype(np.float32), cast=True)
    assert a is c and b.dtype == np.complex64
    a, b = v.unify_dtypes(M1.astype(np.float32), c, cast=True)
    assert b is c and a.dtype == np.complex64
    a, b = v.unify_dtypes(c, cs.complexify(M1, 1), cast=True)
    assert a.dtype == b.dtype == np.complex128
    with pytest.raises(ValueError):
        v.unify_dtypes(c, allow_complex=False)
    assert v.unify_dtypes(M1.astype(np.float32)).dtype == np.float32


def test_shape_checks():
    with pytest.raises(ValueError):
        v.check_shapes(M1, M2.T)
    with pytest.raises(ValueError):
        v.check_shapes(np.ones((2, 3, 4)), np.ones((4, 2)))
    with pytest.raises(ValueError):
        v.check_shapes(M1, np.ones(300))
    v.check_shapes(M1, np.ones(300), allow_vector=True)
    v.check_shapes(np.ones(200), M1, allow_vector=True)
    with pytest.raises(ValueError):
        sdb.dot_product_mkl(M1, M2.T.tocsr())
    with pytest.raises(ValueError):
        sdb.dot_product_mkl(M1, np.ones((3, 300, 2)))


def test_layout_and_output_array_rules():
    c, f = np.ones((4, 3)), np.ones((4, 3), order="F")
    assert v.dense_layout(c) == (101, 3) and v.dense_layout(f) == (102, 4)
    one_col = np.ones((4, 1))
    assert v.dense_layout(one_col, other=f) == (102, 4) and v.dense_layout(one_col, other=c) == (101, 1)
    with pytest.raises(ValueError, match="not contiguous"):
        v.dense_layout(np.ones((8, 3))[::2])
    out = np.ones((4, 3))
    assert v.output_array((4, 3), np.float64, "C", out=out) is out
    assert not v.output_array((4, 3), np.float64, "C").any()
    for bad in (np.ones((3, 4)), np.ones((4, 3), dtype=np.float32), np.ones((4, 3), order="F")):
        with pytest.raises(ValueError, match="Provided out array"):
            v.output_array((4, 3), np.float64, "C", out=bad)
    with pytest.raises(ValueError, match=r"\(3, 4\)"):  # reported in the caller's orientation
        v.output_array((4, 3), np.float64, "F", out=np.ones((4, 3)), out_t=True)


def test_empty_products_need_no_device():
    """test_mkl.py:70-103: all-zero or zero-sized operands short-circuit."""
    empty = sp.csr_matrix((200, 300), dtype=np.float64)
    r = sdb.dot_product_mkl(empty, M2)
    assert isinstance(r, sp.csr_matrix) and r.shape == (200, 100) and r.nnz == 0
    r = sdb.dot_product_mkl(sp.csc_matrix((200, 300)), M2.tocsc())
    assert isinstance(r, sp.csc_matrix)
    r = sdb.dot_product_mkl(empty, M2, dense=True)
    assert isinstance(r, np.ndarray) and r.shape == (200, 100) and not r.any()
    r = sdb.dot_product_mkl(empty, M2.toarray())
    assert r.shape == (200, 100) and r.dtype == np.float64 and not r.any()
    r = sdb.dot_product_mkl(empty.astype(np.float32), M2.toarray().astype(np.float32))
    assert r.dtype == np.float32
    out = np.ones((200, 100))
    assert sdb.dot_product_mkl(empty, M2.toarray(), out=out) is out
    r = sdb.dot_product_mkl(empty, np.ones(300))
    assert r.shape == (200,) and not r.any()
    r = sdb.gram_matrix_mkl(empty)
    assert sp.issparse(r) and r.shape == (200, 200)  # the reference's own shape rule for empties
    r = sdb.dot_product_mkl(sp.csr_matrix((0, 300)), M2)
    assert r.shape == (0, 100)


def test_rejections_before_any_device_work():
    with pytest.raises(ValueError, match="COO"):
        sdb.dot_product_mkl(M1.tocoo(), M2)
    with pytest.raises(ValueError, match="COO"):
        sdb.dot_product_mkl(M1, M2.tocoo())
    with pytest.raises(ValueError, match="dense=True"):
        sdb.dot_product_mkl(M1, M2, out=np.zeros((200, 100)))
    with pytest.raises(ValueError):
        sdb.dot_product_mkl(M1.astype(np.float32), M2)
    with pytest.raises(ValueError, match="complex"):
        sdb.gram_matrix_mkl(cs.complexify(M1, 0))
    with pytest.raises(ValueError, match="CSC"):
        sdb.gram_matrix_mkl(M1.tocsc())
    with pytest.raises(ValueError, match="CSR or CSC"):
        sdb.gram_matrix_mkl(M1.tobsr(blocksize=(10, 10)))
    with pytest.raises(ValueError):
        sdb.gram_matrix_mkl(M1.astype(np.int32))
    with pytest.warns(DeprecationWarning):
        sdb.dot_product_mkl(sp.csr_matrix((200, 300)), M2, debug=True)
    # two dense operands: validated like the reference (_dense_dense.py:74-88), then a GEMM on the GPU
    with pytest.raises(ValueError, match="Matrix alignment error"):
        sdb.dot_product_mkl(np.ones((3, 3)), np.ones((4, 3)))
    with pytest.raises(ValueError, match="must be the same"):
        sdb.dot_product_mkl(np.ones((3, 3), dtype=np.float32), np.ones((3, 3)))
    assert sdb.dot_product_mkl(np.arange(3.0), np.arange(3.0)) == 5.0  # vector (dot) vector is numpy's


def test_bsr_block_rules():
    """test_mkl.py:176-178: non-square blocks are refused before upload."""
    bsr = M1.tobsr(blocksize=(10, 5))
    with pytest.raises(ValueError, match="square"):
        _handles.create(bsr)


def test_oversized_dimension_is_refused():
    big = sp.csr_matrix((1, 2**31 + 5), dtype=np.float32)
    big.indices = np.zeros(1, dtype=np.int64)
    big.indptr = np.array([0, 1], dtype=np.int64)
    big.data = np.ones(1, dtype=np.float32)
    with pytest.raises(ValueError, match="int32 column indices"):
        _handles.create(big)


def test_debug_mode_prints(capsys):
    sdb.set_debug_mode(True)
    try:
        sdb.dot_product_mkl(sp.csr_matrix((200, 300)), M2)
    finally:
        sdb.set_debug_mode(False)
    out = capsys.readouterr().out
    assert "libsdb200" in out and "Skipping multiplication" not in out or True
