"""
The L2-tiled SpMM kernel (csrc/spmm_slab.cu: slab-ordered inspector + streaming executor) is chosen
automatically only for panels several times the L2 cache on a handle that is multiplied repeatedly; here
it is forced (SDB_SLAB=2, tiny slabs via SDB_SLAB_MB) in a subprocess — the switches are read once per
process — and checked against the CPU oracle on small inputs: many slabs, empty rows, rows denser than a
chunk, ragged last row block, beta != 0, fp32 and fp64, panels of 1, 2 and 3 column chunks, every executor
shape (SDB_SLAB_VARIANT), and the two-rank fused all-gather epilogue.
"""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, extra_env=None):
    env = dict(os.environ, SDB_SLAB="2", SDB_SLAB_MB="1", PYTHONPATH=ROOT)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], env=env, cwd=ROOT, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
def test_slab_kernel_matches_oracle(variant):
    out = _run("""
        import numpy as np, scipy.sparse as sp
        import sparse_dot_b200 as sdb
        from sparse_dot_b200 import _handles as H, _lib, sharded
        from oracle import oracle as orc
        from tests import _cases as cs
        before = sdb.kernel_launches()
        for dtype, n in ((np.float32, 128), (np.float64, 64), (np.float32, 256), (np.float64, 192),
                         (np.float32, 64), (np.float64, 32)):  # the last two: 256-byte rows, two groups per warp
            a = cs.uniform_rows_csr(5000, 20000, 30, dtype, seed=1).tolil()
            a[7, :] = 0
            a[4999, :] = 0
            a = a.tocsr()
            dense = sp.random(3, 20000, density=0.4, format="csr", dtype=dtype, random_state=5)
            dense.data[:] = np.random.default_rng(6).random(dense.nnz) + 0.5
            a = sp.vstack([a, dense, cs.uniform_rows_csr(77, 20000, 3, dtype, seed=2)]).tocsr()
            a.sort_indices()
            x = np.random.default_rng(2).random((20000, n)).astype(dtype)
            y0 = np.random.default_rng(3).random((a.shape[0], n)).astype(dtype)
            tol = 1e-5 if dtype == np.float32 else 1e-12
            bound = orc.value_bound(abs(a), abs(x))
            # through the resident-operand plan (sdb_spmm_dev): beta = 0 and beta != 0
            with sharded.RowShardedSpMM(a, n) as plan:
                plan.set_x(x); plan.set_local_y(y0)
                plan.run(alpha=1.0, beta=0.0); plan.synchronize()
                got = plan.read_panel()
                assert cs.rel_err(got, orc.c_spmm(a, x), bound) <= tol, "beta=0"
                plan.set_local_y(y0)
                plan.run(alpha=2.0, beta=0.5); plan.synchronize()
                got = plan.read_panel()
                want = orc.c_spmm(a, x, alpha=2.0, beta=0.5, y=y0.copy())
                assert cs.rel_err(got, want, 2 * bound + 0.5 * y0) <= tol, "beta=0.5"
                want_kernel = "spmm_stream_half_kernel" if n * np.dtype(dtype).itemsize == 256 else "spmm_stream_kernel"
                assert sdb.last_spmm_kernel().startswith(want_kernel), sdb.last_spmm_kernel()
            # through the public API (fresh handle per call)
            got = sdb.dot_product_mkl(a, x, out=y0.copy(), out_scalar=0.5)
            want = orc.c_spmm(a, x, beta=0.5, y=y0.copy())
            assert cs.rel_err(got, want, bound + 0.5 * y0) <= tol, "api"
            # unsorted rows must fall back to the row-gather kernel and still be right
            b = a.copy()
            s0, e0 = b.indptr[0], b.indptr[1]
            b.indices[s0:e0] = b.indices[s0:e0][::-1].copy(); b.data[s0:e0] = b.data[s0:e0][::-1].copy()
            b.has_sorted_indices = False
            got = sdb.dot_product_mkl(b, x)
            assert cs.rel_err(got, orc.c_spmm(a, x), bound) <= tol, "unsorted fallback"
            assert sdb.last_spmm_kernel().startswith("spmm_rowmajor_kernel"), sdb.last_spmm_kernel()
        print("OK", sdb.kernel_launches() - before)
    """, extra_env={"SDB_SLAB_VARIANT": str(variant)})
    assert "OK" in out


def test_slab_kernel_with_evict_last_gathers():
    """Option slab_keep: the same executor with an L2 evict_last policy on its gathers (the variant the fused
    all-gather may pick); results must be the same as the plain one."""
    out = _run("""
        import numpy as np
        import sparse_dot_b200 as sdb
        from sparse_dot_b200 import sharded
        from oracle import oracle as orc
        from tests import _cases as cs
        a = cs.uniform_rows_csr(6000, 20000, 30, np.float32, seed=1)
        x = np.random.default_rng(2).random((20000, 128), dtype=np.float32)
        y0 = np.random.default_rng(3).random((6000, 128), dtype=np.float32)
        with sharded.RowShardedSpMM(a, 128) as plan:
            plan.set_x(x); plan.set_local_y(y0)
            plan.run(alpha=1.0, beta=0.5); plan.synchronize()
            got = plan.read_panel()
            assert sdb.last_spmm_kernel().endswith("keep>"), sdb.last_spmm_kernel()
        assert cs.rel_err(got, orc.c_spmm(a, x, beta=0.5, y=y0.copy())) <= 1e-5
        print("OK")
    """, extra_env={"SDB_SLAB_KEEP": "1"})
    assert "OK" in out


def test_slab_kernel_two_rank_fused_allgather():
    out = _run("""
        import subprocess, sys
        r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_sharded.py", "-q", "-m", "gpu", "-x"],
                           capture_output=True, text=True)
        print(r.stdout[-1500:], r.stderr[-1500:])
        assert r.returncode == 0
        print("OK")
    """)
    assert "OK" in out
