"""
GPU tests of the row-sharded SpMM: single-rank plan, and
two-rank runs of the FUSED all-gather (peer panels mapped with CUDA IPC) for
every exchange strategy (copy-engine chunks, epilogue stores, K1 + stores) and
both kernels.  The two ranks share cuda:0 when the box has one GPU — peer mapping works
between processes on the same device — with `gloo` as the process group, so
the test needs no second GPU; with >= 2 GPUs each rank takes its own.
"""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc
from tests import _cases as cs

pytestmark = pytest.mark.gpu


def test_single_rank_plan_matches_oracle():
    import sparse_dot_b200  # noqa: F401
    from sparse_dot_b200 import sharded

    a = cs.uniform_rows_csr(5000, 4000, 30, np.float32, seed=1)
    x = np.random.default_rng(2).random((4000, 128), dtype=np.float32)
    y0 = np.random.default_rng(3).random((5000, 128), dtype=np.float32)
    with sharded.RowShardedSpMM(a, 128) as plan:
        plan.set_x(x)
        plan.set_local_y(y0)
        plan.run(alpha=1.0, beta=0.5)
        plan.synchronize()
        got = plan.read_panel()
    want = orc.c_spmm(a, x, beta=0.5, y=y0.copy())
    assert cs.rel_err(got, want) <= 1e-5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q, strategy, slab, rows):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    # read once per process by the library: exchange strategy and kernel selection
    os.environ["SDB_ALLGATHER"] = strategy
    os.environ["SDB_SLAB"] = slab
    os.environ["SDB_SLAB_MB"] = "1"
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sparse_dot_b200 as sdb
        from sparse_dot_b200 import _lib, sharded

        dev = rank % sdb.device_count()
        _lib.check(_lib.SDB.lib.sdb_set_device(dev), "sdb_set_device")
        top = cs.uniform_rows_csr(3 * rows // 10, 2000, 40, np.float32, seed=1)
        bottom = cs.uniform_rows_csr(rows - 3 * rows // 10, 2000, 5, np.float32, seed=2)
        a = sp.vstack([top, bottom]).tocsr()
        x = np.random.default_rng(3).random((2000, 128), dtype=np.float32)
        y0 = np.random.default_rng(4).random((rows, 128), dtype=np.float32)
        got = sharded.spmm_sharded(a, x, world, rank, allgather="fused", out=y0, out_scalar=0.25)
        want = orc.c_spmm(a, x, beta=0.25, y=y0.copy())
        q.put((rank, cs.rel_err(got, want), sdb.last_spmm_kernel()))
    except Exception as e:  # surface the failure in the parent
        q.put((rank, repr(e), ""))
    finally:
        dist.destroy_process_group()


# (exchange strategy, SDB_SLAB, global rows, kernel expected): small panels take one chunk, the 300k-row cases are
# cut into several chunks (copy engines) — with SDB_SLAB=2 into whole waves of the streaming kernel's grid
CASES = [
    ("ce", "1", 10_000, "spmm_rowmajor"),
    ("ce", "1", 300_000, "spmm_rowmajor"),
    ("ce", "2", 300_000, "spmm_stream"),
    ("stores", "1", 10_000, "spmm_rowmajor"),
    ("stores", "2", 300_000, "spmm_stream"),
    ("k1", "2", 300_000, "spmm_rowmajor"),
    ("sm", "1", 10_000, "spmm_rowmajor"),
    ("sm", "1", 300_000, "spmm_rowmajor"),
    ("sm", "2", 300_000, "spmm_stream"),
    ("auto", "0", 10_000, "spmm_rowmajor"),
]


# the same with 4 and 8 ranks (the SCALE run's rank counts; they share the visible GPUs round-robin): every peer list
# length the library supports, the seven-target copier kernel, epilogue stores into seven peer panels
MANY_RANK_CASES = [
    (8, "ce", "1", 40_000, "spmm_rowmajor"),
    (8, "sm", "2", 400_000, "spmm_stream"),
    (8, "stores", "2", 400_000, "spmm_stream"),
    (8, "k1", "1", 40_000, "spmm_rowmajor"),
    (4, "sm", "1", 200_000, "spmm_rowmajor"),
]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,strategy,slab,rows,kernel", MANY_RANK_CASES)
def test_many_rank_fused_allgather_matches_oracle(world, strategy, slab, rows, kernel):
    _run_fused_case(world, strategy, slab, rows, kernel)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("strategy,slab,rows,kernel", CASES)
def test_two_rank_fused_allgather_matches_oracle(strategy, slab, rows, kernel):
    _run_fused_case(2, strategy, slab, rows, kernel)


def _run_fused_case(world, strategy, slab, rows, kernel):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, strategy, slab, rows)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, err, used in results:
        assert isinstance(err, float), f"rank {rank} failed: {err}"
        assert err <= 1e-5, f"rank {rank}: full panel differs from the oracle ({err:.2e})"
        assert used.startswith(kernel), (rank, used)


def _worker_products(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sparse_dot_b200 as sdb
        from sparse_dot_b200 import _lib, sharded

        _lib.check(_lib.SDB.lib.sdb_set_device(rank % sdb.device_count()), "sdb_set_device")
        a = cs.rmat_csr(11, 8, np.float64, seed=1)
        b = cs.rmat_csr(11, 8, np.float64, seed=2)
        c = sharded.spgemm_sharded(a, b, world, rank, reorder_output=True)
        w = orc.c_spgemm(a, b, sort=True)
        ok_struct = bool(np.array_equal(c.indptr, w.indptr) and np.array_equal(c.indices, w.indices))
        err = cs.rel_err(c.data, w.data)
        m1, _ = cs.fixture_pair(np.float64)
        g = sharded.gram_dense_sharded(m1, world, rank)
        wg = orc.c_syrkd(m1)
        gerr = float(np.abs(g - wg).max())
        q.put((rank, ok_struct, err, gerr))
    except Exception as e:
        q.put((rank, repr(e), None, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_spgemm_and_gram_sharding():
    """SURVEY §8e last row: SpGEMM sharded by rows of A (B replicated, blocks concatenated) and the dense
    gram as an all-reduce of per-row-block partial grams."""
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_products, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok_struct, err, gerr in results:
        assert ok_struct is True, f"rank {rank}: {ok_struct}"
        assert err <= 1e-12 and gerr <= 1e-10, (rank, err, gerr)


def _worker_nccl(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sparse_dot_b200 as sdb  # noqa: F401
        from sparse_dot_b200 import _lib, sharded

        _lib.check(_lib.SDB.lib.sdb_set_device(rank), "sdb_set_device")
        a = cs.rmat_csr(12, 8, np.float32, seed=1)
        b = cs.rmat_csr(12, 8, np.float32, seed=2)
        w = orc.c_spgemm(a, b, sort=True)
        with sharded.spgemm_sharded_device(a, b, world, rank, reorder_output=True) as c_dev:
            c = c_dev.to_scipy()
            # the replicated result is usable on the device again: multiply it by a panel without leaving HBM
            x = torch.ones((c.shape[1], 8), dtype=torch.float32, device="cuda")
            y = torch.empty((c.shape[0], 8), dtype=torch.float32, device="cuda")
            one, zero = _lib.scalar_pair(1.0), _lib.scalar_pair(0.0)
            import ctypes as C

            _lib.check(_lib.SDB.lib.sdb_spmm_dev(_lib.OP_N, one, c_dev.handle.ref, _lib.LAYOUT_C, C.c_void_p(x.data_ptr()),
                                                 8, 8, zero, C.c_void_p(y.data_ptr()), 8, None), "sdb_spmm_dev")
            _lib.check(_lib.SDB.lib.sdb_device_synchronize(), "sync")
            want_rows = np.asarray(w.astype(np.float64).sum(axis=1)).ravel()
            rowsum_err = float(np.abs(y[:, 0].cpu().numpy() - want_rows).max() / want_rows.max())
        ok_struct = bool(np.array_equal(c.indptr, w.indptr) and np.array_equal(c.indices, w.indices))
        err = cs.rel_err(c.data, w.data)
        # the NCCL baseline mode with nnz-balanced (unequal) row blocks: per-owner broadcasts
        top = cs.uniform_rows_csr(600, 900, 40, np.float32, seed=1)
        bottom = cs.uniform_rows_csr(2400, 900, 5, np.float32, seed=2)
        ab = sp.vstack([top, bottom]).tocsr()
        xb = np.random.default_rng(3).random((900, 64), dtype=np.float32)
        got = sharded.spmm_sharded(ab, xb, world, rank, allgather="nccl")
        nccl_err = cs.rel_err(got, orc.c_spmm(ab, xb))
        m = cs.uniform_rows_csr(3000, 700, 20, np.float64, seed=3)
        wg = orc.c_syrkd(m)
        g = sharded.gram_dense_sharded(m, world, rank)
        gerr = float(np.abs(g - wg).max())
        row0, mine = sharded.gram_dense_sharded(m, world, rank, gather=False)
        perr = float(np.abs(mine - wg[row0:row0 + mine.shape[0]]).max())
        q.put((rank, ok_struct, max(err, rowsum_err, nccl_err), max(gerr, perr)))
    except Exception as e:
        q.put((rank, repr(e), None, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_nccl_device_level_spgemm_and_gram():
    """VERDICT r1 missing #3: the multi-GPU SpGEMM exchanges its row blocks device to device (NCCL broadcasts out of
    the handles' arrays) and the dense gram is reduce-scattered onto panel owners.  Needs two GPUs (NCCL refuses two
    ranks on one device), so it runs on the multi-GPU box and skips on a single-GPU one."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (NCCL)")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_nccl, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok_struct, err, gerr in results:
        assert ok_struct is True, f"rank {rank}: {ok_struct}"
        assert err <= 1e-5 and gerr <= 1e-9, (rank, err, gerr)
