"""
Host-side logic of the multi-GPU path on CPU: two `gloo` processes partition the
rows (sdb_partition_rows), cut their blocks (row_block), agree on the panel
layout (ShardLayout, the same all-gather RowShardedSpMM does) and assemble the
product; the per-block arithmetic is done by the CPU oracle here because there
is no GPU — on the GPU box tests/test_gpu_sharded.py runs the real thing.
"""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        from sparse_dot_b200 import sharded
        from tests import _cases as cs

        # skewed rows: nnz-balanced bounds differ from row-balanced ones
        top = cs.uniform_rows_csr(300, 500, 40, np.float64, seed=1)
        bottom = cs.uniform_rows_csr(900, 500, 5, np.float64, seed=2)
        a = sp.vstack([top, bottom]).tocsr()
        x = np.random.default_rng(3).random((500, 16))
        bounds = sharded.partition_rows(a.indptr, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        blk = sharded.row_block(a, lo, hi)
        assert blk.indptr[0] == 0 and blk.nnz == a.indptr[hi] - a.indptr[lo]
        layout = sharded.ShardLayout(blk.shape[0], world, rank)
        assert layout.row0 == lo and layout.rows_total == a.shape[0]
        assert layout.block(rank) == (lo, hi - lo)
        y_local = orc.c_spmm(blk, x)
        parts = [None] * world
        dist.all_gather_object(parts, (layout.row0, y_local))
        panel = np.empty((layout.rows_total, 16))
        for r0, y in parts:
            panel[r0:r0 + y.shape[0]] = y
        want = orc.c_spmm(a, x)
        nnz_per = [int(a.indptr[bounds[i + 1]] - a.indptr[bounds[i]]) for i in range(world)]
        q.put((rank, bool(np.array_equal(panel, want)), nnz_per, list(map(int, bounds))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_row_sharding_assembles_the_full_product():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, ok, nnz_per, bounds in results:
        assert ok, f"rank {rank}: assembled panel differs from the single-process product"
        assert bounds[0] == 0 and bounds[-1] == 1200
        assert abs(nnz_per[0] - nnz_per[1]) <= 80  # balanced by nnz (within one 40-entry row each side), not by rows
        assert bounds[1] < 600


def test_stack_row_blocks_is_exact():
    """The host-side assembly of a row-sharded sparse product (spgemm_sharded) reproduces the arrays of the
    unsharded matrix exactly, including empty blocks."""
    from sparse_dot_b200 import sharded
    from tests import _cases as cs

    c = cs.rmat_csr(10, 6, np.float64, seed=3)
    bounds = [0, 100, 100, 700, c.shape[0]]
    blocks = [sharded.row_block(c, bounds[i], bounds[i + 1]) for i in range(4)]
    got = sharded.stack_row_blocks(blocks, c.shape[1])
    assert got.shape == c.shape
    assert np.array_equal(got.indptr, c.indptr) and np.array_equal(got.indices, c.indices)
    assert np.array_equal(got.data, c.data)


def test_triangle_row_bounds_balance_the_area():
    """Owners of the reduce-scattered gram panels (sharded.gram_dense_sharded): equal upper-triangle area."""
    from sparse_dot_b200 import sharded

    n, parts = 100_000, 8
    b = sharded.triangle_row_bounds(n, parts)
    assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
    area = [(b[i + 1] - b[i]) * n - (b[i + 1] * (b[i + 1] - 1) - b[i] * (b[i] - 1)) // 2 for i in range(parts)]
    assert max(area) / min(area) < 1.001
    assert sharded.triangle_row_bounds(3, 8)[-1] == 3 and sharded.triangle_row_bounds(0, 4) == [0, 0, 0, 0, 0]
