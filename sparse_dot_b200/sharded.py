"""
Row-sharded SpMM across the GPUs of one node (BASELINE configs[4]).

One process per GPU.  Rank r owns a contiguous block of rows of A (and of Y);
X is replicated; Y = A @ X needs no reduction, only an all-gather of the dense
output panel so that every GPU ends up with all of Y.  Three exchange modes:

  fused  one library call per step (sdb_spmm_dev_allgather) that leaves this
         rank's rows in EVERY rank's full panel through NVLink peer mappings
         (CUDA IPC tokens exchanged over torch.distributed).  Default strategy:
         the shard is cut into a few row chunks and the copy engines push chunk c
         to the peers while chunk c + 1's kernel runs; SDB_ALLGATHER=stores makes
         the kernel's epilogue store into the peer panels itself.
  nccl   kernel into the local block, then ncclAllGather (all_gather_into_tensor; per-owner
         broadcasts when the blocks differ in row count) — the baseline the fused kernel is
         compared with.
  none   no exchange (independent shards).

torch is used here for the process group only (rendezvous, token exchange, the
NCCL baseline); memory and kernels are libsdb200's.  The reference has no
multi-device path at all — this is new surface.
"""
import ctypes as _ct

import numpy as np

from . import _handles as _h
from . import _lib
from ._lib import SDB, check, scalar_pair


def partition_rows(indptr, parts):
    """nnz-balanced contiguous row blocks: bounds[0]=0 <= ... <= bounds[parts]=rows
    (host code in the library: sdb_partition_rows)."""
    indptr = np.ascontiguousarray(indptr)
    if indptr.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
        indptr = indptr.astype(np.int64)
    bounds = (_ct.c_int64 * (parts + 1))()
    check(
        SDB.lib.sdb_partition_rows(indptr.ctypes.data_as(_ct.c_void_p), indptr.itemsize * 8,
                                   indptr.shape[0] - 1, parts, bounds),
        "sdb_partition_rows",
    )
    return np.array(bounds[:], dtype=np.int64)


def row_block(a_csr, lo, hi):
    """Rows [lo, hi) of a CSR matrix as a CSR matrix with rebased offsets (views
    of the index/value arrays, a fresh indptr)."""
    import scipy.sparse as sps

    s, e = int(a_csr.indptr[lo]), int(a_csr.indptr[hi])
    indptr = (a_csr.indptr[lo:hi + 1] - a_csr.indptr[lo]).astype(a_csr.indptr.dtype, copy=False)
    return sps.csr_matrix((a_csr.data[s:e], a_csr.indices[s:e], indptr), shape=(hi - lo, a_csr.shape[1]))


def _all_gather_obj(obj, world, group):
    if world == 1:
        return [obj]
    import torch.distributed as dist

    out = [None] * world
    dist.all_gather_object(out, obj, group=group)
    return out


class ShardLayout:
    """Where this rank's row block sits in the global output panel (pure host
    logic; one all-gather of the local row counts)."""

    def __init__(self, rows_local, world_size=1, rank=0, group=None):
        self.world, self.rank, self.group = int(world_size), int(rank), group
        self.rows_local = int(rows_local)
        self.row_counts = [int(c) for c in _all_gather_obj(self.rows_local, self.world, group)]
        self.row0 = int(sum(self.row_counts[: self.rank]))
        self.rows_total = int(sum(self.row_counts))
        self.uniform = len(set(self.row_counts)) == 1

    def block(self, q):
        """(first row, row count) of rank q's block."""
        return int(sum(self.row_counts[:q])), self.row_counts[q]


class _DevBuf:
    def __init__(self, nbytes):
        self.ptr = _ct.c_void_p()
        self.nbytes = int(nbytes)
        check(SDB.lib.sdb_dev_alloc(_ct.byref(self.ptr), self.nbytes), "sdb_dev_alloc")

    def free(self):
        if self.ptr:
            check(SDB.lib.sdb_dev_free(self.ptr), "sdb_dev_free")
            self.ptr = _ct.c_void_p()

    def at(self, byte_offset):
        return _ct.c_void_p(self.ptr.value + int(byte_offset))


class _CudaView:
    """__cuda_array_interface__ over a raw device pointer (so torch can wrap the
    panel for the NCCL baseline without copying)."""

    def __init__(self, ptr, shape, dtype):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": np.dtype(dtype).str, "data": (int(ptr), False), "version": 3,
            "strides": None,
        }


class RowShardedSpMM:
    """Y_panel[row0 : row0 + rows_local, :] = alpha * A_local @ X + beta * (same rows),
    with the panel replicated on every rank according to ``allgather``."""

    def __init__(self, a_local, n_dense, world_size=1, rank=0, allgather="fused", group=None):
        if allgather not in ("fused", "nccl", "none"):
            raise ValueError("allgather must be 'fused', 'nccl' or 'none'")
        self.world, self.rank, self.group = int(world_size), int(rank), group
        self.mode = allgather if self.world > 1 else "none"
        self.n = int(n_dense)
        self.dtype = np.dtype(a_local.dtype)
        self.es = self.dtype.itemsize
        self.rows_local, self.cols = a_local.shape
        self.handle, _, _ = _h.create(a_local)
        self._peers_opened = []
        # every rank's block position inside the global panel
        self.layout = ShardLayout(self.rows_local, self.world, self.rank, group)
        self.row_counts = self.layout.row_counts
        self.row0 = self.layout.row0
        self.rows_total = self.layout.rows_total
        panel_rows = self.rows_total if self.mode != "none" else self.rows_local
        self.panel_row0 = self.row0 if self.mode != "none" else 0
        self.x = _DevBuf(self.cols * self.n * self.es)
        self.panel = _DevBuf(panel_rows * self.n * self.es)
        self.panel_rows = panel_rows
        self.peer_ptrs = [self.panel.ptr.value]
        if self.mode == "fused":
            token = _ct.create_string_buffer(64)
            check(SDB.lib.sdb_ipc_export(self.panel.ptr, token), "sdb_ipc_export")
            tokens = self._all_gather_obj(token.raw)
            self.peer_ptrs = []
            for q, tok in enumerate(tokens):
                if q == self.rank:
                    self.peer_ptrs.append(self.panel.ptr.value)
                    continue
                p = _ct.c_void_p()
                check(SDB.lib.sdb_ipc_open(_ct.create_string_buffer(tok, 64), _ct.byref(p)), "sdb_ipc_open")
                self._peers_opened.append(p)
                self.peer_ptrs.append(p.value)
        self._torch_panel = None
        if self.mode == "nccl":
            import torch

            self._torch_panel = torch.as_tensor(
                _CudaView(self.panel.ptr.value, (self.panel_rows, self.n), self.dtype), device="cuda")

    # ------------------------------------------------------------------ plumbing
    def _all_gather_obj(self, obj):
        return _all_gather_obj(obj, self.world, self.group)

    def set_x(self, x_host):
        x_host = np.ascontiguousarray(x_host, dtype=self.dtype)
        if x_host.shape != (self.cols, self.n):
            raise ValueError(f"X must be {(self.cols, self.n)}, got {x_host.shape}")
        check(SDB.lib.sdb_memcpy(self.x.ptr, x_host.ctypes.data_as(_ct.c_void_p), x_host.nbytes, 1), "sdb_memcpy")

    def _local_ptr(self):
        return self.panel.at(self.panel_row0 * self.n * self.es)

    def set_local_y(self, y_host):
        y_host = np.ascontiguousarray(y_host, dtype=self.dtype)
        if y_host.shape != (self.rows_local, self.n):
            raise ValueError(f"local Y must be {(self.rows_local, self.n)}, got {y_host.shape}")
        check(SDB.lib.sdb_memcpy(self._local_ptr(), y_host.ctypes.data_as(_ct.c_void_p), y_host.nbytes, 1),
              "sdb_memcpy")

    # ------------------------------------------------------------------ compute
    def run(self, alpha=1.0, beta=0.0, stream_ptr=None):
        """One sharded SpMM step, stream-ordered on ``stream_ptr`` (a cudaStream_t
        as int; None = the library's stream).  Does not synchronise."""
        stream = _ct.c_void_p(stream_ptr) if stream_ptr else None
        a, b = scalar_pair(alpha), scalar_pair(beta)
        if self.mode == "fused":
            ptrs = (_ct.c_void_p * self.world)(*self.peer_ptrs)
            check(
                SDB.lib.sdb_spmm_dev_allgather(a, self.handle.ref, self.x.ptr, self.n, self.n, b, ptrs, self.world,
                                               self.rank, self.row0, self.n, stream),
                "sdb_spmm_dev_allgather",
            )
            return
        check(
            SDB.lib.sdb_spmm_dev(_lib.OP_N, a, self.handle.ref, _lib.LAYOUT_C, self.x.ptr, self.n, self.n, b,
                                 self._local_ptr(), self.n, stream),
            "sdb_spmm_dev",
        )
        if self.mode == "nccl":
            import torch.distributed as dist

            if stream is None:  # NCCL orders after torch's current stream, not the library's
                check(SDB.lib.sdb_device_synchronize(), "sdb_device_synchronize")
            if self.layout.uniform:
                local = self._torch_panel[self.row0:self.row0 + self.rows_local]
                dist.all_gather_into_tensor(self._torch_panel, local, group=self.group)
            else:
                # nnz-balanced blocks differ in row count: ncclAllGather wants equal pieces, so every block is
                # broadcast by its owner instead (same bytes on the wire, one collective per rank)
                for q in range(self.world):
                    first, count = self.layout.block(q)
                    if count:
                        src = dist.get_global_rank(self.group, q) if self.group is not None else q
                        dist.broadcast(self._torch_panel[first:first + count], src=src, group=self.group)

    # ------------------------------------------------------------------ exchange strategy
    STRATEGIES = {"auto": 0, "ce": 1, "stores": 2, "k1": 3, "sm": 4}

    @staticmethod
    def set_exchange(strategy="auto", chunks=0, sms=None, keep=0):
        """Process-wide exchange strategy of the fused mode (sdb_set_allgather): 'auto', 'ce' (copy engines push
        row chunks while the next chunk's kernel runs; ``chunks`` = how many), 'stores' (the kernel's epilogue
        stores into the peer panels), 'k1' (the same with the row-gather kernel)."""
        check(SDB.lib.sdb_set_allgather(RowShardedSpMM.STRATEGIES[strategy], int(chunks)), "sdb_set_allgather")
        if sms:
            check(SDB.lib.sdb_set_allgather_sms(int(sms)), "sdb_set_allgather_sms")
        _lib.set_option("slab_keep", 1 if keep else 0)

    def autotune(self, beta=0.0, stream=None, candidates=None, reps=3):
        """Time the exchange strategies on the live topology (device time, max over ranks) and keep the fastest.
        Which one wins depends on the rank count, the panel width and how the kernel's time compares with
        the NVLink time of the step, so it is measured instead of guessed.  Every candidate computes the same
        product, so the steps count as ordinary steps (``steps_run`` of them are executed).  Collective: every
        rank must call it."""
        import torch
        import torch.distributed as dist

        if self.mode != "fused" or self.world == 1:
            return {"chosen": None, "steps_run": 0}
        if candidates is None:
            # (strategy, chunks, copier SMs, keep): keep = the streaming kernel's gathers carry an L2 evict_last policy
            # (measured at N = 4, profiles/r2_logs/bench_n4_keep.json: keep never wins, so it is not in the default list)
            candidates = [("ce", 5, 0, 0), ("ce", 10, 0, 0), ("ce", 20, 0, 0), ("sm", 10, 8, 0), ("sm", 20, 16, 0),
                          ("stores", 0, 0, 0), ("k1", 0, 0, 0)]
        stream = stream or torch.cuda.current_stream()
        timings, steps_run = [], 0
        for strategy, chunks, sms, keep in candidates:
            self.set_exchange(strategy, chunks, sms, keep)
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                self.run(alpha=1.0, beta=beta, stream_ptr=stream.cuda_stream)  # untimed: first use of this path
                e0.record(stream)
                for _ in range(reps):
                    self.run(alpha=1.0, beta=beta, stream_ptr=stream.cuda_stream)
                e1.record(stream)
            torch.cuda.synchronize()
            steps_run += reps + 1
            t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            timings.append({"strategy": strategy, "chunks": chunks, "sms": sms, "keep": keep,
                            "ms_per_step": float(t.item()), "kernel": _lib.last_spmm_kernel()})
        best = min(timings, key=lambda r: r["ms_per_step"])  # identical on every rank (all-reduced times)
        self.set_exchange(best["strategy"], best["chunks"], best["sms"], best["keep"])
        dist.barrier(group=self.group)
        return {"chosen": best, "candidates": timings, "steps_run": steps_run}

    def synchronize(self):
        check(SDB.lib.sdb_device_synchronize(), "sdb_device_synchronize")
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier(group=self.group)

    # ------------------------------------------------------------------ results
    def read_rows(self, first, count):
        """Rows [first, first+count) of this rank's copy of the panel -> numpy."""
        out = np.empty((count, self.n), dtype=self.dtype)
        check(SDB.lib.sdb_memcpy(out.ctypes.data_as(_ct.c_void_p), self.panel.at(first * self.n * self.es),
                                 out.nbytes, 2), "sdb_memcpy")
        return out

    def read_panel(self):
        return self.read_rows(0, self.panel_rows)

    def close(self):
        for p in self._peers_opened:
            SDB.lib.sdb_ipc_close(p)
        if self._peers_opened and self.world > 1:
            import torch.distributed as dist

            dist.barrier(group=self.group)  # every peer has unmapped before any owner frees
        self._peers_opened = []
        self._torch_panel = None
        self.x.free()
        self.panel.free()
        if self.handle:
            self.handle.destroy()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def spmm_sharded(a_csr, x, world_size, rank, allgather="fused", group=None, out=None, out_scalar=None):
    """Convenience wrapper: every rank passes the SAME global CSR matrix and X;
    rows are split nnz-balanced, each rank computes its block and the full
    product comes back on every rank (``allgather`` != 'none') as a numpy array."""
    bounds = partition_rows(a_csr.indptr, world_size)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    with RowShardedSpMM(row_block(a_csr, lo, hi), x.shape[1], world_size, rank, allgather, group) as plan:
        plan.set_x(x)
        beta = 0.0
        if out is not None:
            plan.set_local_y(out[lo:hi])
            beta = 1.0 if out_scalar is None else out_scalar
        plan.run(alpha=1.0, beta=beta)
        plan.synchronize()
        result = plan.read_panel()
    return result


# --------------------------------------------------------------------------- SpGEMM / gram sharding
def _is_nccl(world_size, group):
    if world_size == 1:
        return False
    import torch.distributed as dist

    return dist.get_backend(group) == "nccl"


class ShardedCSR:
    """A CSR product assembled from the ranks' row blocks and replicated in the HBM of every rank: three torch CUDA
    tensors (int64 row offsets, int32 columns, values) plus a borrowed-array handle over them, so it can be
    multiplied again without leaving the GPU.  ``to_scipy()`` brings it to the host."""

    def __init__(self, indptr, indices, values, shape):
        self.indptr, self.indices, self.values, self.shape = indptr, indices, values, tuple(shape)
        self.nnz = int(indices.shape[0])
        self.dtype = np.dtype(str(values.dtype).replace("torch.", ""))
        ref = _ct.c_void_p()
        check(
            SDB.lib.sdb_create_csr_dev(_ct.byref(ref), self.shape[0], self.shape[1], self.nnz,
                                       _ct.c_void_p(indptr.data_ptr()), _ct.c_void_p(indices.data_ptr()),
                                       _ct.c_void_p(values.data_ptr()), _h._DTYPE_CODE[self.dtype]),
            "sdb_create_csr_dev",
        )
        self.handle = _h.Handle(ref, self.dtype)

    def to_scipy(self):
        import scipy.sparse as sps

        indptr = self.indptr.cpu().numpy()
        it = np.int32 if self.nnz <= np.iinfo(np.int32).max else np.int64
        return sps.csr_matrix((self.values.cpu().numpy(), self.indices.cpu().numpy().astype(it, copy=False),
                               indptr.astype(it)), shape=self.shape)

    def close(self):
        if self.handle:
            self.handle.destroy()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def _device_arrays(handle, rows, nnz, dtype):
    """torch views (no copy) of a handle's interior device arrays (sdb_export_dev)."""
    import torch

    p_ptr, p_idx, p_val = _ct.c_void_p(), _ct.c_void_p(), _ct.c_void_p()
    check(SDB.lib.sdb_export_dev(handle.ref, _ct.byref(p_ptr), _ct.byref(p_idx), _ct.byref(p_val)), "sdb_export_dev")
    indptr = torch.as_tensor(_CudaView(p_ptr.value, (rows + 1,), np.int64), device="cuda")
    if nnz == 0:
        return indptr, None, None
    indices = torch.as_tensor(_CudaView(p_idx.value, (nnz,), np.int32), device="cuda")
    values = torch.as_tensor(_CudaView(p_val.value, (nnz,), dtype), device="cuda")
    return indptr, indices, values


def spgemm_sharded_device(a_csr, b_csr, world_size, rank, group=None, reorder_output=False):
    """C = A @ B with the rows of A split nnz-balanced across ranks and B replicated, entirely on the devices:
    every rank multiplies its row block (sdb_spgemm / sdb_spgemm_ordered; the block stays in HBM), the ranks
    exchange the block sizes (two integers each) and then the blocks themselves with NCCL broadcasts straight out
    of the handles' device arrays into the replicated result — no pickling, no host staging.  Returns a
    ShardedCSR on every rank.  Needs an NCCL process group (or world_size 1)."""
    import torch
    import torch.distributed as dist

    bounds = partition_rows(a_csr.indptr, world_size)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    dtype = np.dtype(a_csr.dtype)
    tdtype = getattr(torch, dtype.name)
    n_cols = b_csr.shape[1]
    block = row_block(a_csr, lo, hi)
    hc = None
    if hi > lo and block.nnz > 0 and b_csr.nnz > 0:
        ha, _, _ = _h.create(block)
        hb, _, _ = _h.create(b_csr)
        with ha, hb:
            ref = _ct.c_void_p()
            fn = "sdb_spgemm_ordered" if reorder_output else "sdb_spgemm"
            check(getattr(SDB.lib, fn)(_lib.OP_N, ha.ref, hb.ref, _ct.byref(ref)), fn)
            hc = _h.Handle(ref, dtype)
    try:
        my_nnz = _h.info(hc)["nnz"] if hc else 0
        sizes = _all_gather_obj((hi - lo, my_nnz), world_size, group)
        rows_total = sum(r for r, _ in sizes)
        nnz_total = sum(z for _, z in sizes)
        row_off = np.cumsum([0] + [r for r, _ in sizes])
        nnz_off = np.cumsum([0] + [z for _, z in sizes])
        indptr = torch.zeros(rows_total + 1, dtype=torch.int64, device="cuda")
        indices = torch.empty(nnz_total, dtype=torch.int32, device="cuda")
        values = torch.empty(nnz_total, dtype=tdtype, device="cuda")
        if hc:
            d_ptr, d_idx, d_val = _device_arrays(hc, hi - lo, my_nnz, dtype)
            indptr[row_off[rank] + 1: row_off[rank + 1] + 1] = d_ptr[1:] + int(nnz_off[rank])
            if my_nnz:
                indices[nnz_off[rank]: nnz_off[rank + 1]] = d_idx
                values[nnz_off[rank]: nnz_off[rank + 1]] = d_val
        else:
            indptr[row_off[rank] + 1: row_off[rank + 1] + 1] = int(nnz_off[rank])
        torch.cuda.synchronize()
        if world_size > 1:
            for q in range(world_size):  # exact-size broadcasts: the blocks differ in size
                r0, r1, z0, z1 = int(row_off[q]), int(row_off[q + 1]), int(nnz_off[q]), int(nnz_off[q + 1])
                if r1 > r0:
                    dist.broadcast(indptr[r0 + 1: r1 + 1], src=dist.get_global_rank(group, q) if group else q, group=group)
                if z1 > z0:
                    src = dist.get_global_rank(group, q) if group else q
                    dist.broadcast(indices[z0:z1], src=src, group=group)
                    dist.broadcast(values[z0:z1], src=src, group=group)
            torch.cuda.synchronize()
    finally:
        if hc:
            hc.destroy()
    return ShardedCSR(indptr, indices, values, (rows_total, n_cols))


def spgemm_sharded(a_csr, b_csr, world_size, rank, group=None, reorder_output=False):
    """C = A @ B with the rows of A split nnz-balanced across ranks and B replicated; every rank returns the full
    product as a scipy matrix.  With an NCCL group the blocks travel device to device (spgemm_sharded_device);
    otherwise (gloo on CPU-only test hosts) every rank multiplies its block through dot_product_mkl and the
    finished blocks are gathered and stacked on the host."""
    import scipy.sparse as sps

    if _is_nccl(world_size, group):
        with spgemm_sharded_device(a_csr, b_csr, world_size, rank, group, reorder_output) as c:
            return c.to_scipy()

    from .api import dot_product_mkl

    bounds = partition_rows(a_csr.indptr, world_size)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    if hi > lo:
        block = dot_product_mkl(row_block(a_csr, lo, hi), b_csr, reorder_output=reorder_output)
    else:
        block = sps.csr_matrix((0, b_csr.shape[1]), dtype=a_csr.dtype)
    blocks = _all_gather_obj(block, world_size, group)
    return stack_row_blocks(blocks, b_csr.shape[1])


def stack_row_blocks(blocks, n_cols):
    """Consecutive CSR row blocks -> one CSR matrix, without touching the entries: data / indices back to
    back, every block's indptr shifted by the running nnz (int32 index arrays when they fit, scipy's rule)."""
    import scipy.sparse as sps

    rows = sum(blk.shape[0] for blk in blocks)
    data = np.concatenate([blk.data for blk in blocks])
    indices = np.concatenate([blk.indices for blk in blocks])
    offsets = np.cumsum([0] + [blk.nnz for blk in blocks])
    indptr = np.concatenate([[0]] + [blk.indptr[1:].astype(np.int64) + offsets[i] for i, blk in enumerate(blocks)])
    it = np.int32 if indptr[-1] <= np.iinfo(np.int32).max and n_cols <= np.iinfo(np.int32).max else np.int64
    return sps.csr_matrix((data, indices.astype(it, copy=False), indptr.astype(it)), shape=(rows, n_cols))


def triangle_row_bounds(n, parts):
    """Row ranges of an n x n upper triangle with about equal AREA: bounds[p] is the first row of owner p
    (rows near the top are longer, so the top owners get fewer rows)."""
    total = n * (n + 1) / 2.0
    bounds = [0]
    for p in range(1, parts):
        # rows [0, r) hold r * n - r (r - 1) / 2 entries; solve for the target share
        target = total * p / parts
        r = int(round((2 * n + 1 - np.sqrt(max(0.0, (2 * n + 1) ** 2 - 8.0 * target))) / 2.0))
        bounds.append(min(n, max(bounds[-1], r)))
    bounds.append(n)
    return bounds


def _download_upper(panel, r0, r1):
    """Rows r0:r1 of the device-resident n x n upper-triangular `panel` as a numpy array.  Only the columns from
    r0 on can be non-zero, so only those cross the bus (sdb_memcpy_2d, staged by the library's copy threads): the
    last of 8 equal-area panels is 35 % of the n columns wide."""
    n = panel.shape[1]
    es = panel.element_size()
    out = np.zeros((r1 - r0, n), dtype=_torch_to_numpy(panel.dtype))
    if r1 > r0:
        check(SDB.lib.sdb_memcpy_2d(_ct.c_void_p(out.ctypes.data + r0 * es), n * es,
                                    _ct.c_void_p(panel.data_ptr() + (r0 * n + r0) * es), n * es,
                                    (n - r0) * es, r1 - r0, 2), "sdb_memcpy_2d")
    return out


def _torch_to_numpy(dtype):
    return {"torch.float32": np.float32, "torch.float64": np.float64}[str(dtype)]


def gram_dense_sharded(a_csr, world_size, rank, group=None, gather=True):
    """Dense upper triangle of A^T A with the ROWS of A split across ranks: A^T A = sum over row blocks of
    A_s^T A_s, so every rank forms the partial gram of its block on its GPU (sdb_syrkd_dev into a zeroed device
    panel) and the partial panels are summed ONTO THEIR OWNERS: the n x n result is cut into row panels of equal
    triangle area (triangle_row_bounds) and each panel is reduced to one rank (NCCL reduce on the device panels) —
    a reduce-scatter, n^2 elements moved in total instead of the 2 n^2 of an all-reduce that leaves everything
    everywhere.  ``gather=True`` then all-gathers the panels so every rank returns the full n x n array (the
    convenience form); ``gather=False`` returns (first_row, panel) — this rank's rows only."""
    n = a_csr.shape[1]
    dtype = np.dtype(a_csr.dtype)
    bounds = partition_rows(a_csr.indptr, world_size)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    block = row_block(a_csr, lo, hi)
    zero, one = scalar_pair(0.0), scalar_pair(1.0)
    owners = triangle_row_bounds(n, world_size)
    if _is_nccl(world_size, group) or world_size == 1:
        import torch
        import torch.distributed as dist

        panel = torch.zeros((n, n), dtype=getattr(torch, dtype.name), device="cuda")  # lower triangle stays zero
        if block.nnz > 0:
            handle, _, _ = _h.create(block)
            with handle:
                check(SDB.lib.sdb_syrkd_dev(_lib.OP_T, handle.ref, one, zero, _ct.c_void_p(panel.data_ptr()),
                                            _lib.LAYOUT_C, n, None), "sdb_syrkd_dev")
                check(SDB.lib.sdb_device_synchronize(), "sdb_device_synchronize")
        if world_size > 1:
            for q in range(world_size):
                r0, r1 = owners[q], owners[q + 1]
                if r1 > r0:
                    dst = dist.get_global_rank(group, q) if group else q
                    dist.reduce(panel[r0:r1], dst=dst, group=group)
            if gather:
                for q in range(world_size):
                    r0, r1 = owners[q], owners[q + 1]
                    if r1 > r0:
                        src = dist.get_global_rank(group, q) if group else q
                        dist.broadcast(panel[r0:r1], src=src, group=group)
        torch.cuda.synchronize()  # the download below runs on the library's stream, not torch's
        if gather or world_size == 1:
            return _download_upper(panel, 0, n)
        return owners[rank], _download_upper(panel, owners[rank], owners[rank + 1])

    # no NCCL (gloo on CPU-only test hosts): per-rank partial gram through the host entry point, summed on the host
    import torch
    import torch.distributed as dist

    host = np.zeros((n, n), dtype=dtype)
    if block.nnz > 0:
        handle, _, _ = _h.create(block)
        with handle:
            check(SDB.lib.sdb_syrkd_new(_lib.OP_T, handle.ref, one, host.ctypes.data_as(_ct.c_void_p), _lib.LAYOUT_C, n),
                  "sdb_syrkd_new")
    t = torch.from_numpy(host)
    dist.all_reduce(t, group=group)
    if gather:
        return host
    return owners[rank], host[owners[rank]:owners[rank + 1]]
