"""
The C-ABI boundary without a GPU: libsdb200.so loads, exports every symbol that
include/sdb200.h declares, the ctypes table covers exactly that set, host-only
entry points work, and compute entry points fail LOUDLY (status != 0 ->
ValueError) instead of falling back to the CPU.
"""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import sparse_dot_b200 as sdb
from sparse_dot_b200 import _handles, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sdb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 30
    lib = ctypes.CDLL(_lib.library_path())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in sdb200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes table and header disagree"


def test_no_oracle_or_cpu_fallback_in_product():
    pkg = os.path.join(ROOT, "sparse_dot_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"
                assert "import torch" not in src or f in ("sharded.py",), f"{f} imports torch"


def test_version_string_and_status_names():
    s = sdb.get_version_string()
    assert "libsdb200" in s and "sm_100a" in s
    assert _lib.STATUS_NAMES[4] == "SPARSE_STATUS_EXECUTION_FAILED"


def test_null_handles_are_value_errors():
    with pytest.raises(ValueError, match="sdb_destroy returned 1"):
        _handles.Handle(ctypes.c_void_p(), np.float64).destroy()
    with pytest.raises(ValueError, match="sdb_order returned 1"):
        _handles.order(_handles.Handle(ctypes.c_void_p(), np.float64))


def test_partition_rows_balances_nnz():
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 50, size=10_000)
    for it in (np.int32, np.int64):
        indptr = np.concatenate([[0], np.cumsum(lens)]).astype(it)
        for parts in (1, 2, 3, 8):
            bounds = (ctypes.c_int64 * (parts + 1))()
            st = _lib.SDB.lib.sdb_partition_rows(indptr.ctypes.data_as(ctypes.c_void_p), indptr.itemsize * 8,
                                                 len(lens), parts, bounds)
            assert st == 0
            b = np.array(bounds[:])
            assert b[0] == 0 and b[-1] == len(lens) and np.all(np.diff(b) >= 0)
            per = np.diff(indptr[b])
            assert per.sum() == indptr[-1]
            assert per.max() - per.min() <= 2 * lens.max()
    empty = np.zeros(11, dtype=np.int64)
    bounds = (ctypes.c_int64 * 5)()
    assert _lib.SDB.lib.sdb_partition_rows(empty.ctypes.data_as(ctypes.c_void_p), 64, 10, 4, bounds) == 0
    assert list(bounds) == [0, 2, 5, 7, 10]
    assert _lib.SDB.lib.sdb_partition_rows(None, 64, 10, 4, bounds) == 3


@pytest.mark.skipif(sdb.device_count() > 0, reason="only meaningful without a GPU")
def test_compute_fails_loudly_without_gpu():
    a = sp.random(20, 30, density=0.2, format="csr", dtype=np.float64, random_state=0)
    b = np.ones((30, 4))
    with pytest.raises(ValueError, match="SPARSE_STATUS_EXECUTION_FAILED"):
        sdb.dot_product_mkl(a, b)
    with pytest.raises(ValueError):
        sdb.gram_matrix_mkl(a)


def test_allgather_switch_validates_its_arguments():
    """sdb_set_allgather is host-only state: usable (and checked) without a GPU."""
    lib = _lib.SDB.lib
    assert lib.sdb_set_allgather(1, 10) == 0
    assert lib.sdb_set_allgather(0, 0) == 0
    assert lib.sdb_set_allgather(9, 0) == 3 and lib.sdb_set_allgather(1, -1) == 3


def test_bench_helpers():
    import importlib.util

    spec = importlib.util.spec_from_file_location("_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.mangled_hint("spmm_stream_kernel<float,6,32,2,2>") == "18spmm_stream_kernelIfLi6ELi32ELi2ELi2EE"
    sha = bench.sass_sha("spmm_stream_kernel")
    assert sha is None or len(sha) == 40
    # gather-model bytes of BASELINE configs[1]: 27.03 GB with beta != 0
    assert abs(bench.algorithmic_bytes(1_000_000, 50_000_000, 128) - 27.032e9) < 1e7


def test_traffic_table_entries_name_kernels_of_the_build():
    """profiles/kernel_traffic.json: every entry's match string must resolve to a kernel of the built objects (else
    bench.py could never quote it), and its fields must be the ones bench.py reads."""
    import importlib.util
    import json

    spec = importlib.util.spec_from_file_location("_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    table = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))["kernels"]
    assert "spmm_stream_kernel<float,6,32,2,2>" in table
    if bench.sass_sha("spmm_stream_kernel") is None:
        pytest.skip("cuobjdump or the objects are not available here")
    for name, entry in table.items():
        assert entry["dram_bytes_per_launch"] > 0 and len(entry["sass_sha1"]) == 40, name
        assert bench.sass_sha(entry["sass_match"]) is not None, f"{name}: no kernel matches {entry['sass_match']}"


def test_host_copy_pool_copies_correctly_and_reports_a_rate():
    """kind 3 of sdb_probe_bandwidth runs without a GPU: the pageable -> staging copies of the host pipeline
    (worker threads, non-temporal stores), verified byte for byte inside the probe."""
    gbs = _lib.probe_bandwidth(3, 64 << 20, 1)
    assert gbs > 0.1
