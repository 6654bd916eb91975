"""pytest configuration: the `gpu` marker, import paths, and a one-time build of
the two native libraries the tests need (libsdb200.so — the product — and
oracle/libsdb_oracle.so — the CPU checker)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("MKL_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # build (no-op when up to date); nvcc and gcc both work without a GPU
    lib = os.path.join(ROOT, "sparse_dot_b200", "libsdb200.so")
    if not os.path.exists(lib):
        # by path: importing the package needs the library this builds
        import importlib.util

        spec = importlib.util.spec_from_file_location(
            "_sdb200_build", os.path.join(ROOT, "sparse_dot_b200", "build.py"))
        _build = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_build)
        _build.build()
    from oracle import oracle as _oracle

    _oracle.build()
    _oracle.build_ref()


def _has_gpu():
    try:
        import sparse_dot_b200 as sdb

        return sdb.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
