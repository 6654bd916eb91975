"""
ctypes binding table over libsdb200.so — the seam where the reference binds
libmkl_rt (sparse_dot_mkl/_mkl_interface/_cfunctions.py:32-184 and
_load_library.py:31-96).  Every prototype here matches a declaration in
include/sdb200.h; nothing else in the package touches ctypes function pointers.

The library is the ONLY compute backend.  If it cannot be loaded the import of
sparse_dot_b200 fails with ImportError, like the reference without libmkl_rt.
"""
import ctypes as _ct
import os as _os

_HERE = _os.path.dirname(_os.path.abspath(__file__))
LIB_ENV = "SDB200_LIB"  # analogue of $MKL_RT (_load_library.py:37-42)
DEFAULT_LIB = _os.path.join(_HERE, "libsdb200.so")

# sparse_status_t names (_mkl_interface/_constants.py:2-10)
STATUS_NAMES = {
    0: "SPARSE_STATUS_SUCCESS",
    1: "SPARSE_STATUS_NOT_INITIALIZED",
    2: "SPARSE_STATUS_ALLOC_FAILED",
    3: "SPARSE_STATUS_INVALID_VALUE",
    4: "SPARSE_STATUS_EXECUTION_FAILED",
    5: "SPARSE_STATUS_INTERNAL_ERROR",
    6: "SPARSE_STATUS_NOT_SUPPORTED",
}

LAYOUT_C, LAYOUT_F = 101, 102
OP_N, OP_T, OP_H = 10, 11, 12
F32, F64, C64, C128 = 0, 1, 2, 3
FMT_CSR, FMT_CSC, FMT_BSR = 0, 1, 2

_vp, _i32, _i64 = _ct.c_void_p, _ct.c_int, _ct.c_int64
_pd = _ct.POINTER(_ct.c_double)
_pvp = _ct.POINTER(_vp)

# name -> (restype, argtypes); mirrors include/sdb200.h one to one
PROTOTYPES = {
    "sdb_create_csr": (_i32, [_pvp, _i64, _i64, _vp, _vp, _i32, _vp, _i32]),
    "sdb_create_csc": (_i32, [_pvp, _i64, _i64, _vp, _vp, _i32, _vp, _i32]),
    "sdb_create_bsr": (_i32, [_pvp, _i64, _i64, _i64, _i32, _vp, _vp, _i32, _vp, _i32]),
    "sdb_create_csr_dev": (_i32, [_pvp, _i64, _i64, _i64, _vp, _vp, _vp, _i32]),
    "sdb_destroy": (_i32, [_vp]),
    "sdb_get_info": (_i32, [_vp, _ct.POINTER(_i32), _ct.POINTER(_i32), _ct.POINTER(_i64), _ct.POINTER(_i64),
                            _ct.POINTER(_i64), _ct.POINTER(_i64), _ct.POINTER(_i32)]),
    "sdb_export": (_i32, [_vp, _vp, _i32, _vp, _i32, _vp]),
    "sdb_export_dev": (_i32, [_vp, _pvp, _pvp, _pvp]),
    "sdb_order": (_i32, [_vp]),
    "sdb_invalidate": (_i32, [_vp]),
    "sdb_convert_csr": (_i32, [_vp, _i32, _pvp]),
    "sdb_spmm": (_i32, [_i32, _pd, _vp, _i32, _vp, _i64, _i64, _pd, _vp, _i64]),
    "sdb_spmm_csr_host": (_i32, [_i64, _i64, _vp, _vp, _i32, _vp, _i32, _pd, _vp, _i64, _i64, _pd, _vp, _i64]),
    "sdb_spmm_dev": (_i32, [_i32, _pd, _vp, _i32, _vp, _i64, _i64, _pd, _vp, _i64, _vp]),
    "sdb_spmm_dev_allgather": (_i32, [_pd, _vp, _vp, _i64, _i64, _pd, _pvp, _i32, _i32, _i64, _i64, _vp]),
    "sdb_set_allgather": (_i32, [_i32, _i32]),
    "sdb_set_allgather_sms": (_i32, [_i32]),
    "sdb_spgemm": (_i32, [_i32, _vp, _vp, _pvp]),
    "sdb_spgemm_ordered": (_i32, [_i32, _vp, _vp, _pvp]),
    "sdb_spgemm_dense": (_i32, [_i32, _vp, _vp, _i32, _vp, _i64]),
    "sdb_spgemm_dense_dev": (_i32, [_i32, _vp, _vp, _i32, _vp, _i64, _vp]),
    "sdb_syrk": (_i32, [_i32, _vp, _pvp]),
    "sdb_syrk_ordered": (_i32, [_i32, _vp, _pvp]),
    "sdb_syrkd": (_i32, [_i32, _vp, _pd, _pd, _vp, _i32, _i64]),
    "sdb_syrkd_new": (_i32, [_i32, _vp, _pd, _vp, _i32, _i64]),
    "sdb_syrkd_dev": (_i32, [_i32, _vp, _pd, _pd, _vp, _i32, _i64, _vp]),
    "sdb_gemm": (_i32, [_i32, _i32, _i32, _i64, _i64, _i64, _pd, _vp, _i64, _vp, _i64, _pd, _vp, _i64, _i32]),
    "sdb_syrk_dense": (_i32, [_i32, _i32, _i32, _i64, _i64, _pd, _vp, _i64, _pd, _vp, _i64, _i32]),
    "sdb_partition_rows": (_i32, [_vp, _i32, _i64, _i32, _ct.POINTER(_i64)]),
    "sdb_host_alloc": (_i32, [_pvp, _ct.c_size_t]),
    "sdb_host_free": (_i32, [_vp]),
    "sdb_dev_alloc": (_i32, [_pvp, _ct.c_size_t]),
    "sdb_dev_free": (_i32, [_vp]),
    "sdb_memcpy": (_i32, [_vp, _vp, _ct.c_size_t, _i32]),
    "sdb_memcpy_2d": (_i32, [_vp, _ct.c_size_t, _vp, _ct.c_size_t, _ct.c_size_t, _ct.c_size_t, _i32]),
    "sdb_ipc_export": (_i32, [_vp, _ct.c_char_p]),
    "sdb_ipc_open": (_i32, [_ct.c_char_p, _pvp]),
    "sdb_ipc_close": (_i32, [_vp]),
    "sdb_device_synchronize": (_i32, []),
    "sdb_device_count": (_i32, [_ct.POINTER(_i32)]),
    "sdb_set_device": (_i32, [_i32]),
    "sdb_get_device": (_i32, [_ct.POINTER(_i32)]),
    "sdb_version_string": (_i32, [_ct.c_char_p, _i32]),
    "sdb_last_error": (_i32, [_ct.c_char_p, _i32]),
    "sdb_set_option": (_i32, [_ct.c_char_p, _i32]),
    "sdb_probe_bandwidth": (_i32, [_i32, _i64, _i32, _pd]),
    "sdb_kernel_launches": (_i64, []),
    "sdb_last_spmm_kernel": (_i32, [_ct.c_char_p, _i32]),
    "sdb_last_timing": (_i32, [_pd]),
}


def library_path():
    return _os.environ.get(LIB_ENV) or DEFAULT_LIB


def _load():
    path = library_path()
    if not _os.path.exists(path):
        raise ImportError(
            f"sparse_dot_b200: {path} not found. Build it with "
            "`python -m sparse_dot_b200.build` (needs nvcc; cross-compiles for sm_100a) "
            f"or point ${LIB_ENV} at a built libsdb200.so. There is no CPU fallback."
        )
    try:
        lib = _ct.CDLL(path)
    except OSError as e:  # pragma: no cover - depends on the host
        raise ImportError(f"sparse_dot_b200: cannot load {path}: {e}") from e
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError(f"sparse_dot_b200: {path} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    return lib


class SDB:
    """Process-wide state: the loaded library and the debug switch
    (reference: class MKL, _cfunctions.py:32-41)."""

    lib = _load()
    DEBUG = False


def last_error():
    buf = _ct.create_string_buffer(512)
    SDB.lib.sdb_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(status, fn_name):
    """status != 0 -> ValueError naming the routine and the status code, the
    reference's convention (_common.py:645-668)."""
    if status != 0:
        msg = f"{fn_name} returned {status} ({STATUS_NAMES.get(status, 'UNKNOWN')})"
        detail = last_error()
        if detail:
            msg += f": {detail}"
        raise ValueError(msg)
    if SDB.DEBUG:
        print(f"{fn_name} returned {status} ({STATUS_NAMES[0]})")


def scalar_pair(value):
    """alpha / beta travel as {re, im} doubles (sdb200.h, sdb_spmm)."""
    z = complex(value)
    return (_ct.c_double * 2)(z.real, z.imag)


def version_string():
    buf = _ct.create_string_buffer(512)
    check(SDB.lib.sdb_version_string(buf, 512), "sdb_version_string")
    return buf.value.decode(errors="replace")


def device_count():
    n = _ct.c_int(0)
    status = SDB.lib.sdb_device_count(_ct.byref(n))
    return n.value if status == 0 else 0


def kernel_launches():
    return int(SDB.lib.sdb_kernel_launches())


def last_spmm_kernel():
    """Name of the SpMM kernel the last sdb_spmm* call on this thread launched."""
    buf = _ct.create_string_buffer(160)
    check(SDB.lib.sdb_last_spmm_kernel(buf, 160), "sdb_last_spmm_kernel")
    return buf.value.decode()


def set_option(name, value):
    """Run-time switch of the library (sdb200.h, sdb_set_option): 'bsr_mma', 'spgemm_wide', 'dense_mode', ..."""
    check(SDB.lib.sdb_set_option(name.encode(), int(value)), "sdb_set_option")


def probe_bandwidth(kind, nbytes, iters=5):
    """GB/s of one of the library's bandwidth probes: 0 HBM read, 1 L2 -> SM read, 2 L2 -> SM 512-byte-row gather."""
    out = _ct.c_double(0.0)
    check(SDB.lib.sdb_probe_bandwidth(kind, int(nbytes), int(iters), _ct.byref(out)), "sdb_probe_bandwidth")
    return out.value


def last_timing_ms():
    """(h2d, kernels, d2h) device milliseconds of the last host-pointer call."""
    buf = (_ct.c_double * 3)()
    check(SDB.lib.sdb_last_timing(buf), "sdb_last_timing")
    return tuple(buf)
