// mkl_shim.cpp — libsdb200_mkl.so: the oneMKL symbol names sparse_dot_mkl binds
// (sparse_dot_mkl/_mkl_interface/_cfunctions.py:43-168), implemented on top of the libsdb200 C-ABI,
// so that the UNMODIFIED reference package runs on the B200 backend:
//
//     MKL_RT=/path/to/libsdb200_mkl.so python -c "import sparse_dot_mkl"
//
// (SURVEY.md §8f rank 1).  The reference loads whatever $MKL_RT names (_load_library.py:37-42), binds 79
// symbols at class-definition time (a missing one is an AttributeError on import) and runs a
// create_csc -> convert_csr -> export_csr self-test to pick its integer width (_mkl_interface/__init__.py:62-125).
//
// Scope: the inspector-executor sparse BLAS the hot path uses is real (create / export / destroy /
// order / convert_csr / mm / mv / spmm / spmmd / syrk / syrkd, s/d/c/z); everything else the
// reference binds (cblas gemm/syrk, sparse QR, PARDISO, RCI CG/FGMRES) exists as a symbol and reports
// "not supported" as loudly as its signature allows — those are out of scope.
//
// Interface: LP64 only (MKL_INT = int32).  The reference's empirical probe tries int64 first; with
// 64-bit index arrays read as 32-bit the row bounds are inconsistent, create fails with
// SPARSE_STATUS_INVALID_VALUE and the reference falls back to int32 — exactly what it does with a
// real LP64 MKL.  Differences from MKL that a caller can observe: create COPIES the arrays to HBM
// (MKL borrows them), so mkl_sparse_order does not reorder the caller's arrays; export hands out
// pointers into host buffers owned by the handle (valid until destroy, as with MKL).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/sdb200.h"

#define SHIM_API extern "C" __attribute__((visibility("default")))

namespace {

typedef int MKL_INT;  // LP64

struct MKL_Complex8 {
    float re, im;
};
struct MKL_Complex16 {
    double re, im;
};
struct matrix_descr {  // _structs.py:13-30, passed by value
    int type, mode, diag;
};
struct MKLVersion {  // _structs.py:66-76
    int MajorVersion, MinorVersion, UpdateVersion;
    char* ProductStatus;
    char* Build;
    char* Processor;
    char* Platform;
};

constexpr uint32_t kShimMagic = 0x4d4b4c35u;

// what a sparse_matrix_t points at
struct ShimMat {
    uint32_t magic = kShimMagic;
    sdb_mat* h = nullptr;
    // host copies handed out by export (interior pointers stay valid until destroy)
    std::vector<MKL_INT> indptr, indices;
    std::vector<unsigned char> values;
};

ShimMat* as_shim(void* p) {
    ShimMat* m = static_cast<ShimMat*>(p);
    return (m != nullptr && m->magic == kShimMagic) ? m : nullptr;
}

int wrap(sdb_mat* h, void** out) {
    ShimMat* m = new (std::nothrow) ShimMat();
    if (!m) {
        sdb_destroy(h);
        return SDB_STATUS_ALLOC_FAILED;
    }
    m->h = h;
    *out = m;
    return SDB_STATUS_SUCCESS;
}

// MKL's 4-array form -> scipy's 3-array form.  The reference always passes rows_start = indptr[:-1],
// rows_end = indptr[1:] (_common.py:310-319); anything else (gaps, 64-bit arrays read as 32-bit) is refused.
int collapse_4array(MKL_INT lines, const MKL_INT* start, const MKL_INT* end, std::vector<MKL_INT>* indptr) {
    if (lines < 0 || (lines > 0 && (!start || !end))) return SDB_STATUS_INVALID_VALUE;
    indptr->assign(size_t(lines) + 1, 0);
    if (lines == 0) return SDB_STATUS_SUCCESS;
    if (start[0] != 0) return SDB_STATUS_INVALID_VALUE;
    for (MKL_INT i = 0; i < lines; ++i) {
        if (end[i] < start[i]) return SDB_STATUS_INVALID_VALUE;
        if (i + 1 < lines && start[i + 1] != end[i]) return SDB_STATUS_INVALID_VALUE;
        (*indptr)[size_t(i)] = start[i];
    }
    (*indptr)[size_t(lines)] = end[lines - 1];
    return SDB_STATUS_SUCCESS;
}

int create_compressed(bool csc, void** A, int indexing, MKL_INT rows, MKL_INT cols, const MKL_INT* start,
                      const MKL_INT* end, const MKL_INT* idx, const void* values, int dtype) {
    if (!A) return SDB_STATUS_INVALID_VALUE;
    *A = nullptr;
    if (indexing != 0) return SDB_STATUS_NOT_SUPPORTED;  // the reference only uses zero-based (_constants.py:27)
    std::vector<MKL_INT> indptr;
    const int st = collapse_4array(csc ? cols : rows, start, end, &indptr);
    if (st != SDB_STATUS_SUCCESS) return st;
    sdb_mat* h = nullptr;
    const int rc = csc ? sdb_create_csc(&h, rows, cols, indptr.data(), idx, 32, values, dtype)
                       : sdb_create_csr(&h, rows, cols, indptr.data(), idx, 32, values, dtype);
    if (rc != SDB_STATUS_SUCCESS) return rc;
    return wrap(h, A);
}

int create_bsr(void** A, int indexing, int block_layout, MKL_INT rows, MKL_INT cols, MKL_INT block,
               const MKL_INT* start, const MKL_INT* end, const MKL_INT* idx, const void* values, int dtype) {
    if (!A) return SDB_STATUS_INVALID_VALUE;
    *A = nullptr;
    if (indexing != 0) return SDB_STATUS_NOT_SUPPORTED;
    std::vector<MKL_INT> indptr;
    const int st = collapse_4array(rows, start, end, &indptr);
    if (st != SDB_STATUS_SUCCESS) return st;
    sdb_mat* h = nullptr;
    const int rc = sdb_create_bsr(&h, rows, cols, block, block_layout, indptr.data(), idx, 32, values, dtype);
    if (rc != SDB_STATUS_SUCCESS) return rc;
    return wrap(h, A);
}

size_t dtype_bytes(int dtype) { return dtype == SDB_F32 ? 4 : (dtype == SDB_C128 ? 16 : 8); }

// download into the handle's host buffers and hand out interior pointers
int export_any(void* A, int want_format, int dtype, int* indexing, int* block_layout, MKL_INT* rows, MKL_INT* cols,
               MKL_INT* block, MKL_INT** start, MKL_INT** end, MKL_INT** idx, void** values) {
    ShimMat* m = as_shim(A);
    if (!m) return SDB_STATUS_NOT_INITIALIZED;
    int fmt = 0, dt = 0, bl = 0;
    int64_t r = 0, c = 0, nnz = 0, bs = 1;
    int rc = sdb_get_info(m->h, &fmt, &dt, &r, &c, &nnz, &bs, &bl);
    if (rc != SDB_STATUS_SUCCESS) return rc;
    if (fmt != want_format || dt != dtype) return SDB_STATUS_INVALID_VALUE;
    if (nnz > std::numeric_limits<MKL_INT>::max() || r > std::numeric_limits<MKL_INT>::max() ||
        c > std::numeric_limits<MKL_INT>::max())
        return SDB_STATUS_ALLOC_FAILED;  // the reference appends its "try ILP64" hint to status 2
    const int64_t lines = fmt == SDB_FMT_CSC ? c : r;
    m->indptr.assign(size_t(lines) + 2, 0);
    m->indices.assign(size_t(nnz) + 1, 0);
    m->values.assign(size_t(nnz) * size_t(bs * bs) * dtype_bytes(dt) + 16, 0);
    rc = sdb_export(m->h, m->indptr.data(), 32, m->indices.data(), 32, m->values.data());
    if (rc != SDB_STATUS_SUCCESS) return rc;
    if (indexing) *indexing = 0;
    if (block_layout) *block_layout = bl;
    if (rows) *rows = MKL_INT(r);
    if (cols) *cols = MKL_INT(c);
    if (block) *block = MKL_INT(bs);
    if (start) *start = m->indptr.data();
    if (end) *end = m->indptr.data() + 1;
    if (idx) *idx = m->indices.data();
    if (values) *values = m->values.data();
    return SDB_STATUS_SUCCESS;
}

int mm_any(int op, double ar, double ai, void* A, int layout, const void* B, MKL_INT n, MKL_INT ldb, double br,
           double bi, void* Cmat, MKL_INT ldc) {
    ShimMat* m = as_shim(A);
    if (!m) return SDB_STATUS_NOT_INITIALIZED;
    const double alpha[2] = {ar, ai}, beta[2] = {br, bi};
    return sdb_spmm(op, alpha, m->h, layout, B, n, ldb, beta, Cmat, ldc);
}

int mv_any(int op, double ar, double ai, void* A, const void* x, double br, double bi, void* y) {
    ShimMat* m = as_shim(A);
    if (!m) return SDB_STATUS_NOT_INITIALIZED;
    const double alpha[2] = {ar, ai}, beta[2] = {br, bi};
    return sdb_spmm(op, alpha, m->h, SDB_LAYOUT_ROW_MAJOR, x, 1, 1, beta, y, 1);
}

int spmmd_any(int op, void* A, void* B, int layout, void* Cmat, MKL_INT ldc) {
    ShimMat *a = as_shim(A), *b = as_shim(B);
    if (!a || !b) return SDB_STATUS_NOT_INITIALIZED;
    return sdb_spgemm_dense(op, a->h, b->h, layout, Cmat, ldc);
}

int syrkd_any(int op, void* A, double ar, double ai, double br, double bi, void* Cmat, int layout, MKL_INT ldc) {
    ShimMat* m = as_shim(A);
    if (!m) return SDB_STATUS_NOT_INITIALIZED;
    const double alpha[2] = {ar, ai}, beta[2] = {br, bi};
    return sdb_syrkd(op, m->h, alpha, beta, Cmat, layout, ldc);
}

void unsupported(const char* name) {
    fprintf(stderr, "libsdb200_mkl: %s is outside the sparse hot path this backend implements (SURVEY.md §2)\n", name);
}

template <typename T> void poison(T* c, long long count) {
    for (long long i = 0; i < count; ++i) c[i] = std::numeric_limits<T>::quiet_NaN();
}

char g_status[128] = "Product", g_build[128] = "sdb200-0.1.0", g_proc[128] = "NVIDIA B200 (sm_100a)",
     g_platform[128] = "CUDA";

}  // namespace

// ------------------------------------------------------------------ create
#define SHIM_CREATE(L, T, DT)                                                                                   \
    SHIM_API int mkl_sparse_##L##_create_csr(void** A, int indexing, MKL_INT rows, MKL_INT cols, MKL_INT* rs,     \
                                             MKL_INT* re, MKL_INT* ci, T* v) {                                    \
        return create_compressed(false, A, indexing, rows, cols, rs, re, ci, v, DT);                             \
    }                                                                                                            \
    SHIM_API int mkl_sparse_##L##_create_csc(void** A, int indexing, MKL_INT rows, MKL_INT cols, MKL_INT* cs,     \
                                             MKL_INT* ce, MKL_INT* ri, T* v) {                                    \
        return create_compressed(true, A, indexing, rows, cols, cs, ce, ri, v, DT);                              \
    }                                                                                                            \
    SHIM_API int mkl_sparse_##L##_create_bsr(void** A, int indexing, int block_layout, MKL_INT rows, MKL_INT cols, \
                                             MKL_INT block, MKL_INT* rs, MKL_INT* re, MKL_INT* ci, T* v) {        \
        return create_bsr(A, indexing, block_layout, rows, cols, block, rs, re, ci, v, DT);                      \
    }                                                                                                            \
    SHIM_API int mkl_sparse_##L##_export_csr(void* A, int* indexing, MKL_INT* rows, MKL_INT* cols, MKL_INT** rs,  \
                                             MKL_INT** re, MKL_INT** ci, T** v) {                                 \
        return export_any(A, SDB_FMT_CSR, DT, indexing, nullptr, rows, cols, nullptr, rs, re, ci,                \
                          reinterpret_cast<void**>(v));                                                          \
    }                                                                                                            \
    SHIM_API int mkl_sparse_##L##_export_csc(void* A, int* indexing, MKL_INT* rows, MKL_INT* cols, MKL_INT** cs,  \
                                             MKL_INT** ce, MKL_INT** ri, T** v) {                                 \
        return export_any(A, SDB_FMT_CSC, DT, indexing, nullptr, rows, cols, nullptr, cs, ce, ri,                \
                          reinterpret_cast<void**>(v));                                                          \
    }                                                                                                            \
    SHIM_API int mkl_sparse_##L##_export_bsr(void* A, int* indexing, int* block_layout, MKL_INT* rows,            \
                                             MKL_INT* cols, MKL_INT* block, MKL_INT** rs, MKL_INT** re,           \
                                             MKL_INT** ci, T** v) {                                               \
        return export_any(A, SDB_FMT_BSR, DT, indexing, block_layout, rows, cols, block, rs, re, ci,             \
                          reinterpret_cast<void**>(v));                                                          \
    }                                                                                                            \
    SHIM_API int mkl_sparse_##L##_spmmd(int op, void* A, void* B, int layout, T* Cm, MKL_INT ldc) {               \
        return spmmd_any(op, A, B, layout, Cm, ldc);                                                             \
    }

SHIM_CREATE(s, float, SDB_F32)
SHIM_CREATE(d, double, SDB_F64)
SHIM_CREATE(c, MKL_Complex8, SDB_C64)
SHIM_CREATE(z, MKL_Complex16, SDB_C128)

// ------------------------------------------------------------------ handle-level
SHIM_API int mkl_sparse_destroy(void* A) {
    ShimMat* m = as_shim(A);
    if (!m) return SDB_STATUS_NOT_INITIALIZED;  // tests/test_mkl.py:128-141 expects an error for NULL
    const int rc = sdb_destroy(m->h);
    m->magic = 0;
    delete m;
    return rc;
}

SHIM_API int mkl_sparse_order(void* A) {
    ShimMat* m = as_shim(A);
    if (!m) return SDB_STATUS_NOT_INITIALIZED;
    return sdb_order(m->h);
}

SHIM_API int mkl_sparse_convert_csr(void* A, int op, void** out) {
    if (!out) return SDB_STATUS_INVALID_VALUE;
    *out = nullptr;
    ShimMat* m = as_shim(A);
    if (!m) return SDB_STATUS_NOT_INITIALIZED;
    sdb_mat* h = nullptr;
    const int rc = sdb_convert_csr(m->h, op, &h);
    if (rc != SDB_STATUS_SUCCESS) return rc;
    return wrap(h, out);
}

SHIM_API int mkl_sparse_spmm(int op, void* A, void* B, void** Cm) {
    if (!Cm) return SDB_STATUS_INVALID_VALUE;
    *Cm = nullptr;
    ShimMat *a = as_shim(A), *b = as_shim(B);
    if (!a || !b) return SDB_STATUS_NOT_INITIALIZED;
    sdb_mat* h = nullptr;
    const int rc = sdb_spgemm(op, a->h, b->h, &h);
    if (rc != SDB_STATUS_SUCCESS) return rc;
    return wrap(h, Cm);
}

SHIM_API int mkl_sparse_syrk(int op, void* A, void** Cm) {
    if (!Cm) return SDB_STATUS_INVALID_VALUE;
    *Cm = nullptr;
    ShimMat* a = as_shim(A);
    if (!a) return SDB_STATUS_NOT_INITIALIZED;
    sdb_mat* h = nullptr;
    // MKL: NON_TRANSPOSE -> A * A^T, (CONJUGATE_)TRANSPOSE -> A^T * A; same convention as sdb_syrk
    const int rc = sdb_syrk(op == SDB_OP_NON_TRANSPOSE ? SDB_OP_NON_TRANSPOSE : SDB_OP_TRANSPOSE, a->h, &h);
    if (rc != SDB_STATUS_SUCCESS) return rc;
    return wrap(h, Cm);
}

// ------------------------------------------------------------------ mm / mv / syrkd (scalars by value)
SHIM_API int mkl_sparse_s_mm(int op, float alpha, void* A, matrix_descr, int layout, const float* B, MKL_INT n,
                             MKL_INT ldb, float beta, float* Cm, MKL_INT ldc) {
    return mm_any(op, alpha, 0, A, layout, B, n, ldb, beta, 0, Cm, ldc);
}
SHIM_API int mkl_sparse_d_mm(int op, double alpha, void* A, matrix_descr, int layout, const double* B, MKL_INT n,
                             MKL_INT ldb, double beta, double* Cm, MKL_INT ldc) {
    return mm_any(op, alpha, 0, A, layout, B, n, ldb, beta, 0, Cm, ldc);
}
SHIM_API int mkl_sparse_c_mm(int op, MKL_Complex8 alpha, void* A, matrix_descr, int layout, const MKL_Complex8* B,
                             MKL_INT n, MKL_INT ldb, MKL_Complex8 beta, MKL_Complex8* Cm, MKL_INT ldc) {
    return mm_any(op, alpha.re, alpha.im, A, layout, B, n, ldb, beta.re, beta.im, Cm, ldc);
}
SHIM_API int mkl_sparse_z_mm(int op, MKL_Complex16 alpha, void* A, matrix_descr, int layout, const MKL_Complex16* B,
                             MKL_INT n, MKL_INT ldb, MKL_Complex16 beta, MKL_Complex16* Cm, MKL_INT ldc) {
    return mm_any(op, alpha.re, alpha.im, A, layout, B, n, ldb, beta.re, beta.im, Cm, ldc);
}

SHIM_API int mkl_sparse_s_mv(int op, float alpha, void* A, matrix_descr, const float* x, float beta, float* y) {
    return mv_any(op, alpha, 0, A, x, beta, 0, y);
}
SHIM_API int mkl_sparse_d_mv(int op, double alpha, void* A, matrix_descr, const double* x, double beta, double* y) {
    return mv_any(op, alpha, 0, A, x, beta, 0, y);
}
SHIM_API int mkl_sparse_c_mv(int op, MKL_Complex8 alpha, void* A, matrix_descr, const MKL_Complex8* x,
                             MKL_Complex8 beta, MKL_Complex8* y) {
    return mv_any(op, alpha.re, alpha.im, A, x, beta.re, beta.im, y);
}
SHIM_API int mkl_sparse_z_mv(int op, MKL_Complex16 alpha, void* A, matrix_descr, const MKL_Complex16* x,
                             MKL_Complex16 beta, MKL_Complex16* y) {
    return mv_any(op, alpha.re, alpha.im, A, x, beta.re, beta.im, y);
}

SHIM_API int mkl_sparse_s_syrkd(int op, void* A, float alpha, float beta, float* Cm, int layout, MKL_INT ldc) {
    return syrkd_any(op, A, alpha, 0, beta, 0, Cm, layout, ldc);
}
SHIM_API int mkl_sparse_d_syrkd(int op, void* A, double alpha, double beta, double* Cm, int layout, MKL_INT ldc) {
    return syrkd_any(op, A, alpha, 0, beta, 0, Cm, layout, ldc);
}
SHIM_API int mkl_sparse_c_syrkd(int op, void* A, MKL_Complex8 alpha, MKL_Complex8 beta, MKL_Complex8* Cm, int layout,
                                MKL_INT ldc) {
    return syrkd_any(op, A, alpha.re, alpha.im, beta.re, beta.im, Cm, layout, ldc);
}
SHIM_API int mkl_sparse_z_syrkd(int op, void* A, MKL_Complex16 alpha, MKL_Complex16 beta, MKL_Complex16* Cm,
                                int layout, MKL_INT ldc) {
    return syrkd_any(op, A, alpha.re, alpha.im, beta.re, beta.im, Cm, layout, ldc);
}

// ------------------------------------------------------------------ service functions
SHIM_API int MKL_Set_Interface_Layer(int) { return 0; }  // LP64 whatever is asked; the reference re-validates
SHIM_API int MKL_Get_Max_Threads(void) {
    int n = 0;
    return sdb_device_count(&n) == SDB_STATUS_SUCCESS && n > 0 ? n : 1;  // "workers" = visible GPUs
}
SHIM_API void MKL_Set_Num_Threads(int) {}
SHIM_API int MKL_Set_Num_Threads_Local(int) { return 0; }
SHIM_API void MKL_Get_Version(MKLVersion* v) {
    if (!v) return;
    v->MajorVersion = 2025;  // the reference warns below 2020 (_mkl_interface/__init__.py:160-163)
    v->MinorVersion = 0;
    v->UpdateVersion = 1;
    v->ProductStatus = g_status;
    v->Build = g_build;
    v->Processor = g_proc;
    v->Platform = g_platform;
}
SHIM_API void MKL_Get_Version_String(char* buf, int len) {
    if (!buf || len <= 0) return;
    char tmp[512];
    if (sdb_version_string(tmp, sizeof(tmp)) != SDB_STATUS_SUCCESS) snprintf(tmp, sizeof(tmp), "libsdb200");
    snprintf(buf, size_t(len), "libsdb200_mkl (oneMKL sparse BLAS names over %.200s)", tmp);
}
SHIM_API void mkl_free_buffers(void) {}

// ------------------------------------------------------------------ cblas gemm / syrk (dense x dense callers)
// void functions in CBLAS: a failure cannot be returned, so it is reported on stderr and the output is poisoned
// with NaNs rather than left looking plausible.
namespace {
void gemm_any(const char* who, int dtype, int layout, int ta, int tb, MKL_INT m, MKL_INT n, MKL_INT k, double ar,
              double ai, const void* a, MKL_INT lda, const void* b, MKL_INT ldb, double br, double bi, void* c,
              MKL_INT ldc) {
    const double alpha[2] = {ar, ai}, beta[2] = {br, bi};
    if (sdb_gemm(layout, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, dtype) != SDB_STATUS_SUCCESS) {
        unsupported(who);
        const long long per = dtype == SDB_C64 || dtype == SDB_C128 ? 2 : 1;
        if (dtype == SDB_F32 || dtype == SDB_C64) poison(static_cast<float*>(c), per * m * n);
        else poison(static_cast<double*>(c), per * m * n);
    }
}
void syrk_any(const char* who, int dtype, int layout, int uplo, int trans, MKL_INT n, MKL_INT k, double ar, double ai,
              const void* a, MKL_INT lda, double br, double bi, void* c, MKL_INT ldc) {
    const double alpha[2] = {ar, ai}, beta[2] = {br, bi};
    if (sdb_syrk_dense(layout, uplo, trans, n, k, alpha, a, lda, beta, c, ldc, dtype) != SDB_STATUS_SUCCESS) {
        unsupported(who);
        const long long per = dtype == SDB_C64 || dtype == SDB_C128 ? 2 : 1;
        if (dtype == SDB_F32 || dtype == SDB_C64) poison(static_cast<float*>(c), per * n * n);
        else poison(static_cast<double*>(c), per * n * n);
    }
}
}  // namespace

SHIM_API void cblas_sgemm(int layout, int ta, int tb, MKL_INT m, MKL_INT n, MKL_INT k, float alpha, const float* a,
                          MKL_INT lda, const float* b, MKL_INT ldb, float beta, float* c, MKL_INT ldc) {
    gemm_any("cblas_sgemm", SDB_F32, layout, ta, tb, m, n, k, alpha, 0, a, lda, b, ldb, beta, 0, c, ldc);
}
SHIM_API void cblas_dgemm(int layout, int ta, int tb, MKL_INT m, MKL_INT n, MKL_INT k, double alpha, const double* a,
                          MKL_INT lda, const double* b, MKL_INT ldb, double beta, double* c, MKL_INT ldc) {
    gemm_any("cblas_dgemm", SDB_F64, layout, ta, tb, m, n, k, alpha, 0, a, lda, b, ldb, beta, 0, c, ldc);
}
SHIM_API void cblas_cgemm(int layout, int ta, int tb, MKL_INT m, MKL_INT n, MKL_INT k, const void* alpha, const void* a,
                          MKL_INT lda, const void* b, MKL_INT ldb, const void* beta, void* c, MKL_INT ldc) {
    const float* al = static_cast<const float*>(alpha);
    const float* be = static_cast<const float*>(beta);
    gemm_any("cblas_cgemm", SDB_C64, layout, ta, tb, m, n, k, al[0], al[1], a, lda, b, ldb, be[0], be[1], c, ldc);
}
SHIM_API void cblas_zgemm(int layout, int ta, int tb, MKL_INT m, MKL_INT n, MKL_INT k, const void* alpha, const void* a,
                          MKL_INT lda, const void* b, MKL_INT ldb, const void* beta, void* c, MKL_INT ldc) {
    const double* al = static_cast<const double*>(alpha);
    const double* be = static_cast<const double*>(beta);
    gemm_any("cblas_zgemm", SDB_C128, layout, ta, tb, m, n, k, al[0], al[1], a, lda, b, ldb, be[0], be[1], c, ldc);
}
SHIM_API void cblas_ssyrk(int layout, int uplo, int trans, MKL_INT n, MKL_INT k, float alpha, const float* a,
                          MKL_INT lda, float beta, float* c, MKL_INT ldc) {
    syrk_any("cblas_ssyrk", SDB_F32, layout, uplo, trans, n, k, alpha, 0, a, lda, beta, 0, c, ldc);
}
SHIM_API void cblas_dsyrk(int layout, int uplo, int trans, MKL_INT n, MKL_INT k, double alpha, const double* a,
                          MKL_INT lda, double beta, double* c, MKL_INT ldc) {
    syrk_any("cblas_dsyrk", SDB_F64, layout, uplo, trans, n, k, alpha, 0, a, lda, beta, 0, c, ldc);
}
SHIM_API void cblas_csyrk(int layout, int uplo, int trans, MKL_INT n, MKL_INT k, const void* alpha, const void* a,
                          MKL_INT lda, const void* beta, void* c, MKL_INT ldc) {
    const float* al = static_cast<const float*>(alpha);
    const float* be = static_cast<const float*>(beta);
    syrk_any("cblas_csyrk", SDB_C64, layout, uplo, trans, n, k, al[0], al[1], a, lda, be[0], be[1], c, ldc);
}
SHIM_API void cblas_zsyrk(int layout, int uplo, int trans, MKL_INT n, MKL_INT k, const void* alpha, const void* a,
                          MKL_INT lda, const void* beta, void* c, MKL_INT ldc) {
    const double* al = static_cast<const double*>(alpha);
    const double* be = static_cast<const double*>(beta);
    syrk_any("cblas_zsyrk", SDB_C128, layout, uplo, trans, n, k, al[0], al[1], a, lda, be[0], be[1], c, ldc);
}

// ------------------------------------------------------------------ bound by the reference, out of scope here
SHIM_API int mkl_sparse_qr_reorder(void*, matrix_descr) { return SDB_STATUS_NOT_SUPPORTED; }
SHIM_API int mkl_sparse_d_qr_factorize(void*, double*) { return SDB_STATUS_NOT_SUPPORTED; }
SHIM_API int mkl_sparse_s_qr_factorize(void*, float*) { return SDB_STATUS_NOT_SUPPORTED; }
SHIM_API int mkl_sparse_d_qr_solve(int, void*, double*, int, MKL_INT, double*, MKL_INT, const double*, MKL_INT) {
    return SDB_STATUS_NOT_SUPPORTED;
}
SHIM_API int mkl_sparse_s_qr_solve(int, void*, float*, int, MKL_INT, float*, MKL_INT, const float*, MKL_INT) {
    return SDB_STATUS_NOT_SUPPORTED;
}

SHIM_API void pardisoinit(void*, const MKL_INT*, MKL_INT*) { unsupported("pardisoinit"); }
SHIM_API void pardiso(void*, const MKL_INT*, const MKL_INT*, const MKL_INT*, const MKL_INT*, const MKL_INT*,
                      const void*, const MKL_INT*, const MKL_INT*, MKL_INT*, const MKL_INT*, MKL_INT*, const MKL_INT*,
                      void*, void*, MKL_INT* error) {
    unsupported("pardiso");
    if (error) *error = -1;  // "input inconsistent"
}

// RCI iterative solvers: (n, x, b, rci_request, ipar, dpar, tmp [, ...]) — report failure through rci_request
#define SHIM_RCI(name)                                                                                  \
    SHIM_API void name(const MKL_INT*, double*, double*, MKL_INT* rci, MKL_INT*, double*, double*) {   \
        unsupported(#name);                                                                             \
        if (rci) *rci = -10000;                                                                         \
    }
SHIM_RCI(dcg_init)
SHIM_RCI(dcg_check)
SHIM_RCI(dcg)
SHIM_RCI(dfgmres_init)
SHIM_RCI(dfgmres_check)
SHIM_RCI(dfgmres)
SHIM_API void dcg_get(const MKL_INT*, double*, double*, MKL_INT* rci, MKL_INT*, double*, double*, MKL_INT* itercount) {
    unsupported("dcg_get");
    if (rci) *rci = -10000;
    if (itercount) *itercount = 0;
}
SHIM_API void dfgmres_get(const MKL_INT*, double*, double*, MKL_INT* rci, MKL_INT*, double*, double*,
                          MKL_INT* itercount) {
    unsupported("dfgmres_get");
    if (rci) *rci = -10000;
    if (itercount) *itercount = 0;
}
// multiple right-hand sides: (n, x, nrhs, b, [method,] rci_request, ipar, dpar, tmp)
SHIM_API void dcgmrhs_init(const MKL_INT*, double*, const MKL_INT*, double*, const MKL_INT*, MKL_INT* rci, MKL_INT*,
                           double*, double*) {
    unsupported("dcgmrhs_init");
    if (rci) *rci = -10000;
}
SHIM_API void dcgmrhs_check(const MKL_INT*, double*, const MKL_INT*, double*, MKL_INT* rci, MKL_INT*, double*,
                            double*) {
    unsupported("dcgmrhs_check");
    if (rci) *rci = -10000;
}
SHIM_API void dcgmrhs(const MKL_INT*, double*, const MKL_INT*, double*, MKL_INT* rci, MKL_INT*, double*, double*) {
    unsupported("dcgmrhs");
    if (rci) *rci = -10000;
}
SHIM_API void dcgmrhs_get(const MKL_INT*, double*, const MKL_INT*, double*, MKL_INT* rci, MKL_INT*, double*, double*,
                          MKL_INT* itercount) {
    unsupported("dcgmrhs_get");
    if (rci) *rci = -10000;
    if (itercount) *itercount = 0;
}
