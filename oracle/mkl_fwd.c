/*
 * mkl_fwd.c -> oracle/_ref/libmkl_fwd.so.  TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Purpose: let the UNMODIFIED reference package (pip-installed under baseline/_ref) run its own SpMM / SpGEMM
 * path on REAL Intel oneMKL on this image, so that `bench.py --impl reference` times
 * `sparse_dot_mkl.dot_product_mkl` itself and tests can compare against the reference's own outputs.
 *
 * The image has no libmkl_rt, but torch's libtorch_cpu.so embeds oneMKL 2024.2 (LP64) and exports 29 of the
 * 79 symbols the reference binds at class-definition time (sparse_dot_mkl/_mkl_interface/_cfunctions.py:43-168;
 * a missing one is an AttributeError on import).  This library is linked against libtorch_cpu.so (DT_NEEDED),
 * so `dlsym(handle-of-this-library, name)` — what ctypes does — finds those 29 REAL routines through the
 * dependency, and finds here only what libtorch_cpu.so lacks:
 *
 *   - mkl_sparse_?_create_csc, mkl_sparse_convert_csr   needed by the reference's import-time self-test
 *     (_mkl_interface/__init__.py:62-105).  Implemented as an index transposition on the host followed by the
 *     REAL mkl_sparse_?_create_csr, so every handle a caller ever sees is a genuine MKL handle.
 *   - mkl_sparse_order        per-row ascending sort in place through the REAL export_csr interior pointers
 *                             (_common.py:683-692; semantics = scipy sort_indices()).
 *   - MKL_Set_Interface_Layer / MKL_Get_Version / MKL_Set_Num_Threads / mkl_free_buffers   service calls.
 *   - everything else the reference binds but the hot path never calls (syrk/syrkd, export_csc/bsr, cblas,
 *     sparse QR, PARDISO, RCI solvers): present as symbols; status routines return 6 (NOT_SUPPORTED), void
 *     routines abort with a message.  No arithmetic of the hot path is implemented here: `?_mm`, `spmm`,
 *     `?_spmmd`, `?_mv`, create/export_csr and destroy are Intel's.
 *
 * To know a handle's value type in the untyped calls (order, convert_csr), create_csr / create_bsr / spmm /
 * destroy are thin wrappers that record it and forward to the real routine.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))
typedef int MKL_INT; /* LP64: the MKL inside libtorch_cpu.so */

enum { ST_OK = 0, ST_NOT_INIT = 1, ST_ALLOC = 2, ST_INVALID = 3, ST_EXEC = 4, ST_INTERNAL = 5, ST_NOSUP = 6 };

/* two routines that only libtorch_cpu.so defines: referencing them keeps the DT_NEEDED entry and lets dladdr find it */
extern int MKL_Get_Max_Threads(void);
extern int MKL_Set_Num_Threads_Local(int);
extern void MKL_Get_Version_String(char*, int);

/* ------------------------------------------------------------------ the real routines, resolved once */
typedef int (*create_csr_fn)(void**, int, MKL_INT, MKL_INT, MKL_INT*, MKL_INT*, MKL_INT*, void*);
typedef int (*create_bsr_fn)(void**, int, int, MKL_INT, MKL_INT, MKL_INT, MKL_INT*, MKL_INT*, MKL_INT*, void*);
typedef int (*export_csr_fn)(void*, int*, MKL_INT*, MKL_INT*, MKL_INT**, MKL_INT**, MKL_INT**, void**);
typedef int (*spmm_fn)(int, void*, void*, void**);
typedef int (*destroy_fn)(void*);

static const char LETTERS[4] = {'s', 'd', 'c', 'z'};
static const size_t VSIZE[4] = {4, 8, 8, 16};
static create_csr_fn real_create_csr[4];
static create_bsr_fn real_create_bsr[4];
static export_csr_fn real_export_csr[4];
static spmm_fn real_spmm;
static destroy_fn real_destroy;
static pthread_once_t once = PTHREAD_ONCE_INIT;

static void resolve(void) {
    Dl_info info;
    void* h = NULL;
    if (dladdr((void*)&MKL_Get_Max_Threads, &info) && info.dli_fname) h = dlopen(info.dli_fname, RTLD_NOW | RTLD_NOLOAD);
    if (!h) {
        fprintf(stderr, "mkl_fwd: cannot find the library that provides MKL_Get_Max_Threads\n");
        abort();
    }
    char name[64];
    for (int t = 0; t < 4; ++t) {
        snprintf(name, sizeof name, "mkl_sparse_%c_create_csr", LETTERS[t]);
        real_create_csr[t] = (create_csr_fn)dlsym(h, name);
        snprintf(name, sizeof name, "mkl_sparse_%c_create_bsr", LETTERS[t]);
        real_create_bsr[t] = (create_bsr_fn)dlsym(h, name);
        snprintf(name, sizeof name, "mkl_sparse_%c_export_csr", LETTERS[t]);
        real_export_csr[t] = (export_csr_fn)dlsym(h, name);
        if (!real_create_csr[t] || !real_create_bsr[t] || !real_export_csr[t]) {
            fprintf(stderr, "mkl_fwd: %s lacks the mkl_sparse_%c_* routines\n", info.dli_fname, LETTERS[t]);
            abort();
        }
    }
    real_spmm = (spmm_fn)dlsym(h, "mkl_sparse_spmm");
    real_destroy = (destroy_fn)dlsym(h, "mkl_sparse_destroy");
    if (!real_spmm || !real_destroy) abort();
}

/* ------------------------------------------------------------------ handle registry: value type + arrays we own */
typedef struct rec {
    void* handle;
    int type;   /* index into LETTERS */
    int is_bsr; /* created through create_bsr: export_csr does not apply */
    void *own_ptr, *own_idx, *own_val;
    struct rec* next;
} rec;
static rec* head;
static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;

static int remember(void* handle, int type, int is_bsr, void* p, void* i, void* v) {
    rec* r = (rec*)calloc(1, sizeof *r);
    if (!r) return ST_ALLOC;
    r->handle = handle, r->type = type, r->is_bsr = is_bsr, r->own_ptr = p, r->own_idx = i, r->own_val = v;
    pthread_mutex_lock(&mu);
    r->next = head;
    head = r;
    pthread_mutex_unlock(&mu);
    return ST_OK;
}
static int lookup(void* handle, int* type, int* is_bsr) {
    int found = 0;
    pthread_mutex_lock(&mu);
    for (rec* r = head; r; r = r->next)
        if (r->handle == handle) {
            *type = r->type, *is_bsr = r->is_bsr, found = 1;
            break;
        }
    pthread_mutex_unlock(&mu);
    return found;
}
static void forget(void* handle) {
    pthread_mutex_lock(&mu);
    for (rec** pp = &head; *pp;) {
        if ((*pp)->handle == handle) {
            rec* r = *pp;
            *pp = r->next;
            free(r->own_ptr), free(r->own_idx), free(r->own_val), free(r);
        } else {
            pp = &(*pp)->next;
        }
    }
    pthread_mutex_unlock(&mu);
}

/* ------------------------------------------------------------------ recorded forwards */
#define DEF_CREATE(L, T)                                                                                                  \
    API int mkl_sparse_##L##_create_csr(void** A, int base, MKL_INT rows, MKL_INT cols, MKL_INT* rs, MKL_INT* re,          \
                                        MKL_INT* ci, void* v) {                                                           \
        pthread_once(&once, resolve);                                                                                     \
        int st = real_create_csr[T](A, base, rows, cols, rs, re, ci, v);                                                  \
        if (st == ST_OK && remember(*A, T, 0, NULL, NULL, NULL) != ST_OK) return ST_ALLOC;                                \
        return st;                                                                                                        \
    }                                                                                                                     \
    API int mkl_sparse_##L##_create_bsr(void** A, int base, int blay, MKL_INT rows, MKL_INT cols, MKL_INT bs, MKL_INT* rs, \
                                        MKL_INT* re, MKL_INT* ci, void* v) {                                              \
        pthread_once(&once, resolve);                                                                                     \
        int st = real_create_bsr[T](A, base, blay, rows, cols, bs, rs, re, ci, v);                                        \
        if (st == ST_OK && remember(*A, T, 1, NULL, NULL, NULL) != ST_OK) return ST_ALLOC;                                \
        return st;                                                                                                        \
    }                                                                                                                     \
    API int mkl_sparse_##L##_create_csc(void** A, int base, MKL_INT rows, MKL_INT cols, MKL_INT* cs, MKL_INT* ce,          \
                                        MKL_INT* ri, void* v) {                                                           \
        return create_csc_any(T, A, base, rows, cols, cs, ce, ri, v);                                                     \
    }

/* CSC (4-array, as the reference passes it: cols_start = indptr[:-1], cols_end = indptr[1:], _common.py:310-319)
 * -> CSR arrays we own -> the REAL create_csr.  Anything that is not a consistent 3-array CSC is refused, which is
 * also what makes the reference's "try int64 first" probe fall back to int32 the way it does on a real LP64 MKL. */
static int create_csc_any(int t, void** A, int base, MKL_INT rows, MKL_INT cols, const MKL_INT* cs, const MKL_INT* ce,
                          const MKL_INT* ri, const void* v) {
    pthread_once(&once, resolve);
    if (!A || rows < 0 || cols < 0 || (base != 0 && base != 1)) return ST_INVALID;
    if (cols > 0 && (!cs || !ce)) return ST_NOT_INIT;
    if (cols > 0 && cs[0] != base) return ST_INVALID;
    for (MKL_INT j = 0; j < cols; ++j) {
        if (ce[j] < cs[j]) return ST_INVALID;
        if (j + 1 < cols && cs[j + 1] != ce[j]) return ST_INVALID;
    }
    const int64_t nnz = cols > 0 ? (int64_t)ce[cols - 1] - base : 0;
    if (nnz > 0 && (!ri || !v)) return ST_NOT_INIT;
    for (int64_t e = 0; e < nnz; ++e)
        if (ri[e] - base < 0 || ri[e] - base >= rows) return ST_INVALID;
    const size_t sv = VSIZE[t];
    MKL_INT* ptr = (MKL_INT*)calloc((size_t)rows + 2, sizeof(MKL_INT));
    MKL_INT* idx = (MKL_INT*)malloc(((size_t)nnz + 1) * sizeof(MKL_INT));
    char* val = (char*)malloc(((size_t)nnz + 1) * sv);
    if (!ptr || !idx || !val) {
        free(ptr), free(idx), free(val);
        return ST_ALLOC;
    }
    for (int64_t e = 0; e < nnz; ++e) ptr[ri[e] - base + 1]++;
    for (MKL_INT r = 0; r < rows; ++r) ptr[r + 1] += ptr[r];
    MKL_INT* cur = (MKL_INT*)malloc(((size_t)rows + 1) * sizeof(MKL_INT));
    if (!cur) {
        free(ptr), free(idx), free(val);
        return ST_ALLOC;
    }
    memcpy(cur, ptr, (size_t)rows * sizeof(MKL_INT));
    for (MKL_INT j = 0; j < cols; ++j)
        for (MKL_INT e = cs[j] - base; e < ce[j] - base; ++e) {
            MKL_INT dst = cur[ri[e] - base]++;
            idx[dst] = j;
            memcpy(val + (size_t)dst * sv, (const char*)v + (size_t)e * sv, sv);
        }
    free(cur);
    int st = real_create_csr[t](A, 0, rows, cols, ptr, ptr + 1, idx, val);
    if (st != ST_OK) {
        free(ptr), free(idx), free(val);
        return st;
    }
    return remember(*A, t, 0, ptr, idx, val);
}

DEF_CREATE(s, 0)
DEF_CREATE(d, 1)
DEF_CREATE(c, 2)
DEF_CREATE(z, 3)

API int mkl_sparse_spmm(int op, void* A, void* B, void** C) {
    pthread_once(&once, resolve);
    int st = real_spmm(op, A, B, C);
    int t = 0, bsr = 0;
    if (st == ST_OK && C && *C && lookup(A, &t, &bsr)) remember(*C, t, bsr, NULL, NULL, NULL);
    return st;
}

API int mkl_sparse_destroy(void* A) {
    pthread_once(&once, resolve);
    int st = real_destroy(A);
    forget(A); /* after destroy: MKL borrowed the arrays we own */
    return st;
}

/* mkl_sparse_convert_csr (_common.py:695-722): a new CSR handle holding a copy.  Only op = NON_TRANSPOSE on a CSR-
 * representable source is needed (CSC sources are already CSR inside, see create_csc_any). */
API int mkl_sparse_convert_csr(void* src, int op, void** dst) {
    pthread_once(&once, resolve);
    int t = 0, bsr = 0;
    if (!src || !dst) return ST_NOT_INIT;
    if (!lookup(src, &t, &bsr)) return ST_INVALID;
    if (bsr || op != 10) return ST_NOSUP;
    int base = 0;
    MKL_INT rows = 0, cols = 0, *rs = NULL, *re = NULL, *ci = NULL;
    void* v = NULL;
    int st = real_export_csr[t](src, &base, &rows, &cols, &rs, &re, &ci, &v);
    if (st != ST_OK) return st;
    const size_t sv = VSIZE[t];
    int64_t nnz = 0;
    for (MKL_INT r = 0; r < rows; ++r) nnz += re[r] - rs[r];
    MKL_INT* ptr = (MKL_INT*)calloc((size_t)rows + 2, sizeof(MKL_INT));
    MKL_INT* idx = (MKL_INT*)malloc(((size_t)nnz + 1) * sizeof(MKL_INT));
    char* val = (char*)malloc(((size_t)nnz + 1) * sv);
    if (!ptr || !idx || !val) {
        free(ptr), free(idx), free(val);
        return ST_ALLOC;
    }
    for (MKL_INT r = 0; r < rows; ++r) {
        const MKL_INT n = re[r] - rs[r];
        for (MKL_INT e = 0; e < n; ++e) idx[ptr[r] + e] = ci[rs[r] - base + e] - base;
        memcpy(val + (size_t)ptr[r] * sv, (const char*)v + (size_t)(rs[r] - base) * sv, (size_t)n * sv);
        ptr[r + 1] = ptr[r] + n;
    }
    st = real_create_csr[t](dst, 0, rows, cols, ptr, ptr + 1, idx, val);
    if (st != ST_OK) {
        free(ptr), free(idx), free(val);
        return st;
    }
    return remember(*dst, t, 0, ptr, idx, val);
}

/* mkl_sparse_order (_common.py:683-692): ascending columns inside every row, values permuted along, in place. */
API int mkl_sparse_order(void* A) {
    pthread_once(&once, resolve);
    int t = 0, bsr = 0;
    if (!A) return ST_NOT_INIT;
    if (!lookup(A, &t, &bsr)) return ST_INVALID;
    if (bsr) return ST_NOSUP;
    int base = 0;
    MKL_INT rows = 0, cols = 0, *rs = NULL, *re = NULL, *ci = NULL;
    void* v = NULL;
    int st = real_export_csr[t](A, &base, &rows, &cols, &rs, &re, &ci, &v);
    if (st != ST_OK) return st;
    const size_t sv = VSIZE[t];
    char tmp[16];
    for (MKL_INT r = 0; r < rows; ++r) {
        MKL_INT* c = ci + (rs[r] - base);
        char* w = (char*)v + (size_t)(rs[r] - base) * sv;
        const MKL_INT n = re[r] - rs[r];
        int sorted = 1;
        for (MKL_INT e = 1; e < n && sorted; ++e) sorted = c[e - 1] <= c[e];
        if (sorted) continue;
        /* shell sort on (column, value) pairs: rows are short and this is not a timed path */
        for (MKL_INT gap = n / 2; gap > 0; gap /= 2)
            for (MKL_INT i = gap; i < n; ++i) {
                MKL_INT key = c[i];
                memcpy(tmp, w + (size_t)i * sv, sv);
                MKL_INT j = i;
                for (; j >= gap && c[j - gap] > key; j -= gap) {
                    c[j] = c[j - gap];
                    memcpy(w + (size_t)j * sv, w + (size_t)(j - gap) * sv, sv);
                }
                c[j] = key;
                memcpy(w + (size_t)j * sv, tmp, sv);
            }
    }
    return ST_OK;
}

/* mkl_sparse_?_export_csc (_common.py:442-451 with the csc function table): handles are CSR inside (see
 * create_csc_any), so transpose the exported CSR into arrays owned by the handle's record (valid until destroy). */
static int export_csc_any(int t, void* A, int* base, MKL_INT* rows, MKL_INT* cols, MKL_INT** cs, MKL_INT** ce, MKL_INT** ri,
                          void** v) {
    pthread_once(&once, resolve);
    int tt = 0, bsr = 0;
    if (!A) return ST_NOT_INIT;
    if (!lookup(A, &tt, &bsr) || tt != t) return ST_INVALID;
    if (bsr) return ST_NOSUP;
    int b0 = 0;
    MKL_INT m = 0, n = 0, *rs = NULL, *re = NULL, *ci = NULL;
    void* val = NULL;
    int st = real_export_csr[t](A, &b0, &m, &n, &rs, &re, &ci, &val);
    if (st != ST_OK) return st;
    const size_t sv = VSIZE[t];
    int64_t nnz = 0;
    for (MKL_INT r = 0; r < m; ++r) nnz += re[r] - rs[r];
    MKL_INT* ptr = (MKL_INT*)calloc((size_t)n + 2, sizeof(MKL_INT));
    MKL_INT* idx = (MKL_INT*)malloc(((size_t)nnz + 1) * sizeof(MKL_INT));
    char* out = (char*)malloc(((size_t)nnz + 1) * sv);
    MKL_INT* cur = (MKL_INT*)malloc(((size_t)n + 1) * sizeof(MKL_INT));
    if (!ptr || !idx || !out || !cur) {
        free(ptr), free(idx), free(out), free(cur);
        return ST_ALLOC;
    }
    for (MKL_INT r = 0; r < m; ++r)
        for (MKL_INT e = rs[r] - b0; e < re[r] - b0; ++e) ptr[ci[e] - b0 + 1]++;
    for (MKL_INT j = 0; j < n; ++j) ptr[j + 1] += ptr[j];
    memcpy(cur, ptr, (size_t)n * sizeof(MKL_INT));
    for (MKL_INT r = 0; r < m; ++r)
        for (MKL_INT e = rs[r] - b0; e < re[r] - b0; ++e) {
            MKL_INT dst = cur[ci[e] - b0]++;
            idx[dst] = r;
            memcpy(out + (size_t)dst * sv, (const char*)val + (size_t)e * sv, sv);
        }
    free(cur);
    /* hang the arrays on a second record of the same handle: freed by forget() at destroy */
    st = remember(A, t, 0, ptr, idx, out);
    if (st != ST_OK) {
        free(ptr), free(idx), free(out);
        return st;
    }
    *base = 0, *rows = m, *cols = n, *cs = ptr, *ce = ptr + 1, *ri = idx, *v = out;
    return ST_OK;
}
#define DEF_EXPORT_CSC(L, T)                                                                                        \
    API int mkl_sparse_##L##_export_csc(void* A, int* base, MKL_INT* rows, MKL_INT* cols, MKL_INT** cs, MKL_INT** ce, \
                                        MKL_INT** ri, void** v) {                                                    \
        return export_csc_any(T, A, base, rows, cols, cs, ce, ri, v);                                                \
    }
DEF_EXPORT_CSC(s, 0)
DEF_EXPORT_CSC(d, 1)
DEF_EXPORT_CSC(c, 2)
DEF_EXPORT_CSC(z, 3)

/* ------------------------------------------------------------------ service calls libtorch_cpu.so lacks */
API int MKL_Set_Interface_Layer(int code) {
    (void)code;
    return 0; /* MKL_INTERFACE_LP64: the embedded MKL is LP64 whatever is asked */
}
API void MKL_Set_Num_Threads(int n) { MKL_Set_Num_Threads_Local(n); }
API void mkl_free_buffers(void) {}

typedef struct {
    int MajorVersion, MinorVersion, UpdateVersion;
    char *ProductStatus, *Build, *Processor, *Platform;
} MKLVersion;
API void MKL_Get_Version(MKLVersion* v) {
    static char status[128] = "Product", build[128] = "", proc[128] = "Intel(R) 64 architecture", plat[128] = "libtorch_cpu";
    char s[256] = {0};
    MKL_Get_Version_String(s, 255);
    int major = 0, update = 0;
    const char* p = strstr(s, "Version ");
    if (p) sscanf(p, "Version %d.%d", &major, &update);
    p = strstr(s, "Build ");
    if (p) sscanf(p, "Build %127s", build);
    v->MajorVersion = major, v->MinorVersion = 0, v->UpdateVersion = update;
    v->ProductStatus = status, v->Build = build, v->Processor = proc, v->Platform = plat;
}

/* cblas_?gemm (_dense_dense.py:9-12,53-66; not on the hot path): libtorch_cpu.so exports MKL's Fortran ?gemm_ but not
 * the cblas wrappers.  Row-major C = op(A) op(B) is column-major C^T = op(B)^T op(A)^T, so swap the operands. */
extern void sgemm_(const char*, const char*, const MKL_INT*, const MKL_INT*, const MKL_INT*, const void*, const void*,
                   const MKL_INT*, const void*, const MKL_INT*, const void*, void*, const MKL_INT*);
extern void dgemm_(const char*, const char*, const MKL_INT*, const MKL_INT*, const MKL_INT*, const void*, const void*,
                   const MKL_INT*, const void*, const MKL_INT*, const void*, void*, const MKL_INT*);
extern void cgemm_(const char*, const char*, const MKL_INT*, const MKL_INT*, const MKL_INT*, const void*, const void*,
                   const MKL_INT*, const void*, const MKL_INT*, const void*, void*, const MKL_INT*);
extern void zgemm_(const char*, const char*, const MKL_INT*, const MKL_INT*, const MKL_INT*, const void*, const void*,
                   const MKL_INT*, const void*, const MKL_INT*, const void*, void*, const MKL_INT*);
static char trans_char(int t) { return t == 111 ? 'N' : (t == 112 ? 'T' : 'C'); }
#define GEMM_BODY(F, ALPHA, BETA)                                                                 \
    char ta = trans_char(transa), tb = trans_char(transb);                                        \
    if (layout == 102) F(&ta, &tb, &m, &n, &k, ALPHA, a, &lda, b, &ldb, BETA, c, &ldc);           \
    else F(&tb, &ta, &n, &m, &k, ALPHA, b, &ldb, a, &lda, BETA, c, &ldc);
API void cblas_sgemm(int layout, int transa, int transb, MKL_INT m, MKL_INT n, MKL_INT k, float alpha, const void* a,
                     MKL_INT lda, const void* b, MKL_INT ldb, float beta, void* c, MKL_INT ldc) {
    GEMM_BODY(sgemm_, &alpha, &beta)
}
API void cblas_dgemm(int layout, int transa, int transb, MKL_INT m, MKL_INT n, MKL_INT k, double alpha, const void* a,
                     MKL_INT lda, const void* b, MKL_INT ldb, double beta, void* c, MKL_INT ldc) {
    GEMM_BODY(dgemm_, &alpha, &beta)
}
API void cblas_cgemm(int layout, int transa, int transb, MKL_INT m, MKL_INT n, MKL_INT k, const void* alpha, const void* a,
                     MKL_INT lda, const void* b, MKL_INT ldb, const void* beta, void* c, MKL_INT ldc) {
    GEMM_BODY(cgemm_, alpha, beta)
}
API void cblas_zgemm(int layout, int transa, int transb, MKL_INT m, MKL_INT n, MKL_INT k, const void* alpha, const void* a,
                     MKL_INT lda, const void* b, MKL_INT ldb, const void* beta, void* c, MKL_INT ldc) {
    GEMM_BODY(zgemm_, alpha, beta)
}

/* ------------------------------------------------------------------ bound by the reference, absent from this MKL subset */
#define NOSUP(name) \
    API int name() { return ST_NOSUP; }
#define ABSENT(name)                                                                                         \
    API void name() {                                                                                        \
        fprintf(stderr, "mkl_fwd: " #name " is not part of the oneMKL subset embedded in libtorch_cpu.so\n"); \
        abort();                                                                                             \
    }
NOSUP(mkl_sparse_s_export_bsr) NOSUP(mkl_sparse_d_export_bsr) NOSUP(mkl_sparse_c_export_bsr) NOSUP(mkl_sparse_z_export_bsr)
NOSUP(mkl_sparse_syrk)
NOSUP(mkl_sparse_s_syrkd) NOSUP(mkl_sparse_d_syrkd) NOSUP(mkl_sparse_c_syrkd) NOSUP(mkl_sparse_z_syrkd)
NOSUP(mkl_sparse_qr_reorder)
NOSUP(mkl_sparse_s_qr_factorize) NOSUP(mkl_sparse_d_qr_factorize) NOSUP(mkl_sparse_s_qr_solve) NOSUP(mkl_sparse_d_qr_solve)
ABSENT(cblas_ssyrk) ABSENT(cblas_dsyrk) ABSENT(cblas_csyrk) ABSENT(cblas_zsyrk)
ABSENT(pardisoinit) ABSENT(pardiso)
ABSENT(dcg_init) ABSENT(dcg_check) ABSENT(dcg) ABSENT(dcg_get)
ABSENT(dcgmrhs_init) ABSENT(dcgmrhs_check) ABSENT(dcgmrhs) ABSENT(dcgmrhs_get)
ABSENT(dfgmres_init) ABSENT(dfgmres_check) ABSENT(dfgmres) ABSENT(dfgmres_get)
