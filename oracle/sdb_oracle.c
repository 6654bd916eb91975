/*
 * sdb_oracle.c — CPU restatement of the sparse-matmul hot path of sparse_dot_mkl.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sparse_dot_b200/ may import, link or
 * execute this file; it is the checker the CUDA path is compared against in
 * tests/, in __graft_entry__.smoke() and in bench.py's cpu_baseline leg.
 *
 * The reference (pure Python) forwards every floating-point operation to Intel
 * oneMKL, a closed-source dependency that is not under /root/reference and is
 * not pinned by it (setup.py:29 lists only numpy/scipy; CI does `pip install mkl`,
 * .github/workflows/python-package.yml:37).  What is restated here is therefore
 * the *published* semantics of the MKL inspector-executor routines exactly as
 * the reference calls them; each function cites that call site.  The oracle is
 * PINNED (tests/test_oracle.py) against
 *   (1) the reference's own known-answer tests: seeded fixtures compared with
 *       numpy/scipy products (sparse_dot_mkl/tests/test_mkl.py:27-67), and
 *   (2) outputs of real oneMKL 2024.2 (embedded in torch's libtorch_cpu.so)
 *       driven through the reference's exact call sequence (oracle/mkl_ref.py),
 *       committed as tests/golden/ (npz files) by oracle/gen_golden.py.
 *
 * Index convention: int64 row offsets, int32 column indices, zero based.
 * Plain C99 + OpenMP; one function per value type via the T macro trick.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_BAD 3
#define ORC_NOMEM 2

#define LAYOUT_ROW 101
#define LAYOUT_COL 102
#define OP_N 10
#define OP_T 11

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------- */
/* mkl_sparse_order as used at _mkl_interface/_common.py:683-692: ascending   */
/* column index inside every row, values carried along.  Insertion sort for  */
/* short rows, heap-free merge sort otherwise; stable, so duplicate columns  */
/* keep their relative order.                                                */
/* ------------------------------------------------------------------------- */
#define DEFINE_ORDER(SUF, T)                                                              \
    static void msort_##SUF(int32_t* c, T* v, int32_t* tc, T* tv, int64_t n) {            \
        if (n <= 24) {                                                                    \
            for (int64_t i = 1; i < n; ++i) {                                             \
                int32_t ck = c[i];                                                        \
                T vk = v[i];                                                              \
                int64_t j = i - 1;                                                        \
                while (j >= 0 && c[j] > ck) {                                             \
                    c[j + 1] = c[j];                                                      \
                    v[j + 1] = v[j];                                                      \
                    --j;                                                                  \
                }                                                                         \
                c[j + 1] = ck;                                                            \
                v[j + 1] = vk;                                                            \
            }                                                                             \
            return;                                                                       \
        }                                                                                 \
        int64_t h = n / 2;                                                                \
        msort_##SUF(c, v, tc, tv, h);                                                     \
        msort_##SUF(c + h, v + h, tc, tv, n - h);                                         \
        memcpy(tc, c, (size_t)h * sizeof(int32_t));                                       \
        memcpy(tv, v, (size_t)h * sizeof(T));                                             \
        int64_t a = 0, b = h, o = 0;                                                      \
        while (a < h && b < n) {                                                          \
            if (c[b] < tc[a]) {                                                           \
                c[o] = c[b];                                                              \
                v[o++] = v[b++];                                                          \
            } else {                                                                      \
                c[o] = tc[a];                                                             \
                v[o++] = tv[a++];                                                         \
            }                                                                             \
        }                                                                                 \
        while (a < h) {                                                                   \
            c[o] = tc[a];                                                                 \
            v[o++] = tv[a++];                                                             \
        }                                                                                 \
    }                                                                                     \
    int orc_order_##SUF(int64_t rows, const int64_t* indptr, int32_t* indices, T* values) { \
        int64_t longest = 0;                                                              \
        for (int64_t i = 0; i < rows; ++i)                                                \
            if (indptr[i + 1] - indptr[i] > longest) longest = indptr[i + 1] - indptr[i]; \
        int bad = 0;                                                                      \
        _Pragma("omp parallel")                                                           \
        {                                                                                 \
            int32_t* tc = (int32_t*)malloc((size_t)(longest / 2 + 1) * sizeof(int32_t));  \
            T* tv = (T*)malloc((size_t)(longest / 2 + 1) * sizeof(T));                    \
            if (!tc || !tv) {                                                             \
                _Pragma("omp atomic write") bad = 1;                                      \
            } else {                                                                      \
                _Pragma("omp for schedule(dynamic, 256)")                                 \
                for (int64_t i = 0; i < rows; ++i)                                        \
                    msort_##SUF(indices + indptr[i], values + indptr[i], tc, tv,          \
                                indptr[i + 1] - indptr[i]);                               \
            }                                                                             \
            free(tc);                                                                     \
            free(tv);                                                                     \
        }                                                                                 \
        return bad ? ORC_NOMEM : ORC_OK;                                                  \
    }

DEFINE_ORDER(f32, float)
DEFINE_ORDER(f64, double)

/* ------------------------------------------------------------------------- */
/* mkl_sparse_convert_csr on a CSC handle (_common.py:695-722): the CSC       */
/* arrays of A are the CSR arrays of A^T, so conversion is a transpose.       */
/* Counting sort by column; entries of an output row come out in ascending   */
/* source-row order (so a sorted input gives a sorted output).               */
/* in: CSR(rows x cols) -> out: CSR(cols x rows) of the transpose.           */
/* ------------------------------------------------------------------------- */
#define DEFINE_TRANSPOSE(SUF, T)                                                          \
    int orc_transpose_##SUF(int64_t rows, int64_t cols, const int64_t* indptr,            \
                            const int32_t* indices, const T* values, int64_t* t_indptr,   \
                            int32_t* t_indices, T* t_values) {                            \
        int64_t nnz = indptr[rows];                                                       \
        memset(t_indptr, 0, (size_t)(cols + 1) * sizeof(int64_t));                        \
        for (int64_t p = 0; p < nnz; ++p) {                                               \
            if (indices[p] < 0 || indices[p] >= cols) return ORC_BAD;                     \
            t_indptr[indices[p] + 1]++;                                                   \
        }                                                                                 \
        for (int64_t c = 0; c < cols; ++c) t_indptr[c + 1] += t_indptr[c];                \
        int64_t* cursor = (int64_t*)malloc((size_t)(cols > 0 ? cols : 1) * sizeof(int64_t)); \
        if (!cursor) return ORC_NOMEM;                                                    \
        memcpy(cursor, t_indptr, (size_t)cols * sizeof(int64_t));                         \
        for (int64_t i = 0; i < rows; ++i)                                                \
            for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) {                         \
                int64_t q = cursor[indices[p]]++;                                         \
                t_indices[q] = (int32_t)i;                                                \
                t_values[q] = values[p];                                                  \
            }                                                                             \
        free(cursor);                                                                     \
        return ORC_OK;                                                                    \
    }

DEFINE_TRANSPOSE(f32, float)
DEFINE_TRANSPOSE(f64, double)

/* ------------------------------------------------------------------------- */
/* mkl_sparse_?_mm as called at _sparse_dense.py:111-123:                     */
/*     Y := alpha * op(A) * X + beta * Y                                      */
/* A is CSR (rows x cols), descr GENERAL; layout 101 = row-major X/Y with     */
/* leading dimensions ldx/ldy, 102 = column-major.  n = columns of X and Y.   */
/* beta == 0 overwrites Y without reading it (BLAS convention).               */
/* op = N: one output row per CSR row (row-parallel AXPYs of X rows).         */
/* op = T: Y (cols x n) is scaled first, then every CSR row i scatters        */
/*         a_ij * X[i,:] into Y[j,:] (serial: the scatter races otherwise).   */
/* ------------------------------------------------------------------------- */
#define XAT(r, c) (*(layout == LAYOUT_ROW ? &X[(r) * ldx + (c)] : &X[(c) * ldx + (r)]))
#define YAT(r, c) (*(layout == LAYOUT_ROW ? &Y[(r) * ldy + (c)] : &Y[(c) * ldy + (r)]))

#define DEFINE_SPMM(SUF, T)                                                               \
    int orc_spmm_##SUF(int op, double alpha_d, int64_t rows, int64_t cols,                \
                       const int64_t* indptr, const int32_t* indices, const T* values,    \
                       int layout, const T* X, int64_t n, int64_t ldx, double beta_d,     \
                       T* Y, int64_t ldy) {                                               \
        if (layout != LAYOUT_ROW && layout != LAYOUT_COL) return ORC_BAD;                 \
        const T alpha = (T)alpha_d, beta = (T)beta_d;                                     \
        if (op == OP_N) {                                                                 \
            _Pragma("omp parallel for schedule(dynamic, 64)")                             \
            for (int64_t i = 0; i < rows; ++i) {                                          \
                if (layout == LAYOUT_ROW) {                                               \
                    T* y = Y + i * ldy;                                                   \
                    if (beta == (T)0)                                                     \
                        for (int64_t c = 0; c < n; ++c) y[c] = (T)0;                      \
                    else                                                                  \
                        for (int64_t c = 0; c < n; ++c) y[c] *= beta;                     \
                    for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) {                 \
                        const T a = alpha * values[p];                                    \
                        const T* x = X + (int64_t)indices[p] * ldx;                       \
                        for (int64_t c = 0; c < n; ++c) y[c] += a * x[c];                 \
                    }                                                                     \
                } else {                                                                  \
                    for (int64_t c = 0; c < n; ++c) {                                     \
                        T acc = (T)0;                                                     \
                        for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p)               \
                            acc += values[p] * X[c * ldx + indices[p]];                   \
                        T* y = Y + c * ldy + i;                                           \
                        *y = (beta == (T)0 ? (T)0 : beta * *y) + alpha * acc;             \
                    }                                                                     \
                }                                                                         \
            }                                                                             \
            return ORC_OK;                                                                \
        }                                                                                 \
        if (op != OP_T) return ORC_BAD;                                                   \
        for (int64_t r = 0; r < cols; ++r)                                                \
            for (int64_t c = 0; c < n; ++c) YAT(r, c) = (beta == (T)0 ? (T)0 : beta * YAT(r, c)); \
        for (int64_t i = 0; i < rows; ++i)                                                \
            for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) {                         \
                const T a = alpha * values[p];                                            \
                const int64_t j = indices[p];                                             \
                for (int64_t c = 0; c < n; ++c) YAT(j, c) += a * XAT(i, c);               \
            }                                                                             \
        return ORC_OK;                                                                    \
    }

DEFINE_SPMM(f32, float)
DEFINE_SPMM(f64, double)

/* ------------------------------------------------------------------------- */
/* mkl_sparse_spmm as called at _sparse_sparse.py:35-40 (op = 10): C = A * B, */
/* all CSR.  Gustavson row-by-row with a dense marker/accumulator per thread; */
/* two passes (count, then fill) so the caller allocates C between them — the */
/* same symbolic/numeric split the CUDA path uses.  MKL's convention, probed  */
/* on real MKL: an output entry exists for every structural      */
/* product, even when the values cancel to 0.0, and column order inside a row */
/* is unspecified.  Here: order of first touch.  `upper` != 0 keeps only      */
/* entries with col >= row (mkl_sparse_syrk, _gram_matrix.py:70-74).          */
/* ------------------------------------------------------------------------- */
int orc_spgemm_count(int64_t m, int64_t n, const int64_t* a_ptr, const int32_t* a_idx,
                     const int64_t* b_ptr, const int32_t* b_idx, int upper, int64_t* c_ptr) {
    int bad = 0;
    c_ptr[0] = 0;
#pragma omp parallel
    {
        int64_t* mark = (int64_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
        if (!mark) {
#pragma omp atomic write
            bad = 1;
        } else {
            for (int64_t j = 0; j < n; ++j) mark[j] = -1;
#pragma omp for schedule(dynamic, 128)
            for (int64_t i = 0; i < m; ++i) {
                int64_t cnt = 0;
                for (int64_t p = a_ptr[i]; p < a_ptr[i + 1]; ++p) {
                    const int64_t k = a_idx[p];
                    for (int64_t q = b_ptr[k]; q < b_ptr[k + 1]; ++q) {
                        const int64_t j = b_idx[q];
                        if (upper && j < i) continue;
                        if (mark[j] != i) {
                            mark[j] = i;
                            ++cnt;
                        }
                    }
                }
                c_ptr[i + 1] = cnt;
            }
        }
        free(mark);
    }
    if (bad) return ORC_NOMEM;
    for (int64_t i = 0; i < m; ++i) c_ptr[i + 1] += c_ptr[i];
    return ORC_OK;
}

#define DEFINE_SPGEMM_FILL(SUF, T)                                                        \
    int orc_spgemm_fill_##SUF(int64_t m, int64_t n, const int64_t* a_ptr,                 \
                              const int32_t* a_idx, const T* a_val, const int64_t* b_ptr, \
                              const int32_t* b_idx, const T* b_val, int upper,            \
                              const int64_t* c_ptr, int32_t* c_idx, T* c_val) {           \
        int bad = 0;                                                                      \
        _Pragma("omp parallel")                                                           \
        {                                                                                 \
            int64_t* slot = (int64_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));  \
            if (!slot) {                                                                  \
                _Pragma("omp atomic write") bad = 1;                                      \
            } else {                                                                      \
                for (int64_t j = 0; j < n; ++j) slot[j] = -1;                             \
                _Pragma("omp for schedule(dynamic, 128)")                                 \
                for (int64_t i = 0; i < m; ++i) {                                         \
                    const int64_t base = c_ptr[i];                                        \
                    int64_t fill = base;                                                  \
                    for (int64_t p = a_ptr[i]; p < a_ptr[i + 1]; ++p) {                   \
                        const int64_t k = a_idx[p];                                       \
                        const T a = a_val[p];                                             \
                        for (int64_t q = b_ptr[k]; q < b_ptr[k + 1]; ++q) {               \
                            const int64_t j = b_idx[q];                                   \
                            if (upper && j < i) continue;                                 \
                            if (slot[j] < base) {                                         \
                                slot[j] = fill;                                           \
                                c_idx[fill] = (int32_t)j;                                 \
                                c_val[fill] = a * b_val[q];                               \
                                ++fill;                                                   \
                            } else {                                                      \
                                c_val[slot[j]] += a * b_val[q];                           \
                            }                                                             \
                        }                                                                 \
                    }                                                                     \
                }                                                                         \
            }                                                                             \
            free(slot);                                                                   \
        }                                                                                 \
        return bad ? ORC_NOMEM : ORC_OK;                                                  \
    }

DEFINE_SPGEMM_FILL(f32, float)
DEFINE_SPGEMM_FILL(f64, double)

/* ------------------------------------------------------------------------- */
/* mkl_sparse_?_spmmd as called at _sparse_sparse.py:94-101: dense row-major  */
/* (or column-major) C = A * B, OVERWRITING C — there is no beta; garbage in  */
/* the caller's `out` must not leak (tests/test_sparse_sparse.py:286-297).    */
/* ------------------------------------------------------------------------- */
#define DEFINE_SPMMD(SUF, T)                                                              \
    int orc_spmmd_##SUF(int64_t m, int64_t n, const int64_t* a_ptr, const int32_t* a_idx, \
                        const T* a_val, const int64_t* b_ptr, const int32_t* b_idx,       \
                        const T* b_val, int layout, T* C, int64_t ldc) {                  \
        if (layout != LAYOUT_ROW && layout != LAYOUT_COL) return ORC_BAD;                 \
        _Pragma("omp parallel for schedule(dynamic, 64)")                                 \
        for (int64_t i = 0; i < m; ++i) {                                                 \
            for (int64_t j = 0; j < n; ++j) {                                             \
                if (layout == LAYOUT_ROW) C[i * ldc + j] = (T)0;                          \
                else C[j * ldc + i] = (T)0;                                               \
            }                                                                             \
            for (int64_t p = a_ptr[i]; p < a_ptr[i + 1]; ++p) {                           \
                const int64_t k = a_idx[p];                                               \
                const T a = a_val[p];                                                     \
                for (int64_t q = b_ptr[k]; q < b_ptr[k + 1]; ++q) {                       \
                    const int64_t j = b_idx[q];                                           \
                    if (layout == LAYOUT_ROW) C[i * ldc + j] += a * b_val[q];             \
                    else C[j * ldc + i] += a * b_val[q];                                  \
                }                                                                         \
            }                                                                             \
        }                                                                                 \
        return ORC_OK;                                                                    \
    }

DEFINE_SPMMD(f32, float)
DEFINE_SPMMD(f64, double)

/* ------------------------------------------------------------------------- */
/* mkl_sparse_?_syrkd as called at _gram_matrix.py:149-157:                   */
/*     C := alpha * G + beta * C   on the upper triangle (j >= i) only,       */
/* G = A^T A (op = 11) or A A^T (op = 10), C dense n x n.  The strict lower   */
/* triangle is never written (the reference zeroes it itself in one case,     */
/* _gram_matrix.py:168-169).  `at_*` is the CSR of A^T, supplied by the       */
/* caller (orc_transpose_*), so both ops are "row i of L times R restricted   */
/* to j >= i" with (L, R) = (A^T, A) or (A, A^T).                             */
/* ------------------------------------------------------------------------- */
#define DEFINE_SYRKD(SUF, T)                                                              \
    int orc_syrkd_##SUF(int64_t n, const int64_t* l_ptr, const int32_t* l_idx,            \
                        const T* l_val, const int64_t* r_ptr, const int32_t* r_idx,       \
                        const T* r_val, double alpha_d, double beta_d, int layout, T* C,  \
                        int64_t ldc) {                                                    \
        if (layout != LAYOUT_ROW && layout != LAYOUT_COL) return ORC_BAD;                 \
        const T alpha = (T)alpha_d, beta = (T)beta_d;                                     \
        int bad = 0;                                                                      \
        _Pragma("omp parallel")                                                           \
        {                                                                                 \
            T* acc = (T*)malloc((size_t)(n > 0 ? n : 1) * sizeof(T));                     \
            if (!acc) {                                                                   \
                _Pragma("omp atomic write") bad = 1;                                      \
            } else {                                                                      \
                _Pragma("omp for schedule(dynamic, 16)")                                  \
                for (int64_t i = 0; i < n; ++i) {                                         \
                    for (int64_t j = i; j < n; ++j) acc[j] = (T)0;                        \
                    for (int64_t p = l_ptr[i]; p < l_ptr[i + 1]; ++p) {                   \
                        const int64_t k = l_idx[p];                                       \
                        const T a = l_val[p];                                             \
                        for (int64_t q = r_ptr[k]; q < r_ptr[k + 1]; ++q)                 \
                            if (r_idx[q] >= i) acc[r_idx[q]] += a * r_val[q];             \
                    }                                                                     \
                    for (int64_t j = i; j < n; ++j) {                                     \
                        T* c = layout == LAYOUT_ROW ? C + i * ldc + j : C + j * ldc + i;  \
                        *c = (beta == (T)0 ? (T)0 : beta * *c) + alpha * acc[j];          \
                    }                                                                     \
                }                                                                         \
            }                                                                             \
            free(acc);                                                                    \
        }                                                                                 \
        return bad ? ORC_NOMEM : ORC_OK;                                                  \
    }

DEFINE_SYRKD(f32, float)
DEFINE_SYRKD(f64, double)

/* ------------------------------------------------------------------------- */
/* mkl_sparse_convert_csr on a BSR handle (_common.py:695-722, exercised by   */
/* tests/test_mkl.py:251-268): every stored block expands to b*b CSR entries  */
/* (zeros inside a block stay as explicit entries).  block_layout 101 = the   */
/* b x b block is row-major.  Caller sizes the outputs: rows = mb*b,          */
/* nnz = nblocks*b*b.                                                         */
/* ------------------------------------------------------------------------- */
#define DEFINE_BSR2CSR(SUF, T)                                                            \
    int orc_bsr_to_csr_##SUF(int64_t mb, int64_t b, int block_layout,                     \
                             const int64_t* bptr, const int32_t* bidx, const T* bval,     \
                             int64_t* indptr, int32_t* indices, T* values) {              \
        indptr[0] = 0;                                                                    \
        for (int64_t I = 0; I < mb; ++I) {                                                \
            const int64_t nb = bptr[I + 1] - bptr[I];                                     \
            for (int64_t r = 0; r < b; ++r) {                                             \
                const int64_t row = I * b + r;                                            \
                int64_t o = bptr[I] * b * b + r * nb * b;                                 \
                indptr[row + 1] = o + nb * b;                                             \
                for (int64_t q = bptr[I]; q < bptr[I + 1]; ++q)                           \
                    for (int64_t c = 0; c < b; ++c) {                                     \
                        indices[o] = (int32_t)((int64_t)bidx[q] * b + c);                 \
                        values[o] = block_layout == LAYOUT_ROW ? bval[q * b * b + r * b + c] \
                                                               : bval[q * b * b + c * b + r]; \
                        ++o;                                                              \
                    }                                                                     \
            }                                                                             \
        }                                                                                 \
        return ORC_OK;                                                                    \
    }

DEFINE_BSR2CSR(f32, float)
DEFINE_BSR2CSR(f64, double)
