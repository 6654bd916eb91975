// dense.cu — the two dense BLAS-3 calls the reference makes when NO operand is sparse:
//   sdb_gemm        C := alpha * op(A) * op(B) + beta * C      cblas_?gemm  (_dense_dense.py:55-68)
//   sdb_syrk_dense  C := alpha * A^T A (or A A^T) + beta * C   cblas_?syrk, upper triangle (_gram_matrix.py:233-245)
// They exist so that code written against sparse_dot_mkl keeps working when it hands dot_product_mkl /
// gram_matrix_mkl two dense arrays; they are not part of the sparse hot path and are a plain
// shared-memory tiled kernel (64 x 64 x 16 tiles, 4 x 4 outputs per thread, any strides, all four dtypes), not
// a tuned GEMM: one formulation serves row- / column-major and every transpose flag by working on element
// strides.  Host pointers in, host pointers out (H2D / D2H inside), like every non-_dev entry point.
#include "common.h"
#include "types.cuh"

namespace sdb {

namespace {

constexpr int kTile = 64, kTileK = 16, kGemmThreads = 256;

// C[i, j] (row-major, ldc) = alpha * sum_k A(i, k) * B(k, j) + beta * C[i, j];
// A(i, k) = a[i * a_rs + k * a_cs] (conjugated when conj_a), B(k, j) = b[k * b_rs + j * b_cs].
// upper: only tiles / entries with j >= i are computed and written (syrk).
template <typename T>
__global__ void __launch_bounds__(kGemmThreads)
    gemm_tiled_kernel(int64_t m, int64_t n, int64_t k, T alpha, const T* __restrict__ a, int64_t a_rs, int64_t a_cs,
                      bool conj_a, const T* __restrict__ b, int64_t b_rs, int64_t b_cs, bool conj_b, T beta,
                      T* __restrict__ c, int64_t c_rs, int64_t c_cs, bool upper) {
    __shared__ T sa[kTileK][kTile + 1];
    __shared__ T sb[kTileK][kTile + 1];
    const int64_t i0 = int64_t(blockIdx.y) * kTile, j0 = int64_t(blockIdx.x) * kTile;
    if (upper && j0 + kTile <= i0) return;  // tile strictly below the diagonal
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    T acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = Num<T>::zero();
    for (int64_t k0 = 0; k0 < k; k0 += kTileK) {
        for (int e = threadIdx.x; e < kTile * kTileK; e += kGemmThreads) {
            // walk the faster-varying index of each operand first so that the loads coalesce where they can
            const int ka = a_cs == 1 ? e % kTileK : e / kTile, ia = a_cs == 1 ? e / kTileK : e % kTile;
            T va = Num<T>::zero();
            if (i0 + ia < m && k0 + ka < k) {
                va = a[(i0 + ia) * a_rs + (k0 + ka) * a_cs];
                if (conj_a) va = conj_(va);
            }
            sa[ka][ia] = va;
            const int kb = b_cs == 1 ? e / kTile : e % kTileK, jb = b_cs == 1 ? e % kTile : e / kTileK;
            T vb = Num<T>::zero();
            if (j0 + jb < n && k0 + kb < k) {
                vb = b[(k0 + kb) * b_rs + (j0 + jb) * b_cs];
                if (conj_b) vb = conj_(vb);
            }
            sb[kb][jb] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kTileK; ++kk) {
            T ar[4], br[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) ar[r] = sa[kk][ty * 4 + r];
#pragma unroll
            for (int q = 0; q < 4; ++q) br[q] = sb[kk][tx * 4 + q];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[r][q] = madd(ar[r], br[q], acc[r][q]);
        }
        __syncthreads();
    }
    const bool beta_zero = Num<T>::is_zero(beta);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t i = i0 + ty * 4 + r, j = j0 + tx * 4 + q;
            if (i >= m || j >= n || (upper && j < i)) continue;
            T* out = c + i * c_rs + j * c_cs;
            const T v = mul(alpha, acc[r][q]);
            *out = beta_zero ? v : madd(beta, *out, v);
        }
}

template <typename T>
sdb_status launch_gemm(cudaStream_t s, int64_t m, int64_t n, int64_t k, const double* alpha, const void* a,
                       int64_t a_rs, int64_t a_cs, bool conj_a, const void* b, int64_t b_rs, int64_t b_cs, bool conj_b,
                       const double* beta, void* c, int64_t c_rs, int64_t c_cs, bool upper) {
    const int64_t gx = (n + kTile - 1) / kTile, gy = (m + kTile - 1) / kTile;
    SDB_REQUIRE(gx < (int64_t(1) << 31) && gy < 65536, SDB_STATUS_NOT_SUPPORTED, "gemm: grid too large");
    SDB_LAUNCH(gemm_tiled_kernel<T>, dim3(unsigned(gx), unsigned(gy)), kGemmThreads, 0, s, m, n, k,
               Num<T>::make(alpha[0], alpha[1]), static_cast<const T*>(a), a_rs, a_cs, conj_a, static_cast<const T*>(b),
               b_rs, b_cs, conj_b, Num<T>::make(beta[0], beta[1]), static_cast<T*>(c), c_rs, c_cs, upper);
    return SDB_STATUS_SUCCESS;
}

// element strides (row, column) of a stored `rows x cols` matrix in `layout` with leading dimension ld
void strides_of(int layout, int64_t ld, int64_t* rs, int64_t* cs) {
    if (layout == SDB_LAYOUT_ROW_MAJOR) {
        *rs = ld;
        *cs = 1;
    } else {
        *rs = 1;
        *cs = ld;
    }
}

constexpr int kNoTrans = 111, kTrans = 112, kConjTrans = 113, kUpper = 121;

}  // namespace
}  // namespace sdb

using namespace sdb;

extern "C" {

sdb_status sdb_gemm(int layout, int transa, int transb, int64_t m, int64_t n, int64_t k, const double* alpha,
                    const void* A, int64_t lda, const void* B, int64_t ldb, const double* beta, void* C, int64_t ldc,
                    int dtype) {
    SDB_REQUIRE(layout == SDB_LAYOUT_ROW_MAJOR || layout == SDB_LAYOUT_COL_MAJOR, SDB_STATUS_INVALID_VALUE,
                "gemm: bad layout %d", layout);
    SDB_REQUIRE(transa >= kNoTrans && transa <= kConjTrans && transb >= kNoTrans && transb <= kConjTrans,
                SDB_STATUS_INVALID_VALUE, "gemm: bad transpose flag");
    SDB_REQUIRE(m >= 0 && n >= 0 && k >= 0 && alpha && beta, SDB_STATUS_INVALID_VALUE, "gemm: bad arguments");
    const size_t es = dtype_size(dtype);
    SDB_REQUIRE(es != 0, SDB_STATUS_NOT_SUPPORTED, "gemm: unknown dtype %d", dtype);
    if (m == 0 || n == 0) return SDB_STATUS_SUCCESS;
    SDB_REQUIRE(C != nullptr && (k == 0 || (A && B)), SDB_STATUS_INVALID_VALUE, "gemm: null matrix");
    // stored shapes: op(A) is m x k, op(B) is k x n
    const int64_t a_rows = transa == kNoTrans ? m : k, a_cols = transa == kNoTrans ? k : m;
    const int64_t b_rows = transb == kNoTrans ? k : n, b_cols = transb == kNoTrans ? n : k;
    const bool rm = layout == SDB_LAYOUT_ROW_MAJOR;
    SDB_REQUIRE(lda >= (rm ? a_cols : a_rows) && ldb >= (rm ? b_cols : b_rows) && ldc >= (rm ? n : m),
                SDB_STATUS_INVALID_VALUE, "gemm: leading dimension too small");
    Context* ctx;
    SDB_TRY(get_context(&ctx));
    cudaStream_t s = ctx->stream;
    const bool beta_zero = beta[0] == 0.0 && beta[1] == 0.0;
    // device copies keep the host pitches: lines x ld elements
    const size_t a_bytes = size_t(rm ? a_rows : a_cols) * size_t(lda) * es;
    const size_t b_bytes = size_t(rm ? b_rows : b_cols) * size_t(ldb) * es;
    const size_t c_lines = size_t(rm ? m : n), c_run = size_t(rm ? n : m);
    DevBuf da, db, dc;
    SDB_TRY(da.alloc(a_bytes, s));
    SDB_TRY(db.alloc(b_bytes, s));
    SDB_TRY(dc.alloc(c_lines * c_run * es, s));
    if (k > 0) {
        SDB_TRY(h2d(ctx, da.p, A, a_bytes));
        SDB_TRY(h2d(ctx, db.p, B, b_bytes));
    }
    if (!beta_zero) SDB_TRY(h2d_2d(ctx, dc.p, c_run * es, C, size_t(ldc) * es, c_run * es, c_lines));
    int64_t ars, acs, brs, bcs, crs, ccs;
    strides_of(layout, lda, &ars, &acs);
    strides_of(layout, ldb, &brs, &bcs);
    strides_of(layout, int64_t(c_run), &crs, &ccs);
    if (transa != kNoTrans) std::swap(ars, acs);
    if (transb != kNoTrans) std::swap(brs, bcs);
    SDB_TRY(SDB_DISPATCH_DTYPE(dtype, T, [&]() -> sdb_status {
        return launch_gemm<T>(s, m, n, k, alpha, da.p, ars, acs, transa == kConjTrans, db.p, brs, bcs,
                              transb == kConjTrans, beta, dc.p, crs, ccs, false);
    }));
    SDB_TRY(d2h_2d(ctx, C, size_t(ldc) * es, dc.p, c_run * es, c_run * es, c_lines));
    SDB_CUDA(cudaStreamSynchronize(s));
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_syrk_dense(int layout, int uplo, int trans, int64_t n, int64_t k, const double* alpha, const void* A,
                          int64_t lda, const double* beta, void* C, int64_t ldc, int dtype) {
    SDB_REQUIRE(layout == SDB_LAYOUT_ROW_MAJOR || layout == SDB_LAYOUT_COL_MAJOR, SDB_STATUS_INVALID_VALUE,
                "syrk_dense: bad layout %d", layout);
    SDB_REQUIRE(uplo == kUpper, SDB_STATUS_NOT_SUPPORTED, "syrk_dense: only the upper triangle (the reference's call)");
    SDB_REQUIRE(trans >= kNoTrans && trans <= kConjTrans, SDB_STATUS_INVALID_VALUE, "syrk_dense: bad transpose flag");
    SDB_REQUIRE(n >= 0 && k >= 0 && alpha && beta, SDB_STATUS_INVALID_VALUE, "syrk_dense: bad arguments");
    const size_t es = dtype_size(dtype);
    SDB_REQUIRE(es != 0, SDB_STATUS_NOT_SUPPORTED, "syrk_dense: unknown dtype %d", dtype);
    if (n == 0) return SDB_STATUS_SUCCESS;
    SDB_REQUIRE(C != nullptr && (k == 0 || A), SDB_STATUS_INVALID_VALUE, "syrk_dense: null matrix");
    // NoTrans: C = A A^T with A n x k;  Trans: C = A^T A with A k x n
    const int64_t a_rows = trans == kNoTrans ? n : k, a_cols = trans == kNoTrans ? k : n;
    const bool rm = layout == SDB_LAYOUT_ROW_MAJOR;
    SDB_REQUIRE(lda >= (rm ? a_cols : a_rows) && ldc >= n, SDB_STATUS_INVALID_VALUE,
                "syrk_dense: leading dimension too small");
    Context* ctx;
    SDB_TRY(get_context(&ctx));
    cudaStream_t s = ctx->stream;
    const size_t a_bytes = size_t(rm ? a_rows : a_cols) * size_t(lda) * es;
    DevBuf da, dc;
    SDB_TRY(da.alloc(a_bytes, s));
    SDB_TRY(dc.alloc(size_t(n) * size_t(n) * es, s));
    if (k > 0) SDB_TRY(h2d(ctx, da.p, A, a_bytes));
    // the other triangle of the caller's array must survive the whole-panel copy back
    SDB_TRY(h2d_2d(ctx, dc.p, size_t(n) * es, C, size_t(ldc) * es, size_t(n) * es, size_t(n)));
    int64_t rs, cs, crs, ccs;
    strides_of(layout, lda, &rs, &cs);
    strides_of(layout, n, &crs, &ccs);
    // left operand L(i, k) and right operand R(k, j) = L(j, k): L = A (NoTrans) or A^T (Trans)
    int64_t l_rs = rs, l_cs = cs;
    if (trans != kNoTrans) std::swap(l_rs, l_cs);
    SDB_TRY(SDB_DISPATCH_DTYPE(dtype, T, [&]() -> sdb_status {
        return launch_gemm<T>(s, n, n, k, alpha, da.p, l_rs, l_cs, false, da.p, l_cs, l_rs, false, beta, dc.p, crs, ccs,
                              true);
    }));
    SDB_TRY(d2h_2d(ctx, C, size_t(ldc) * es, dc.p, size_t(n) * es, size_t(n) * es, size_t(n)));
    SDB_CUDA(cudaStreamSynchronize(s));
    return SDB_STATUS_SUCCESS;
}

}  // extern "C"
