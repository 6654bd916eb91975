// spmm_bsr.cu — Y := alpha * A * X + beta * Y with A in BSR (dense b x b blocks),
// X / Y row-major.  The native kernel behind mkl_sparse_?_mm on a BSR handle
// (_common.py:327-384 create_bsr -> _sparse_dense.py:111-123), BASELINE configs[4]
// "BSR(blocksize 16) variant".
//
// Roofline: HBM.  Per stored block: b*b*sv (block) + 4 (index) + b*n*sv (the b
// rows of X it multiplies) bytes and 2*b*b*n FLOPs — 7.5 FLOP/B at b = 16,
// n = 256 fp32, just under the FP32 FFMA ridge, so the FMA loop is register
// tiled (8 rows x 4 columns per thread) to stay off the shared-memory pipe.
//
// Data movement is TMA bulk copies into a shared-memory ring: one elected
// thread per CTA issues, per pipeline stage, ONE cp.async.bulk for the block
// (b*b values, contiguous) and b cp.async.bulk for the X rows of that block
// column (each a contiguous chunk of the X row), all completing on the stage's
// mbarrier (complete_tx).  Consumers spin on the mbarrier parity, run the FMA
// tile out of shared memory and release the stage through a second mbarrier.
// Tensor cores are NOT used: fp32 parity (1e-5) rules out single-pass TF32 and
// the kernel is HBM-bound with FFMA (see DESIGN.md §3 K2 for the measurement).
#include <cstdlib>

#include "common.h"
#include "tma.cuh"
#include "types.cuh"

namespace sdb {

namespace {


using namespace tma;

template <typename T> struct Vec16;  // 16-byte vector of T
template <> struct Vec16<float> {
    static constexpr int N = 4;
    using type = float4;
};
template <> struct Vec16<double> {
    static constexpr int N = 2;
    using type = double2;
};

}  // namespace

// One CTA = one block row x CW columns.  Threads: (CW / VEC) column groups x 2 row halves;
// each thread owns B/2 rows x VEC columns of the output tile.
template <typename T, int B, int CW, bool COL_MAJOR_BLOCKS, int kBsrStages>
__global__ void __launch_bounds__(2 * CW / Vec16<T>::N)
    spmm_bsr_kernel(int64_t block_rows, const int64_t* __restrict__ bptr, const int32_t* __restrict__ bidx,
                    const T* __restrict__ bval, const T* __restrict__ X, int64_t ldx, int64_t n, T alpha, T beta,
                    T* __restrict__ Y, int64_t ldy) {
    constexpr int VEC = Vec16<T>::N;
    constexpr int TX = CW / VEC;       // column groups
    constexpr int RH = B / 2;          // rows per thread
    constexpr int kThreads = 2 * TX;
    using V = typename Vec16<T>::type;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* s_blk = reinterpret_cast<T*>(smem_raw);                            // [stages][B*B]
    T* s_x = s_blk + kBsrStages * B * B;                                  // [stages][B][CW]
    uint64_t* full = reinterpret_cast<uint64_t*>(s_x + kBsrStages * B * CW);  // [stages]
    uint64_t* empty = full + kBsrStages;                                  // [stages]

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int64_t brow = blockIdx.x;
    const int64_t c0 = int64_t(blockIdx.y) * CW;
    const int cw = int(min(int64_t(CW), n - c0));  // live columns of this chunk (multiple of VEC)
    const int64_t q0 = bptr[brow], q1 = bptr[brow + 1];
    const int nblk = int(q1 - q0);

    if (tid == 0) {
        for (int s = 0; s < kBsrStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, kThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const uint32_t stage_bytes = uint32_t(B * B * sizeof(T) + B * cw * sizeof(T));
    auto issue = [&](int i) {  // tid 0 only: stream block i of this block row into its ring slot
        const int s = i % kBsrStages;
        const int64_t q = q0 + i;
        const int64_t bc = bidx[q];
        mbar_expect_tx(full + s, stage_bytes);
        bulk_g2s(s_blk + s * B * B, bval + q * (B * B), uint32_t(B * B * sizeof(T)), full + s);
        const T* xrow = X + (bc * B) * ldx + c0;
#pragma unroll 4
        for (int k = 0; k < B; ++k)
            bulk_g2s(s_x + (s * B + k) * CW, xrow + k * ldx, uint32_t(cw * sizeof(T)), full + s);
    };

    if (tid == 0) {
        const int pre = nblk < kBsrStages - 1 ? nblk : kBsrStages - 1;
        for (int i = 0; i < pre; ++i) issue(i);
    }

    T acc[RH][VEC];
#pragma unroll
    for (int r = 0; r < RH; ++r)
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc[r][c] = Num<T>::zero();

    const bool col_live = tx * VEC < cw;
    for (int i = 0; i < nblk; ++i) {
        const int s = i % kBsrStages;
        if (tid == 0) {
            const int nxt = i + kBsrStages - 1;
            if (nxt < nblk) {
                // the slot of block nxt was last used by block nxt - stages (= i - 1)
                if (nxt >= kBsrStages) mbar_wait(empty + nxt % kBsrStages, uint32_t((nxt / kBsrStages - 1) & 1));
                issue(nxt);
            }
        }
        __syncwarp();  // warp 0 reconverges before the FMA tile
        mbar_wait(full + s, uint32_t((i / kBsrStages) & 1));
        const T* blk = s_blk + s * B * B;
        const T* xs = s_x + s * B * CW + tx * VEC;
        if (col_live) {
#pragma unroll
            for (int k0 = 0; k0 < B; k0 += 4) {
                V xv[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) xv[kk] = *reinterpret_cast<const V*>(xs + (k0 + kk) * CW);
#pragma unroll
                for (int r = 0; r < RH; ++r) {
                    const int row = ty * RH + r;
                    T a[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        a[kk] = COL_MAJOR_BLOCKS ? blk[(k0 + kk) * B + row] : blk[row * B + k0 + kk];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const T* xe = reinterpret_cast<const T*>(&xv[kk]);
#pragma unroll
                        for (int c = 0; c < VEC; ++c) acc[r][c] = madd(a[kk], xe[c], acc[r][c]);
                    }
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty + s);
    }

    if (!col_live) return;
    const bool beta_zero = Num<T>::is_zero(beta);
#pragma unroll
    for (int r = 0; r < RH; ++r) {
        const int64_t row = brow * B + ty * RH + r;
        T* yp = Y + row * ldy + c0 + tx * VEC;
        V out;
        T* oe = reinterpret_cast<T*>(&out);
        if (beta_zero) {
#pragma unroll
            for (int c = 0; c < VEC; ++c) oe[c] = mul(alpha, acc[r][c]);
        } else {
            const V old = *reinterpret_cast<const V*>(yp);
            const T* pe = reinterpret_cast<const T*>(&old);
#pragma unroll
            for (int c = 0; c < VEC; ++c) oe[c] = madd(alpha, acc[r][c], mul(beta, pe[c]));
        }
        *reinterpret_cast<V*>(yp) = out;
    }
}

template <typename T, int B, int CW, int kBsrStages>
static sdb_status launch_bsr_s(cudaStream_t s, const sdb_mat* a, const T* X, int64_t ldx, int64_t n, T alpha, T beta,
                               T* Y, int64_t ldy) {
    constexpr int VEC = Vec16<T>::N;
    constexpr int kThreads = 2 * CW / VEC;
    const size_t smem = size_t(kBsrStages) * (B * B + B * CW) * sizeof(T) + 2 * kBsrStages * sizeof(uint64_t);
    const int64_t gy = (n + CW - 1) / CW;
    SDB_REQUIRE(a->rows < (int64_t(1) << 31) && gy < 65536, SDB_STATUS_NOT_SUPPORTED, "spmm_bsr: grid too large");
    const dim3 grid(unsigned(a->rows), unsigned(gy));
    note_spmm_kernel("spmm_bsr_kernel<%s,%d,%d,%d,%d>", dtype_cname(Num<T>::dtype), B, CW,
                     a->block_layout == SDB_LAYOUT_COL_MAJOR ? 1 : 0, kBsrStages);
    if (a->block_layout == SDB_LAYOUT_COL_MAJOR) {
        SDB_CUDA(cudaFuncSetAttribute(spmm_bsr_kernel<T, B, CW, true, kBsrStages>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(smem)));
        SDB_LAUNCH((spmm_bsr_kernel<T, B, CW, true, kBsrStages>), grid, kThreads, smem, s, a->rows, a->indptr, a->indices,
                   static_cast<const T*>(a->values), X, ldx, n, alpha, beta, Y, ldy);
    } else {
        SDB_CUDA(cudaFuncSetAttribute(spmm_bsr_kernel<T, B, CW, false, kBsrStages>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(smem)));
        SDB_LAUNCH((spmm_bsr_kernel<T, B, CW, false, kBsrStages>), grid, kThreads, smem, s, a->rows, a->indptr, a->indices,
                   static_cast<const T*>(a->values), X, ldx, n, alpha, beta, Y, ldy);
    }
    return SDB_STATUS_SUCCESS;
}

// Ring depth: short block rows (a handful of blocks) finish before a deep ring pays off and a shallow
// ring lets more CTAs share the SM; long block rows want the deeper prefetch.
static int ring_depth(const sdb_mat* a);
constexpr int64_t kBsrMmaFromBlock = 32;  // smallest block size that takes the tensor-core kernel by default

template <typename T, int B, int CW>
static sdb_status launch_bsr(cudaStream_t s, const sdb_mat* a, const T* X, int64_t ldx, int64_t n, T alpha, T beta,
                             T* Y, int64_t ldy) {
    const int stages = ring_depth(a);
    if (stages <= 2) return launch_bsr_s<T, B, CW, 2>(s, a, X, ldx, n, alpha, beta, Y, ldy);
    if (stages == 3) return launch_bsr_s<T, B, CW, 3>(s, a, X, ldx, n, alpha, beta, Y, ldy);
    return launch_bsr_s<T, B, CW, 4>(s, a, X, ldx, n, alpha, beta, Y, ldy);
}

template <typename T, int B>
static sdb_status pick_cw(cudaStream_t s, const sdb_mat* a, const T* X, int64_t ldx, int64_t n, T alpha, T beta,
                          T* Y, int64_t ldy) {
    constexpr int VEC = Vec16<T>::N;
    if (n > 32 * VEC) return launch_bsr<T, B, 64 * VEC>(s, a, X, ldx, n, alpha, beta, Y, ldy);
    if (n > 16 * VEC) return launch_bsr<T, B, 32 * VEC>(s, a, X, ldx, n, alpha, beta, Y, ldy);
    return launch_bsr<T, B, 16 * VEC>(s, a, X, ldx, n, alpha, beta, Y, ldy);
}

// True when the native kernel covers this call; otherwise the caller uses the CSR expansion.
bool spmm_bsr_supported(const sdb_mat* a, int op, int layout, const void* dX, int64_t n, int64_t ldx, const void* dY,
                        int64_t ldy) {
    if (a->format != SDB_FMT_BSR || op != SDB_OP_NON_TRANSPOSE || layout != SDB_LAYOUT_ROW_MAJOR) return false;
    if (a->dtype != SDB_F32 && a->dtype != SDB_F64) return false;
    if (a->block != 4 && a->block != 8 && a->block != 16 && a->block != 32) return false;
    const int64_t vec = 16 / int64_t(dtype_size(a->dtype));
    if (n < vec || n % vec || ldx % vec || ldy % vec) return false;
    if ((reinterpret_cast<uintptr_t>(dX) | reinterpret_cast<uintptr_t>(dY)) & 15u) return false;
    return a->rows > 0;
}

static int ring_depth(const sdb_mat* a) {
    static const int forced = [] {
        const char* e = getenv("SDB_BSR_STAGES");
        return e ? atoi(e) : 0;
    }();
    const double mean_blocks = a->rows > 0 ? double(a->nnz) / double(a->rows) : 0.0;
    return forced ? forced : (mean_blocks <= 6.0 ? 2 : 4);
}

sdb_status spmm_bsr_device(cudaStream_t s, const sdb_mat* a, const double* alpha, const double* beta, const void* dX,
                           int64_t n, int64_t ldx, void* dY, int64_t ldy) {
    // tensor cores (spmm_bsr_mma.cu) when switched on and the shape is covered; see DESIGN.md K2 for the measurements
    // Automatic rule, from profiles/r2_bsr_mma_table.json: with 16 x 16 blocks the FMA kernel is already at the HBM
    // roof (fp64 0.93 of peak either way; fp32 0.94 ms against 1.09 ms with 3xTF32, whose operand splits cost more
    // than the MMAs save) and 8 x 8 blocks half-fill an m16 tile; with 32 x 32 blocks the FMA loop is compute-bound
    // and the tensor cores win (fp64 5.8 vs 7.7 ms, fp32 3.3 vs 3.6 ms at N = 512).  Only fp64 takes them by
    // default: DMMA is exact, while the fp32 accumulation inside an HMMA truncates, so the 3xTF32 error grows with
    // the number of terms of a row (6e-6 at 512 terms, 1.1e-5 at 928: over the 1e-5 bar on long block rows) —
    // fp32 stays opt-in ("bsr_mma" = 1) for callers who know their rows are short.
    const int mma = get_option(kOptBsrMma);
    if ((mma == 1 || (mma < 0 && a->block >= kBsrMmaFromBlock && a->dtype == SDB_F64)) && spmm_bsr_mma_supported(a, n))
        return spmm_bsr_mma_device(s, a, alpha, beta, dX, n, ldx, dY, ldy, ring_depth(a));
#define SDB_BSR_CASE(T, B)                                                                                    \
    return pick_cw<T, B>(s, a, static_cast<const T*>(dX), ldx, n, Num<T>::make(alpha[0], alpha[1]),           \
                         Num<T>::make(beta[0], beta[1]), static_cast<T*>(dY), ldy)
    if (a->dtype == SDB_F32) {
        switch (a->block) {
            case 4: SDB_BSR_CASE(float, 4);
            case 8: SDB_BSR_CASE(float, 8);
            case 16: SDB_BSR_CASE(float, 16);
            case 32: SDB_BSR_CASE(float, 32);
        }
    } else if (a->dtype == SDB_F64) {
        switch (a->block) {
            case 4: SDB_BSR_CASE(double, 4);
            case 8: SDB_BSR_CASE(double, 8);
            case 16: SDB_BSR_CASE(double, 16);
            case 32: SDB_BSR_CASE(double, 32);
        }
    }
#undef SDB_BSR_CASE
    set_error("spmm_bsr: unsupported dtype/block combination");
    return SDB_STATUS_NOT_SUPPORTED;
}

}  // namespace sdb
