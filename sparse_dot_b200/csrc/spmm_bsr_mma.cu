// spmm_bsr_mma.cu — BSR x dense on the tensor cores: Y := alpha * A * X + beta * Y with A in BSR (dense
// b x b blocks), X / Y row-major; the same contract and the same TMA-staged ring as spmm_bsr.cu
// (mkl_sparse_?_mm on a BSR handle, _common.py:327-384 create_bsr -> _sparse_dense.py:111-123), with the
// register-tiled FFMA / DFMA loop replaced by warp-level MMAs.  This is the one path of the library where a
// stored block is a real dense tile (BASELINE.json north_star).
//
//   fp64  mma.sync m8n8k4 f64 (DMMA): exact fp64 products and accumulation — same parity bar as DFMA (1e-12).
//   fp32  3xTF32: every operand is split into a TF32 head and a TF32 tail (x = hi + lo, |lo| <= 2^-11 |x|)
//         and each tile product is three mma.sync m16n8k8 tf32 with fp32 accumulation:
//         lo*hi + hi*lo + hi*hi.  The dropped lo*lo term is 2^-22 relative, so the result matches the fp32
//         FFMA kernel to ~1e-6 (the parity bar is 1e-5); single-pass TF32 (2^-11) would not.
//
// Mapping: one CTA = one block row x CW columns, 8 warps; a warp owns a strip of CW / 8 columns (n-tiles of 8)
// and all B rows (m-tiles of 16 / 8).  A fragments come from the block in shared memory (read once per
// k-step, reused by every n-tile), B fragments from the staged X rows, whose row pitch is padded by 32 bytes so
// that the four k-rows a fragment load touches fall into different banks.
// Roofline: HBM, as for spmm_bsr.cu (b*b*sv + 4 + b*n*sv bytes per block); the point of the tensor cores is to
// take the FMA issue pressure away (sm__throughput 67 % with FFMA at 0.81 of the HBM peak) so that the kernel
// stays on the memory roof for b >= 16 and wide panels.
#include <cstdlib>

#include "common.h"
#include "tma.cuh"
#include "types.cuh"

namespace sdb {

namespace {

using namespace tma;

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
// D(16x8) += A(16x8, row) * B(8x8, col), TF32 inputs, fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D(8x8) += A(8x4, row) * B(4x8, col), fp64
__device__ __forceinline__ void mma_f64(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

template <typename T> struct MmaShape;
template <> struct MmaShape<float> {
    static constexpr int M = 16, K = 8, PAD = 8;  // m16n8k8; X row pitch padded by 8 floats
};
template <> struct MmaShape<double> {
    static constexpr int M = 8, K = 4, PAD = 4;  // m8n8k4; padded by 4 doubles
};

}  // namespace

template <typename T, int B, int CW, bool COL_MAJOR_BLOCKS, int STAGES>
__global__ void __launch_bounds__((CW / 8 < 8 ? CW / 8 : 8) * 32)
    spmm_bsr_mma_kernel(int64_t block_rows, const int64_t* __restrict__ bptr, const int32_t* __restrict__ bidx,
                        const T* __restrict__ bval, const T* __restrict__ X, int64_t ldx, int64_t n, T alpha, T beta,
                        T* __restrict__ Y, int64_t ldy) {
    constexpr int WARPS = CW / 8 < 8 ? CW / 8 : 8;
    constexpr int kThreads = WARPS * 32;
    constexpr int WCOLS = CW / WARPS;  // columns per warp
    constexpr int NT = WCOLS / 8;      // n-tiles per warp
    constexpr int M = MmaShape<T>::M, K = MmaShape<T>::K;
    constexpr int MT = (B + M - 1) / M;  // m-tiles (B = 8 with m16: one half-filled tile)
    constexpr int KS = B / K;            // k-steps
    constexpr int XS = CW + MmaShape<T>::PAD;  // padded pitch of a staged X row
    static_assert(B % K == 0 && NT >= 1, "block size / chunk width not covered by the MMA shape");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* s_blk = reinterpret_cast<T*>(smem_raw);                           // [stages][B*B]
    T* s_x = s_blk + STAGES * B * B;                                     // [stages][B][XS]
    uint64_t* full = reinterpret_cast<uint64_t*>(s_x + STAGES * B * XS);  // [stages]
    uint64_t* empty = full + STAGES;                                     // [stages]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;  // groupID / threadID_in_group of the MMA fragment layouts
    const int64_t brow = blockIdx.x;
    const int64_t c0 = int64_t(blockIdx.y) * CW;
    const int cw = int(min(int64_t(CW), n - c0));  // live columns of this chunk (a multiple of 8)
    const int64_t q0 = bptr[brow], q1 = bptr[brow + 1];
    const int nblk = int(q1 - q0);
    (void)block_rows;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const uint32_t stage_bytes = uint32_t(B * B * sizeof(T) + B * cw * sizeof(T));
    auto issue = [&](int i) {  // tid 0 only: stream block i of this block row into its ring slot
        const int s = i % STAGES;
        const int64_t q = q0 + i;
        const int64_t bc = bidx[q];
        mbar_expect_tx(full + s, stage_bytes);
        bulk_g2s(s_blk + s * B * B, bval + q * (B * B), uint32_t(B * B * sizeof(T)), full + s);
        const T* xrow = X + (bc * B) * ldx + c0;
#pragma unroll 4
        for (int k = 0; k < B; ++k)
            bulk_g2s(s_x + (s * B + k) * XS, xrow + k * ldx, uint32_t(cw * sizeof(T)), full + s);
    };
    if (tid == 0) {
        const int pre = nblk < STAGES - 1 ? nblk : STAGES - 1;
        for (int i = 0; i < pre; ++i) issue(i);
    }

    constexpr int CREGS = sizeof(T) == 4 ? 4 : 2;  // accumulator registers per (m-tile, n-tile)
    T acc[MT][NT][CREGS];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int r = 0; r < CREGS; ++r) acc[mt][nt][r] = Num<T>::zero();

    const int wcol0 = warp * WCOLS;  // first column of this warp's strip inside the chunk
    for (int i = 0; i < nblk; ++i) {
        const int s = i % STAGES;
        if (tid == 0) {
            const int nxt = i + STAGES - 1;
            if (nxt < nblk) {
                if (nxt >= STAGES) mbar_wait(empty + nxt % STAGES, uint32_t((nxt / STAGES - 1) & 1));
                issue(nxt);
            }
        }
        __syncwarp();
        mbar_wait(full + s, uint32_t((i / STAGES) & 1));
        const T* blk = s_blk + s * B * B;
        const T* xs = s_x + s * B * XS + wcol0;
        auto a_at = [&](int r, int k) -> T { return COL_MAJOR_BLOCKS ? blk[k * B + r] : blk[r * B + k]; };
        if constexpr (sizeof(T) == 4) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const int r0 = mt * 16 + g, r1 = r0 + 8, k0 = ks * 8 + tig, k1 = k0 + 4;
                    const bool low_half_only = r1 >= B;  // B = 8: rows 8..15 of the tile do not exist
                    split_tf32(a_at(r0, k0), ahi[mt][0], alo[mt][0]);
                    split_tf32(a_at(r0, k1), ahi[mt][2], alo[mt][2]);
                    if (low_half_only) {
                        ahi[mt][1] = alo[mt][1] = ahi[mt][3] = alo[mt][3] = 0u;
                    } else {
                        split_tf32(a_at(r1, k0), ahi[mt][1], alo[mt][1]);
                        split_tf32(a_at(r1, k1), ahi[mt][3], alo[mt][3]);
                    }
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    if (wcol0 + nt * 8 < cw) {  // warp-uniform: dead n-tiles were never staged
                        uint32_t bh0, bl0, bh1, bl1;
                        split_tf32(xs[(ks * 8 + tig) * XS + nt * 8 + g], bh0, bl0);
                        split_tf32(xs[(ks * 8 + tig + 4) * XS + nt * 8 + g], bh1, bl1);
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            mma_tf32(acc[mt][nt], alo[mt], bh0, bh1);
                            mma_tf32(acc[mt][nt], ahi[mt], bl0, bl1);
                            mma_tf32(acc[mt][nt], ahi[mt], bh0, bh1);
                        }
                    }
                }
            }
        } else {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                T a[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) a[mt] = a_at(mt * 8 + g, ks * 4 + tig);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    if (wcol0 + nt * 8 < cw) {
                        const T b = xs[(ks * 4 + tig) * XS + nt * 8 + g];
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) mma_f64(acc[mt][nt], a[mt], b);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }

    // epilogue: every lane holds two adjacent columns of one (fp64) or two (fp32) rows per tile
    const bool beta_zero = Num<T>::is_zero(beta);
    auto store2 = [&](int64_t row, int64_t col, T v0, T v1) {
        T* yp = Y + row * ldy + col;
        if (!beta_zero) {
            v0 = madd(alpha, v0, mul(beta, yp[0]));
            v1 = madd(alpha, v1, mul(beta, yp[1]));
        } else {
            v0 = mul(alpha, v0);
            v1 = mul(alpha, v1);
        }
        if constexpr (sizeof(T) == 4) *reinterpret_cast<float2*>(yp) = make_float2(v0, v1);
        else *reinterpret_cast<double2*>(yp) = make_double2(v0, v1);
    };
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            if (wcol0 + nt * 8 >= cw) continue;
            const int64_t col = c0 + wcol0 + nt * 8 + 2 * tig;
            if constexpr (sizeof(T) == 4) {
                const int r0 = mt * 16 + g;
                store2(brow * B + r0, col, acc[mt][nt][0], acc[mt][nt][1]);
                if (r0 + 8 < B) store2(brow * B + r0 + 8, col, acc[mt][nt][2], acc[mt][nt][3]);
            } else {
                store2(brow * B + mt * 8 + g, col, acc[mt][nt][0], acc[mt][nt][1]);
            }
        }
}

template <typename T, int B, int CW, int STAGES>
static sdb_status launch_mma_s(cudaStream_t s, const sdb_mat* a, const T* X, int64_t ldx, int64_t n, T alpha, T beta,
                               T* Y, int64_t ldy) {
    constexpr int WARPS = CW / 8 < 8 ? CW / 8 : 8;
    constexpr int XS = CW + MmaShape<T>::PAD;
    const size_t smem = size_t(STAGES) * (B * B + B * XS) * sizeof(T) + 2 * STAGES * sizeof(uint64_t);
    const int64_t gy = (n + CW - 1) / CW;
    SDB_REQUIRE(a->rows < (int64_t(1) << 31) && gy < 65536, SDB_STATUS_NOT_SUPPORTED, "spmm_bsr_mma: grid too large");
    const dim3 grid(unsigned(a->rows), unsigned(gy));
    const bool colmaj = a->block_layout == SDB_LAYOUT_COL_MAJOR;
    note_spmm_kernel("spmm_bsr_mma_kernel<%s,%d,%d,%d,%d>", dtype_cname(Num<T>::dtype), B, CW, colmaj ? 1 : 0, STAGES);
    if (colmaj) {
        auto kernel = spmm_bsr_mma_kernel<T, B, CW, true, STAGES>;
        SDB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        SDB_LAUNCH(kernel, grid, WARPS * 32, smem, s, a->rows, a->indptr, a->indices, static_cast<const T*>(a->values), X,
                   ldx, n, alpha, beta, Y, ldy);
    } else {
        auto kernel = spmm_bsr_mma_kernel<T, B, CW, false, STAGES>;
        SDB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        SDB_LAUNCH(kernel, grid, WARPS * 32, smem, s, a->rows, a->indptr, a->indices, static_cast<const T*>(a->values), X,
                   ldx, n, alpha, beta, Y, ldy);
    }
    return SDB_STATUS_SUCCESS;
}

template <typename T, int B, int CW>
static sdb_status launch_mma(cudaStream_t s, const sdb_mat* a, const T* X, int64_t ldx, int64_t n, T alpha, T beta, T* Y,
                             int64_t ldy, int stages) {
    if (stages <= 2) return launch_mma_s<T, B, CW, 2>(s, a, X, ldx, n, alpha, beta, Y, ldy);
    if (stages == 3) return launch_mma_s<T, B, CW, 3>(s, a, X, ldx, n, alpha, beta, Y, ldy);
    return launch_mma_s<T, B, CW, 4>(s, a, X, ldx, n, alpha, beta, Y, ldy);
}

template <typename T, int B>
static sdb_status pick_cw_mma(cudaStream_t s, const sdb_mat* a, const T* X, int64_t ldx, int64_t n, T alpha, T beta,
                              T* Y, int64_t ldy, int stages) {
    constexpr int kWide = sizeof(T) == 4 ? 256 : 128;  // same chunk widths as the FMA kernel
    if (n > kWide / 2) return launch_mma<T, B, kWide>(s, a, X, ldx, n, alpha, beta, Y, ldy, stages);
    if (n > kWide / 4) return launch_mma<T, B, kWide / 2>(s, a, X, ldx, n, alpha, beta, Y, ldy, stages);
    return launch_mma<T, B, kWide / 4>(s, a, X, ldx, n, alpha, beta, Y, ldy, stages);
}

// The tensor-core kernel covers block sizes 8 / 16 / 32 and panels whose width is a multiple of 8.
bool spmm_bsr_mma_supported(const sdb_mat* a, int64_t n) {
    return (a->dtype == SDB_F32 || a->dtype == SDB_F64) && (a->block == 8 || a->block == 16 || a->block == 32) &&
           n % 8 == 0 && n >= 8;
}

sdb_status spmm_bsr_mma_device(cudaStream_t s, const sdb_mat* a, const double* alpha, const double* beta,
                               const void* dX, int64_t n, int64_t ldx, void* dY, int64_t ldy, int stages) {
#define SDB_MMA_CASE(T, B)                                                                                       \
    return pick_cw_mma<T, B>(s, a, static_cast<const T*>(dX), ldx, n, Num<T>::make(alpha[0], alpha[1]),          \
                             Num<T>::make(beta[0], beta[1]), static_cast<T*>(dY), ldy, stages)
    if (a->dtype == SDB_F32) {
        switch (a->block) {
            case 8: SDB_MMA_CASE(float, 8);
            case 16: SDB_MMA_CASE(float, 16);
            case 32: SDB_MMA_CASE(float, 32);
        }
    } else if (a->dtype == SDB_F64) {
        switch (a->block) {
            case 8: SDB_MMA_CASE(double, 8);
            case 16: SDB_MMA_CASE(double, 16);
            case 32: SDB_MMA_CASE(double, 32);
        }
    }
#undef SDB_MMA_CASE
    set_error("spmm_bsr_mma: unsupported dtype/block combination");
    return SDB_STATUS_NOT_SUPPORTED;
}

}  // namespace sdb
