// prims.cu — prefix sum, index width conversion, dense transposes, zero fill.
#include "prims.h"

namespace sdb {

// ------------------------------------------------------------------ scan
// Three launches: per-block totals, one-block scan of the totals, per-block
// rescan with the block offset.  A block covers kScanTile items; each warp
// walks a contiguous 1/8 of the tile 32 items at a time (coalesced), carrying
// its running sum in a register.
constexpr int kScanThreads = 256;
constexpr int kScanPerWarp = 512;                           // items per warp
constexpr int kScanTile = kScanPerWarp * (kScanThreads / 32);  // 4096

__device__ __forceinline__ int64_t warp_incl_scan(int64_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int64_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_totals(const int32_t* __restrict__ in, int64_t n,
                                                                 int64_t* __restrict__ totals) {
    __shared__ int64_t warp_sum[kScanThreads / 32];
    const int64_t base = int64_t(blockIdx.x) * kScanTile;
    int64_t acc = 0;
    for (int i = threadIdx.x; i < kScanTile; i += kScanThreads) {
        int64_t idx = base + i;
        if (idx < n) acc += in[idx];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += warp_sum[w];
        totals[blockIdx.x] = t;
    }
}

// exclusive scan of totals[0..m) in place by ONE block; writes the grand total to *grand
__global__ void __launch_bounds__(1024) scan_totals_inplace(int64_t* __restrict__ totals, int64_t m,
                                                            int64_t* __restrict__ grand) {
    __shared__ int64_t warp_sum[32];
    __shared__ int64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < m; base += 1024) {
        int64_t idx = base + threadIdx.x;
        int64_t v = idx < m ? totals[idx] : 0;
        int64_t inc = warp_incl_scan(v, lane);
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int64_t w = warp_sum[lane];
            int64_t winc = warp_incl_scan(w, lane);
            warp_sum[lane] = winc - w;  // exclusive offsets of the 32 warps
        }
        __syncthreads();
        const int64_t carry = carry_s;
        int64_t excl = carry + warp_sum[warp] + inc - v;
        if (idx < m) totals[idx] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand = carry_s;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_apply(const int32_t* __restrict__ in, int64_t n,
                                                                const int64_t* __restrict__ tile_offset,
                                                                int64_t* __restrict__ out) {
    __shared__ int64_t warp_sum[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = int64_t(blockIdx.x) * kScanTile + int64_t(warp) * kScanPerWarp;
    constexpr int kIters = kScanPerWarp / 32;
    int64_t excl[kIters];
    int64_t carry = 0;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
        int64_t idx = base + it * 32 + lane;
        int64_t v = idx < n ? in[idx] : 0;
        int64_t inc = warp_incl_scan(v, lane);
        excl[it] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) warp_sum[warp] = carry;
    __syncthreads();
    int64_t off = tile_offset[blockIdx.x];
    for (int w = 0; w < warp; ++w) off += warp_sum[w];
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
        int64_t idx = base + it * 32 + lane;
        if (idx < n) out[idx] = off + excl[it];
    }
}

sdb_status exclusive_scan_i32_to_i64(cudaStream_t s, const int32_t* in, int64_t* out, int64_t n) {
    if (n <= 0) {
        SDB_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), s));
        return SDB_STATUS_SUCCESS;
    }
    const int64_t tiles = (n + kScanTile - 1) / kScanTile;
    DevBuf totals;
    SDB_TRY(totals.alloc(size_t(tiles) * sizeof(int64_t), s));
    SDB_LAUNCH(scan_tile_totals, unsigned(tiles), kScanThreads, 0, s, in, n, totals.as<int64_t>());
    SDB_LAUNCH(scan_totals_inplace, 1, 1024, 0, s, totals.as<int64_t>(), tiles, out + n);
    SDB_LAUNCH(scan_tile_apply, unsigned(tiles), kScanThreads, 0, s, in, n, totals.as<int64_t>(), out);
    return SDB_STATUS_SUCCESS;
}

// ------------------------------------------------------------ index width
template <typename S, typename D>
__global__ void convert_index_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (; i < n; i += stride) dst[i] = D(src[i]);
}

static unsigned grid_for(int64_t n, int block, int cap = 148 * 16) {
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return unsigned(g);
}

sdb_status widen_i32_to_i64(cudaStream_t s, const int32_t* src, int64_t* dst, int64_t n) {
    if (n <= 0) return SDB_STATUS_SUCCESS;
    SDB_LAUNCH((convert_index_kernel<int32_t, int64_t>), grid_for(n, 256), 256, 0, s, src, dst, n);
    return SDB_STATUS_SUCCESS;
}

sdb_status narrow_i64_to_i32(cudaStream_t s, const int64_t* src, int32_t* dst, int64_t n) {
    if (n <= 0) return SDB_STATUS_SUCCESS;
    SDB_LAUNCH((convert_index_kernel<int64_t, int32_t>), grid_for(n, 256), 256, 0, s, src, dst, n);
    return SDB_STATUS_SUCCESS;
}

// ------------------------------------------------------- dense transpose
// 32x32 tile through padded shared memory; both the read and the write are
// coalesced.  E is an opaque element of 4, 8 or 16 bytes.
template <typename E>
__global__ void __launch_bounds__(256) transpose_tile_kernel(const E* __restrict__ src, int64_t ld_src,
                                                             E* __restrict__ dst, int64_t ld_dst,
                                                             int64_t rows, int64_t cols,
                                                             int64_t tiles_c) {
    __shared__ E tile[32][33];
    const int64_t tile_r = int64_t(blockIdx.x) / tiles_c, tile_c = int64_t(blockIdx.x) % tiles_c;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t r0 = tile_r * 32, c0 = tile_c * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        int64_t r = r0 + ty + j, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + j][tx] = src[r * ld_src + c];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        int64_t c = c0 + ty + j, r = r0 + tx;
        if (r < rows && c < cols) dst[c * ld_dst + r] = tile[tx][ty + j];
    }
}

template <typename E>
static sdb_status transpose_typed(cudaStream_t s, const void* src, int64_t ld_src, void* dst,
                                  int64_t ld_dst, int64_t rows, int64_t cols) {
    const int64_t tr = (rows + 31) / 32, tc = (cols + 31) / 32;
    SDB_REQUIRE(tr * tc < (int64_t(1) << 31), SDB_STATUS_NOT_SUPPORTED, "dense panel too large to transpose");
    SDB_LAUNCH(transpose_tile_kernel<E>, unsigned(tr * tc), 256, 0, s, static_cast<const E*>(src), ld_src,
               static_cast<E*>(dst), ld_dst, rows, cols, tc);
    return SDB_STATUS_SUCCESS;
}

sdb_status transpose_dense(cudaStream_t s, const void* src, int64_t ld_src, void* dst, int64_t ld_dst,
                           int64_t rows, int64_t cols, int elem_bytes) {
    if (rows <= 0 || cols <= 0) return SDB_STATUS_SUCCESS;
    switch (elem_bytes) {
        case 4: return transpose_typed<uint32_t>(s, src, ld_src, dst, ld_dst, rows, cols);
        case 8: return transpose_typed<uint64_t>(s, src, ld_src, dst, ld_dst, rows, cols);
        case 16: return transpose_typed<uint4>(s, src, ld_src, dst, ld_dst, rows, cols);
        default:
            set_error("transpose_dense: element size %d", elem_bytes);
            return SDB_STATUS_INTERNAL_ERROR;
    }
}

sdb_status fill_zero(cudaStream_t s, void* p, size_t bytes) {
    if (bytes) SDB_CUDA(cudaMemsetAsync(p, 0, bytes, s));
    return SDB_STATUS_SUCCESS;
}

}  // namespace sdb
