// probe.cu — bandwidth probes measured ON THE BOX the numbers are quoted on, so that bench.py's
// roofline denominators are measurements, not figures from a guide:
//   kind 0  HBM read        sequential 16-byte loads over a buffer much larger than L2
//   kind 1  L2 -> SM read   the same loop over a buffer that stays L2-resident (L1 bypassed)
//   kind 2  L2 -> SM gather every warp reads whole 512-byte rows at random positions of an
//                           L2-resident buffer, one 16-byte read-only load per lane — the access
//                           shape of the SpMM gathers (spmm.cu / spmm_slab.cu), whose ceiling it is
// No reference counterpart (the reference has no device); test/bench infrastructure of the product.
#include <algorithm>
#include <chrono>
#include <cstdlib>

#include "common.h"

namespace sdb {
namespace {

constexpr int kProbeThreads = 256;
constexpr int kProbeUnroll = 8;

__device__ __forceinline__ uint32_t fold(const uint4& v) { return v.x ^ v.y ^ v.z ^ v.w; }

// sequential: the grid walks the buffer `passes` times, kProbeUnroll independent loads per thread in flight
__global__ void __launch_bounds__(kProbeThreads) probe_stream_kernel(const uint4* __restrict__ buf, size_t n_vec,
                                                                     int passes, uint32_t* __restrict__ sink) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    uint32_t acc = 0;
    for (int p = 0; p < passes; ++p) {
        size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
        for (; i + (kProbeUnroll - 1) * stride < n_vec; i += kProbeUnroll * stride) {
            uint4 v[kProbeUnroll];
#pragma unroll
            for (int u = 0; u < kProbeUnroll; ++u) v[u] = __ldcg(buf + i + u * stride);
#pragma unroll
            for (int u = 0; u < kProbeUnroll; ++u) acc ^= fold(v[u]);
        }
        for (; i < n_vec; i += stride) acc ^= fold(__ldcg(buf + i));
    }
    if (acc == 0x9e3779b9u) *sink = acc;  // keeps the loads alive; practically never true
}

// gather: a warp reads `rows_per_warp` rows of 512 bytes at pseudo-random positions, kProbeUnroll in flight
__global__ void __launch_bounds__(kProbeThreads) probe_gather_kernel(const char* __restrict__ buf, uint32_t n_rows,
                                                                     int rows_per_warp, uint32_t* __restrict__ sink) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t state = warp * 0x9E3779B1u + 0x7F4A7C15u;
    uint32_t acc = 0;
    const char* lane_base = buf + lane * 16;
    for (int r = 0; r < rows_per_warp; r += kProbeUnroll) {
        uint4 v[kProbeUnroll];
#pragma unroll
        for (int u = 0; u < kProbeUnroll; ++u) {
            state = state * 1664525u + 1013904223u;  // warp-uniform LCG
            const uint32_t row = uint32_t((uint64_t(state) * n_rows) >> 32);
            v[u] = __ldg(reinterpret_cast<const uint4*>(lane_base + size_t(row) * 512));
        }
#pragma unroll
        for (int u = 0; u < kProbeUnroll; ++u) acc ^= fold(v[u]);
    }
    if (acc == 0x9e3779b9u) *sink = acc;
}

}  // namespace
}  // namespace sdb

using namespace sdb;

// kind 3: the host side of the pageable-memory pipeline — pageable -> page-locked copies in 16 MiB slots by
// the library's copy threads (runtime.cu), no GPU involved (falls back to plain memory without a device)
static sdb_status probe_host_copy(int64_t bytes, int iters, double* gbs) {
    constexpr size_t kSlot = size_t(16) << 20;
    char* src = static_cast<char*>(malloc(size_t(bytes)));
    SDB_REQUIRE(src != nullptr, SDB_STATUS_ALLOC_FAILED, "probe: out of host memory");
    memset(src, 1, size_t(bytes));
    void* slots[4] = {nullptr, nullptr, nullptr, nullptr};
    bool pinned = true;
    for (auto& p : slots)
        if (cudaHostAlloc(&p, kSlot, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            pinned = false;
            p = malloc(kSlot);
            memset(p, 0, kSlot);
        }
    host_copy(slots[0], src, kSlot);  // starts the pool
    const auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < iters; ++it) {
        int k = 0;
        for (size_t off = 0; off < size_t(bytes); off += kSlot, k = (k + 1) & 3)
            host_copy(slots[k], src + off, std::min(kSlot, size_t(bytes) - off));
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *gbs = double(iters) * double(bytes) / sec / 1e9;
    // the copies are checked too: an odd-sized, misaligned piece and the last full slot
    bool ok = true;
    {
        const size_t odd = (size_t(5) << 20) + 12345;
        for (size_t i = 0; i < odd; ++i) src[7 + i] = char(i * 131u + (i >> 9));
        host_copy(static_cast<char*>(slots[1]) + 3, src + 7, odd);
        ok = memcmp(static_cast<char*>(slots[1]) + 3, src + 7, odd) == 0;
    }
    if (ok && size_t(bytes) >= (size_t(8) << 20)) {
        // strided panels (host_copy_2d): 997 rows of 5003 bytes, pitch 8191 -> packed, and back out to pitch 6007
        const size_t rows = 997, rb = 5003, sp = 8191, dp = 6007;
        char* packed = static_cast<char*>(slots[2]);
        char* spread = static_cast<char*>(slots[3]);
        memset(spread, 0x33, rows * dp);
        host_copy_2d(packed, rb, src + 11, sp, rb, rows);
        host_copy_2d(spread + 5, dp, packed, rb, rb, rows);
        for (size_t r = 0; r < rows && ok; ++r) {
            ok = memcmp(packed + r * rb, src + 11 + r * sp, rb) == 0 && memcmp(spread + 5 + r * dp, src + 11 + r * sp, rb) == 0;
            for (size_t g = rb; g < dp && ok && r + 1 < rows; ++g) ok = spread[5 + r * dp + g] == 0x33;  // gaps untouched
        }
    }
    for (auto& p : slots) {
        if (pinned) cudaFreeHost(p);
        else free(p);
    }
    free(src);
    SDB_REQUIRE(ok, SDB_STATUS_INTERNAL_ERROR, "probe: the copy pool produced a wrong copy");
    return SDB_STATUS_SUCCESS;
}

extern "C" sdb_status sdb_probe_bandwidth(int kind, int64_t bytes, int iters, double* gbs) {
    SDB_REQUIRE(gbs != nullptr, SDB_STATUS_INVALID_VALUE, "probe: null output");
    SDB_REQUIRE(kind >= 0 && kind <= 3, SDB_STATUS_INVALID_VALUE, "probe: kind must be 0..3");
    SDB_REQUIRE(bytes >= (int64_t(1) << 20) && iters >= 1, SDB_STATUS_INVALID_VALUE, "probe: bad size / iterations");
    *gbs = 0.0;
    if (kind == 3) return probe_host_copy(bytes, iters, gbs);
    Context* ctx;
    SDB_TRY(get_context(&ctx));
    cudaStream_t s = ctx->stream;
    bytes &= ~int64_t(511);
    DevBuf buf, sink;
    SDB_TRY(buf.alloc(size_t(bytes), s));
    SDB_TRY(sink.alloc(16, s));
    SDB_CUDA(cudaMemsetAsync(buf.p, 0x5a, size_t(bytes), s));
    cudaEvent_t e0, e1;
    SDB_CUDA(cudaEventCreate(&e0));
    SDB_CUDA(cudaEventCreate(&e1));
    const unsigned grid = unsigned(ctx->sm_count) * (2048 / kProbeThreads);
    double moved = 0.0;
    sdb_status st = [&]() -> sdb_status {
        if (kind == 2) {
            const uint32_t n_rows = uint32_t(bytes / 512);
            const int rows_per_warp = 4096;
            // warm-up launch brings the buffer into L2
            SDB_LAUNCH(probe_gather_kernel, grid, kProbeThreads, 0, s, static_cast<const char*>(buf.p), n_rows,
                       rows_per_warp, sink.as<uint32_t>());
            SDB_CUDA(cudaEventRecord(e0, s));
            for (int i = 0; i < iters; ++i)
                SDB_LAUNCH(probe_gather_kernel, grid, kProbeThreads, 0, s, static_cast<const char*>(buf.p), n_rows,
                           rows_per_warp, sink.as<uint32_t>());
            SDB_CUDA(cudaEventRecord(e1, s));
            moved = double(iters) * double(grid) * (kProbeThreads / 32) * rows_per_warp * 512.0;
        } else {
            const size_t n_vec = size_t(bytes) / 16;
            // L2 probe: several passes inside one launch so that launch overhead does not count
            const int passes = kind == 1 ? 64 : 1;
            SDB_LAUNCH(probe_stream_kernel, grid, kProbeThreads, 0, s, static_cast<const uint4*>(buf.p), n_vec, 1,
                       sink.as<uint32_t>());
            SDB_CUDA(cudaEventRecord(e0, s));
            for (int i = 0; i < iters; ++i)
                SDB_LAUNCH(probe_stream_kernel, grid, kProbeThreads, 0, s, static_cast<const uint4*>(buf.p), n_vec,
                           passes, sink.as<uint32_t>());
            SDB_CUDA(cudaEventRecord(e1, s));
            moved = double(iters) * double(passes) * double(n_vec) * 16.0;
        }
        SDB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SDB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms > 0.f) *gbs = moved / (double(ms) * 1e-3) / 1e9;
        return SDB_STATUS_SUCCESS;
    }();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return st;
}
