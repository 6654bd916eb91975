// spmv_tile.cu — SpMV with x staged in shared memory (inspector + executor), for handles that are multiplied
// by a vector again and again (the CG / FGMRES callers of mkl_sparse_?_mv, _sparse_vector.py:84-92).
//
// Why: y = A x with columns spread over a long x is not bound by bytes on this machine.  Every stored entry
// gathers one element of x, i.e. one 32-byte sector through the L1 / L2 request path: spmv_wide_kernel spends
// 1.57 cycles per entry and SM there and reaches 0.23 of the HBM roofline on BASELINE's configs[1] matrix
// (DESIGN.md, K1v).  Shared memory serves 32 scattered 4-byte reads in a few cycles.
//
// Inspector (once per handle, cached like the SpMM slab copy): the matrix is cut into row blocks of equal WORK
// (stored entries + mean row length per row, so neither a block of long rows nor a block of many short rows is
// heavier than twice the mean; at most 128 KiB of partial sums each) and column slabs of S columns
// (S * sizeof(T) = 64 KiB of x); every row block splits its slabs into groups of equal entry counts.  The entries
// of a tile (row block x slab) are
// stored together as (local row << 15 | local column, value), padded to a multiple of four so that every tile
// is read with 16-byte loads; inside a tile consecutive entries belong to different rows (lane-per-row scatter),
// so the lanes of a warp rarely add into the same partial sum.
//
// Executor: one CTA per (row block, group of column slabs), sm_count of them in one wave.  For each of its
// slabs the CTA loads the x slab into shared memory, streams the tile's entries from HBM (coalesced, each read
// once) and accumulates value * x into the row's partial sum in shared memory (atomic adds on shared memory:
// a compare-and-swap loop in SASS, cheap while lanes hit different addresses).  At the end the partial sums are
// added into y with global reductions (y was scaled by beta beforehand); with one slab group per row block
// that is a plain coalesced update.
//
// Accumulation order inside a row is not fixed (atomics), so results can differ in the last bits from call to
// call, within the tolerance the parity tests state — like a threaded CPU SpMV with a dynamic schedule.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.h"
#include "prims.h"
#include "types.cuh"

namespace sdb {
namespace {

constexpr int kTileColBits = 15;  // local column: S <= 32768
constexpr int kTileThreads = 1024;
constexpr size_t kTileXBytes = size_t(64) << 10;
constexpr size_t kTileYBytes = size_t(128) << 10;
constexpr int64_t kTileMaxSlabs = 8192;  // shared-memory counters of the inspector

// ------------------------------------------------------------------ inspector
// A row block is cut into sub-blocks of kTileThreads rows, one CTA each, one row per thread.
// sub_counts[(rb * n_slabs + s) * subs + sub] = entries of that sub-block inside column slab s.
__global__ void __launch_bounds__(kTileThreads) tile_count_kernel(const int64_t* __restrict__ rb_start, int subs,
                                                                 int slab_cols, int n_slabs,
                                                                 const int64_t* __restrict__ indptr,
                                                                 const int32_t* __restrict__ indices,
                                                                 int32_t* __restrict__ sub_counts) {
    extern __shared__ int32_t cnt[];
    for (int i = threadIdx.x; i < n_slabs; i += kTileThreads) cnt[i] = 0;
    __syncthreads();
    const int64_t rb = blockIdx.x / subs;
    const int sub = blockIdx.x % subs;
    const int64_t local = int64_t(sub) * kTileThreads + threadIdx.x;
    const int64_t r = rb_start[rb] + local;
    if (r < rb_start[rb + 1])
        for (int64_t p = indptr[r], e = indptr[r + 1]; p < e; ++p) atomicAdd(&cnt[__ldg(indices + p) / slab_cols], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < n_slabs; i += kTileThreads)
        sub_counts[(rb * n_slabs + i) * subs + sub] = cnt[i];
}

// every tile padded to a multiple of four entries: the padding goes to the tile's last sub-block
__global__ void tile_pad_kernel(int64_t n_tiles, int subs, int32_t* __restrict__ sub_counts) {
    const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    int32_t sum = 0;
    for (int i = 0; i < subs; ++i) sum += sub_counts[t * subs + i];
    sub_counts[t * subs + subs - 1] += (-sum) & 3;
}

__global__ void tile_ptr_kernel(int64_t n_tiles, int subs, const int64_t* __restrict__ sub_base,
                                int64_t* __restrict__ tile_ptr) {
    const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t <= n_tiles) tile_ptr[t] = sub_base[t * subs];
}

// Row-block boundaries of equal work: W(r) = entries before row r + r * mean row length; block b starts at the first
// row with W(r) >= b * W(rows) / n_rb.
__global__ void tile_bounds_kernel(int64_t rows, int64_t n_rb, const int64_t* __restrict__ indptr,
                                   int64_t* __restrict__ rb_start) {
    const int64_t b = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (b > n_rb) return;
    if (b == n_rb) {
        rb_start[b] = rows;
        return;
    }
    const double nnz = double(indptr[rows] - indptr[0]);
    const double mean = nnz / double(rows);
    const double target = 2.0 * nnz * double(b) / double(n_rb);
    int64_t lo = 0, hi = rows;  // smallest r with W(r) >= target
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (double(indptr[mid] - indptr[0]) + double(mid) * mean >= target) hi = mid;
        else lo = mid + 1;
    }
    rb_start[b] = lo;
}

// One row per thread: at every step a warp places one entry of 32 different rows, so neighbours inside a tile
// belong to different rows.  Padding slots keep (0, 0.0) from the memset.
template <typename T>
__global__ void __launch_bounds__(kTileThreads) tile_scatter_kernel(const int64_t* __restrict__ rb_start, int subs,
                                                                   int slab_cols, int n_slabs,
                                                                   const int64_t* __restrict__ indptr,
                                                                   const int32_t* __restrict__ indices,
                                                                   const T* __restrict__ values,
                                                                   const int64_t* __restrict__ sub_base,
                                                                   uint32_t* __restrict__ rc, T* __restrict__ val) {
    extern __shared__ int32_t cur[];
    for (int i = threadIdx.x; i < n_slabs; i += kTileThreads) cur[i] = 0;
    __syncthreads();
    const int64_t rb = blockIdx.x / subs;
    const int sub = blockIdx.x % subs;
    const int64_t local = int64_t(sub) * kTileThreads + threadIdx.x;
    const int64_t r = rb_start[rb] + local;
    if (r >= rb_start[rb + 1]) return;
    const uint32_t tag = uint32_t(local) << kTileColBits;
    for (int64_t p = indptr[r], e = indptr[r + 1]; p < e; ++p) {
        const int32_t c = __ldg(indices + p);
        const int s = c / slab_cols;
        const int64_t at = sub_base[(rb * n_slabs + s) * subs + sub] + atomicAdd(&cur[s], 1);
        rc[at] = tag | uint32_t(c - s * slab_cols);
        val[at] = values[p];
    }
}

// ------------------------------------------------------------------ executor
__device__ __forceinline__ void tile_load4(const float* p, float (&v)[4]) {
    const float4 q = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
}
__device__ __forceinline__ void tile_load4(const double* p, double (&v)[4]) {
    const double2 q0 = __ldcs(reinterpret_cast<const double2*>(p)), q1 = __ldcs(reinterpret_cast<const double2*>(p) + 1);
    v[0] = q0.x, v[1] = q0.y, v[2] = q1.x, v[3] = q1.y;
}

template <typename T>
__global__ void __launch_bounds__(kTileThreads, 1) spmv_tile_kernel(int64_t cols, const int64_t* __restrict__ rb_start,
                                                                  const int32_t* __restrict__ grp_start,
                                                                  int slab_cols, int n_slabs, int n_groups,
                                                                  const int64_t* __restrict__ tile_ptr,
                                                                  const uint32_t* __restrict__ rc,
                                                                  const T* __restrict__ val, const T* __restrict__ x,
                                                                  T alpha, T* __restrict__ y) {
    extern __shared__ __align__(16) unsigned char tile_smem[];
    T* xs = reinterpret_cast<T*>(tile_smem);
    T* ys = xs + slab_cols;
    const int tid = threadIdx.x;
    const int64_t rb = blockIdx.x / n_groups;
    const int g = blockIdx.x % n_groups;
    const int64_t row0 = rb_start[rb];
    const int nr = int(rb_start[rb + 1] - row0);
    for (int i = tid; i < nr; i += kTileThreads) ys[i] = T(0);
    const int s0 = grp_start[rb * (n_groups + 1) + g], s1 = grp_start[rb * (n_groups + 1) + g + 1];
    constexpr uint32_t kMask = (1u << kTileColBits) - 1;
    bool touched = false;
    for (int s = s0; s < s1; ++s) {
        const int64_t b = tile_ptr[rb * n_slabs + s], e = tile_ptr[rb * n_slabs + s + 1];
        if (b == e) continue;  // the same for every thread of the CTA
        touched = true;
        // first pack of this tile in flight while the x slab is loaded
        int64_t p = b + 4 * tid;
        bool have = p < e;
        uint4 w = make_uint4(0, 0, 0, 0);
        T v[4] = {T(0), T(0), T(0), T(0)};
        if (have) {
            w = __ldcs(reinterpret_cast<const uint4*>(rc + p));
            tile_load4(val + p, v);
        }
        __syncthreads();  // the previous tile has finished reading xs (and ys is zeroed)
        const int64_t c0 = int64_t(s) * slab_cols;
        const int nc = int(min(int64_t(slab_cols), cols - c0));
        for (int i = tid; i < nc; i += kTileThreads) xs[i] = ldg(x + c0 + i);
        __syncthreads();
        while (have) {
            const int64_t pn = p + 4 * kTileThreads;
            const bool hn = pn < e;
            uint4 wn = make_uint4(0, 0, 0, 0);
            T vn[4] = {T(0), T(0), T(0), T(0)};
            if (hn) {
                wn = __ldcs(reinterpret_cast<const uint4*>(rc + pn));
                tile_load4(val + pn, vn);
            }
            // atomicAdd on shared memory = ATOMS.CAST.SPIN loops in SASS.  Running the four compare-and-swap chains
            // of a pack side by side by hand (plain ATOMS.CAS) measured slower: 0.211 vs 0.146 ms on configs[1].
            atomicAdd(&ys[w.x >> kTileColBits], v[0] * xs[w.x & kMask]);
            atomicAdd(&ys[w.y >> kTileColBits], v[1] * xs[w.y & kMask]);
            atomicAdd(&ys[w.z >> kTileColBits], v[2] * xs[w.z & kMask]);
            atomicAdd(&ys[w.w >> kTileColBits], v[3] * xs[w.w & kMask]);
            w = wn;
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = vn[i];
            p = pn;
            have = hn;
        }
    }
    __syncthreads();
    if (touched)
        for (int i = tid; i < nr; i += kTileThreads) atomicAdd(y + row0 + i, alpha * ys[i]);
}

// y = beta * y (beta == 0 overwrites: NaNs in y do not survive, as with MKL)
template <typename T> __global__ void scale_vector_kernel(int64_t n, T beta, T* __restrict__ y) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) y[i] = beta == T(0) ? T(0) : beta * y[i];
}

// Everything the executor needs, cached on the handle (sdb_mat::vt_cache)
struct TileCache {
    int64_t n_rb = 0, max_rb_rows = 0, entries = 0;
    int slab_cols = 0, n_slabs = 0, n_groups = 1;
    int64_t* rb_start = nullptr;   // [n_rb + 1]
    int32_t* grp_start = nullptr;  // [n_rb * (n_groups + 1)] first slab of every group
    int64_t* tile_ptr = nullptr;   // [n_rb * n_slabs + 1]
    void* rc = nullptr;
    void* val = nullptr;
    void release(cudaStream_t s) {
        for (void* p : {static_cast<void*>(rb_start), static_cast<void*>(grp_start), static_cast<void*>(tile_ptr), rc, val})
            if (p) cudaFreeAsync(p, s);
        rb_start = nullptr, grp_start = nullptr, tile_ptr = nullptr, rc = val = nullptr;
    }
};

int slab_cols_for(size_t es) {
    // SDB_SPMV_TILE_XKB: KiB of x per slab (sweeps; 16..96, default 64)
    static const size_t x_bytes = [] {
        const char* e = getenv("SDB_SPMV_TILE_XKB");
        const int kb = e ? atoi(e) : 0;
        return kb >= 16 && kb <= 96 ? size_t(kb) << 10 : kTileXBytes;
    }();
    return int(x_bytes / es);
}

sdb_status read_back(cudaStream_t s, void* dst, const void* src, size_t bytes, const char* what) {
    if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
        set_error("spmv tiles: reading %s failed: %s", what, cudaGetErrorString(cudaGetLastError()));
        return SDB_STATUS_EXECUTION_FAILED;
    }
    return SDB_STATUS_SUCCESS;
}

// Row blocks: one wave of (row blocks) x (slab groups) = sm_count CTAs with as few row blocks as the partial sums
// allow (every row block reads all of x once); if even sm_count blocks are too tall, several waves.
sdb_status plan_row_blocks(Context* ctx, const CsrView& a, size_t es, TileCache* tc) {
    cudaStream_t ls = ctx->stream;
    const int64_t rb_max = int64_t(kTileYBytes / es);
    const int sm = ctx->sm_count;
    std::vector<std::pair<int64_t, int>> tries;  // (row blocks, slab groups)
    for (int groups : {4, 2, 1})
        if (sm % groups == 0 && groups <= tc->n_slabs) tries.push_back({int64_t(sm / groups), groups});
    for (int64_t waves = 2; waves <= 64; waves *= 2) tries.push_back({int64_t(sm) * waves, 1});
    for (const auto& t : tries) {
        const int64_t n_rb = std::min<int64_t>(t.first, a.rows);
        if ((a.rows + n_rb - 1) / n_rb > rb_max) continue;  // even equal-height blocks would not fit
        DevBuf bounds;
        SDB_TRY(bounds.alloc(size_t(n_rb + 1) * 8, ls));
        SDB_LAUNCH(tile_bounds_kernel, unsigned((n_rb + 256) / 256), 256, 0, ls, a.rows, n_rb, a.indptr,
                   bounds.as<int64_t>());
        std::vector<int64_t> host(size_t(n_rb) + 1);
        SDB_TRY(read_back(ls, host.data(), bounds.p, host.size() * 8, "the row-block boundaries"));
        int64_t tallest = 0;
        for (int64_t b = 0; b < n_rb; ++b) tallest = std::max(tallest, host[size_t(b) + 1] - host[size_t(b)]);
        if (tallest > rb_max) continue;
        tc->n_rb = n_rb;
        tc->n_groups = t.second;
        tc->max_rb_rows = tallest;
        tc->rb_start = static_cast<int64_t*>(bounds.release());
        return SDB_STATUS_SUCCESS;
    }
    return SDB_STATUS_NOT_SUPPORTED;
}

template <typename T>
sdb_status build_tiles(Context* ctx, const CsrView& a, TileCache* tc, bool check_balance) {
    cudaStream_t ls = ctx->stream;
    tc->slab_cols = slab_cols_for(sizeof(T));
    tc->n_slabs = int((a.cols + tc->slab_cols - 1) / tc->slab_cols);
    SDB_TRY(plan_row_blocks(ctx, a, sizeof(T), tc));
    const int64_t n_tiles = tc->n_rb * tc->n_slabs;
    // Every row block streams all of x through shared memory, slab by slab: that only pays while a tile holds at
    // least as many bytes of entries as the slab of x it needs (R-MAT scale 22, edge factor 4: 219 entries per
    // 64 KiB slab, and the kernel would spend its time loading x — measured 1.41 ms against 0.05 ms of entries).
    if (check_balance && double(a.nnz) * double(4 + sizeof(T)) < double(n_tiles) * double(tc->slab_cols) * sizeof(T))
        return SDB_STATUS_NOT_SUPPORTED;
    const int subs = int(std::max<int64_t>(1, (tc->max_rb_rows + kTileThreads - 1) / kTileThreads));
    const int64_t n_sub = n_tiles * subs;
    DevBuf counts, sub_base;
    SDB_TRY(counts.alloc(size_t(n_sub) * 4, ls));
    SDB_TRY(sub_base.alloc(size_t(n_sub + 1) * 8, ls));
    SDB_TRY(dev_alloc(reinterpret_cast<void**>(&tc->tile_ptr), size_t(n_tiles + 1) * 8, ls));
    const size_t smem = size_t(tc->n_slabs) * 4;
    const unsigned igrid = unsigned(tc->n_rb * subs);
    SDB_LAUNCH(tile_count_kernel, igrid, kTileThreads, smem, ls, tc->rb_start, subs, tc->slab_cols, tc->n_slabs,
               a.indptr, a.indices, counts.as<int32_t>());
    SDB_LAUNCH(tile_pad_kernel, unsigned((n_tiles + 255) / 256), 256, 0, ls, n_tiles, subs, counts.as<int32_t>());
    SDB_TRY(exclusive_scan_i32_to_i64(ls, counts.as<int32_t>(), sub_base.as<int64_t>(), n_sub));
    SDB_LAUNCH(tile_ptr_kernel, unsigned((n_tiles + 256) / 256), 256, 0, ls, n_tiles, subs, sub_base.as<int64_t>(),
               tc->tile_ptr);
    std::vector<int64_t> host(size_t(n_tiles) + 1);
    SDB_TRY(read_back(ls, host.data(), tc->tile_ptr, host.size() * 8, "the tile offsets"));
    tc->entries = host[size_t(n_tiles)];
    // slab groups of equal entry counts inside every row block; one CTA per (row block, group)
    const int G = tc->n_groups;
    std::vector<int32_t> grp(size_t(tc->n_rb) * size_t(G + 1));
    int64_t worst = 0;
    for (int64_t rb = 0; rb < tc->n_rb; ++rb) {
        const int64_t* row = host.data() + rb * tc->n_slabs;
        const int64_t base = row[0], tot = row[tc->n_slabs] - base;
        int32_t* out = grp.data() + rb * (G + 1);
        out[0] = 0;
        for (int g = 1; g < G; ++g) {
            const int64_t target = base + tot * g / G;
            out[g] = std::max<int32_t>(out[g - 1], int32_t(std::lower_bound(row, row + tc->n_slabs + 1, target) - row));
        }
        out[G] = tc->n_slabs;
        for (int g = 0; g < G; ++g) worst = std::max(worst, row[out[g + 1]] - row[out[g]]);
    }
    // the slowest CTA sets the time: decline when one item holds more than three times its share (a row is never
    // split, so a single giant row does that)
    if (check_balance && double(worst) > 3.0 * double(tc->entries) / double(tc->n_rb * G)) return SDB_STATUS_NOT_SUPPORTED;
    SDB_TRY(dev_alloc(reinterpret_cast<void**>(&tc->grp_start), grp.size() * 4, ls));
    SDB_CUDA(cudaMemcpyAsync(tc->grp_start, grp.data(), grp.size() * 4, cudaMemcpyHostToDevice, ls));
    const size_t slots = size_t(std::max<int64_t>(tc->entries, 4));
    SDB_TRY(dev_alloc(&tc->rc, slots * 4, ls));
    SDB_TRY(dev_alloc(&tc->val, slots * sizeof(T), ls));
    SDB_CUDA(cudaMemsetAsync(tc->rc, 0, slots * 4, ls));
    SDB_CUDA(cudaMemsetAsync(tc->val, 0, slots * sizeof(T), ls));
    SDB_LAUNCH((tile_scatter_kernel<T>), igrid, kTileThreads, smem, ls, tc->rb_start, subs, tc->slab_cols, tc->n_slabs,
               a.indptr, a.indices, static_cast<const T*>(a.values), sub_base.as<int64_t>(),
               static_cast<uint32_t*>(tc->rc), static_cast<T*>(tc->val));
    SDB_CUDA(cudaStreamSynchronize(ls));  // grp (host vector) must outlive its upload
    trace(ls, "spmv tiles: %lld row blocks (tallest %lld rows) x %d column slabs in %d groups, %lld entries (%lld stored)",
          (long long)tc->n_rb, (long long)tc->max_rb_rows, tc->n_slabs, G, (long long)tc->entries, (long long)a.nnz);
    return SDB_STATUS_SUCCESS;
}

}  // namespace

void drop_spmv_tiles(sdb_mat* m, cudaStream_t s) {
    if (m->vt_cache) {
        TileCache* tc = static_cast<TileCache*>(m->vt_cache);
        tc->release(s);
        delete tc;
    }
    m->vt_cache = nullptr;
    m->vt_state = 0;
    m->spmv_calls = 0;
    // the long-row list of the gather SpMV (spmm.cu) depends on the same structure
    for (int slot = 0; slot < 2; ++slot) {
        if (m->long_rows[slot]) cudaFreeAsync(m->long_rows[slot], s);
        m->long_rows[slot] = nullptr;
        m->n_long[slot] = 0;
        m->long_state[slot] = 0;
    }
}

// Policy ("spmv_tile": 0 automatic, 1 never, 2 whenever the shape allows).  Automatic: the handle has been
// multiplied by a vector before (a matrix used once never pays the inspector), x is far larger than an L1, the
// matrix is large enough to fill the machine, and the inspector could balance it.
bool spmv_tile_wanted(const CsrView& a, int dtype, int64_t incx, int64_t incy) {
    const int mode = get_option(kOptSpmvTile);
    if (mode == 1 || a.owner == nullptr || !a.owner->owns) return false;
    if (dtype != SDB_F32 && dtype != SDB_F64) return false;
    if (incx != 1 || incy != 1 || a.sub_rows >= 0) return false;
    if (a.rows <= 0 || a.nnz <= 0 || a.cols <= 0) return false;
    const size_t es = dtype_size(dtype);
    if (a.cols / int64_t((size_t(16) << 10) / es) >= kTileMaxSlabs) return false;
    if (a.owner->vt_state == -1) return false;
    if (mode == 2) return true;
    const int uses = a.owner->spmv_calls++;
    return uses >= 1 && a.nnz >= (int64_t(1) << 22) && size_t(a.cols) * es >= (size_t(1) << 20) &&
           a.rows >= (int64_t(1) << 15);
}

sdb_status spmv_tile_device(Context* ctx, cudaStream_t s, const CsrView& a, int dtype, const double* alpha,
                            const double* beta, const void* dX, void* dY) {
    sdb_mat* m = a.owner;
    SDB_REQUIRE(m != nullptr, SDB_STATUS_NOT_SUPPORTED, "spmv tiles: ad-hoc view");
    const size_t es = dtype_size(dtype);
    TileCache* tc = nullptr;
    {
        std::lock_guard<std::mutex> cache_lock(g_companion_mutex);  // the tiles are a per-handle cache
        if (m->vt_state == -1) return SDB_STATUS_NOT_SUPPORTED;
        if (m->vt_state != 1) {
            drop_spmv_tiles(m, ctx->stream);
            m->spmv_calls = 2;
            tc = new TileCache();
            const bool check = get_option(kOptSpmvTile) != 2;
            const sdb_status st = dtype == SDB_F32 ? build_tiles<float>(ctx, a, tc, check)
                                                   : build_tiles<double>(ctx, a, tc, check);
            if (st != SDB_STATUS_SUCCESS) {
                tc->release(ctx->stream);
                delete tc;
                if (st == SDB_STATUS_NOT_SUPPORTED) m->vt_state = -1;  // the gather kernel serves this matrix
                return st;
            }
            m->vt_cache = tc;
            m->vt_state = 1;
        }
        tc = static_cast<TileCache*>(m->vt_cache);
    }
    const size_t smem = size_t(tc->slab_cols) * es + size_t(std::max<int64_t>(tc->max_rb_rows, 1)) * es;
    const unsigned grid = unsigned(tc->n_rb * tc->n_groups);
    const unsigned sgrid = unsigned((a.rows + 255) / 256);
    if (dtype == SDB_F32) {
        SDB_LAUNCH(scale_vector_kernel<float>, sgrid, 256, 0, s, a.rows, float(beta[0]), static_cast<float*>(dY));
        SDB_CUDA(cudaFuncSetAttribute(spmv_tile_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        SDB_LAUNCH(spmv_tile_kernel<float>, grid, kTileThreads, smem, s, a.cols, tc->rb_start, tc->grp_start,
                   tc->slab_cols, tc->n_slabs, tc->n_groups, tc->tile_ptr, static_cast<const uint32_t*>(tc->rc),
                   static_cast<const float*>(tc->val), static_cast<const float*>(dX), float(alpha[0]),
                   static_cast<float*>(dY));
    } else {
        SDB_LAUNCH(scale_vector_kernel<double>, sgrid, 256, 0, s, a.rows, beta[0], static_cast<double*>(dY));
        SDB_CUDA(cudaFuncSetAttribute(spmv_tile_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        SDB_LAUNCH(spmv_tile_kernel<double>, grid, kTileThreads, smem, s, a.cols, tc->rb_start, tc->grp_start,
                   tc->slab_cols, tc->n_slabs, tc->n_groups, tc->tile_ptr, static_cast<const uint32_t*>(tc->rc),
                   static_cast<const double*>(tc->val), static_cast<const double*>(dX), alpha[0],
                   static_cast<double*>(dY));
    }
    note_spmm_kernel("spmv_tile_kernel<%s>", dtype == SDB_F32 ? "float" : "double");
    return SDB_STATUS_SUCCESS;
}

}  // namespace sdb
