// spmv_tile.cu — SpMV with x staged in shared memory (inspector + executor), for handles that are multiplied
// by a vector again and again (the CG / FGMRES callers of mkl_sparse_?_mv, _sparse_vector.py:84-92).
//
// Why: y = A x with columns spread over a long x is not bound by bytes on this machine.  Every stored entry
// gathers one element of x, i.e. one 32-byte sector through the L1 / L2 request path: spmv_wide_kernel spends
// 1.57 cycles per entry and SM there and reaches 0.23 of the HBM roofline on BASELINE's configs[1] matrix
// (DESIGN.md, K1v).  Shared memory serves 32 scattered 4-byte reads in a few cycles.
//
// Inspector (once per handle, cached like the SpMM slab copy): the matrix is cut into tiles of RB rows x S
// columns (S * sizeof(T) = 64 KiB of x, RB * sizeof(T) <= 128 KiB of partial sums).  The entries of a tile are
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
__global__ void __launch_bounds__(kTileThreads) tile_count_kernel(int64_t rows, int64_t rb_rows, int subs, int slab_cols,
                                                                 int n_slabs, const int64_t* __restrict__ indptr,
                                                                 const int32_t* __restrict__ indices,
                                                                 int32_t* __restrict__ sub_counts) {
    extern __shared__ int32_t cnt[];
    for (int i = threadIdx.x; i < n_slabs; i += kTileThreads) cnt[i] = 0;
    __syncthreads();
    const int64_t rb = blockIdx.x / subs;
    const int sub = blockIdx.x % subs;
    const int64_t local = int64_t(sub) * kTileThreads + threadIdx.x;
    const int64_t r = rb * rb_rows + local;
    if (local < rb_rows && r < rows)
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

// One row per thread: at every step a warp places one entry of 32 different rows, so neighbours inside a tile
// belong to different rows.  Padding slots keep (0, 0.0) from the memset.
template <typename T>
__global__ void __launch_bounds__(kTileThreads) tile_scatter_kernel(int64_t rows, int64_t rb_rows, int subs,
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
    const int64_t r = rb * rb_rows + local;
    if (local >= rb_rows || r >= rows) return;
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
__global__ void __launch_bounds__(kTileThreads, 1) spmv_tile_kernel(int64_t rows, int64_t cols, int64_t rb_rows,
                                                                  int slab_cols, int n_slabs, int n_groups,
                                                                  int slabs_per_group,
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
    const int64_t row0 = rb * rb_rows;
    const int nr = int(min(rb_rows, rows - row0));
    for (int i = tid; i < nr; i += kTileThreads) ys[i] = T(0);
    const int s0 = g * slabs_per_group, s1 = min(n_slabs, s0 + slabs_per_group);
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

struct TilePlan {
    int64_t rb_rows = 0, n_rb = 0;
    int slab_cols = 0, n_slabs = 0, n_groups = 1, slabs_per_group = 0;
};

TilePlan make_plan(int64_t rows, int64_t cols, size_t es, int sm_count) {
    TilePlan pl;
    // SDB_SPMV_TILE_XKB: KiB of x per slab (sweeps; 16..96, default 64)
    static const size_t x_bytes = [] {
        const char* e = getenv("SDB_SPMV_TILE_XKB");
        const int kb = e ? atoi(e) : 0;
        return kb >= 16 && kb <= 96 ? size_t(kb) << 10 : kTileXBytes;
    }();
    pl.slab_cols = int(x_bytes / es);
    const int64_t rb_max = int64_t(kTileYBytes / es);
    pl.n_slabs = int((cols + pl.slab_cols - 1) / pl.slab_cols);
    // one wave: (row blocks) x (slab groups) = sm_count, as few row blocks as the partial sums allow (every row
    // block reads all of x once)
    int64_t n_rb = 0;
    for (int groups : {4, 2, 1}) {
        if (sm_count % groups != 0 || groups > pl.n_slabs) continue;
        const int64_t blocks = sm_count / groups;
        if ((rows + blocks - 1) / blocks <= rb_max) {
            pl.n_groups = groups;
            n_rb = blocks;
            break;
        }
    }
    if (n_rb == 0) {  // several waves of sm_count row blocks
        pl.n_groups = 1;
        const int64_t waves = (rows + int64_t(sm_count) * rb_max - 1) / (int64_t(sm_count) * rb_max);
        n_rb = int64_t(sm_count) * waves;
    }
    pl.rb_rows = ((rows + n_rb - 1) / n_rb + 31) & ~int64_t(31);
    pl.n_rb = (rows + pl.rb_rows - 1) / pl.rb_rows;
    pl.slabs_per_group = (pl.n_slabs + pl.n_groups - 1) / pl.n_groups;
    return pl;
}

template <typename T>
sdb_status build_tiles(Context* ctx, const CsrView& a, sdb_mat* m, const TilePlan& pl, bool check_balance) {
    cudaStream_t ls = ctx->stream;
    const int64_t n_tiles = pl.n_rb * pl.n_slabs;
    const int subs = int((pl.rb_rows + kTileThreads - 1) / kTileThreads);
    const int64_t n_sub = n_tiles * subs;
    DevBuf counts, sub_base;
    SDB_TRY(counts.alloc(size_t(n_sub) * 4, ls));
    SDB_TRY(sub_base.alloc(size_t(n_sub + 1) * 8, ls));
    int64_t* tile_ptr = nullptr;
    SDB_TRY(dev_alloc(reinterpret_cast<void**>(&tile_ptr), size_t(n_tiles + 1) * 8, ls));
    auto fail = [&](sdb_status st) {
        cudaFreeAsync(tile_ptr, ls);
        return st;
    };
    const size_t smem = size_t(pl.n_slabs) * 4;
    const unsigned igrid = unsigned(pl.n_rb * subs);
    SDB_LAUNCH(tile_count_kernel, igrid, kTileThreads, smem, ls, a.rows, pl.rb_rows, subs, pl.slab_cols, pl.n_slabs,
               a.indptr, a.indices, counts.as<int32_t>());
    SDB_LAUNCH(tile_pad_kernel, unsigned((n_tiles + 255) / 256), 256, 0, ls, n_tiles, subs, counts.as<int32_t>());
    sdb_status st = exclusive_scan_i32_to_i64(ls, counts.as<int32_t>(), sub_base.as<int64_t>(), n_sub);
    if (st != SDB_STATUS_SUCCESS) return fail(st);
    SDB_LAUNCH(tile_ptr_kernel, unsigned((n_tiles + 256) / 256), 256, 0, ls, n_tiles, subs, sub_base.as<int64_t>(),
               tile_ptr);
    std::vector<int64_t> host(size_t(n_tiles) + 1);
    if (cudaMemcpyAsync(host.data(), tile_ptr, host.size() * 8, cudaMemcpyDeviceToHost, ls) != cudaSuccess ||
        cudaStreamSynchronize(ls) != cudaSuccess) {
        set_error("spmv tiles: reading the tile offsets failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(SDB_STATUS_EXECUTION_FAILED);
    }
    const int64_t total = host[size_t(n_tiles)];
    if (check_balance) {
        // one CTA per (row block, slab group): the slowest item sets the time
        int64_t worst = 0;
        for (int64_t rb = 0; rb < pl.n_rb; ++rb)
            for (int g = 0; g < pl.n_groups; ++g) {
                const int64_t s0 = std::min<int64_t>(pl.n_slabs, int64_t(g) * pl.slabs_per_group);
                const int64_t s1 = std::min<int64_t>(pl.n_slabs, s0 + pl.slabs_per_group);
                worst = std::max(worst, host[size_t(rb * pl.n_slabs + s1)] - host[size_t(rb * pl.n_slabs + s0)]);
            }
        const double mean = double(total) / double(pl.n_rb * pl.n_groups);
        if (double(worst) > 1.35 * mean) {
            m->vt_state = -1;  // skewed: the gather kernel balances better
            return fail(SDB_STATUS_NOT_SUPPORTED);
        }
    }
    void *rc = nullptr, *val = nullptr;
    st = dev_alloc(&rc, size_t(std::max<int64_t>(total, 4)) * 4, ls);
    if (st != SDB_STATUS_SUCCESS) return fail(st);
    st = dev_alloc(&val, size_t(std::max<int64_t>(total, 4)) * sizeof(T), ls);
    if (st != SDB_STATUS_SUCCESS) {
        cudaFreeAsync(rc, ls);
        return fail(st);
    }
    cudaMemsetAsync(rc, 0, size_t(std::max<int64_t>(total, 4)) * 4, ls);
    cudaMemsetAsync(val, 0, size_t(std::max<int64_t>(total, 4)) * sizeof(T), ls);
    SDB_LAUNCH((tile_scatter_kernel<T>), igrid, kTileThreads, smem, ls, a.rows, pl.rb_rows, subs, pl.slab_cols,
               pl.n_slabs, a.indptr, a.indices, static_cast<const T*>(a.values), sub_base.as<int64_t>(),
               static_cast<uint32_t*>(rc), static_cast<T*>(val));
    trace(ls, "spmv tiles: %lld row blocks x %d column slabs, %lld entries (%lld stored)", (long long)pl.n_rb, pl.n_slabs,
          (long long)total, (long long)a.nnz);
    m->vt_rc = rc;
    m->vt_val = val;
    m->vt_ptr = tile_ptr;
    m->vt_rb_rows = pl.rb_rows;
    m->vt_entries = total;
    m->vt_state = 1;
    return SDB_STATUS_SUCCESS;
}

}  // namespace

void drop_spmv_tiles(sdb_mat* m, cudaStream_t s) {
    if (m->vt_rc) cudaFreeAsync(m->vt_rc, s);
    if (m->vt_val) cudaFreeAsync(m->vt_val, s);
    if (m->vt_ptr) cudaFreeAsync(m->vt_ptr, s);
    m->vt_rc = m->vt_val = nullptr;
    m->vt_ptr = nullptr;
    m->vt_state = 0;
    m->vt_entries = 0;
    m->spmv_calls = 0;
}

// Policy ("spmv_tile": 0 automatic, 1 never, 2 whenever the shape allows).  Automatic: the handle has been
// multiplied by a vector before (a matrix used once never pays the inspector), x is far larger than an L1, the
// matrix is large enough to fill the machine, and the inspector has not found it too skewed.
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
           a.rows >= 148 * 256;
}

sdb_status spmv_tile_device(Context* ctx, cudaStream_t s, const CsrView& a, int dtype, const double* alpha,
                            const double* beta, const void* dX, void* dY) {
    sdb_mat* m = a.owner;
    SDB_REQUIRE(m != nullptr, SDB_STATUS_NOT_SUPPORTED, "spmv tiles: ad-hoc view");
    const size_t es = dtype_size(dtype);
    const TilePlan pl = make_plan(a.rows, a.cols, es, ctx->sm_count);
    {
        std::lock_guard<std::mutex> cache_lock(g_companion_mutex);  // the tiles are a per-handle cache
        if (m->vt_state == -1) return SDB_STATUS_NOT_SUPPORTED;
        if (m->vt_state != 1 || m->vt_rb_rows != pl.rb_rows) {
            drop_spmv_tiles(m, ctx->stream);
            m->spmv_calls = 2;
            const bool check = get_option(kOptSpmvTile) != 2;
            const sdb_status st = dtype == SDB_F32 ? build_tiles<float>(ctx, a, m, pl, check)
                                                   : build_tiles<double>(ctx, a, m, pl, check);
            if (st != SDB_STATUS_SUCCESS) return st;
            if (s != ctx->stream) SDB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    const size_t smem = size_t(pl.slab_cols) * es + size_t(pl.rb_rows) * es;
    const unsigned grid = unsigned(pl.n_rb * pl.n_groups);
    const unsigned sgrid = unsigned((a.rows + 255) / 256);
    if (dtype == SDB_F32) {
        SDB_LAUNCH(scale_vector_kernel<float>, sgrid, 256, 0, s, a.rows, float(beta[0]), static_cast<float*>(dY));
        SDB_CUDA(cudaFuncSetAttribute(spmv_tile_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        SDB_LAUNCH(spmv_tile_kernel<float>, grid, kTileThreads, smem, s, a.rows, a.cols, pl.rb_rows, pl.slab_cols,
                   pl.n_slabs, pl.n_groups, pl.slabs_per_group, m->vt_ptr, static_cast<const uint32_t*>(m->vt_rc),
                   static_cast<const float*>(m->vt_val), static_cast<const float*>(dX), float(alpha[0]),
                   static_cast<float*>(dY));
    } else {
        SDB_LAUNCH(scale_vector_kernel<double>, sgrid, 256, 0, s, a.rows, beta[0], static_cast<double*>(dY));
        SDB_CUDA(cudaFuncSetAttribute(spmv_tile_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        SDB_LAUNCH(spmv_tile_kernel<double>, grid, kTileThreads, smem, s, a.rows, a.cols, pl.rb_rows, pl.slab_cols,
                   pl.n_slabs, pl.n_groups, pl.slabs_per_group, m->vt_ptr, static_cast<const uint32_t*>(m->vt_rc),
                   static_cast<const double*>(m->vt_val), static_cast<const double*>(dX), alpha[0],
                   static_cast<double*>(dY));
    }
    note_spmm_kernel("spmv_tile_kernel<%s>", dtype == SDB_F32 ? "float" : "double");
    return SDB_STATUS_SUCCESS;
}

}  // namespace sdb
