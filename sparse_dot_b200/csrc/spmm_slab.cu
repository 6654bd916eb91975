// spmm_slab.cu — L2-tiled CSR x dense SpMM for panels much larger than the L2 cache.
//
// The row-gather kernel (spmm.cu) reads one n-wide row of X per stored entry; when X is several
// times the 126 MB L2 and the columns are scattered, ~85 % of those gathers come from HBM and the
// kernel sits on the DRAM roofline at ~22 GB of traffic for a problem whose unique bytes are
// ~2 GB (profiles/README.md).  This kernel cuts the traffic instead of chasing bandwidth:
//
//   * the columns are cut into S slabs of W rows of X, W chosen so that one slab (W * n * sv
//     bytes) stays L2-resident;
//   * a persistent grid of one CTA per SM walks row blocks; every warp owns 32 rows whose
//     accumulators live in SHARED MEMORY (32 x n x sv bytes per warp) for the whole sweep;
//   * all CTAs sweep the slabs in the same order, so at any moment the whole chip gathers from
//     the same L2-resident slab: X is read from HBM once per wave of row blocks
//     (rows / (SMs * rows per CTA) times) instead of once per stored entry;
//   * per slab a warp flattens the entries of its 32 rows that fall into the slab (a per-row
//     offset table built once per matrix and cached on the handle gives the segment bounds),
//     issues the X-row gathers 8 at a time and adds each row's partial sum into shared memory once.
//
// The bound moves from DRAM (22 GB) to L2 bandwidth (every stored entry still pulls n * sv bytes
// from L2 into an SM).  Requirements: row-major panels, n * sv = 512 bytes per X row (one 16-byte
// pack per lane), rows strictly ascending (the slab table is a binary search per row and slab).
#include <cstdlib>
#include <type_traits>

#include "common.h"
#include "types.cuh"

namespace sdb {

namespace {

constexpr int kSlabRowsPerCta = 416;         // 416 rows x 512 B = 208 KB of accumulators per CTA
constexpr int kSlabUnroll = 8;
constexpr int kSlabMaxPeers = 8;

template <typename T> struct SlabPeers {
    T* y[kSlabMaxPeers];
};

// off[s * rows + r] = number of entries of row r with column < s * width   (s = 0 .. S)
__global__ void __launch_bounds__(256) slab_offsets_kernel(int64_t rows, const int64_t* __restrict__ indptr,
                                                           const int32_t* __restrict__ indices, int S,
                                                           int64_t width, int32_t* __restrict__ off) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    const int64_t b = indptr[r];
    const int len = int(indptr[r + 1] - b);
    for (int s = lane; s <= S; s += 32) {
        const int64_t key = int64_t(s) * width;  // first entry with column >= key
        int lo = 0, hi = len;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (int64_t(indices[b + mid]) < key) lo = mid + 1;
            else hi = mid;
        }
        off[int64_t(s) * rows + r] = s == S ? len : lo;
    }
}

template <typename T> struct alignas(16) Pack16 {
    static constexpr int N = 16 / int(sizeof(T));
    T v[N];
};

template <typename T> __device__ __forceinline__ Pack16<T> ldg16(const T* p) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    Pack16<T> r;
    *reinterpret_cast<float4*>(&r) = q;
    return r;
}

// RPW rows per warp: fewer rows per warp = more warps per CTA (416 / RPW) = more independent chains
template <typename T, int RPW>
__global__ void __launch_bounds__((kSlabRowsPerCta / RPW) * 32, 1)
    spmm_slab_kernel(int64_t rows, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                     const T* __restrict__ values, bool conj_a, const int32_t* __restrict__ slab_off, int S,
                     const T* __restrict__ X, int64_t ldx, T alpha, T beta, T* __restrict__ y_self,
                     SlabPeers<T> peers, int n_peers, int self, int64_t row0, int64_t ldy) {
    constexpr int VEC = Pack16<T>::N;
    constexpr unsigned kFull = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // this warp's accumulators: [RPW rows][32 lanes] packs of 16 bytes
    Pack16<T>* acc = reinterpret_cast<Pack16<T>*>(smem_raw) + warp * RPW * 32;
    const T* xlane = X + lane * VEC;
    const int64_t n_blocks = (rows + kSlabRowsPerCta - 1) / kSlabRowsPerCta;

    for (int64_t rb = blockIdx.x; rb < n_blocks; rb += gridDim.x) {
        const int64_t row_base = rb * kSlabRowsPerCta + int64_t(warp) * RPW;
        Pack16<T> zero;
#pragma unroll
        for (int i = 0; i < VEC; ++i) zero.v[i] = Num<T>::zero();
#pragma unroll 8
        for (int r = 0; r < RPW; ++r) acc[r * 32 + lane] = zero;
        __syncwarp();

        const int64_t my_row = row_base + lane;
        const bool valid = lane < RPW && my_row < rows;
        const int64_t rstart = valid ? indptr[my_row] : 0;
        int off_prev = 0;  // entries with column < 0

        int off_ahead = valid ? __ldg(slab_off + rows + my_row) : 0;  // boundary of slab 0 | 1, loaded one slab ahead
        for (int s = 0; s < S; ++s) {
            const int off_next = off_ahead;
            if (s + 1 < S && valid) off_ahead = __ldg(slab_off + int64_t(s + 2) * rows + my_row);
            const int cnt = off_next - off_prev;      // this row's entries inside slab s
            const int64_t seg = rstart + off_prev;    // where they start
            off_prev = off_next;
            // inclusive prefix of cnt over the 32 rows of the warp
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += o;
            }
            const int total = __shfl_sync(kFull, incl, 31);
            if (total == 0) continue;

            // (column, value, row) of flat entry f0 + lane of this slab
            auto fetch = [&](int f0, int32_t& c_out, T& v_out, int& r_out) {
                const int f = f0 + lane;
                int lo = 0, hi = 31;  // first row whose inclusive prefix exceeds f
#pragma unroll
                for (int it = 0; it < 5; ++it) {
                    const int mid = (lo + hi) >> 1;
                    const int pm = __shfl_sync(kFull, incl, mid);
                    if (f >= pm) lo = mid + 1;
                    else hi = mid;
                }
                r_out = lo;  // meaningful when f < total
                const int r_incl = __shfl_sync(kFull, incl, lo);
                const int r_cnt = __shfl_sync(kFull, cnt, lo);
                const int64_t r_seg = __shfl_sync(kFull, seg, lo);
                c_out = 0;
                v_out = Num<T>::zero();
                if (f < total) {
                    const int64_t p = r_seg + (f - (r_incl - r_cnt));
                    c_out = __ldg(indices + p);
                    v_out = ldg(values + p);
                    if (conj_a) v_out = conj_(v_out);
                }
            };

            Pack16<T> cur = zero;
            int cur_row = -1;
            int32_t c, c_nx = 0;
            T v, v_nx = Num<T>::zero();
            int my_r, my_r_nx = 0;
            fetch(0, c, v, my_r);
            for (int f0 = 0; f0 < total; f0 += 32) {
                // the next chunk's entries are requested before this chunk's gathers are consumed
                if (f0 + 32 < total) fetch(f0 + 32, c_nx, v_nx, my_r_nx);
                const int batch = min(32, total - f0);
                // bit u set = flat entry u of this chunk starts a new row (relative to the entry before it)
                const int prev_r = __shfl_up_sync(kFull, my_r, 1);
                const unsigned starts = __ballot_sync(kFull, lane == 0 ? my_r != cur_row : my_r != prev_r);
                // one batch of kSlabUnroll gathers: loads first, then the FMAs with a row-change test per entry
                auto run_batch = [&](int u0, auto full) {
                    constexpr bool kFullBatch = decltype(full)::value;
                    Pack16<T> x[kSlabUnroll];
                    T a[kSlabUnroll];
#pragma unroll
                    for (int u = 0; u < kSlabUnroll; ++u) {
                        const int32_t cj = __shfl_sync(kFull, c, u0 + u);
                        a[u] = shfl(kFull, v, u0 + u, 32);
                        if (kFullBatch || u0 + u < batch) x[u] = ldg16<T>(xlane + int64_t(cj) * ldx);
                    }
                    const unsigned sb = starts >> u0;
#pragma unroll
                    for (int u = 0; u < kSlabUnroll; ++u) {
                        if (kFullBatch || u0 + u < batch) {  // warp-uniform
                            if (sb & (1u << u)) {
                                if (cur_row >= 0) {
                                    Pack16<T> t = acc[cur_row * 32 + lane];
#pragma unroll
                                    for (int i = 0; i < VEC; ++i) t.v[i] = add(t.v[i], cur.v[i]);
                                    acc[cur_row * 32 + lane] = t;
                                }
                                cur = zero;
                                cur_row = __shfl_sync(kFull, my_r, u0 + u);
                            }
#pragma unroll
                            for (int i = 0; i < VEC; ++i) cur.v[i] = madd(a[u], x[u].v[i], cur.v[i]);
                        }
                    }
                };
                int u0 = 0;
                for (; u0 + kSlabUnroll <= batch; u0 += kSlabUnroll) run_batch(u0, std::true_type{});
                for (; u0 < batch; ++u0) {  // ragged tail of the last chunk: one entry at a time
                    const int32_t cj = __shfl_sync(kFull, c, u0);
                    const T aj = shfl(kFull, v, u0, 32);
                    const Pack16<T> xj = ldg16<T>(xlane + int64_t(cj) * ldx);
                    if ((starts >> u0) & 1u) {
                        if (cur_row >= 0) {
                            Pack16<T> t = acc[cur_row * 32 + lane];
#pragma unroll
                            for (int i = 0; i < VEC; ++i) t.v[i] = add(t.v[i], cur.v[i]);
                            acc[cur_row * 32 + lane] = t;
                        }
                        cur = zero;
                        cur_row = __shfl_sync(kFull, my_r, u0);
                    }
#pragma unroll
                    for (int i = 0; i < VEC; ++i) cur.v[i] = madd(aj, xj.v[i], cur.v[i]);
                }
                c = c_nx;
                v = v_nx;
                my_r = my_r_nx;
            }
            if (cur_row >= 0) {
                Pack16<T> t = acc[cur_row * 32 + lane];
#pragma unroll
                for (int i = 0; i < VEC; ++i) t.v[i] = add(t.v[i], cur.v[i]);
                acc[cur_row * 32 + lane] = t;
            }
        }
        __syncwarp();

        // epilogue: y = alpha * acc + beta * y, one 512-byte row per iteration
        const bool beta_zero = Num<T>::is_zero(beta);
        const int live_rows = int(max(int64_t(0), min(int64_t(RPW), rows - row_base)));
        for (int r = 0; r < live_rows; ++r) {
            const int64_t o = (row0 + row_base + r) * ldy + lane * VEC;
            const Pack16<T> t = acc[r * 32 + lane];
            Pack16<T> out;
            if (beta_zero) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) out.v[i] = mul(alpha, t.v[i]);
            } else {
                Pack16<T> old;
                *reinterpret_cast<float4*>(&old) = *reinterpret_cast<const float4*>(y_self + o);
#pragma unroll
                for (int i = 0; i < VEC; ++i) out.v[i] = madd(alpha, t.v[i], mul(beta, old.v[i]));
            }
            *reinterpret_cast<float4*>(y_self + o) = *reinterpret_cast<const float4*>(&out);
            if (n_peers > 1) {
#pragma unroll
                for (int q = 0; q < kSlabMaxPeers; ++q)
                    if (q < n_peers && q != self)
                        *reinterpret_cast<float4*>(peers.y[q] + o) = *reinterpret_cast<const float4*>(&out);
            }
        }
        __syncwarp();
    }
}

size_t slab_target_bytes() {
    static const size_t v = [] {
        const char* e = getenv("SDB_SLAB_MB");
        return size_t(e ? atoi(e) : 24) << 20;
    }();
    return v;
}

int slab_mode() {  // 0 / 1 = off (default), 2 = whenever the shape allows
    static const int v = [] {
        const char* e = getenv("SDB_SLAB");
        return e ? atoi(e) : 0;
    }();
    return v;
}

}  // namespace

template <typename T, int RPW>
static sdb_status launch_slab_rpw(cudaStream_t s, const CsrView& a, bool conj_a, const int32_t* slab_off, int S,
                                  const T* X, int64_t ldx, T alpha, T beta, const SlabPeers<T>& peers, int n_peers,
                                  int self, int64_t row0, int64_t ldy, unsigned grid, size_t smem) {
    SDB_CUDA(cudaFuncSetAttribute(spmm_slab_kernel<T, RPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    SDB_LAUNCH((spmm_slab_kernel<T, RPW>), grid, (kSlabRowsPerCta / RPW) * 32, smem, s, a.rows, a.indptr, a.indices,
               static_cast<const T*>(a.values), conj_a, slab_off, S, X, ldx, alpha, beta, peers.y[self], peers,
               n_peers, self, row0, ldy);
    return SDB_STATUS_SUCCESS;
}

template <typename T>
static sdb_status launch_slab(cudaStream_t s, const CsrView& a, bool conj_a, const int32_t* slab_off, int S,
                              const T* X, int64_t ldx, T alpha, T beta, void* const* dY_peers, int n_peers, int self,
                              int64_t row0, int64_t ldy, unsigned grid, size_t smem) {
    static const int rpw = [] {
        const char* e = getenv("SDB_SLAB_RPW");
        return e ? atoi(e) : 13;
    }();
    SlabPeers<T> peers;
    for (int q = 0; q < kSlabMaxPeers; ++q) peers.y[q] = q < n_peers ? static_cast<T*>(dY_peers[q]) : nullptr;
    if (rpw == 32)
        return launch_slab_rpw<T, 32>(s, a, conj_a, slab_off, S, X, ldx, alpha, beta, peers, n_peers, self, row0, ldy,
                                      grid, smem);
    if (rpw == 13)
        return launch_slab_rpw<T, 13>(s, a, conj_a, slab_off, S, X, ldx, alpha, beta, peers, n_peers, self, row0, ldy,
                                      grid, smem);
    return launch_slab_rpw<T, 16>(s, a, conj_a, slab_off, S, X, ldx, alpha, beta, peers, n_peers, self, row0, ldy,
                                  grid, smem);
}

bool spmm_slab_wanted(const CsrView& a, int dtype, int64_t n, int64_t ldx) {
    if (slab_mode() == 1 || a.owner == nullptr) return false;
    if (dtype != SDB_F32 && dtype != SDB_F64) return false;
    const size_t sv = dtype_size(dtype);
    if (size_t(n) * sv != 512 || ldx != n) return false;
    if (a.owner->strict_sorted == -1) return false;
    // Measured on B200 (profiles/README.md, round 1c): on configs[1] this kernel cuts DRAM traffic from 22.2 GB to
    // 10.1 GB as designed, but it becomes instruction-issue bound (68 % issue slots busy, IPC 2.7) at 3.99 ms
    // against 3.54 ms for the row-gather kernel on the DRAM roofline — so it is opt-in (SDB_SLAB=2) until the
    // per-entry instruction count comes down.
    return slab_mode() == 2 && a.rows > 0 && a.nnz > 0;
}

sdb_status spmm_slab_device(Context* ctx, cudaStream_t s, const CsrView& a, int dtype, bool conj_a,
                            const double* alpha, const double* beta, const void* dX, int64_t n, int64_t ldx,
                            void* const* dY_peers, int n_peers, int self, int64_t row0, int64_t ldy) {
    sdb_mat* m = a.owner;
    SDB_REQUIRE(m != nullptr, SDB_STATUS_NOT_SUPPORTED, "spmm_slab: ad-hoc view");
    if (m->strict_sorted == 0) {
        // the check runs on the library stream; the caller's stream must not race with it
        SDB_TRY(ensure_strict_flag(ctx, m));
    }
    if (m->strict_sorted != 1) return SDB_STATUS_NOT_SUPPORTED;
    const size_t sv = dtype_size(dtype);
    int64_t width = std::max<int64_t>(1024, int64_t(slab_target_bytes() / (size_t(n) * sv)));
    const int S = int((a.cols + width - 1) / width);
    if (S < 2 || S > 4096) return SDB_STATUS_NOT_SUPPORTED;
    if (m->slab_off == nullptr || m->slab_count != S || m->slab_width != width) {
        cudaStream_t ls = ctx->stream;
        if (m->slab_off) cudaFreeAsync(m->slab_off, ls);
        m->slab_off = nullptr;
        SDB_TRY(dev_alloc(reinterpret_cast<void**>(&m->slab_off), size_t(S + 1) * size_t(a.rows) * 4, ls));
        SDB_LAUNCH(slab_offsets_kernel, unsigned((a.rows * 32 + 255) / 256), 256, 0, ls, a.rows, a.indptr, a.indices,
                   S, width, m->slab_off);
        m->slab_count = S;
        m->slab_width = width;
        if (s != ls) SDB_CUDA(cudaStreamSynchronize(ls));
    }
    const size_t smem = size_t(kSlabRowsPerCta) * 512;
    const int64_t n_blocks = (a.rows + kSlabRowsPerCta - 1) / kSlabRowsPerCta;
    const unsigned grid = unsigned(std::min<int64_t>(n_blocks, ctx->sm_count));
    return SDB_DISPATCH_DTYPE(dtype, T, [&]() -> sdb_status {
        if constexpr (sizeof(T) > 8) {
            return SDB_STATUS_NOT_SUPPORTED;
        } else {
            return launch_slab<T>(s, a, conj_a, m->slab_off, S, static_cast<const T*>(dX), ldx,
                                  Num<T>::make(alpha[0], alpha[1]), Num<T>::make(beta[0], beta[1]), dY_peers, n_peers,
                                  self, row0, ldy, grid, smem);
        }
    });
}

}  // namespace sdb
