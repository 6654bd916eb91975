// spmm_slab.cu — L2-tiled CSR x dense SpMM for panels much larger than the L2 cache
// (same contract as spmm.cu: Y := alpha * A * X + beta * Y, mkl_sparse_?_mm as
// sparse_dot_mkl/_sparse_dense.py:111-123 calls it).
//
// The row-gather kernel (spmm.cu) reads one n-wide row of X per stored entry; when X is several
// times the 126 MB L2 and the columns are scattered, ~85 % of those gathers come from HBM and the
// kernel sits on the DRAM roofline at ~22 GB of traffic for a problem whose unique bytes are
// ~2 GB (profiles/README.md).  This kernel cuts the traffic instead of chasing bandwidth, with an
// inspector / executor split (the analogue of mkl_sparse_optimize, which the reference never calls):
//
// Inspector (once per handle, cached; `slab_permute_kernel`): rows are taken in groups of RPW
// consecutive rows (one group = one warp of the executor).  The stored entries of a group — which
// are contiguous in CSR — are re-ordered by (column slab, row, column), where a slab is a range of
// `width` columns whose X rows (width * n * sv bytes, 24 MB by default) fit L2 comfortably, and
// written as packed (local row << 27 | column) words plus values.  Nothing else is kept: a group's
// range is still [indptr[g * RPW], indptr[(g + 1) * RPW]).
//
// Executor (`spmm_stream_kernel`): a persistent grid of one 1024-thread CTA per SM; a CTA owns
// RPW * 32 rows at a time and keeps their accumulators in SHARED MEMORY (RPW * 32 rows x 512 B =
// 208-224 KB).  Each warp simply streams its group's re-ordered entries, 32 per coalesced load,
// gathers the 512-byte X rows 8 at a time (one 16-byte load per lane) and accumulates in registers;
// when the row changes (entries of one row inside one slab are adjacent) the running sums are
// swapped with the new row's accumulators in shared memory.  Because every warp of the chip walks
// its entries in slab order at about the same pace, at any moment the whole chip gathers from the
// same few slabs of X, which therefore stay L2-resident: X comes from HBM once per wave of row
// blocks (rows / (SMs * RPW * 32) times) instead of once per stored entry.  There is no barrier
// and no slab table in the hot loop — the ordering alone provides the locality, and a matrix with
// uneven rows merely loses some of it.
//
// The bound moves from DRAM (22 GB) to L2 -> SM bandwidth (every stored entry still pulls n * sv
// bytes from L2).  Requirements: row-major panels with n * sv = 512 bytes per row (one 16-byte
// pack per lane), fp32 / fp64, strictly ascending rows, fewer than 2^27 columns.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.h"
#include "types.cuh"

namespace sdb {

namespace {

constexpr int kSlabMaxPeers = 8;
constexpr int kColBits = 27;
constexpr uint32_t kColMask = (1u << kColBits) - 1u;

template <typename T> struct SlabPeers {
    T* y[kSlabMaxPeers];
};

template <typename T> struct alignas(16) Pack16 {
    static constexpr int N = 16 / int(sizeof(T));
    T v[N];
};

// ---------------------------------------------------------------- inspector
// One warp per group of `rpw` rows, lane = local row.  Every lane walks its own (ascending) row
// once; per slab the lanes' counts are scanned and each lane copies its run behind the runs of
// the lower rows.  One-time cost, a few passes over A.
template <typename T>
__global__ void __launch_bounds__(256) slab_permute_kernel(int64_t rows, int rpw, const int64_t* __restrict__ indptr,
                                                           const int32_t* __restrict__ indices,
                                                           const T* __restrict__ values, int64_t width,
                                                           uint32_t* __restrict__ ent_rc, T* __restrict__ ent_val) {
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t row_first = g * rpw;
    if (row_first >= rows) return;
    const int64_t row = row_first + lane;
    const bool valid = lane < rpw && row < rows;
    const int64_t rbeg = valid ? indptr[row] : 0;
    const int len = valid ? int(indptr[row + 1] - rbeg) : 0;
    int64_t cursor = indptr[row_first];
    const int64_t gend = indptr[min(row_first + rpw, rows)];
    int p = 0;
    while (cursor < gend) {  // warp-uniform
        // the next slab that holds an entry of this group
        int64_t mine = p < len ? int64_t(indices[rbeg + p]) / width : INT64_MAX;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) mine = min(mine, __shfl_xor_sync(kFull, mine, d));
        const int64_t bound = (mine + 1) * width;
        const int start = p;
        while (p < len && int64_t(indices[rbeg + p]) < bound) ++p;
        const int cnt = p - start;
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += o;
        }
        const int total = __shfl_sync(kFull, incl, 31);
        const int64_t dst = cursor + (incl - cnt);
        for (int e = 0; e < cnt; ++e) {
            ent_rc[dst + e] = (uint32_t(lane) << kColBits) | uint32_t(indices[rbeg + start + e]);
            ent_val[dst + e] = values[rbeg + start + e];
        }
        cursor += total;
    }
}

// ---------------------------------------------------------------- executor
template <typename T> __device__ __forceinline__ Pack16<T> gather16(const char* xlane, uint32_t col, uint32_t row_bytes) {
    // 32 x 32 -> 64-bit multiply-add on the base pointer: one IMAD.WIDE.U32
    const float4 q = __ldg(reinterpret_cast<const float4*>(xlane + uint64_t(col) * row_bytes));
    Pack16<T> r;
    *reinterpret_cast<float4*>(&r) = q;
    return r;
}
// The same gather with an L2 evict_last policy: X rows of the active slabs are kept in L2 in preference to lines
// without a policy — the output panel rows peers write into this GPU during a fused all-gather (SDB_SLAB_KEEP=1).
template <typename T>
__device__ __forceinline__ Pack16<T> gather16_keep(const char* xlane, uint32_t col, uint32_t row_bytes, uint64_t policy) {
    float4 q;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                 : "l"(xlane + uint64_t(col) * row_bytes), "l"(policy));
    Pack16<T> r;
    *reinterpret_cast<float4*>(&r) = q;
    return r;
}

// One staged entry: packed (local row << 27 | column) and the value; 8 bytes (fp32) or 16 bytes (fp64), so a
// warp-uniform read of it is ONE shared-memory wavefront (broadcast) where two shuffles would be two.
template <typename T> struct alignas(sizeof(T) == 4 ? 8 : 16) StagedEntry {
    uint32_t rc;
    T v;
};

// RPW rows per warp, WARPS warps per CTA (RPW * WARPS * 512 B of accumulators + WARPS * 32 staged entries),
// U gathers in flight per lane, CTAS resident CTAs per SM (register budget = 65536 / (CTAS * WARPS * 32)).
//
// What bounds this kernel is the L1 data pipe (one 128-byte wavefront per clock per SM; ncu:
// l1tex__data_pipe_lsu_wavefronts), so the hot loop is written to spend as few wavefronts per stored entry as
// possible: 4 for the 512-byte gather (irreducible), 1 for the staged-entry broadcast, and 8 per row change
// (store the running sums, load the next row's) amortised over the run of entries a row has inside one slab.
template <typename T, int RPW, int WARPS, int U, int CTAS, bool KEEP = false>
__global__ void __launch_bounds__(WARPS * 32, CTAS)
    spmm_stream_kernel(int64_t rows, const int64_t* __restrict__ indptr, const uint32_t* __restrict__ ent_rc,
                       const T* __restrict__ ent_val, const T* __restrict__ X, uint32_t row_bytes, T alpha, T beta,
                       T* __restrict__ y_self, SlabPeers<T> peers, int n_peers, int self, int64_t row0, int64_t ldy) {
    constexpr int VEC = Pack16<T>::N;
    constexpr int kRows = RPW * WARPS;  // rows per CTA
    using Ent = StagedEntry<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // this warp's accumulators: [RPW rows][32 lanes] packs of 16 bytes
    Pack16<T>* acc = reinterpret_cast<Pack16<T>*>(smem_raw) + warp * RPW * 32 + lane;
    // this warp's staging buffer: the 32 entries of the chunk being consumed
    Ent* stage = reinterpret_cast<Ent*>(smem_raw + size_t(kRows) * 512) + warp * 32;
    const char* xlane = reinterpret_cast<const char*>(X + lane * VEC);
    const int64_t n_blocks = (rows + kRows - 1) / kRows;
    Pack16<T> zero;
#pragma unroll
    for (int i = 0; i < VEC; ++i) zero.v[i] = Num<T>::zero();
    uint64_t keep_policy = 0;
    if constexpr (KEEP) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_policy));
    auto gather = [&](uint32_t col) {
        if constexpr (KEEP) return gather16_keep<T>(xlane, col, row_bytes, keep_policy);
        else return gather16<T>(xlane, col, row_bytes);
    };

    for (int64_t rb = blockIdx.x; rb < n_blocks; rb += gridDim.x) {
        const int64_t row_base = rb * kRows + int64_t(warp) * RPW;
        if (row_base >= rows) continue;  // warp-uniform; no CTA-wide barrier anywhere
        const int live_rows = int(min(int64_t(RPW), rows - row_base));
        const int64_t beg = indptr[row_base];
        // 32-bit counters relative to the group start keep the hot loop small (a group never holds 2^31 entries)
        const int n_ent = int(min(indptr[row_base + live_rows] - beg, int64_t(INT32_MAX)));
        const uint32_t* __restrict__ erc = ent_rc + beg;
        const T* __restrict__ eva = ent_val + beg;
#pragma unroll
        for (int r = 0; r < RPW; ++r) acc[r * 32] = zero;

        // invariant: `cur` is the live value of acc[cur_row]
        Pack16<T> cur = zero;
        uint32_t cur_row = 0;
        // one entry: swap accumulators when the row changes (rare: the entries of a row inside a slab are
        // adjacent), then VEC FMAs.  Every lane works on the same entry, so the branch is warp-uniform.
        auto consume = [&](const Ent& e, const Pack16<T>& x) {
            const uint32_t r = e.rc >> kColBits;
            if (r != cur_row) {
                acc[cur_row * 32] = cur;
                cur_row = r;
                cur = acc[r * 32];
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) cur.v[i] = madd(e.v, x.v[i], cur.v[i]);
        };

        uint32_t rc_nx = 0;
        T v_nx = Num<T>::zero();
        if (lane < n_ent) {
            rc_nx = __ldcs(erc + lane);
            v_nx = ldcs(eva + lane);
        }
        for (int f = 0; f < n_ent; f += 32) {
            __syncwarp();  // the previous chunk has been consumed by every lane
            Ent mine;
            mine.rc = rc_nx;
            mine.v = v_nx;
            stage[lane] = mine;
            __syncwarp();
            // the next chunk is requested before this chunk's gathers are consumed
            if (f + 32 + lane < n_ent) {
                rc_nx = __ldcs(erc + f + 32 + lane);
                v_nx = ldcs(eva + f + 32 + lane);
            }
            const int cnt = min(32, n_ent - f);
            if (cnt == 32) {
                // rolling ring of U gathers: the slot an entry has just been consumed from is refilled at once,
                // so ~U 16-byte loads per lane stay in flight for the whole chunk (all indices are compile-time).
                // NOTE for whoever edits this kernel: at 32 registers ptxas's schedule of this loop is sensitive to
                // unrelated code (an A/B on one box: moving the peer stores of the epilogue out of line made ptxas
                // pair the refills — two gathers back to back, then two consumes — and the same loop ran 3.91 ms
                // instead of 2.56 ms).  Check the SASS: the LDG.E.128 of the chunk must be evenly spaced (13
                // instructions apart), see profiles/README.md round 1e.
                Ent e[U];
                Pack16<T> x[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    e[u] = stage[u];
                    x[u] = gather(e[u].rc & kColMask);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    consume(e[j % U], x[j % U]);
                    if (j + U < 32) {
                        e[j % U] = stage[j + U];
                        x[j % U] = gather(e[j % U].rc & kColMask);
                    }
                }
            } else {
                // ragged last chunk of the group: one entry at a time
                for (int j = 0; j < cnt; ++j) {
                    const Ent ej = stage[j];
                    const Pack16<T> xj = gather(ej.rc & kColMask);
                    consume(ej, xj);
                }
            }
        }
        acc[cur_row * 32] = cur;

        // epilogue: y = alpha * acc + beta * y, one 512-byte row per iteration (each lane reads back
        // only what it wrote itself, so no warp barrier is needed)
        const bool beta_zero = Num<T>::is_zero(beta);
        for (int r = 0; r < live_rows; ++r) {
            const int64_t o = (row0 + row_base + r) * ldy + lane * VEC;
            const Pack16<T> t = acc[r * 32];
            Pack16<T> out;
            if (beta_zero) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) out.v[i] = mul(alpha, t.v[i]);
            } else {
                Pack16<T> old;
                *reinterpret_cast<float4*>(&old) = __ldcs(reinterpret_cast<const float4*>(y_self + o));
#pragma unroll
                for (int i = 0; i < VEC; ++i) out.v[i] = madd(alpha, t.v[i], mul(beta, old.v[i]));
            }
            __stcs(reinterpret_cast<float4*>(y_self + o), *reinterpret_cast<const float4*>(&out));
            if (n_peers > 1) {
#pragma unroll
                for (int q = 0; q < kSlabMaxPeers; ++q)
                    if (q < n_peers && q != self)
                        __stcs(reinterpret_cast<float4*>(peers.y[q] + o), *reinterpret_cast<const float4*>(&out));
            }
        }
    }
}

// ---------------------------------------------------------------- executor for 256-byte rows
// Panels of n * sv = 256 bytes per row (N = 64 fp32, N = 32 fp64): a row is 16 lanes x 16 bytes, so a warp runs
// TWO independent streams, one per half-warp.  Half h of warp w owns group 2 w + h of the CTA (RPW rows, its own
// accumulators, its own 16-entry staging buffer, its own running sums); the halves never touch each other's state,
// they only share the instruction stream.  Everything that synchronises (__syncwarp) sits at warp-uniform points:
// the chunk loop runs to the longer of the two streams and each half masks what it has run out of.  The common
// case — both halves hold a full chunk of 16 entries — takes an unguarded unrolled path with the same rolling ring
// of two gathers as the 512-byte kernel.  Same inspector (slab-ordered copy per group of RPW rows), same epilogue.
template <typename T, int RPW, int WARPS, int CTAS>
__global__ void __launch_bounds__(WARPS * 32, CTAS)
    spmm_stream_half_kernel(int64_t rows, const int64_t* __restrict__ indptr, const uint32_t* __restrict__ ent_rc,
                            const T* __restrict__ ent_val, const T* __restrict__ X, uint32_t row_bytes, T alpha, T beta,
                            T* __restrict__ y_self, SlabPeers<T> peers, int n_peers, int self, int64_t row0,
                            int64_t ldy) {
    constexpr int VEC = Pack16<T>::N;
    constexpr int kGroups = WARPS * 2;     // groups (half-warps) per CTA
    constexpr int kRows = RPW * kGroups;   // rows per CTA
    constexpr unsigned kFull = 0xffffffffu;
    using Ent = StagedEntry<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, sub = lane & 15;
    const int group = warp * 2 + half;
    Pack16<T>* acc = reinterpret_cast<Pack16<T>*>(smem_raw) + group * RPW * 16 + sub;  // [RPW rows][16 lanes]
    Ent* stage = reinterpret_cast<Ent*>(smem_raw + size_t(kRows) * 256) + group * 16;
    const char* xlane = reinterpret_cast<const char*>(X + sub * VEC);
    const int64_t n_blocks = (rows + kRows - 1) / kRows;
    Pack16<T> zero;
#pragma unroll
    for (int i = 0; i < VEC; ++i) zero.v[i] = Num<T>::zero();

    for (int64_t rb = blockIdx.x; rb < n_blocks; rb += gridDim.x) {
        const int64_t row_base = rb * kRows + int64_t(group) * RPW;
        const bool live = row_base < rows;  // per half; the warp stays together
        const int live_rows = live ? int(min(int64_t(RPW), rows - row_base)) : 0;
        int64_t beg = 0;
        int n_ent = 0;
        if (live) {
            beg = indptr[row_base];
            n_ent = int(min(indptr[row_base + live_rows] - beg, int64_t(INT32_MAX)));
        }
        const uint32_t* __restrict__ erc = ent_rc + beg;
        const T* __restrict__ eva = ent_val + beg;
#pragma unroll
        for (int r = 0; r < RPW; ++r) acc[r * 16] = zero;
        Pack16<T> cur = zero;
        uint32_t cur_row = 0;
        auto consume = [&](const Ent& e, const Pack16<T>& x) {
            const uint32_t r = e.rc >> kColBits;
            if (r != cur_row) {  // uniform inside a half-warp
                acc[cur_row * 16] = cur;
                cur_row = r;
                cur = acc[r * 16];
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) cur.v[i] = madd(e.v, x.v[i], cur.v[i]);
        };
        // the longer stream of the two halves bounds the chunk loop (warp-uniform)
        const int n_other = __shfl_xor_sync(kFull, n_ent, 16);
        const int n_max = max(n_ent, n_other);
        uint32_t rc_nx = 0;
        T v_nx = Num<T>::zero();
        if (sub < n_ent) {
            rc_nx = __ldcs(erc + sub);
            v_nx = ldcs(eva + sub);
        }
        for (int f = 0; f < n_max; f += 16) {
            __syncwarp();  // the previous chunk has been consumed by every lane
            Ent mine;
            mine.rc = rc_nx;
            mine.v = v_nx;
            stage[sub] = mine;
            __syncwarp();
            if (f + 16 + sub < n_ent) {
                rc_nx = __ldcs(erc + f + 16 + sub);
                v_nx = ldcs(eva + f + 16 + sub);
            }
            const int cnt = max(0, min(16, n_ent - f));  // this half's entries in the chunk
            if (__all_sync(kFull, cnt == 16)) {
                Ent e[2];
                Pack16<T> x[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    e[u] = stage[u];
                    x[u] = gather16<T>(xlane, e[u].rc & kColMask, row_bytes);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    consume(e[j % 2], x[j % 2]);
                    if (j + 2 < 16) {
                        e[j % 2] = stage[j + 2];
                        x[j % 2] = gather16<T>(xlane, e[j % 2].rc & kColMask, row_bytes);
                    }
                }
            } else {
                // ragged end of either stream: one entry at a time, each half masking what it does not have
                for (int j = 0; j < 16; ++j) {
                    if (j < cnt) {
                        const Ent ej = stage[j];
                        const Pack16<T> xj = gather16<T>(xlane, ej.rc & kColMask, row_bytes);
                        consume(ej, xj);
                    }
                }
            }
        }
        acc[cur_row * 16] = cur;

        const bool beta_zero = Num<T>::is_zero(beta);
        for (int r = 0; r < live_rows; ++r) {
            const int64_t o = (row0 + row_base + r) * ldy + sub * VEC;
            const Pack16<T> t = acc[r * 16];
            Pack16<T> out;
            if (beta_zero) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) out.v[i] = mul(alpha, t.v[i]);
            } else {
                Pack16<T> old;
                *reinterpret_cast<float4*>(&old) = __ldcs(reinterpret_cast<const float4*>(y_self + o));
#pragma unroll
                for (int i = 0; i < VEC; ++i) out.v[i] = madd(alpha, t.v[i], mul(beta, old.v[i]));
            }
            __stcs(reinterpret_cast<float4*>(y_self + o), *reinterpret_cast<const float4*>(&out));
            if (n_peers > 1) {
#pragma unroll
                for (int q = 0; q < kSlabMaxPeers; ++q)
                    if (q < n_peers && q != self)
                        __stcs(reinterpret_cast<float4*>(peers.y[q] + o), *reinterpret_cast<const float4*>(&out));
            }
        }
    }
}

size_t slab_target_bytes() {
    static const size_t v = [] {
        const char* e = getenv("SDB_SLAB_MB");
        return size_t(e ? std::max(1, atoi(e)) : 24) << 20;
    }();
    return v;
}

int slab_mode() {  // 0 = automatic (default), 1 = off, 2 = whenever the shape allows (tests)
    static const int v = [] {
        const char* e = getenv("SDB_SLAB");
        return e ? atoi(e) : 0;
    }();
    return v;
}

// Executor shapes (SDB_SLAB_VARIANT picks; 0 is the default); shared memory per SM = accumulators + staging.
// Measured on configs[1] (profiles/r1_logs/stream_sweep*.log): what matters most is the number of resident warps.
//   0: 2 CTAs x 32 warps x 6 rows (384 rows / SM), 2 gathers in flight per lane   (32 registers)
//   1: 2 CTAs x 32 warps x 6 rows,                 3 gathers
//   2: 2 CTAs x 24 warps x 8 rows (384 rows / SM), 3 gathers                      (40 registers)
//   3: 2 CTAs x 24 warps x 8 rows,                 4 gathers
//   4: 3 CTAs x 16 warps x 8 rows (384 rows / SM), 3 gathers                      (40 registers)
//   5: 1 CTA  x 32 warps x 13 rows (416 rows / SM), 4 gathers                     (64 registers)
int slab_variant() {
    static const int v = [] {
        const char* e = getenv("SDB_SLAB_VARIANT");
        const int r = e ? atoi(e) : 0;
        return r >= 0 && r <= 5 ? r : 0;
    }();
    return v;
}
int slab_rpw() {
    const int v = slab_variant();
    return v <= 1 ? 6 : (v <= 4 ? 8 : 13);
}
int slab_rows_per_sm() { return slab_variant() == 5 ? 416 : 384; }

// SMs the persistent grid leaves free (the "sm" exchange strategy runs its copier CTAs there); per host thread,
// set around one sdb_spmm_dev_allgather call
thread_local int t_reserved_sms = 0;

}  // namespace

void spmm_slab_reserve_sms(int sms) { t_reserved_sms = sms < 0 ? 0 : sms; }

template <typename T, int RPW, int WARPS, int U, int CTAS, bool KEEP = false>
static sdb_status launch_stream(cudaStream_t s, const CsrView& a, const sdb_mat* m, const T* X, int64_t ldx, T alpha,
                                T beta, void* const* dY_peers, int n_peers, int self, int64_t row0, int64_t ldy,
                                int sm_count, int col_chunks) {
    constexpr int kRows = RPW * WARPS;
    constexpr size_t kSmem = size_t(kRows) * 512 + size_t(WARPS) * 32 * sizeof(StagedEntry<T>);
    static_assert(RPW <= 32 && (kSmem + 1024) * CTAS <= 233472, "shared memory per SM");
    const int64_t sub_rows = a.sub_rows < 0 ? a.rows : a.sub_rows;  // row sub-range of the view (default: all)
    SDB_REQUIRE(a.sub_begin % kRows == 0, SDB_STATUS_INVALID_VALUE, "spmm_slab: sub-range not aligned to the row groups");
    const int64_t* sub_indptr = a.indptr + a.sub_begin;
    row0 += a.sub_begin;
    const int64_t n_blocks = (sub_rows + kRows - 1) / kRows;
    const int live_sms = std::max(1, sm_count - t_reserved_sms);
    const unsigned grid = unsigned(std::min<int64_t>(n_blocks, int64_t(live_sms) * CTAS));
    SDB_CUDA(cudaFuncSetAttribute(spmm_stream_kernel<T, RPW, WARPS, U, CTAS, KEEP>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmem)));
    note_spmm_kernel(KEEP ? "spmm_stream_kernel<%s,%d,%d,%d,%d,keep>" : "spmm_stream_kernel<%s,%d,%d,%d,%d>",
                     dtype_cname(Num<T>::dtype), RPW, WARPS, U, CTAS);
    // panels wider than 512 bytes per row: one sweep per 512-byte column chunk (the accumulator tile is 512 B wide)
    constexpr int kChunkElems = 512 / int(sizeof(T));
    for (int c = 0; c < col_chunks; ++c) {
        SlabPeers<T> peers;
        for (int q = 0; q < kSlabMaxPeers; ++q)
            peers.y[q] = q < n_peers ? static_cast<T*>(dY_peers[q]) + int64_t(c) * kChunkElems : nullptr;
        SDB_LAUNCH((spmm_stream_kernel<T, RPW, WARPS, U, CTAS, KEEP>), grid, WARPS * 32, kSmem, s, sub_rows, sub_indptr,
                   static_cast<const uint32_t*>(m->slab_rc), static_cast<const T*>(m->slab_val),
                   X + int64_t(c) * kChunkElems, uint32_t(ldx * int64_t(sizeof(T))), alpha, beta, peers.y[self], peers,
                   n_peers, self, row0, ldy);
    }
    return SDB_STATUS_SUCCESS;
}

// 256-byte rows: two groups per warp (spmm_stream_half_kernel); rows per SM as for the 512-byte kernel
template <typename T>
static sdb_status launch_half(cudaStream_t s, const CsrView& a, const sdb_mat* m, const T* X, int64_t ldx, T alpha, T beta,
                              void* const* dY_peers, int n_peers, int self, int64_t row0, int64_t ldy, int sm_count) {
    constexpr int RPW = 6, WARPS = 16, CTAS = 4;  // 4 CTAs x 16 warps x 2 groups x 6 rows = 768 rows of 256 B per SM
    constexpr int kRows = RPW * WARPS * 2;
    constexpr size_t kSmem = size_t(kRows) * 256 + size_t(WARPS) * 32 * sizeof(StagedEntry<T>);
    static_assert((kSmem + 1024) * CTAS <= 233472, "shared memory per SM");
    const int64_t sub_rows = a.sub_rows < 0 ? a.rows : a.sub_rows;
    SDB_REQUIRE(a.sub_begin % kRows == 0, SDB_STATUS_INVALID_VALUE, "spmm_slab: sub-range not aligned to the row groups");
    const int64_t* sub_indptr = a.indptr + a.sub_begin;
    row0 += a.sub_begin;
    const int64_t n_blocks = (sub_rows + kRows - 1) / kRows;
    const int live_sms = std::max(1, sm_count - t_reserved_sms);
    const unsigned grid = unsigned(std::min<int64_t>(n_blocks, int64_t(live_sms) * CTAS));
    auto kernel = spmm_stream_half_kernel<T, RPW, WARPS, CTAS>;
    SDB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmem)));
    note_spmm_kernel("spmm_stream_half_kernel<%s,%d,%d,%d>", dtype_cname(Num<T>::dtype), RPW, WARPS, CTAS);
    SlabPeers<T> peers;
    for (int q = 0; q < kSlabMaxPeers; ++q) peers.y[q] = q < n_peers ? static_cast<T*>(dY_peers[q]) : nullptr;
    SDB_LAUNCH(kernel, grid, WARPS * 32, kSmem, s, sub_rows, sub_indptr, static_cast<const uint32_t*>(m->slab_rc),
               static_cast<const T*>(m->slab_val), X, uint32_t(ldx * int64_t(sizeof(T))), alpha, beta, peers.y[self], peers,
               n_peers, self, row0, ldy);
    return SDB_STATUS_SUCCESS;
}

#define SDB_SLAB_ARGS s, a, m, X, ldx, alpha, beta, dY_peers, n_peers, self, row0, ldy, sm_count, col_chunks
template <typename T>
static sdb_status launch_variant(cudaStream_t s, const CsrView& a, const sdb_mat* m, const T* X, int64_t ldx, T alpha,
                                 T beta, void* const* dY_peers, int n_peers, int self, int64_t row0, int64_t ldy,
                                 int sm_count, int col_chunks) {
    switch (slab_variant()) {
        case 1: return launch_stream<T, 6, 32, 3, 2>(SDB_SLAB_ARGS);
        case 2: return launch_stream<T, 8, 24, 3, 2>(SDB_SLAB_ARGS);
        case 3: return launch_stream<T, 8, 24, 4, 2>(SDB_SLAB_ARGS);
        case 4: return launch_stream<T, 8, 16, 3, 3>(SDB_SLAB_ARGS);
        case 5: return launch_stream<T, 13, 32, 4, 1>(SDB_SLAB_ARGS);
        default: {
            // option "slab_keep" (SDB_SLAB_KEEP=1): the gathers carry an L2 evict_last policy, so the X rows of the
            // active slabs outlive the lines peers write into this GPU's panel during a fused all-gather
            if (get_option(kOptSlabKeep) == 1) return launch_stream<T, 6, 32, 2, 2, true>(SDB_SLAB_ARGS);
            return launch_stream<T, 6, 32, 2, 2>(SDB_SLAB_ARGS);
        }
    }
}
#undef SDB_SLAB_ARGS

// Does this call qualify?  Shape rules always; in automatic mode also the size rules under which the
// re-ordered copy pays for itself (it costs nnz * (4 + sv) bytes of HBM and a few passes over A):
//   * X several times larger than L2 (otherwise the row-gather kernel already hits L2),
//   * enough rows to fill the persistent grid,
//   * expected reuse of an X row inside one wave of row blocks >= 1.5
//     (rows in flight * mean row length / columns),
//   * the handle has been multiplied before (a matrix used once never pays the inspector).
static bool slab_wanted(const CsrView& a, int dtype, int64_t n, int64_t ldx, bool count_call);

bool spmm_slab_wanted(const CsrView& a, int dtype, int64_t n, int64_t ldx) {
    // a row sub-range is one piece of a call that has been counted already (spmm_slab_wave_rows)
    return slab_wanted(a, dtype, n, ldx, a.sub_rows < 0);
}

int64_t spmm_slab_wave_rows(const Context* ctx, const CsrView& a, int dtype, int64_t n, int64_t ldx, bool count_call) {
    if (!slab_wanted(a, dtype, n, ldx, count_call)) return 0;
    if (a.owner->strict_sorted == -1) return 0;
    return int64_t(std::max(1, ctx->sm_count - t_reserved_sms)) * slab_rows_per_sm();
}

static bool slab_wanted(const CsrView& a, int dtype, int64_t n, int64_t ldx, bool count_call) {
    const int mode = slab_mode();
    if (mode == 1 || a.owner == nullptr) return false;
    // borrowed device arrays (sdb_create_csr_dev) may be rewritten by their owner between calls: never cache a copy
    if (!a.owner->owns) return false;
    if (dtype != SDB_F32 && dtype != SDB_F64) return false;
    const size_t sv = dtype_size(dtype);
    // 1..8 column chunks of 512 bytes, or one row of 256 bytes (two groups per warp)
    const size_t row_b = size_t(n > 0 ? n : 0) * sv;
    if (n <= 0 || !(row_b == 256 || (row_b % 512 == 0 && row_b <= 4096))) return false;
    if ((size_t(ldx) * sv) % 16 != 0 || size_t(ldx) * sv >= (size_t(1) << 31)) return false;
    if (a.cols >= (int64_t(1) << kColBits) || a.rows <= 0 || a.nnz <= 0) return false;
    if (a.owner->strict_sorted == -1) return false;
    if (mode == 2) return true;
    const int uses = count_call ? a.owner->spmm_calls++ : a.owner->spmm_calls - 1;
    Context* ctx = nullptr;
    const int sms = get_context(&ctx) == SDB_STATUS_SUCCESS ? ctx->sm_count : 148;
    const int64_t wave = int64_t(sms) * slab_rows_per_sm();  // rows one wave of the persistent grid holds
    const double reuse = double(std::min<int64_t>(a.rows, wave)) * (double(a.nnz) / double(a.rows)) / double(a.cols);
    return uses >= 1 && size_t(a.cols) * size_t(n) * sv >= (size_t(192) << 20) && a.rows >= wave && reuse >= 1.5;
}

sdb_status spmm_slab_device(Context* ctx, cudaStream_t s, const CsrView& a, int dtype, bool conj_a,
                            const double* alpha, const double* beta, const void* dX, int64_t n, int64_t ldx,
                            void* const* dY_peers, int n_peers, int self, int64_t row0, int64_t ldy) {
    (void)conj_a;  // real dtypes only
    sdb_mat* m = a.owner;
    SDB_REQUIRE(m != nullptr, SDB_STATUS_NOT_SUPPORTED, "spmm_slab: ad-hoc view");
    std::unique_lock<std::mutex> cache_lock(g_companion_mutex);  // the inspector's output is a per-handle cache
    if (m->strict_sorted == 0) {
        // the check runs on the library stream; the caller's stream must not race with it
        SDB_TRY(ensure_strict_flag(ctx, m));
    }
    if (m->strict_sorted != 1) return SDB_STATUS_NOT_SUPPORTED;
    const size_t sv = dtype_size(dtype);
    const bool half_rows = size_t(n) * sv == 256;
    const int64_t width = std::max<int64_t>(64, int64_t(slab_target_bytes() / (half_rows ? 256 : 512)));
    const int rpw = half_rows ? 6 : slab_rpw();
    if (m->slab_rc == nullptr || m->slab_width != width || m->slab_rpw != rpw) {
        // inspector: the slab-ordered copy of A, cached on the handle until sdb_order / destroy
        cudaStream_t ls = ctx->stream;
        if (m->slab_rc) cudaFreeAsync(m->slab_rc, ls);
        if (m->slab_val) cudaFreeAsync(m->slab_val, ls);
        m->slab_rc = m->slab_val = nullptr;
        SDB_TRY(dev_alloc(&m->slab_rc, size_t(a.nnz) * 4, ls));
        SDB_TRY(dev_alloc(&m->slab_val, size_t(a.nnz) * sv, ls));
        const int64_t groups = (a.rows + rpw - 1) / rpw;
        const unsigned grid = unsigned((groups * 32 + 255) / 256);
        if (dtype == SDB_F32) {
            SDB_LAUNCH(slab_permute_kernel<float>, grid, 256, 0, ls, a.rows, rpw, a.indptr, a.indices,
                       static_cast<const float*>(a.values), width, static_cast<uint32_t*>(m->slab_rc),
                       static_cast<float*>(m->slab_val));
        } else {
            SDB_LAUNCH(slab_permute_kernel<double>, grid, 256, 0, ls, a.rows, rpw, a.indptr, a.indices,
                       static_cast<const double*>(a.values), width, static_cast<uint32_t*>(m->slab_rc),
                       static_cast<double*>(m->slab_val));
        }
        m->slab_width = width;
        m->slab_rpw = rpw;
        if (s != ls) SDB_CUDA(cudaStreamSynchronize(ls));
    }
    cache_lock.unlock();
    if (half_rows) {
        if (dtype == SDB_F32)
            return launch_half<float>(s, a, m, static_cast<const float*>(dX), ldx, float(alpha[0]), float(beta[0]),
                                      dY_peers, n_peers, self, row0, ldy, ctx->sm_count);
        return launch_half<double>(s, a, m, static_cast<const double*>(dX), ldx, alpha[0], beta[0], dY_peers, n_peers,
                                   self, row0, ldy, ctx->sm_count);
    }
    const int col_chunks = int(size_t(n) * sv / 512);
    if (dtype == SDB_F32)
        return launch_variant<float>(s, a, m, static_cast<const float*>(dX), ldx, float(alpha[0]), float(beta[0]),
                                     dY_peers, n_peers, self, row0, ldy, ctx->sm_count, col_chunks);
    return launch_variant<double>(s, a, m, static_cast<const double*>(dX), ldx, alpha[0], beta[0], dY_peers, n_peers,
                                  self, row0, ldy, ctx->sm_count, col_chunks);
}

}  // namespace sdb
