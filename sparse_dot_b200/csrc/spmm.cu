// spmm.cu — Y := alpha * op(A) * X + beta * Y with A sparse (CSR view) and X, Y
// dense.  Replaces mkl_sparse_{s,d,c,z}_mm as _sparse_dense_matmul drives it
// (sparse_dot_mkl/_sparse_dense.py:34-132, the call itself at :111-123).
//
// Roofline: HBM.  Every stored entry of A gathers one n-wide row of X, so the
// algorithmic traffic per entry is (4 + sizeof(T)) + n*sizeof(T) bytes, plus
// n*sizeof(T)*(1 + [beta != 0]) per output row (gather model).
// Arithmetic intensity is ~0.5 FLOP/B: no tensor cores, the job is to keep
// enough 16-byte gathers in flight.
//
// Mapping (row-major X / Y): a group of LANES lanes owns one output row and a
// (LANES * VEC)-column chunk of it; each lane keeps VEC consecutive columns in
// registers.  The group reads LANES (column, value) pairs of the CSR row with
// one coalesced load each, then every pair is broadcast by shuffle and each
// lane issues ONE 16-byte load of its slice of X[column, :] — the group's
// loads cover a contiguous LANES*16 bytes of that X row, i.e. whole 128-byte
// lines.  kUnroll gathers are issued back to back before the first FMA so each
// warp keeps kUnroll * 512 B in flight.  The alpha/beta epilogue is fused and
// the finished row slice is stored once to every peer panel (n_peers = 1
// normally; > 1 is the fused all-gather over NVLink peer mappings, §8e).
#include <algorithm>
#include <cstdlib>

#include <mutex>

#include "common.h"
#include "prims.h"
#include "types.cuh"

namespace sdb {

constexpr int kMaxPeers = 8;
constexpr int kSpmmWarps = 8;  // warps per CTA
constexpr int kUnroll = 8;

template <typename T> struct PeerPanels {
    T* y[kMaxPeers];
};

template <typename T, int VEC> struct alignas(sizeof(T) * VEC) Pack {
    T v[VEC];
};

// read-only path; sizeof(Pack) is 4, 8 or 16 -> LDG.32 / .64 / .128
template <typename T, int VEC> __device__ __forceinline__ Pack<T, VEC> load_pack(const T* p);
template <> __device__ __forceinline__ Pack<float, 4> load_pack<float, 4>(const float* p) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    Pack<float, 4> r;
    r.v[0] = q.x, r.v[1] = q.y, r.v[2] = q.z, r.v[3] = q.w;
    return r;
}
template <> __device__ __forceinline__ Pack<double, 2> load_pack<double, 2>(const double* p) {
    const double2 q = __ldg(reinterpret_cast<const double2*>(p));
    Pack<double, 2> r;
    r.v[0] = q.x, r.v[1] = q.y;
    return r;
}
template <> __device__ __forceinline__ Pack<cf32, 2> load_pack<cf32, 2>(const cf32* p) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    Pack<cf32, 2> r;
    r.v[0] = cf32{q.x, q.y}, r.v[1] = cf32{q.z, q.w};
    return r;
}
template <> __device__ __forceinline__ Pack<cf64, 1> load_pack<cf64, 1>(const cf64* p) {
    const double2 q = __ldg(reinterpret_cast<const double2*>(p));
    Pack<cf64, 1> r;
    r.v[0] = cf64{q.x, q.y};
    return r;
}
template <> __device__ __forceinline__ Pack<float, 1> load_pack<float, 1>(const float* p) {
    Pack<float, 1> r;
    r.v[0] = __ldg(p);
    return r;
}
template <> __device__ __forceinline__ Pack<double, 1> load_pack<double, 1>(const double* p) {
    Pack<double, 1> r;
    r.v[0] = __ldg(p);
    return r;
}
template <> __device__ __forceinline__ Pack<cf32, 1> load_pack<cf32, 1>(const cf32* p) {
    const float2 q = __ldg(reinterpret_cast<const float2*>(p));
    Pack<cf32, 1> r;
    r.v[0] = cf32{q.x, q.y};
    return r;
}

template <typename T, int VEC> __device__ __forceinline__ void store_pack(T* p, const Pack<T, VEC>& a) {
    // Y is written once and not re-read by this kernel: streaming stores (evict-first)
    if constexpr (sizeof(Pack<T, VEC>) == 16) {
        __stcs(reinterpret_cast<float4*>(p), *reinterpret_cast<const float4*>(&a));
    } else if constexpr (sizeof(Pack<T, VEC>) == 8) {
        __stcs(reinterpret_cast<float2*>(p), *reinterpret_cast<const float2*>(&a));
    } else {
        *reinterpret_cast<Pack<T, VEC>*>(p) = a;
    }
}

template <typename T, int VEC, int LANES, int UNROLL = kUnroll, int MINB = 1, bool LONG = false>
__global__ void __launch_bounds__(kSpmmWarps * 32, MINB)
    spmm_rowmajor_kernel(int64_t rows, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                         const T* __restrict__ values, bool conj_a, const T* __restrict__ X, int64_t ldx, int64_t n,
                         T alpha, T beta, T* __restrict__ y_self, PeerPanels<T> out, int n_peers, int self, int64_t row0,
                         int64_t ldy, int long_row) {
    constexpr int kRowsPerWarp = 32 / LANES;
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int kU = LANES < UNROLL ? LANES : UNROLL;  // gathers in flight per lane
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES;  // position inside the row group
    const int64_t warp = int64_t(blockIdx.x) * kSpmmWarps + (threadIdx.x >> 5);
    const int64_t row = warp * kRowsPerWarp + lane / LANES;
    const int64_t col0 = (int64_t(blockIdx.y) * LANES + sub) * VEC;
    const bool row_ok = row < rows;
    const bool col_ok = col0 < n;  // n % VEC == 0 on the vector path, VEC == 1 otherwise

    // 32-bit counters relative to the row start keep the hot loop small (a row never holds 2^31 entries)
    int len = 0;
    const int32_t* ci = indices;
    const T* cv = values;
    if (row_ok) {
        const int64_t start = indptr[row];
        len = int(min(indptr[row + 1] - start, int64_t(INT32_MAX)));
        ci += start;
        cv += start;
    }
    // LONG (a separate instantiation: the default one keeps its schedule): rows longer than long_row are left to
    // spmm_long_rows_kernel, which spreads their entries over a whole CTA
    const bool skipped = LONG && len > long_row;
    if (LONG && skipped) len = 0;
    // the warp iterates together: shuffles need every lane, groups may differ in length
    int maxlen = len;
    if (LANES < 32) {
#pragma unroll
        for (int d = 16; d >= LANES; d >>= 1) maxlen = max(maxlen, __shfl_xor_sync(kFull, maxlen, d));
    }

    Pack<T, VEC> acc;
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc.v[i] = Num<T>::zero();

    const T* xcol = X + col0;
    for (int base = 0; base < maxlen; base += LANES) {
        const int mine = base + sub;
        int32_t c = 0;
        T v = Num<T>::zero();
        if (mine < len) {
            // A is streamed once: evict-first, so it does not push X rows out of L2
            c = __ldcs(ci + mine);
            v = ldcs(cv + mine);
            if (conj_a) v = conj_(v);
        }
        const int batch = min(LANES, maxlen - base);  // warp-uniform
        for (int j0 = 0; j0 < batch; j0 += kU) {
            Pack<T, VEC> x[kU];
            T a[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int j = j0 + u;  // may run past LANES: shuffles wrap, the predicate masks it
                const int32_t cj = __shfl_sync(kFull, c, j, LANES);
                a[u] = shfl(kFull, v, j, LANES);
                const bool live = col_ok && j < LANES && (base + j) < len;
                if (live) {
                    x[u] = load_pack<T, VEC>(xcol + int64_t(cj) * ldx);
                } else {
                    a[u] = Num<T>::zero();
#pragma unroll
                    for (int i = 0; i < VEC; ++i) x[u].v[i] = Num<T>::zero();
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc.v[i] = madd(a[u], x[u].v[i], acc.v[i]);
        }
    }

    if (!(row_ok && col_ok)) return;
    if (LONG && skipped) return;
    const int64_t off = (row0 + row) * ldy + col0;
    Pack<T, VEC> y;
    if (Num<T>::is_zero(beta)) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) y.v[i] = mul(alpha, acc.v[i]);
    } else {
        const Pack<T, VEC> old = *reinterpret_cast<const Pack<T, VEC>*>(y_self + off);
#pragma unroll
        for (int i = 0; i < VEC; ++i) y.v[i] = madd(alpha, acc.v[i], mul(beta, old.v[i]));
    }
    store_pack<T, VEC>(y_self + off, y);
    if (n_peers > 1) {
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)  // static indices keep the pointers in the constant bank
            if (q < n_peers && q != self) store_pack<T, VEC>(out.y[q] + off, y);
    }
}

// rows far longer than the rest (defined next to the SpMV kernels below)
constexpr int kSpmmLongRow = 1024;
static sdb_status long_rows_of(cudaStream_t s, const CsrView& a, int slot, int64_t long_row, const int32_t** list,
                               int32_t* n);
template <typename T>
static sdb_status spmm_long_rows_launch(cudaStream_t s, const CsrView& a, const int32_t* list, int32_t n_long,
                                        bool conj_a, const T* X, int64_t ldx, int64_t n, T alpha, T beta, T* Y,
                                        int64_t row0, int64_t ldy);

template <typename T, int VEC, int LANES>
static sdb_status launch_rowmajor(cudaStream_t s, const CsrView& a, bool conj_a, const T* X, int64_t ldx, int64_t n,
                                  T alpha, T beta, const PeerPanels<T>& out, int n_peers, int self, int64_t row0,
                                  int64_t ldy) {
    constexpr int kRowsPerCta = kSpmmWarps * (32 / LANES);
    // single-GPU whole-matrix products: rows beyond kSpmmLongRow entries (power-law matrices) go to their own kernel
    const int32_t* long_list = nullptr;
    int32_t n_long = 0;
    if (n_peers == 1 && a.sub_rows < 0) SDB_TRY(long_rows_of(s, a, 1, kSpmmLongRow, &long_list, &n_long));
    const int64_t sub_rows = a.sub_rows < 0 ? a.rows : a.sub_rows;  // row sub-range of the view (default: all)
    const int64_t* sub_indptr = a.indptr + a.sub_begin;
    row0 += a.sub_begin;
    const int64_t gx = (sub_rows + kRowsPerCta - 1) / kRowsPerCta;
    const int64_t gy = (n + int64_t(LANES) * VEC - 1) / (int64_t(LANES) * VEC);
    SDB_REQUIRE(gx < (int64_t(1) << 31) && gy < 65536, SDB_STATUS_NOT_SUPPORTED, "spmm: grid too large");
#define SDB_SPMM_LAUNCH(U, MB)                                                                                  \
    note_spmm_kernel("spmm_rowmajor_kernel<%s,%d,%d,%d,%d>", dtype_cname(Num<T>::dtype), VEC, LANES, U, MB);                    \
    SDB_LAUNCH((spmm_rowmajor_kernel<T, VEC, LANES, U, MB>), dim3(unsigned(gx), unsigned(gy)), kSpmmWarps * 32, 0, s, \
               sub_rows, sub_indptr, a.indices, static_cast<const T*>(a.values), conj_a, X, ldx, n, alpha, beta,    \
               out.y[self], out, n_peers, self, row0, ldy, 0)
#define SDB_SPMM_LAUNCH_LONG(U, MB)                                                                             \
    note_spmm_kernel("spmm_rowmajor_kernel<%s,%d,%d,%d,%d,long>", dtype_cname(Num<T>::dtype), VEC, LANES, U, MB);   \
    SDB_LAUNCH((spmm_rowmajor_kernel<T, VEC, LANES, U, MB, true>), dim3(unsigned(gx), unsigned(gy)), kSpmmWarps * 32, \
               0, s, sub_rows, sub_indptr, a.indices, static_cast<const T*>(a.values), conj_a, X, ldx, n, alpha,    \
               beta, out.y[self], out, n_peers, self, row0, ldy, kSpmmLongRow);                                     \
    return spmm_long_rows_launch<T>(s, a, long_list, n_long, conj_a, X, ldx, n, alpha, beta, out.y[self], row0, ldy)
    if (LANES == 32 && VEC * sizeof(T) == 16) {
        // full-warp rows (the headline shape): tuning variants selectable for experiments
        static const int tune = [] {
            const char* e = getenv("SDB_SPMM_TUNE");
            return e ? atoi(e) : 0;
        }();
        switch (tune) {
            case 1: SDB_SPMM_LAUNCH(4, 4); return SDB_STATUS_SUCCESS;
            case 2: SDB_SPMM_LAUNCH(8, 4); return SDB_STATUS_SUCCESS;
            case 3: SDB_SPMM_LAUNCH(16, 2); return SDB_STATUS_SUCCESS;
            case 4: SDB_SPMM_LAUNCH(4, 6); return SDB_STATUS_SUCCESS;
            case 5: SDB_SPMM_LAUNCH(2, 8); return SDB_STATUS_SUCCESS;
            case 6: SDB_SPMM_LAUNCH(1, 8); return SDB_STATUS_SUCCESS;
            case 7: SDB_SPMM_LAUNCH(2, 6); return SDB_STATUS_SUCCESS;
            case 8: SDB_SPMM_LAUNCH(4, 5); return SDB_STATUS_SUCCESS;
            case 9: SDB_SPMM_LAUNCH(3, 8); return SDB_STATUS_SUCCESS;
            case 10: SDB_SPMM_LAUNCH(3, 6); return SDB_STATUS_SUCCESS;
            case 11: SDB_SPMM_LAUNCH(kUnroll, 1); return SDB_STATUS_SUCCESS;
            default: break;
        }
        if (n_long > 0) {
            SDB_SPMM_LAUNCH_LONG(2, 8);
        }
        // measured on B200 (profiles/README.md, round 1): full occupancy (32 registers, 8 CTAs x 8 warps per
        // SM) with 2 gathers in flight per lane beats deeper unrolling at 3 CTAs/SM by 22 %
        SDB_SPMM_LAUNCH(2, 8);
        return SDB_STATUS_SUCCESS;
    }
    if (n_long > 0) {
        SDB_SPMM_LAUNCH_LONG(kUnroll, 1);
    }
    SDB_SPMM_LAUNCH(kUnroll, 1);
#undef SDB_SPMM_LAUNCH
#undef SDB_SPMM_LAUNCH_LONG
    return SDB_STATUS_SUCCESS;
}

template <typename T, int VEC>
static sdb_status pick_lanes(cudaStream_t s, const CsrView& a, bool conj_a, const T* X, int64_t ldx, int64_t n,
                             T alpha, T beta, const PeerPanels<T>& out, int n_peers, int self, int64_t row0,
                             int64_t ldy) {
    const int64_t packs = (n + VEC - 1) / VEC;
#define SDB_GO(L) return launch_rowmajor<T, VEC, L>(s, a, conj_a, X, ldx, n, alpha, beta, out, n_peers, self, row0, ldy)
    if (packs > 16) SDB_GO(32);
    if (packs > 8) SDB_GO(16);
    if (packs > 4) SDB_GO(8);
    if (packs > 2) SDB_GO(4);
    if (packs > 1) SDB_GO(2);
    SDB_GO(1);
#undef SDB_GO
}

// ------------------------------------------------------------------ SpMV (n == 1)
// mkl_sparse_?_mv (_sparse_vector.py:20-25,84-92): y = alpha * op(A) * x + beta * y.
// A group of LANES lanes reduces one row: lanes stride the row's entries (coalesced index / value
// loads, gathered x), then a shuffle tree adds the partial sums.  LANES follows the mean row length.
template <typename T, int LANES, bool LONG>
__global__ void __launch_bounds__(256) spmv_kernel(int64_t rows, const int64_t* __restrict__ indptr,
                                                   const int32_t* __restrict__ indices,
                                                   const T* __restrict__ values, bool conj_a,
                                                   const T* __restrict__ x, int64_t incx, T alpha, T beta,
                                                   T* __restrict__ y, int64_t incy, int64_t long_row) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES;
    int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) / LANES;
    // LONG: rows longer than long_row belong to spmv_long_rows_kernel (a separate instantiation: the check costs
    // registers, and with them occupancy, that matrices without such rows should not pay)
    if (LONG && row < rows && indptr[row + 1] - indptr[row] > long_row) row = rows;
    T acc = Num<T>::zero();
    if (row < rows) {
        const int64_t e = indptr[row + 1];
        for (int64_t p = indptr[row] + sub; p < e; p += LANES) {
            T v = ldg(values + p);
            if (conj_a) v = conj_(v);
            acc = madd(v, ldg(x + int64_t(__ldg(indices + p)) * incx), acc);
        }
    }
#pragma unroll
    for (int d = LANES / 2; d > 0; d >>= 1) acc = add(acc, shfl(0xffffffffu, acc, (lane ^ d), 32));
    if (row < rows && sub == 0) {
        T* out = y + row * incy;
        *out = Num<T>::is_zero(beta) ? mul(alpha, acc) : madd(alpha, acc, mul(beta, *out));
    }
}

// Four consecutive stored values / column indices with 16-byte streaming loads (read exactly once).  `p` must be
// 16-byte aligned: entry index a multiple of 4 from a 16-byte aligned base.
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    const float4 q = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
}
__device__ __forceinline__ void load4(const double* p, double (&v)[4]) {
    const double2 q0 = __ldcs(reinterpret_cast<const double2*>(p)), q1 = __ldcs(reinterpret_cast<const double2*>(p) + 1);
    v[0] = q0.x, v[1] = q0.y, v[2] = q1.x, v[3] = q1.y;
}
__device__ __forceinline__ void load4(const cf32* p, cf32 (&v)[4]) {
    const float4 q0 = __ldcs(reinterpret_cast<const float4*>(p)), q1 = __ldcs(reinterpret_cast<const float4*>(p) + 1);
    v[0] = cf32{q0.x, q0.y}, v[1] = cf32{q0.z, q0.w}, v[2] = cf32{q1.x, q1.y}, v[3] = cf32{q1.z, q1.w};
}
__device__ __forceinline__ void load4(const cf64* p, cf64 (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double2 q = __ldcs(reinterpret_cast<const double2*>(p) + i);
        v[i] = cf64{q.x, q.y};
    }
}

// The same product with 16-byte loads of A: every lane takes four consecutive entries per step (one int4 of column
// indices, 16-64 bytes of values), so a warp has 4x the bytes in flight of spmv_kernel for the same occupancy — what
// an HBM-bound stream of A wants.  Rows start anywhere: the entries before the first multiple-of-4 position and
// after the last whole pack (at most 3 + 3) are taken one per lane.
template <typename T, int LANES, bool LONG>
__global__ void __launch_bounds__(256) spmv_wide_kernel(int64_t rows, const int64_t* __restrict__ indptr,
                                                        const int32_t* __restrict__ indices,
                                                        const T* __restrict__ values, bool conj_a,
                                                        const T* __restrict__ x, int64_t incx, T alpha, T beta,
                                                        T* __restrict__ y, int64_t incy, int64_t long_row) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES;
    int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) / LANES;
    if (LONG && row < rows && indptr[row + 1] - indptr[row] > long_row) row = rows;
    T acc = Num<T>::zero();
    if (row < rows) {
        const int64_t b = indptr[row], e = indptr[row + 1];
        const int64_t body = min(e, (b + 3) & ~int64_t(3));
        const int64_t tail = body + ((e - body) & ~int64_t(3));
        for (int64_t p = body + 4 * sub; p < tail; p += 4 * LANES) {
            const int4 c = __ldcs(reinterpret_cast<const int4*>(indices + p));
            T v[4];
            load4(values + p, v);
            if (conj_a) {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = conj_(v[i]);
            }
            acc = madd(v[0], ldg(x + int64_t(c.x) * incx), acc);
            acc = madd(v[1], ldg(x + int64_t(c.y) * incx), acc);
            acc = madd(v[2], ldg(x + int64_t(c.z) * incx), acc);
            acc = madd(v[3], ldg(x + int64_t(c.w) * incx), acc);
        }
        for (int64_t p = b + sub; p < body; p += LANES) {
            T v = ldg(values + p);
            if (conj_a) v = conj_(v);
            acc = madd(v, ldg(x + int64_t(__ldg(indices + p)) * incx), acc);
        }
        for (int64_t p = tail + sub; p < e; p += LANES) {
            T v = ldg(values + p);
            if (conj_a) v = conj_(v);
            acc = madd(v, ldg(x + int64_t(__ldg(indices + p)) * incx), acc);
        }
    }
#pragma unroll
    for (int d = LANES / 2; d > 0; d >>= 1) acc = add(acc, shfl(0xffffffffu, acc, (lane ^ d), 32));
    if (row < rows && sub == 0) {
        T* out = y + row * incy;
        *out = Num<T>::is_zero(beta) ? mul(alpha, acc) : madd(alpha, acc, mul(beta, *out));
    }
}

// ---- rows far longer than the rest (power-law matrices: R-MAT scale 22 has rows of 10^5 entries next to a mean of 4)
// A group of 2-32 lanes would walk such a row alone while the rest of the machine idles (measured: 1.89 ms for 16.6 M
// entries).  They are listed once per handle and reduced by one CTA each; the group kernels skip them.
constexpr int kLongThreads = 1024;

__global__ void find_long_rows_kernel(int64_t rows, const int64_t* __restrict__ indptr, int64_t long_row,
                                      int32_t capacity, int32_t* __restrict__ count, int32_t* __restrict__ list) {
    const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r < rows && indptr[r + 1] - indptr[r] > long_row) {
        const int32_t at = atomicAdd(count, 1);
        if (at < capacity) list[at] = int32_t(r);
    }
}

template <typename T>
__global__ void __launch_bounds__(kLongThreads) spmv_long_rows_kernel(const int32_t* __restrict__ list, int32_t n_long,
                                                                    const int64_t* __restrict__ indptr,
                                                                    const int32_t* __restrict__ indices,
                                                                    const T* __restrict__ values, bool conj_a,
                                                                    const T* __restrict__ x, int64_t incx, T alpha,
                                                                    T beta, T* __restrict__ y, int64_t incy) {
    __shared__ T part[kLongThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int32_t i = blockIdx.x; i < n_long; i += gridDim.x) {
        const int64_t row = list[i];
        const int64_t b = indptr[row], e = indptr[row + 1];
        T acc0 = Num<T>::zero(), acc1 = Num<T>::zero();
        int64_t p = b + threadIdx.x;
        for (; p + kLongThreads < e; p += 2 * kLongThreads) {  // two independent chains per thread
            T v0 = ldg(values + p), v1 = ldg(values + p + kLongThreads);
            if (conj_a) v0 = conj_(v0), v1 = conj_(v1);
            acc0 = madd(v0, ldg(x + int64_t(__ldg(indices + p)) * incx), acc0);
            acc1 = madd(v1, ldg(x + int64_t(__ldg(indices + p + kLongThreads)) * incx), acc1);
        }
        if (p < e) {
            T v = ldg(values + p);
            if (conj_a) v = conj_(v);
            acc0 = madd(v, ldg(x + int64_t(__ldg(indices + p)) * incx), acc0);
        }
        T acc = add(acc0, acc1);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc = add(acc, shfl(0xffffffffu, acc, lane ^ d, 32));
        if (lane == 0) part[warp] = acc;
        __syncthreads();
        if (warp == 0) {
            acc = part[lane];  // kLongThreads / 32 == 32 partial sums
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) acc = add(acc, shfl(0xffffffffu, acc, lane ^ d, 32));
            if (lane == 0) {
                T* out = y + row * incy;
                *out = Num<T>::is_zero(beta) ? mul(alpha, acc) : madd(alpha, acc, mul(beta, *out));
            }
        }
        __syncthreads();
    }
}

// The handle's list of rows longer than `long_row` (built on the first product with a vector; dropped with the
// other per-handle caches).  Returns the count, 0 when there is none or the view has no owning handle.
// slot 0: the SpMV kernels' threshold, slot 1: the SpMM kernel's
static sdb_status long_rows_of(cudaStream_t s, const CsrView& a, int slot, int64_t long_row, const int32_t** list,
                               int32_t* n) {
    *list = nullptr;
    *n = 0;
    sdb_mat* m = a.owner;
    if (m == nullptr || !m->owns || a.sub_rows >= 0 || a.rows >= (int64_t(1) << 31)) return SDB_STATUS_SUCCESS;
    std::lock_guard<std::mutex> cache_lock(g_companion_mutex);
    if (m->long_state[slot] != 1 || m->long_threshold[slot] != long_row) {
        if (m->long_rows[slot]) cudaFreeAsync(m->long_rows[slot], s);
        m->long_rows[slot] = nullptr;
        m->n_long[slot] = 0;
        const int64_t capacity = std::min<int64_t>(a.nnz / (long_row + 1) + 1, int64_t(1) << 30);
        DevBuf count;
        SDB_TRY(count.alloc(4, s));
        SDB_CUDA(cudaMemsetAsync(count.p, 0, 4, s));
        int32_t* d_list = nullptr;
        SDB_TRY(dev_alloc(reinterpret_cast<void**>(&d_list), size_t(capacity) * 4, s));
        SDB_LAUNCH(find_long_rows_kernel, unsigned((a.rows + 255) / 256), 256, 0, s, a.rows, a.indptr, long_row,
                   int32_t(capacity), count.as<int32_t>(), d_list);
        int32_t found = 0;
        if (cudaMemcpyAsync(&found, count.p, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) {
            cudaFreeAsync(d_list, s);
            set_error("spmv: reading the long-row count failed: %s", cudaGetErrorString(cudaGetLastError()));
            return SDB_STATUS_EXECUTION_FAILED;
        }
        if (found == 0) {
            cudaFreeAsync(d_list, s);
            d_list = nullptr;
        }
        m->long_rows[slot] = d_list;
        m->n_long[slot] = int32_t(std::min<int64_t>(found, capacity));
        m->long_threshold[slot] = long_row;
        m->long_state[slot] = 1;
    }
    *list = m->long_rows[slot];
    *n = m->n_long[slot];
    return SDB_STATUS_SUCCESS;
}

// SpMM on a long row: the CTA's threads form G groups of cw columns; group g takes entries g, g + G, ... of the row
// (four gathers in flight per thread, X rows read coalesced across the group), the partial rows are added up in
// shared memory.  Panels wider than cw columns are swept in chunks.
template <typename T>
__global__ void __launch_bounds__(kLongThreads) spmm_long_rows_kernel(const int32_t* __restrict__ list, int32_t n_long,
                                                                    const int64_t* __restrict__ indptr,
                                                                    const int32_t* __restrict__ indices,
                                                                    const T* __restrict__ values, bool conj_a,
                                                                    const T* __restrict__ X, int64_t ldx, int64_t n,
                                                                    int cw, T alpha, T beta, T* __restrict__ Y,
                                                                    int64_t row0, int64_t ldy) {
    __shared__ T partial[kLongThreads];
    const int groups = kLongThreads / cw;
    const int c = threadIdx.x % cw, g = threadIdx.x / cw;
    for (int32_t i = blockIdx.x; i < n_long; i += gridDim.x) {
        const int64_t row = list[i];
        const int64_t b = indptr[row], e = indptr[row + 1];
        for (int64_t cb = 0; cb < n; cb += cw) {
            const int64_t col = cb + c;
            const bool live = col < n && g < groups;
            T acc[4] = {Num<T>::zero(), Num<T>::zero(), Num<T>::zero(), Num<T>::zero()};
            if (live) {
                int64_t p = b + g;
                for (; p + 3 * int64_t(groups) < e; p += 4 * int64_t(groups)) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        T v = ldg(values + p + u * int64_t(groups));
                        if (conj_a) v = conj_(v);
                        acc[u] = madd(v, ldg(X + int64_t(__ldg(indices + p + u * int64_t(groups))) * ldx + col), acc[u]);
                    }
                }
                for (; p < e; p += groups) {
                    T v = ldg(values + p);
                    if (conj_a) v = conj_(v);
                    acc[0] = madd(v, ldg(X + int64_t(__ldg(indices + p)) * ldx + col), acc[0]);
                }
            }
            partial[threadIdx.x] = add(add(acc[0], acc[1]), add(acc[2], acc[3]));
            __syncthreads();
            if (g == 0 && col < n) {
                T sum = partial[c];
                for (int k = 1; k < groups; ++k) sum = add(sum, partial[k * cw + c]);
                T* out = Y + (row0 + row) * ldy + col;
                *out = Num<T>::is_zero(beta) ? mul(alpha, sum) : madd(alpha, sum, mul(beta, *out));
            }
            __syncthreads();
        }
    }
}

template <typename T>
static sdb_status spmm_long_rows_launch(cudaStream_t s, const CsrView& a, const int32_t* list, int32_t n_long,
                                        bool conj_a, const T* X, int64_t ldx, int64_t n, T alpha, T beta, T* Y,
                                        int64_t row0, int64_t ldy) {
    if (n_long <= 0) return SDB_STATUS_SUCCESS;
    const int cw = int(std::min<int64_t>(128, (n + 31) / 32 * 32));
    SDB_LAUNCH((spmm_long_rows_kernel<T>), unsigned(std::min<int32_t>(n_long, 1024)), kLongThreads, 0, s, list, n_long,
               a.indptr, a.indices, static_cast<const T*>(a.values), conj_a, X, ldx, n, cw, alpha, beta, Y, row0, ldy);
    return SDB_STATUS_SUCCESS;
}

constexpr int kLongPerLane = 512;  // a row is "long" beyond this many entries per lane of its group

template <typename T>
static sdb_status long_rows_launch(cudaStream_t s, const CsrView& a, const int32_t* list, int32_t n_long, bool conj_a,
                                   const void* dX, int64_t incx, T alpha, T beta, void* dY, int64_t incy) {
    if (n_long <= 0) return SDB_STATUS_SUCCESS;
    SDB_LAUNCH((spmv_long_rows_kernel<T>), unsigned(std::min<int32_t>(n_long, 1024)), kLongThreads, 0, s, list, n_long,
               a.indptr, a.indices, static_cast<const T*>(a.values), conj_a, static_cast<const T*>(dX), incx, alpha,
               beta, static_cast<T*>(dY), incy);
    return SDB_STATUS_SUCCESS;
}

template <typename T>
static sdb_status spmv(cudaStream_t s, const CsrView& a, bool conj_a, const double* alpha_d, const double* beta_d,
                       const void* dX, int64_t incx, void* dY, int64_t incy) {
    const T alpha = Num<T>::make(alpha_d[0], alpha_d[1]), beta = Num<T>::make(beta_d[0], beta_d[1]);
    const double mean = a.rows > 0 ? double(a.nnz) / double(a.rows) : 0.0;
    // rows long enough to fill packs of four, arrays aligned for 16-byte loads: the wide kernel ("spmv_wide" 1 = never)
    const bool packs_ok = ((reinterpret_cast<uintptr_t>(a.indices) | reinterpret_cast<uintptr_t>(a.values)) & 15u) == 0;
    if (mean > 6 && packs_ok && get_option(kOptSpmvWide) != 1) {
#define SDB_SPMV_WIDE(L)                                                                                       \
    do {                                                                                                       \
        const int64_t blocks = (a.rows * L + 255) / 256;                                                       \
        SDB_REQUIRE(blocks < (int64_t(1) << 31), SDB_STATUS_NOT_SUPPORTED, "spmv: grid too large");             \
        const int64_t long_row = int64_t(kLongPerLane) * L;                                                    \
        const int32_t* long_list = nullptr;                                                                    \
        int32_t n_long = 0;                                                                                    \
        SDB_TRY(long_rows_of(s, a, 0, long_row, &long_list, &n_long));                                            \
        if (n_long > 0)                                                                                        \
            SDB_LAUNCH((spmv_wide_kernel<T, L, true>), unsigned(blocks), 256, 0, s, a.rows, a.indptr, a.indices, \
                       static_cast<const T*>(a.values), conj_a, static_cast<const T*>(dX), incx, alpha, beta,  \
                       static_cast<T*>(dY), incy, long_row);                                                   \
        else                                                                                                   \
            SDB_LAUNCH((spmv_wide_kernel<T, L, false>), unsigned(blocks), 256, 0, s, a.rows, a.indptr, a.indices, \
                       static_cast<const T*>(a.values), conj_a, static_cast<const T*>(dX), incx, alpha, beta,  \
                       static_cast<T*>(dY), incy, 0);                                                          \
        note_spmm_kernel("spmv_wide_kernel<%s,%d,%d>", dtype_cname(Num<T>::dtype), L, n_long > 0 ? 1 : 0);     \
        return long_rows_launch<T>(s, a, long_list, n_long, conj_a, dX, incx, alpha, beta, dY, incy);                                                                             \
    } while (0)
        if (mean > 96) SDB_SPMV_WIDE(32);
        if (mean > 48) SDB_SPMV_WIDE(16);
        if (mean > 24) SDB_SPMV_WIDE(8);
        if (mean > 12) SDB_SPMV_WIDE(4);
        SDB_SPMV_WIDE(2);
#undef SDB_SPMV_WIDE
    }
#define SDB_SPMV(L)                                                                                            \
    do {                                                                                                       \
        const int64_t blocks = (a.rows * L + 255) / 256;                                                       \
        SDB_REQUIRE(blocks < (int64_t(1) << 31), SDB_STATUS_NOT_SUPPORTED, "spmv: grid too large");             \
        const int64_t long_row = int64_t(kLongPerLane) * L;                                                    \
        const int32_t* long_list = nullptr;                                                                    \
        int32_t n_long = 0;                                                                                    \
        SDB_TRY(long_rows_of(s, a, 0, long_row, &long_list, &n_long));                                            \
        if (n_long > 0)                                                                                        \
            SDB_LAUNCH((spmv_kernel<T, L, true>), unsigned(blocks), 256, 0, s, a.rows, a.indptr, a.indices,    \
                       static_cast<const T*>(a.values), conj_a, static_cast<const T*>(dX), incx, alpha, beta,  \
                       static_cast<T*>(dY), incy, long_row);                                                   \
        else                                                                                                   \
            SDB_LAUNCH((spmv_kernel<T, L, false>), unsigned(blocks), 256, 0, s, a.rows, a.indptr, a.indices,   \
                       static_cast<const T*>(a.values), conj_a, static_cast<const T*>(dX), incx, alpha, beta,  \
                       static_cast<T*>(dY), incy, 0);                                                          \
        note_spmm_kernel("spmv_kernel<%s,%d,%d>", dtype_cname(Num<T>::dtype), L, n_long > 0 ? 1 : 0);          \
        return long_rows_launch<T>(s, a, long_list, n_long, conj_a, dX, incx, alpha, beta, dY, incy);                                                                             \
    } while (0)
    if (mean > 48) SDB_SPMV(32);
    if (mean > 24) SDB_SPMV(16);
    if (mean > 12) SDB_SPMV(8);
    if (mean > 6) SDB_SPMV(4);
    SDB_SPMV(2);
#undef SDB_SPMV
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
static sdb_status spmm_rowmajor(cudaStream_t s, const CsrView& a, bool conj_a, const double* alpha_d,
                                const double* beta_d, const void* dX, int64_t n, int64_t ldx, void* const* dY_peers,
                                int n_peers, int self, int64_t row0, int64_t ldy) {
    const T alpha = Num<T>::make(alpha_d[0], alpha_d[1]);
    const T beta = Num<T>::make(beta_d[0], beta_d[1]);
    PeerPanels<T> out;
    bool vec_ok = aligned16(dX);
    for (int q = 0; q < kMaxPeers; ++q) {
        out.y[q] = q < n_peers ? static_cast<T*>(dY_peers[q]) : nullptr;
        if (q < n_peers) vec_ok = vec_ok && aligned16(dY_peers[q]);
    }
    constexpr int kVec = 16 / int(sizeof(T));
    vec_ok = vec_ok && (n % kVec == 0) && (ldx % kVec == 0) && (ldy % kVec == 0);
    if (kVec > 1 && vec_ok)
        return pick_lanes<T, kVec>(s, a, conj_a, static_cast<const T*>(dX), ldx, n, alpha, beta, out, n_peers, self,
                                   row0, ldy);
    return pick_lanes<T, 1>(s, a, conj_a, static_cast<const T*>(dX), ldx, n, alpha, beta, out, n_peers, self, row0,
                            ldy);
}

// Y (m x n), X (k x n), both in `layout`.  Column-major panels are turned
// row-major on the device (tiled transposes, coalesced both ways), run through
// the same gather kernel and turned back: the gather wants X rows contiguous.
sdb_status spmm_device(Context* ctx, cudaStream_t s, const CsrView& a, int dtype, bool conj_a, const double* alpha,
                       const double* beta, int layout, const void* dX, int64_t n, int64_t ldx,
                       void* const* dY_peers, int n_peers, int self, int64_t row0, int64_t ldy) {
    SDB_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers && self >= 0 && self < n_peers, SDB_STATUS_INVALID_VALUE,
                "spmm: bad peer configuration (%d peers, self %d)", n_peers, self);
    if (a.rows == 0 || n == 0) return SDB_STATUS_SUCCESS;
    if (n == 1 && n_peers == 1 && row0 == 0 && a.sub_rows < 0) {
        // one column: the panel layout only decides the element strides
        const bool rm = layout == SDB_LAYOUT_ROW_MAJOR;
        // repeated products with a vector: x staged in shared memory (spmv_tile.cu) unless the inspector declines
        if (!conj_a && spmv_tile_wanted(a, dtype, rm ? ldx : 1, rm ? ldy : 1)) {
            const sdb_status st = spmv_tile_device(ctx, s, a, dtype, alpha, beta, dX, dY_peers[0]);
            if (st != SDB_STATUS_NOT_SUPPORTED) return st;
        }
        return SDB_DISPATCH_DTYPE(dtype, T, [&]() -> sdb_status {
            return spmv<T>(s, a, conj_a, alpha, beta, dX, rm ? ldx : 1, dY_peers[0], rm ? ldy : 1);
        });
    }
    if (layout == SDB_LAYOUT_ROW_MAJOR) {
        SDB_REQUIRE(ldx >= n && ldy >= n, SDB_STATUS_INVALID_VALUE, "spmm: leading dimension smaller than n");
        // panels far larger than L2: the slab-tiled kernel (spmm_slab.cu) when the shape qualifies
        if (spmm_slab_wanted(a, dtype, n, ldx) && aligned16(dX) && (ldy * int64_t(dtype_size(dtype))) % 16 == 0) {
            bool ok = true;
            for (int q = 0; q < n_peers; ++q) ok = ok && aligned16(dY_peers[q]);
            if (ok) {
                const sdb_status st = spmm_slab_device(ctx, s, a, dtype, conj_a, alpha, beta, dX, n, ldx, dY_peers,
                                                       n_peers, self, row0, ldy);
                if (st != SDB_STATUS_NOT_SUPPORTED) return st;
            }
        }
        return SDB_DISPATCH_DTYPE(dtype, T, [&]() -> sdb_status {
            return spmm_rowmajor<T>(s, a, conj_a, alpha, beta, dX, n, ldx, dY_peers, n_peers, self, row0, ldy);
        });
    }
    SDB_REQUIRE(layout == SDB_LAYOUT_COL_MAJOR, SDB_STATUS_INVALID_VALUE, "spmm: bad layout %d", layout);
    SDB_REQUIRE(n_peers == 1 && row0 == 0, SDB_STATUS_NOT_SUPPORTED, "spmm: peer panels must be row-major");
    SDB_REQUIRE(ldx >= a.cols && ldy >= a.rows, SDB_STATUS_INVALID_VALUE,
                "spmm: leading dimension smaller than the column length");
    const int es = int(dtype_size(dtype));
    const bool beta_zero = beta[0] == 0.0 && beta[1] == 0.0;
    DevBuf xr, yr;
    SDB_TRY(xr.alloc(size_t(a.cols) * size_t(n) * es, s));
    SDB_TRY(yr.alloc(size_t(a.rows) * size_t(n) * es, s));
    // X col-major (k x n, ld) is a row-major (n x k, ld) array: transpose it to (k x n, n)
    SDB_TRY(transpose_dense(s, dX, ldx, xr.p, n, n, a.cols, es));
    if (!beta_zero) SDB_TRY(transpose_dense(s, dY_peers[0], ldy, yr.p, n, n, a.rows, es));
    void* yp[1] = {yr.p};
    SDB_TRY(SDB_DISPATCH_DTYPE(dtype, T, [&]() -> sdb_status {
        return spmm_rowmajor<T>(s, a, conj_a, alpha, beta, xr.p, n, n, yp, 1, 0, 0, n);
    }));
    return transpose_dense(s, yr.p, n, dY_peers[0], ldy, a.rows, n, es);
}

// Exchange strategy "sm": finished rows are pushed to the peers by a few COPIER CTAs instead of the copy engines:
// every 16-byte pack of the chunk is read from the local panel ONCE and stored into each of the n_dst peer
// panels (the copy engines read it once per peer), with streaming loads / stores so neither side's L2 keeps it.
// The streaming SpMM kernel leaves `reserve` SMs free for these CTAs (spmm_slab.cu, spmm_slab_reserve_sms).
struct PushTargets {
    uint4* dst[kMaxPeers];
};
constexpr int kPushThreads = 512, kPushUnroll = 4;
__global__ void __launch_bounds__(kPushThreads) push_rows_kernel(const uint4* __restrict__ src, PushTargets t, int n_dst,
                                                                 size_t n_vec) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; i + (kPushUnroll - 1) * stride < n_vec; i += kPushUnroll * stride) {
        uint4 v[kPushUnroll];
#pragma unroll
        for (int u = 0; u < kPushUnroll; ++u) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q) {
            if (q < n_dst) {
#pragma unroll
                for (int u = 0; u < kPushUnroll; ++u) __stcs(t.dst[q] + i + u * stride, v[u]);
            }
        }
    }
    for (; i < n_vec; i += stride) {
        const uint4 v = __ldcs(src + i);
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)
            if (q < n_dst) __stcs(t.dst[q] + i, v);
    }
}

}  // namespace sdb

using namespace sdb;

static sdb_status check_spmm_args(int op, const double* alpha, const sdb_mat* A, int layout, const void* X,
                                  const double* beta, const void* Y) {
    SDB_REQUIRE(A != nullptr, SDB_STATUS_NOT_INITIALIZED, "spmm: null handle");
    SDB_REQUIRE(valid(A), SDB_STATUS_INVALID_VALUE, "spmm: not a live sdb_mat handle");
    SDB_REQUIRE(op == SDB_OP_NON_TRANSPOSE || op == SDB_OP_TRANSPOSE || op == SDB_OP_CONJUGATE_TRANSPOSE,
                SDB_STATUS_INVALID_VALUE, "spmm: bad operation %d", op);
    SDB_REQUIRE(layout == SDB_LAYOUT_ROW_MAJOR || layout == SDB_LAYOUT_COL_MAJOR, SDB_STATUS_INVALID_VALUE,
                "spmm: bad layout %d", layout);
    SDB_REQUIRE(alpha && beta && X && Y, SDB_STATUS_INVALID_VALUE, "spmm: null argument");
    return SDB_STATUS_SUCCESS;
}

extern "C" {

sdb_status sdb_spmm_dev(int op, const double* alpha, const sdb_mat* A, int layout, const void* dX, int64_t n,
                        int64_t ldx, const double* beta, void* dY, int64_t ldy, void* stream) {
    SDB_TRY(check_spmm_args(op, alpha, A, layout, dX, beta, dY));
    SDB_REQUIRE(n >= 0, SDB_STATUS_INVALID_VALUE, "spmm: negative n");
    Context* ctx;
    SDB_TRY(get_context(&ctx));
    if (spmm_bsr_supported(A, op, layout, dX, n, ldx, dY, ldy)) {
        SDB_REQUIRE(ldx >= n && ldy >= n, SDB_STATUS_INVALID_VALUE, "spmm: leading dimension smaller than n");
        return spmm_bsr_device(stream ? static_cast<cudaStream_t>(stream) : ctx->stream, A, alpha, beta, dX, n, ldx,
                               dY, ldy);
    }
    CsrView v;
    SDB_TRY(csr_view(ctx, A, op != SDB_OP_NON_TRANSPOSE, &v));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    if (s != ctx->stream) {
        // companions (transpose / expansion) are built on the library stream
        SDB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    const bool conj_a = op == SDB_OP_CONJUGATE_TRANSPOSE;
    void* yp[1] = {dY};
    return spmm_device(ctx, s, v, A->dtype, conj_a, alpha, beta, layout, dX, n, ldx, yp, 1, 0, 0, ldy);
}

// How the finished rows reach the peers (SDB_ALLGATHER overrides the automatic choice): "ce" = chunk-pipelined
// copy-engine exchange, "stores" = the kernel's epilogue stores every finished 16-byte slice into every peer panel
// itself, "k1" = the same with the row-gather kernel K1 even where the streaming kernel would qualify.
enum { kXchgAuto = 0, kXchgCopyEngines = 1, kXchgStores = 2, kXchgStoresRowGather = 3, kXchgSmCopier = 4 };
// process-wide; sdb_set_allgather overrides the environment (RowShardedSpMM.autotune measures the strategies on
// the live topology during warm-up and keeps the fastest)
static std::atomic<int> g_xchg_strategy{-1};
static std::atomic<int> g_xchg_chunks{0};  // target number of row chunks of the "ce" / "sm" pipelines (0 = about five)
static std::atomic<int> g_xchg_sms{12};    // SMs the "sm" strategy keeps free of SpMM CTAs for its copier CTAs
static int allgather_strategy() {
    int v = g_xchg_strategy.load(std::memory_order_relaxed);
    if (v >= 0) return v;
    const char* e = getenv("SDB_ALLGATHER");
    v = kXchgAuto;
    if (e && e[0] == 'c') v = kXchgCopyEngines;      // "ce"
    if (e && e[0] == 's') v = e[1] == 'm' ? kXchgSmCopier : kXchgStores;  // "sm" / "stores"
    if (e && e[0] == 'k') v = kXchgStoresRowGather;  // "k1": epilogue stores, row-gather kernel K1
    g_xchg_strategy.store(v, std::memory_order_relaxed);
    return v;
}

static sdb_status ensure_exchange_streams(Context* ctx, int n_peers) {
    for (int q = 0; q < n_peers; ++q) {
        if (!ctx->xchg_stream[q]) SDB_CUDA(cudaStreamCreateWithFlags(&ctx->xchg_stream[q], cudaStreamNonBlocking));
        if (!ctx->xchg_done[q]) SDB_CUDA(cudaEventCreateWithFlags(&ctx->xchg_done[q], cudaEventDisableTiming));
    }
    for (int e = 0; e < Context::kExchangeEvents; ++e)
        if (!ctx->xchg_event[e]) SDB_CUDA(cudaEventCreateWithFlags(&ctx->xchg_event[e], cudaEventDisableTiming));
    return SDB_STATUS_SUCCESS;
}

// Row-sharded SpMM + all-gather of the output panel as ONE stream-ordered call.
//
// Strategy "ce": the shard's rows are cut into a few chunks (whole waves of the streaming kernel's persistent
// grid, so no chunk ends in a partial wave except the last); chunk c's kernel writes only the local panel, and
// as soon as it has finished (event) the N-1 copy engines push its rows into the peers' panels over NVLink on
// one stream per peer while chunk c+1's kernel runs.  No SM ever waits for NVLink, and only the last chunk's push
// is exposed.  The caller's stream waits for the pushes at the end, so "the stream has reached this point" still
// means "my rows are in every panel".
// Strategy "stores": one kernel, peer stores in its epilogue (fine while the kernel is much longer than the
// exchange; with the streaming kernel all warps reach their epilogues together and the bursts stall the gathers).
sdb_status sdb_spmm_dev_allgather(const double* alpha, const sdb_mat* A, const void* dX, int64_t n, int64_t ldx,
                                  const double* beta, void* const* dY_peers, int n_peers, int self, int64_t row0,
                                  int64_t ldy, void* stream) {
    SDB_REQUIRE(dY_peers != nullptr && n_peers >= 1 && n_peers <= kMaxPeers && self >= 0 && self < n_peers,
                SDB_STATUS_INVALID_VALUE, "spmm_allgather: bad peer list");
    SDB_TRY(check_spmm_args(SDB_OP_NON_TRANSPOSE, alpha, A, SDB_LAYOUT_ROW_MAJOR, dX, beta, dY_peers[self]));
    for (int q = 0; q < n_peers; ++q)
        SDB_REQUIRE(dY_peers[q] != nullptr, SDB_STATUS_INVALID_VALUE, "spmm_allgather: null peer panel %d", q);
    SDB_REQUIRE(n >= 0 && row0 >= 0, SDB_STATUS_INVALID_VALUE, "spmm_allgather: negative size");
    Context* ctx;
    SDB_TRY(get_context(&ctx));
    CsrView v;
    SDB_TRY(csr_view(ctx, A, false, &v));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    if (s != ctx->stream) SDB_CUDA(cudaStreamSynchronize(ctx->stream));
    // Automatic choice, from the measurements in DESIGN.md §5: up to 4 ranks the copy engines keep up (N = 2: 2.78,
    // N = 4: 3.87 ms/step against 3.13 / 4.99 with epilogue stores); at 8 ranks they do not (8.74 ms/step for the
    // 3.6 GB every rank receives), while the row-gather kernel K1 with epilogue stores — whose CTAs finish
    // continuously, so its peer stores are spread over the whole kernel — ran the same exchange in 5.52 ms/step.
    int strategy = allgather_strategy();
    if (strategy == kXchgAuto) strategy = n_peers <= 4 ? kXchgCopyEngines : kXchgStoresRowGather;
    const bool chunked = strategy == kXchgCopyEngines || strategy == kXchgSmCopier;
    if (n_peers == 1 || !chunked || v.rows == 0 || n == 0) {
        if (n_peers > 1 && strategy == kXchgStoresRowGather) v.owner = nullptr;  // ad-hoc view: never the streaming kernel
        return spmm_device(ctx, s, v, A->dtype, false, alpha, beta, SDB_LAYOUT_ROW_MAJOR, dX, n, ldx, dY_peers,
                           n_peers, self, row0, ldy);
    }

    SDB_TRY(ensure_exchange_streams(ctx, n_peers));
    const size_t row_bytes = size_t(n) * dtype_size(A->dtype);
    const size_t pitch = size_t(ldy) * dtype_size(A->dtype);
    // the copier kernel moves whole packed rows in 16-byte packs; anything else goes through the copy engines
    bool sm_push = strategy == kXchgSmCopier && pitch == row_bytes && row_bytes % 16 == 0;
    for (int q = 0; q < n_peers && sm_push; ++q) sm_push = aligned16(dY_peers[q]);
    const int reserve = sm_push ? std::max(1, std::min(ctx->sm_count / 2, g_xchg_sms.load(std::memory_order_relaxed))) : 0;
    spmm_slab_reserve_sms(reserve);
    struct Unreserve {
        ~Unreserve() { spmm_slab_reserve_sms(0); }
    } unreserve;
    // chunking: whole waves of the streaming kernel when it will run, else quarters of the shard
    int64_t chunk = spmm_slab_wave_rows(ctx, v, A->dtype, n, ldx, /*count_call=*/true);
    const int want_chunks = g_xchg_chunks.load(std::memory_order_relaxed);
    if (chunk > 0) {
        const int64_t waves = (v.rows + chunk - 1) / chunk;
        const int64_t target = want_chunks > 0 ? want_chunks : 5;  // about five chunks per step by default
        chunk *= std::max<int64_t>(1, (waves + target / 2) / target);
    } else {
        const int64_t target = want_chunks > 0 ? want_chunks : 4;
        chunk = ((v.rows + target - 1) / target + 63) / 64 * 64;
    }
    if (size_t(v.rows) * row_bytes < (size_t(16) << 20)) chunk = v.rows;  // too small to be worth pipelining
    void* local[1] = {dY_peers[self]};
    for (int64_t r0 = 0; r0 < v.rows; r0 += chunk) {
        const int64_t r1 = std::min(v.rows, r0 + chunk);
        CsrView part = v;
        if (r0 > 0 || r1 < v.rows) {
            part.sub_begin = r0;
            part.sub_rows = r1 - r0;
        }
        SDB_TRY(spmm_device(ctx, s, part, A->dtype, false, alpha, beta, SDB_LAYOUT_ROW_MAJOR, dX, n, ldx, local, 1, 0,
                            row0, ldy));
        cudaEvent_t done = ctx->xchg_event[ctx->xchg_next_event];
        ctx->xchg_next_event = (ctx->xchg_next_event + 1) % Context::kExchangeEvents;
        SDB_CUDA(cudaEventRecord(done, s));
        const size_t off = size_t(row0 + r0) * pitch;
        if (sm_push) {
            cudaStream_t cs = ctx->xchg_stream[self];  // one stream: the copier kernels of successive chunks queue up
            SDB_CUDA(cudaStreamWaitEvent(cs, done, 0));
            PushTargets t;
            int n_dst = 0;
            for (int q = 0; q < n_peers; ++q)
                if (q != self) t.dst[n_dst++] = reinterpret_cast<uint4*>(static_cast<char*>(dY_peers[q]) + off);
            for (int q = n_dst; q < kMaxPeers; ++q) t.dst[q] = nullptr;
            const size_t n_vec = size_t(r1 - r0) * row_bytes / 16;
            SDB_LAUNCH(push_rows_kernel, unsigned(reserve) * 2, kPushThreads, 0, cs,
                       reinterpret_cast<const uint4*>(static_cast<const char*>(dY_peers[self]) + off), t, n_dst, n_vec);
            continue;
        }
        for (int q = 0; q < n_peers; ++q) {
            if (q == self) continue;
            cudaStream_t cs = ctx->xchg_stream[q];
            SDB_CUDA(cudaStreamWaitEvent(cs, done, 0));
            char* dst = static_cast<char*>(dY_peers[q]) + off;
            const char* src = static_cast<const char*>(dY_peers[self]) + off;
            if (pitch == row_bytes)
                SDB_CUDA(cudaMemcpyAsync(dst, src, size_t(r1 - r0) * row_bytes, cudaMemcpyDeviceToDevice, cs));
            else
                SDB_CUDA(cudaMemcpy2DAsync(dst, pitch, src, pitch, row_bytes, size_t(r1 - r0),
                                           cudaMemcpyDeviceToDevice, cs));
        }
    }
    for (int q = 0; q < n_peers; ++q) {  // the call is complete on `s` once every push has landed
        if (sm_push ? q != self : q == self) continue;
        SDB_CUDA(cudaEventRecord(ctx->xchg_done[q], ctx->xchg_stream[q]));
        SDB_CUDA(cudaStreamWaitEvent(s, ctx->xchg_done[q], 0));
    }
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_set_allgather(int strategy, int chunks) {
    SDB_REQUIRE(strategy >= kXchgAuto && strategy <= kXchgSmCopier && chunks >= 0 && chunks <= 64,
                SDB_STATUS_INVALID_VALUE, "sdb_set_allgather: strategy 0..4, chunks 0..64");
    g_xchg_strategy.store(strategy, std::memory_order_relaxed);
    g_xchg_chunks.store(chunks, std::memory_order_relaxed);
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_set_allgather_sms(int sms) {
    SDB_REQUIRE(sms >= 1 && sms <= 64, SDB_STATUS_INVALID_VALUE, "sdb_set_allgather_sms: 1..64");
    g_xchg_sms.store(sms, std::memory_order_relaxed);
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_spmm(int op, const double* alpha, const sdb_mat* A, int layout, const void* X, int64_t n, int64_t ldx,
                    const double* beta, void* Y, int64_t ldy) {
    SDB_TRY(check_spmm_args(op, alpha, A, layout, X, beta, Y));
    SDB_REQUIRE(n >= 0, SDB_STATUS_INVALID_VALUE, "spmm: negative n");
    Context* ctx;
    SDB_TRY(get_context(&ctx));
    cudaStream_t s = ctx->stream;
    PhaseTimer timer;
    SDB_TRY(timer.init(s));
    SDB_TRY(timer.mark(0));
    const size_t es = dtype_size(A->dtype);
    // BSR with a natively supported block size: no CSR expansion (packed device panels are 16-byte aligned)
    const bool native_bsr = n > 0 && spmm_bsr_supported(A, op, layout, reinterpret_cast<const void*>(uintptr_t(16)), n,
                                                       n, reinterpret_cast<const void*>(uintptr_t(16)), n);
    CsrView v;
    if (native_bsr) {
        v.rows = A->rows * A->block;
        v.cols = A->cols * A->block;
        v.nnz = A->nnz * A->block * A->block;
        v.indptr = nullptr;
        v.indices = nullptr;
        v.values = nullptr;
    } else {
        SDB_TRY(csr_view(ctx, A, op != SDB_OP_NON_TRANSPOSE, &v));
    }
    const bool row_major = layout == SDB_LAYOUT_ROW_MAJOR;
    const bool beta_zero = beta[0] == 0.0 && beta[1] == 0.0;
    // host panels: `lines` contiguous runs of `run` elements, pitch ld
    const int64_t x_lines = row_major ? v.cols : n, x_run = row_major ? n : v.cols;
    const int64_t y_lines = row_major ? v.rows : n, y_run = row_major ? n : v.rows;
    SDB_REQUIRE(ldx >= x_run && ldy >= y_run, SDB_STATUS_INVALID_VALUE, "spmm: leading dimension too small");
    DevBuf dx, dy;
    SDB_TRY(dx.alloc(size_t(x_lines) * size_t(x_run) * es, s));
    SDB_TRY(dy.alloc(size_t(y_lines) * size_t(y_run) * es, s));
    SDB_TRY(h2d_2d(ctx, dx.p, size_t(x_run) * es, X, size_t(ldx) * es, size_t(x_run) * es, size_t(x_lines)));
    if (!beta_zero)
        SDB_TRY(h2d_2d(ctx, dy.p, size_t(y_run) * es, Y, size_t(ldy) * es, size_t(y_run) * es, size_t(y_lines)));
    SDB_TRY(timer.mark(1));
    void* yp[1] = {dy.p};
    if (native_bsr) {
        SDB_TRY(spmm_bsr_device(s, A, alpha, beta, dx.p, n, x_run, dy.p, y_run));
    } else {
        SDB_TRY(spmm_device(ctx, s, v, A->dtype, op == SDB_OP_CONJUGATE_TRANSPOSE, alpha, beta, layout, dx.p, n, x_run,
                            yp, 1, 0, 0, y_run));
    }
    SDB_TRY(timer.mark(2));
    SDB_TRY(d2h_2d(ctx, Y, size_t(ldy) * es, dy.p, size_t(y_run) * es, size_t(y_run) * es, size_t(y_lines)));
    SDB_TRY(timer.mark(3));
    SDB_CUDA(cudaStreamSynchronize(s));
    timer.finish(ctx);
    return SDB_STATUS_SUCCESS;
}

}  // extern "C"
