// format.cu — structure-changing operations on a device-resident sparse matrix:
//   * per-row column sort        (mkl_sparse_order,        _common.py:683-692)
//   * compressed-axis transpose  (mkl_sparse_convert_csr on a CSC handle, and
//                                 op = TRANSPOSE for every kernel; _common.py:695-722)
//   * BSR -> CSR expansion       (mkl_sparse_convert_csr on a BSR handle;
//                                 tests/test_mkl.py:251-268)
// All integer/byte work: HBM-bound, coalesced streams, no tensor cores.
#include <mutex>

#include "common.h"
#include "prims.h"

namespace sdb {

// ============================================================ row sorting
// Rows are binned by length: <= 32 entries sort inside one warp's registers,
// <= kSmemSortMax inside one CTA's shared memory, longer rows by one CTA
// working in global memory (L2 resident) with the small strides of every merge
// stage done tile by tile in shared memory.  All three run the same
// ascending-only bitonic network (first step of a merge stage pairs i with
// i ^ (2k-1), the rest are half-cleaners i ^ j), which sorts any length n
// without padding because a virtual +inf tail never has to move.
// The sort key is (column, original position): unique, so the result is the
// stable order whatever the network does.

constexpr int kSmemSortMax = 4096;  // entries per CTA-sorted row (32 KB of keys)
constexpr int kSortThreads = 256;

__global__ void __launch_bounds__(256) classify_rows_kernel(int64_t rows, const int64_t* __restrict__ indptr,
                                                            const int32_t* __restrict__ indices,
                                                            int32_t* __restrict__ short_rows,
                                                            int32_t* __restrict__ med_rows,
                                                            int32_t* __restrict__ long_rows,
                                                            unsigned* __restrict__ counters /*[4]: short, med, long, unsorted*/) {
    // one warp per row: lanes compare neighbouring entries 32 at a time; unsorted rows are listed by size class
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    const int64_t b = indptr[r], e = indptr[r + 1];
    const int64_t len = e - b;
    bool unsorted = false;
    for (int64_t p0 = b + 1; p0 < e && !unsorted; p0 += 32) {
        const int64_t p = p0 + lane;
        const bool bad = p < e && indices[p] < indices[p - 1];
        unsorted = __any_sync(0xffffffffu, bad);
    }
    if (!unsorted || lane != 0) return;
    atomicAdd(&counters[3], 1u);
    if (len > kSmemSortMax) long_rows[atomicAdd(&counters[2], 1u)] = int32_t(r);
    else if (len > 32) med_rows[atomicAdd(&counters[1], 1u)] = int32_t(r);
    else short_rows[atomicAdd(&counters[0], 1u)] = int32_t(r);
}

// one warp per row of <= 32 entries; writes sorted columns in place and the
// source position of every output entry into perm
__global__ void __launch_bounds__(256) sort_rows_warp_kernel(const int32_t* __restrict__ row_list, unsigned n_list,
                                                             const int64_t* __restrict__ indptr,
                                                             int32_t* __restrict__ indices,
                                                             int32_t* __restrict__ perm) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_list) return;
    const int64_t r = row_list[w];
    const int64_t b = indptr[r];
    const int len = int(min(int64_t(33), indptr[r + 1] - b));
    if (len > 32 || len == 0) return;  // warp-uniform
    uint64_t key = lane < len ? (uint64_t(uint32_t(indices[b + lane])) << 32) | uint32_t(lane) : ~uint64_t(0);
    // full 32-wide bitonic network (padding keys are +inf and end up at the top)
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int partner = (j == (k >> 1)) ? (lane ^ (k - 1)) : (lane ^ j);
            const uint64_t other = __shfl_sync(0xffffffffu, key, partner);
            const bool keep_min = lane < partner;
            key = keep_min ? (key < other ? key : other) : (key < other ? other : key);
        }
    }
    if (lane < len) {
        indices[b + lane] = int32_t(key >> 32);
        perm[b + lane] = int32_t(uint32_t(key));
    }
}

// one CTA per listed row of <= kSmemSortMax entries
__global__ void __launch_bounds__(kSortThreads) sort_rows_cta_kernel(const int32_t* __restrict__ row_list,
                                                                     const int64_t* __restrict__ indptr,
                                                                     int32_t* __restrict__ indices,
                                                                     int32_t* __restrict__ perm) {
    __shared__ uint64_t keys[kSmemSortMax];
    const int64_t r = row_list[blockIdx.x];
    const int64_t b = indptr[r];
    const int n = int(indptr[r + 1] - b);
    for (int i = threadIdx.x; i < n; i += kSortThreads)
        keys[i] = (uint64_t(uint32_t(indices[b + i])) << 32) | uint32_t(i);
    __syncthreads();
    int p2 = 1;
    while (p2 < n) p2 <<= 1;
    for (int k = 2; k <= p2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool flip = j == (k >> 1);
            for (int t = threadIdx.x; t < (p2 >> 1); t += kSortThreads) {
                // t enumerates the lower element of every pair at stride j
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int hi = flip ? (lo ^ (k - 1)) : (lo | j);
                if (hi < n) {
                    const uint64_t a = keys[lo], c = keys[hi];
                    if (a > c) {
                        keys[lo] = c;
                        keys[hi] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += kSortThreads) {
        indices[b + i] = int32_t(keys[i] >> 32);
        perm[b + i] = int32_t(uint32_t(keys[i]));
    }
}

// one CTA per listed long row; keys live in a global scratch array (row-aligned
// with indices), strides < kSmemSortMax are finished tile by tile in shared memory
__global__ void __launch_bounds__(1024) sort_rows_global_kernel(const int32_t* __restrict__ row_list,
                                                                const int64_t* __restrict__ indptr,
                                                                int32_t* __restrict__ indices,
                                                                int32_t* __restrict__ perm,
                                                                uint64_t* __restrict__ scratch) {
    __shared__ uint64_t tile[kSmemSortMax];
    const int64_t r = row_list[blockIdx.x];
    const int64_t b = indptr[r];
    const int64_t n = indptr[r + 1] - b;
    uint64_t* keys = scratch + b;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
        keys[i] = (uint64_t(uint32_t(indices[b + i])) << 32) | uint32_t(i);
    __syncthreads();
    int64_t p2 = 1;
    while (p2 < n) p2 <<= 1;
    // stages k <= kSmemSortMax never leave an aligned tile: run them all per tile load
    for (int64_t t0 = 0; t0 < n; t0 += kSmemSortMax) {
        const int len = int(min(int64_t(kSmemSortMax), n - t0));
        for (int i = threadIdx.x; i < len; i += blockDim.x) tile[i] = keys[t0 + i];
        __syncthreads();
        for (int k = 2; k <= kSmemSortMax; k <<= 1) {
            for (int jj = k >> 1; jj > 0; jj >>= 1) {
                const bool flip = jj == (k >> 1);
                for (int t = threadIdx.x; t < (kSmemSortMax >> 1); t += blockDim.x) {
                    const int lo = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
                    const int hi = flip ? (lo ^ (k - 1)) : (lo | jj);
                    if (hi < len) {
                        const uint64_t a = tile[lo], c = tile[hi];
                        if (a > c) {
                            tile[lo] = c;
                            tile[hi] = a;
                        }
                    }
                }
                __syncthreads();
            }
        }
        for (int i = threadIdx.x; i < len; i += blockDim.x) keys[t0 + i] = tile[i];
        __syncthreads();
    }
    for (int64_t k = int64_t(kSmemSortMax) << 1; k <= p2; k <<= 1) {
        int64_t j = k >> 1;
        // large strides: straight in global memory
        for (; j >= kSmemSortMax; j >>= 1) {
            const bool flip = j == (k >> 1);
            for (int64_t t = threadIdx.x; t < (p2 >> 1); t += blockDim.x) {
                const int64_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int64_t hi = flip ? (lo ^ (k - 1)) : (lo | j);
                if (hi < n) {
                    const uint64_t a = keys[lo], c = keys[hi];
                    if (a > c) {
                        keys[lo] = c;
                        keys[hi] = a;
                    }
                }
            }
            __syncthreads();
        }
        // remaining half-cleaner strides j < kSmemSortMax stay inside aligned tiles
        for (int64_t t0 = 0; t0 < n; t0 += kSmemSortMax) {
            const int len = int(min(int64_t(kSmemSortMax), n - t0));
            for (int i = threadIdx.x; i < len; i += blockDim.x) tile[i] = keys[t0 + i];
            __syncthreads();
            for (int jj = int(j); jj > 0; jj >>= 1) {
                for (int t = threadIdx.x; t < (kSmemSortMax >> 1); t += blockDim.x) {
                    const int lo = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
                    const int hi = lo | jj;
                    if (hi < len) {
                        const uint64_t a = tile[lo], c = tile[hi];
                        if (a > c) {
                            tile[lo] = c;
                            tile[hi] = a;
                        }
                    }
                }
                __syncthreads();
            }
            for (int i = threadIdx.x; i < len; i += blockDim.x) keys[t0 + i] = tile[i];
            __syncthreads();
        }
    }
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        indices[b + i] = int32_t(keys[i] >> 32);
        perm[b + i] = int32_t(uint32_t(keys[i]));
    }
}

// For every LISTED row: dst entry p = src entry (row start + perm[p]); one entry is `words_per_entry`
// 4-byte words (values move as raw words so every dtype / block size shares this).  Rows that were
// already in order are not listed and are never touched.
__global__ void __launch_bounds__(256) permute_rows_kernel(const int32_t* __restrict__ row_list, unsigned n_list,
                                                           const int64_t* __restrict__ indptr,
                                                           const int32_t* __restrict__ perm,
                                                           const uint32_t* __restrict__ src,
                                                           uint32_t* __restrict__ dst, int64_t words_per_entry) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_list) return;
    const int64_t r = row_list[w];
    const int64_t b = indptr[r], e = indptr[r + 1];
    if (words_per_entry == 1) {
        for (int64_t p = b + lane; p < e; p += 32) dst[p] = src[b + perm[p]];
    } else if (words_per_entry <= 4) {
        for (int64_t p = b + lane; p < e; p += 32) {
            const int64_t s = (b + perm[p]) * words_per_entry, d = p * words_per_entry;
            for (int64_t k = 0; k < words_per_entry; ++k) dst[d + k] = src[s + k];
        }
    } else {  // blocks: the whole warp moves one entry at a time
        for (int64_t p = b; p < e; ++p) {
            const int64_t s = (b + perm[p]) * words_per_entry, d = p * words_per_entry;
            for (int64_t k = lane; k < words_per_entry; k += 32) dst[d + k] = src[s + k];
        }
    }
}

__global__ void __launch_bounds__(256) copy_rows_kernel(const int32_t* __restrict__ row_list, unsigned n_list,
                                                        const int64_t* __restrict__ indptr,
                                                        const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                        int64_t words_per_entry) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_list) return;
    const int64_t r = row_list[w];
    const int64_t b = indptr[r] * words_per_entry, e = indptr[r + 1] * words_per_entry;
    for (int64_t p = b + lane; p < e; p += 32) dst[p] = src[p];
}

static unsigned blocks_for(int64_t n, int per_block) { return unsigned((n + per_block - 1) / per_block); }

__global__ void __launch_bounds__(256) strict_rows_kernel(int64_t rows, const int64_t* __restrict__ indptr,
                                                          const int32_t* __restrict__ indices,
                                                          unsigned* __restrict__ violations) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    const int64_t b = indptr[r], e = indptr[r + 1];
    bool bad = false;
    for (int64_t p = b + 1 + lane; p < e; p += 32) bad |= indices[p] <= indices[p - 1];
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicAdd(violations, 1u);
}

// One pass over freshly uploaded compressed arrays (sdb_create_*): flags[0] counts lines whose offsets or
// indices are out of range (offsets must be non-decreasing inside [0, nnz], indices inside [0, minor)),
// flags[1] counts lines that are not strictly ascending.  A malformed scipy matrix thus becomes
// SDB_STATUS_INVALID_VALUE at create instead of an out-of-bounds access in a later kernel.
__global__ void __launch_bounds__(256) validate_lines_kernel(int64_t lines, int64_t minor, int64_t nnz,
                                                             const int64_t* __restrict__ indptr,
                                                             const int32_t* __restrict__ indices,
                                                             unsigned* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (r >= lines) return;
    const int64_t b = indptr[r], e = indptr[r + 1];
    if (b < 0 || e > nnz || b > e) {
        if (lane == 0) atomicAdd(&flags[0], 1u);
        return;
    }
    bool range_bad = false, order_bad = false;
    for (int64_t p = b + lane; p < e; p += 32) {
        const int32_t c = indices[p];
        range_bad |= c < 0 || int64_t(c) >= minor;
        if (p > b) order_bad |= c <= indices[p - 1];
    }
    if (__any_sync(0xffffffffu, range_bad) && lane == 0) atomicAdd(&flags[0], 1u);
    if (__any_sync(0xffffffffu, order_bad) && lane == 0) atomicAdd(&flags[1], 1u);
}

sdb_status validate_compressed(Context* ctx, sdb_mat* m) {
    const int64_t lines = major_dim(m);
    if (lines <= 0) return SDB_STATUS_SUCCESS;
    cudaStream_t s = ctx->stream;
    DevBuf f;
    SDB_TRY(f.alloc(2 * sizeof(unsigned), s));
    SDB_CUDA(cudaMemsetAsync(f.p, 0, 2 * sizeof(unsigned), s));
    SDB_LAUNCH(validate_lines_kernel, blocks_for(lines * 32, 256), 256, 0, s, lines, minor_dim(m), m->nnz, m->indptr,
               m->indices, f.as<unsigned>());
    unsigned h[2] = {0, 0};
    SDB_CUDA(cudaMemcpyAsync(h, f.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    SDB_CUDA(cudaStreamSynchronize(s));
    SDB_REQUIRE(h[0] == 0, SDB_STATUS_INVALID_VALUE,
                "create: %u line(s) with row offsets that decrease or leave [0, nnz], or indices outside [0, %lld)", h[0],
                (long long)minor_dim(m));
    m->strict_sorted = h[1] == 0 ? 1 : -1;
    return SDB_STATUS_SUCCESS;
}

// every line strictly ascending (sorted, no duplicate index)?  cached on the handle
sdb_status ensure_strict_flag(Context* ctx, sdb_mat* m) {
    if (m->strict_sorted != 0) return SDB_STATUS_SUCCESS;
    cudaStream_t s = ctx->stream;
    const int64_t lines = major_dim(m);
    unsigned h = 0;
    if (lines > 0 && m->nnz > 0) {
        DevBuf v;
        SDB_TRY(v.alloc(sizeof(unsigned), s));
        SDB_CUDA(cudaMemsetAsync(v.p, 0, sizeof(unsigned), s));
        SDB_LAUNCH(strict_rows_kernel, blocks_for(lines * 32, 256), 256, 0, s, lines, m->indptr, m->indices,
                   v.as<unsigned>());
        SDB_CUDA(cudaMemcpyAsync(&h, v.p, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
        SDB_CUDA(cudaStreamSynchronize(s));
    }
    m->strict_sorted = h == 0 ? 1 : -1;
    return SDB_STATUS_SUCCESS;
}

sdb_status rows_sorted(Context* ctx, int64_t rows, const int64_t* indptr, const int32_t* indices, bool* sorted) {
    *sorted = true;
    if (rows <= 0) return SDB_STATUS_SUCCESS;
    cudaStream_t s = ctx->stream;
    DevBuf counters, dummy;
    SDB_TRY(counters.alloc(4 * sizeof(unsigned), s));
    SDB_TRY(dummy.alloc(size_t(rows) * sizeof(int32_t), s));
    SDB_CUDA(cudaMemsetAsync(counters.p, 0, 4 * sizeof(unsigned), s));
    SDB_LAUNCH(classify_rows_kernel, blocks_for(rows * 32, 256), 256, 0, s, rows, indptr, indices, dummy.as<int32_t>(),
               dummy.as<int32_t>(), dummy.as<int32_t>(), counters.as<unsigned>());
    unsigned h[4];
    SDB_CUDA(cudaMemcpyAsync(h, counters.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    SDB_CUDA(cudaStreamSynchronize(s));
    *sorted = h[3] == 0;
    return SDB_STATUS_SUCCESS;
}

sdb_status sort_rows(Context* ctx, int dtype, int64_t rows, const int64_t* indptr, int32_t* indices,
                     void* values, int64_t elems_per_entry, int32_t* extra) {
    if (rows <= 0) return SDB_STATUS_SUCCESS;
    cudaStream_t s = ctx->stream;
    SDB_REQUIRE(rows < (int64_t(1) << 31), SDB_STATUS_NOT_SUPPORTED, "sort_rows: too many rows");
    DevBuf counters, shortl, med, lng;
    SDB_TRY(counters.alloc(4 * sizeof(unsigned), s));
    SDB_TRY(shortl.alloc(size_t(rows) * sizeof(int32_t), s));
    SDB_TRY(med.alloc(size_t(rows) * sizeof(int32_t), s));
    SDB_TRY(lng.alloc(size_t(rows) * sizeof(int32_t), s));
    SDB_CUDA(cudaMemsetAsync(counters.p, 0, 4 * sizeof(unsigned), s));
    SDB_LAUNCH(classify_rows_kernel, blocks_for(rows * 32, 256), 256, 0, s, rows, indptr, indices,
               shortl.as<int32_t>(), med.as<int32_t>(), lng.as<int32_t>(), counters.as<unsigned>());
    unsigned h[4];
    int64_t nnz = 0;
    SDB_CUDA(cudaMemcpyAsync(h, counters.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    SDB_CUDA(cudaMemcpyAsync(&nnz, indptr + rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    SDB_CUDA(cudaStreamSynchronize(s));
    trace(s, "sort_rows: %lld rows, %u unsorted (%u in a warp, %u in shared memory, %u in global memory)",
          (long long)rows, h[3], h[0], h[1], h[2]);
    if (h[3] == 0 || nnz == 0) return SDB_STATUS_SUCCESS;  // already in order: nothing moves

    // perm / tmp are row-aligned with the matrix but only the listed (unsorted) rows are ever touched
    DevBuf perm, tmp;
    SDB_TRY(perm.alloc(size_t(nnz) * sizeof(int32_t), s));
    if (h[0] > 0)
        SDB_LAUNCH(sort_rows_warp_kernel, blocks_for(int64_t(h[0]) * 32, 256), 256, 0, s, shortl.as<int32_t>(), h[0],
                   indptr, indices, perm.as<int32_t>());
    if (h[1] > 0)
        SDB_LAUNCH(sort_rows_cta_kernel, h[1], kSortThreads, 0, s, med.as<int32_t>(), indptr, indices,
                   perm.as<int32_t>());
    if (h[2] > 0) {
        DevBuf scratch;
        SDB_TRY(scratch.alloc(size_t(nnz) * sizeof(uint64_t), s));
        SDB_LAUNCH(sort_rows_global_kernel, h[2], 1024, 0, s, lng.as<int32_t>(), indptr, indices,
                   perm.as<int32_t>(), scratch.as<uint64_t>());
    }
    trace(s, "sort_rows: columns sorted");
    const int32_t* lists[3] = {shortl.as<int32_t>(), med.as<int32_t>(), lng.as<int32_t>()};
    auto move_payload = [&](void* payload, int64_t words) -> sdb_status {
        SDB_TRY(tmp.alloc(size_t(nnz) * size_t(words) * 4, s));
        for (int c = 0; c < 3; ++c) {
            if (h[c] == 0) continue;
            SDB_LAUNCH(permute_rows_kernel, blocks_for(int64_t(h[c]) * 32, 256), 256, 0, s, lists[c], h[c], indptr,
                       perm.as<int32_t>(), static_cast<const uint32_t*>(payload), tmp.as<uint32_t>(), words);
        }
        for (int c = 0; c < 3; ++c) {
            if (h[c] == 0) continue;
            SDB_LAUNCH(copy_rows_kernel, blocks_for(int64_t(h[c]) * 32, 256), 256, 0, s, lists[c], h[c], indptr,
                       tmp.as<uint32_t>(), static_cast<uint32_t*>(payload), words);
        }
        return SDB_STATUS_SUCCESS;
    };
    if (values != nullptr) SDB_TRY(move_payload(values, int64_t(dtype_size(dtype) * size_t(elems_per_entry) / 4)));
    if (extra != nullptr) SDB_TRY(move_payload(extra, 1));  // a second per-entry payload (one 4-byte word)
    trace(s, "sort_rows: values permuted");
    return SDB_STATUS_SUCCESS;
}

// ============================================================ transpose
// CSR(A) -> CSR(A^T): column histogram, prefix sum, scatter through per-column
// cursors, then the row sort above puts every output row in ascending order
// (the scatter order is whatever the atomics gave; the sort key includes the
// source position only within a row, so sort on the source ROW id instead:
// the scattered "column" of A^T *is* the source row, unique per output row when
// A has no duplicate entries, and stable enough otherwise).
__global__ void __launch_bounds__(256) count_columns_kernel(int64_t nnz, const int32_t* __restrict__ indices,
                                                            int32_t* __restrict__ counts) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (; i < nnz; i += stride) atomicAdd(&counts[indices[i]], 1);
}

__global__ void __launch_bounds__(256) scatter_transpose_kernel(int64_t rows, const int64_t* __restrict__ indptr,
                                                                const int32_t* __restrict__ indices,
                                                                const uint32_t* __restrict__ values,
                                                                int words_per_entry,
                                                                unsigned long long* __restrict__ cursor,
                                                                int32_t* __restrict__ t_indices,
                                                                uint32_t* __restrict__ t_values,
                                                                int32_t* __restrict__ t_pos) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    const int64_t b = indptr[r], e = indptr[r + 1];
    for (int64_t p = b + lane; p < e; p += 32) {
        const unsigned long long q = atomicAdd(&cursor[indices[p]], 1ull);
        t_indices[q] = int32_t(r);
        if (t_pos) t_pos[q] = int32_t(p - b);
        for (int w = 0; w < words_per_entry; ++w) t_values[q * words_per_entry + w] = values[p * words_per_entry + w];
    }
}

// after the companion is final: a.pos[source entry] = position of its image inside the companion line
__global__ void __launch_bounds__(256) back_positions_kernel(int64_t t_lines, const int64_t* __restrict__ t_ptr,
                                                             const int32_t* __restrict__ t_idx,
                                                             const int32_t* __restrict__ t_pos,
                                                             const int64_t* __restrict__ a_ptr,
                                                             int32_t* __restrict__ a_pos) {
    const int lane = threadIdx.x & 31;
    const int64_t k = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (k >= t_lines) return;
    const int64_t b = t_ptr[k], e = t_ptr[k + 1];
    for (int64_t q = b + lane; q < e; q += 32) a_pos[a_ptr[t_idx[q]] + t_pos[q]] = int32_t(q - b);
}

sdb_status transpose_compressed(Context* ctx, sdb_mat* a, sdb_mat** out, bool with_pos) {
    // `a` is read as a compressed matrix with major_dim(a) lines over minor_dim(a) indices
    cudaStream_t s = ctx->stream;
    const int64_t major = major_dim(a), minor = minor_dim(a);
    SDB_REQUIRE(a->block == 1, SDB_STATUS_NOT_SUPPORTED, "transpose: BSR must be expanded first");
    sdb_mat* t;
    // the result has `minor` lines; describe it as a CSR (minor x major) matrix
    SDB_TRY(new_handle(&t, SDB_FMT_CSR, a->dtype, minor, major, a->nnz, 1, SDB_LAYOUT_ROW_MAJOR, s));
    if (with_pos) {
        SDB_TRY(ensure_strict_flag(ctx, a));
        with_pos = a->strict_sorted == 1 && a->nnz > 0;
    }
    sdb_status st = [&]() -> sdb_status {
        DevBuf counts, cursor;
        if (with_pos) {
            SDB_TRY(dev_alloc(reinterpret_cast<void**>(&t->pos), size_t(a->nnz) * 4, s));
            if (!a->pos) SDB_TRY(dev_alloc(reinterpret_cast<void**>(&a->pos), size_t(a->nnz) * 4, s));
        }
        SDB_TRY(counts.alloc(size_t(minor + 1) * sizeof(int32_t), s));
        SDB_CUDA(cudaMemsetAsync(counts.p, 0, size_t(minor + 1) * sizeof(int32_t), s));
        if (a->nnz > 0) {
            unsigned g = unsigned(std::min<int64_t>((a->nnz + 255) / 256, int64_t(ctx->sm_count) * 16));
            SDB_LAUNCH(count_columns_kernel, g, 256, 0, s, a->nnz, a->indices, counts.as<int32_t>());
        }
        SDB_TRY(exclusive_scan_i32_to_i64(s, counts.as<int32_t>(), t->indptr, minor));
        if (a->nnz == 0) return SDB_STATUS_SUCCESS;
        SDB_TRY(cursor.alloc(size_t(minor) * sizeof(int64_t), s));
        SDB_CUDA(cudaMemcpyAsync(cursor.p, t->indptr, size_t(minor) * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
        SDB_LAUNCH(scatter_transpose_kernel, blocks_for(major * 32, 256), 256, 0, s, major, a->indptr, a->indices,
                   static_cast<const uint32_t*>(a->values), int(dtype_size(a->dtype) / 4),
                   cursor.as<unsigned long long>(), t->indices, static_cast<uint32_t*>(t->values), t->pos);
        SDB_TRY(sort_rows(ctx, t->dtype, minor, t->indptr, t->indices, t->values, 1, t->pos));
        if (with_pos) {
            // the source has no duplicates, so the companion's lines are strictly ascending too
            t->strict_sorted = 1;
            SDB_LAUNCH(back_positions_kernel, blocks_for(minor * 32, 256), 256, 0, s, minor, t->indptr, t->indices,
                       t->pos, a->indptr, a->pos);
        }
        return SDB_STATUS_SUCCESS;
    }();
    if (st != SDB_STATUS_SUCCESS) {
        free_handle(t);
        return st;
    }
    *out = t;
    return SDB_STATUS_SUCCESS;
}

// ============================================================ BSR -> CSR
// Row (I*b + r) of the expansion lists, for every stored block q of block row
// I in order, the b columns bidx[q]*b .. +b-1.  One warp per output row.
__global__ void __launch_bounds__(256) expand_bsr_kernel(int64_t block_rows, int b, int col_major_blocks,
                                                         const int64_t* __restrict__ bptr,
                                                         const int32_t* __restrict__ bidx,
                                                         const uint32_t* __restrict__ bval, int words,
                                                         int64_t* __restrict__ indptr, int32_t* __restrict__ indices,
                                                         uint32_t* __restrict__ values) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= block_rows * b) return;
    const int64_t I = row / b;
    const int r = int(row - I * b);
    const int64_t qb = bptr[I], nb = bptr[I + 1] - qb;
    const int64_t o0 = qb * b * b + int64_t(r) * nb * b;
    if (lane == 0) {
        indptr[row + 1] = o0 + nb * b;
        if (row == 0) indptr[0] = 0;
    }
    for (int64_t e = lane; e < nb * b; e += 32) {
        const int64_t q = qb + e / b;
        const int c = int(e % b);
        indices[o0 + e] = int32_t(int64_t(bidx[q]) * b + c);
        const int64_t src = q * b * b + (col_major_blocks ? int64_t(c) * b + r : int64_t(r) * b + c);
        for (int w = 0; w < words; ++w) values[(o0 + e) * words + w] = bval[src * words + w];
    }
}

sdb_status expand_bsr(Context* ctx, const sdb_mat* m, sdb_mat** out_csr) {
    cudaStream_t s = ctx->stream;
    const int64_t b = m->block;
    sdb_mat* c;
    SDB_TRY(new_handle(&c, SDB_FMT_CSR, m->dtype, m->rows * b, m->cols * b, m->nnz * b * b, 1,
                       SDB_LAYOUT_ROW_MAJOR, s));
    if (c->rows == 0) {
        cudaMemsetAsync(c->indptr, 0, sizeof(int64_t), s);
    } else {
        expand_bsr_kernel<<<blocks_for(c->rows * 32, 256), 256, 0, s>>>(
            m->rows, int(b), m->block_layout == SDB_LAYOUT_COL_MAJOR, m->indptr, m->indices,
            static_cast<const uint32_t*>(m->values), int(dtype_size(m->dtype) / 4), c->indptr, c->indices,
            static_cast<uint32_t*>(c->values));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            free_handle(c);
            return cuda_fail(e, "expand_bsr_kernel", __FILE__, __LINE__);
        }
    }
    *out_csr = c;
    return SDB_STATUS_SUCCESS;
}

// ============================================================ CSR -> BSR
// Inverse of expand_bsr for a CSR matrix whose rows come in groups of b with an
// identical, sorted pattern made of whole b-wide column blocks (what a product
// of two expanded BSR matrices looks like).  Used by BSR x BSR SpGEMM.
__global__ void __launch_bounds__(256) bsr_row_blocks_kernel(int64_t block_rows, int b,
                                                             const int64_t* __restrict__ indptr,
                                                             int32_t* __restrict__ nblk) {
    const int64_t I = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (I >= block_rows) return;
    nblk[I] = int32_t((indptr[I * b + 1] - indptr[I * b]) / b);
}

__global__ void __launch_bounds__(256) compress_bsr_kernel(int64_t block_rows, int b,
                                                           const int64_t* __restrict__ indptr,
                                                           const int32_t* __restrict__ indices,
                                                           const uint32_t* __restrict__ values, int words,
                                                           const int64_t* __restrict__ bptr,
                                                           int32_t* __restrict__ bidx, uint32_t* __restrict__ bval) {
    const int lane = threadIdx.x & 31;
    const int64_t I = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (I >= block_rows) return;
    const int64_t q0 = bptr[I], nb = bptr[I + 1] - q0;
    const int64_t first = indptr[I * b];
    for (int64_t t = lane; t < nb; t += 32) bidx[q0 + t] = indices[first + t * b] / b;
    const int64_t per_block = int64_t(b) * b;
    for (int64_t e = lane; e < nb * per_block; e += 32) {
        const int64_t t = e / per_block;
        const int rc = int(e - t * per_block);
        const int r = rc / b, c = rc - r * b;
        const int64_t src = indptr[I * b + r] + t * b + c;
        for (int w = 0; w < words; ++w) bval[(q0 * per_block + e) * words + w] = values[src * words + w];
    }
}

sdb_status compress_to_bsr(Context* ctx, const sdb_mat* csr, int64_t b, sdb_mat** out) {
    cudaStream_t s = ctx->stream;
    SDB_REQUIRE(b >= 1 && csr->rows % b == 0 && csr->cols % b == 0 && csr->nnz % (b * b) == 0,
                SDB_STATUS_INTERNAL_ERROR, "compress_to_bsr: matrix is not made of whole %lld-blocks", (long long)b);
    const int64_t block_rows = csr->rows / b;
    sdb_mat* m;
    SDB_TRY(new_handle(&m, SDB_FMT_BSR, csr->dtype, block_rows, csr->cols / b, csr->nnz / (b * b), b,
                       SDB_LAYOUT_ROW_MAJOR, s));
    sdb_status st = [&]() -> sdb_status {
        DevBuf nblk;
        SDB_TRY(nblk.alloc(size_t(block_rows + 1) * 4, s));
        if (block_rows > 0)
            SDB_LAUNCH(bsr_row_blocks_kernel, blocks_for(block_rows, 256), 256, 0, s, block_rows, int(b), csr->indptr,
                       nblk.as<int32_t>());
        SDB_TRY(exclusive_scan_i32_to_i64(s, nblk.as<int32_t>(), m->indptr, block_rows));
        if (block_rows > 0 && m->nnz > 0)
            SDB_LAUNCH(compress_bsr_kernel, blocks_for(block_rows * 32, 256), 256, 0, s, block_rows, int(b),
                       csr->indptr, csr->indices, static_cast<const uint32_t*>(csr->values),
                       int(dtype_size(csr->dtype) / 4), m->indptr, m->indices, static_cast<uint32_t*>(m->values));
        return SDB_STATUS_SUCCESS;
    }();
    if (st != SDB_STATUS_SUCCESS) {
        free_handle(m);
        return st;
    }
    *out = m;
    return SDB_STATUS_SUCCESS;
}

// ============================================================ CSR view of op(A)
// Companions (transposed form, BSR expansion, cross positions) are built lazily and cached on the handle; one
// process-wide mutex serialises building them, so two threads multiplying with the same handle cannot build
// (or free) a companion twice.  A companion is only ever REPLACED by the positions request of a triangular
// product (syrk / syrkd): those must not run concurrently with other calls on the same handle (sdb200.h).
std::mutex g_companion_mutex;

sdb_status csr_view(Context* ctx, const sdb_mat* m_in, bool transpose, CsrView* v, bool want_pos) {
    sdb_mat* m = const_cast<sdb_mat*>(m_in);  // companions are a cache, not a logical mutation
    std::lock_guard<std::mutex> companion_lock(g_companion_mutex);
    if (m->format == SDB_FMT_BSR) {
        if (!m->expanded) SDB_TRY(expand_bsr(ctx, m, &m->expanded));
        m = m->expanded;
    }
    // the stored arrays list lines of A (CSR) or of A^T (CSC)
    const bool stored_is_transposed = m->format == SDB_FMT_CSC;
    v->pos = nullptr;
    if (want_pos) {
        // positions need the companion, built WITH positions; a companion cached without them is rebuilt
        if (m->transposed && !m->transposed->pos) {
            SDB_TRY(ensure_strict_flag(ctx, m));
            if (m->strict_sorted == 1 && m->nnz > 0) {
                free_handle(m->transposed);
                m->transposed = nullptr;
            }
        }
        if (!m->transposed) SDB_TRY(transpose_compressed(ctx, m, &m->transposed, true));
    }
    if (stored_is_transposed == transpose) {
        if (want_pos && m->transposed && m->transposed->pos) v->pos = m->pos;
        v->owner = m;
        v->rows = major_dim(m);
        v->cols = minor_dim(m);
        v->nnz = m->nnz;
        v->indptr = m->indptr;
        v->indices = m->indices;
        v->values = m->values;
        return SDB_STATUS_SUCCESS;
    }
    if (!m->transposed) SDB_TRY(transpose_compressed(ctx, m, &m->transposed));
    sdb_mat* t = m->transposed;
    if (want_pos) v->pos = t->pos;
    v->owner = t;
    v->rows = t->rows;
    v->cols = t->cols;
    v->nnz = t->nnz;
    v->indptr = t->indptr;
    v->indices = t->indices;
    v->values = t->values;
    return SDB_STATUS_SUCCESS;
}

}  // namespace sdb

using namespace sdb;

extern "C" {

sdb_status sdb_order(sdb_mat* m) {
    SDB_REQUIRE(m != nullptr, SDB_STATUS_NOT_INITIALIZED, "order: null handle");
    SDB_REQUIRE(valid(m), SDB_STATUS_INVALID_VALUE, "order: not a live sdb_mat handle");
    Context* ctx;
    SDB_TRY(get_context(&ctx));
    SDB_TRY(sort_rows(ctx, m->dtype, major_dim(m), m->indptr, m->indices, m->values, m->block * m->block));
    // companions were derived from the old entry order; they stay valid as
    // matrices (same entries) but drop them so exports of derived handles are
    // reproducible from the ordered arrays
    if (m->transposed) {
        free_handle(m->transposed);
        m->transposed = nullptr;
    }
    if (m->expanded) {
        free_handle(m->expanded);
        m->expanded = nullptr;
    }
    if (m->pos) {
        cudaFreeAsync(m->pos, ctx->stream);
        m->pos = nullptr;
    }
    if (m->slab_rc) cudaFreeAsync(m->slab_rc, ctx->stream);
    if (m->slab_val) cudaFreeAsync(m->slab_val, ctx->stream);
    m->slab_rc = m->slab_val = nullptr;
    drop_spmv_tiles(m, ctx->stream);
    m->strict_sorted = 0;
    SDB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_invalidate(sdb_mat* m) {
    SDB_REQUIRE(m != nullptr, SDB_STATUS_NOT_INITIALIZED, "invalidate: null handle");
    SDB_REQUIRE(valid(m), SDB_STATUS_INVALID_VALUE, "invalidate: not a live sdb_mat handle");
    Context* ctx;
    SDB_TRY(get_context(&ctx));
    std::lock_guard<std::mutex> companion_lock(g_companion_mutex);
    if (m->transposed) free_handle(m->transposed);
    if (m->expanded) free_handle(m->expanded);
    m->transposed = m->expanded = nullptr;
    if (m->pos) cudaFreeAsync(m->pos, ctx->stream);
    if (m->slab_rc) cudaFreeAsync(m->slab_rc, ctx->stream);
    if (m->slab_val) cudaFreeAsync(m->slab_val, ctx->stream);
    m->pos = nullptr;
    m->slab_rc = m->slab_val = nullptr;
    drop_spmv_tiles(m, ctx->stream);
    m->strict_sorted = 0;
    m->spmm_calls = 0;
    SDB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_convert_csr(const sdb_mat* m, int op, sdb_mat** out) {
    SDB_REQUIRE(out != nullptr, SDB_STATUS_INVALID_VALUE, "convert_csr: null output handle");
    *out = nullptr;
    SDB_REQUIRE(m != nullptr, SDB_STATUS_NOT_INITIALIZED, "convert_csr: null handle");
    SDB_REQUIRE(valid(m), SDB_STATUS_INVALID_VALUE, "convert_csr: not a live sdb_mat handle");
    SDB_REQUIRE(op == SDB_OP_NON_TRANSPOSE || op == SDB_OP_TRANSPOSE, SDB_STATUS_NOT_SUPPORTED,
                "convert_csr: op %d not supported", op);
    Context* ctx;
    SDB_TRY(get_context(&ctx));
    cudaStream_t s = ctx->stream;
    CsrView v;
    SDB_TRY(csr_view(ctx, m, op == SDB_OP_TRANSPOSE, &v));
    // hand back an independent copy: the caller destroys both handles separately
    sdb_mat* c;
    SDB_TRY(new_handle(&c, SDB_FMT_CSR, m->dtype, v.rows, v.cols, v.nnz, 1, SDB_LAYOUT_ROW_MAJOR, s));
    cudaError_t e = cudaMemcpyAsync(c->indptr, v.indptr, size_t(v.rows + 1) * 8, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && v.nnz > 0)
        e = cudaMemcpyAsync(c->indices, v.indices, size_t(v.nnz) * 4, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && v.nnz > 0)
        e = cudaMemcpyAsync(c->values, v.values, size_t(v.nnz) * dtype_size(m->dtype), cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        free_handle(c);
        return cuda_fail(e, "convert_csr copy", __FILE__, __LINE__);
    }
    *out = c;
    return SDB_STATUS_SUCCESS;
}

}  // extern "C"
