// prims.h — device-wide building blocks shared by the sparse kernels: prefix
// sums, index widening, dense panel transposes.  All stream-ordered.
#pragma once

#include "common.h"

namespace sdb {

// out[0] = 0, out[i+1] = sum_{j<=i} in[j]   (in: int32[n], out: int64[n+1]).
sdb_status exclusive_scan_i32_to_i64(cudaStream_t s, const int32_t* in, int64_t* out, int64_t n);

// dst[i] = src[i] widened / narrowed between int32 and int64 on the device.
sdb_status widen_i32_to_i64(cudaStream_t s, const int32_t* src, int64_t* dst, int64_t n);
sdb_status narrow_i64_to_i32(cudaStream_t s, const int64_t* src, int32_t* dst, int64_t n);

// Dense (rows x cols) panel: dst[c * ld_dst + r] = src[r * ld_src + c]
// (element size elem_bytes in {4, 8, 16}).
sdb_status transpose_dense(cudaStream_t s, const void* src, int64_t ld_src, void* dst, int64_t ld_dst,
                           int64_t rows, int64_t cols, int elem_bytes);

sdb_status fill_zero(cudaStream_t s, void* p, size_t bytes);

}  // namespace sdb
