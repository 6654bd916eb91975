// handle.cu — lifetime of the device-resident sparse matrix: upload (create),
// shape query, download (export), destroy.  Replaces mkl_sparse_?_create_* /
// _export_* / mkl_sparse_destroy as the reference drives them from
// sparse_dot_mkl/_mkl_interface/_common.py:245-384, 387-642, 671-680.
#include "common.h"
#include "prims.h"

namespace sdb {

sdb_status new_handle(sdb_mat** out, int format, int dtype, int64_t rows, int64_t cols, int64_t nnz,
                      int64_t block, int block_layout, cudaStream_t s) {
    *out = nullptr;
    sdb_mat* m = static_cast<sdb_mat*>(calloc(1, sizeof(sdb_mat)));
    SDB_REQUIRE(m != nullptr, SDB_STATUS_ALLOC_FAILED, "out of host memory for a handle");
    m->magic = kMagic;
    m->format = format;
    m->dtype = dtype;
    m->rows = rows;
    m->cols = cols;
    m->nnz = nnz;
    m->block = block;
    m->block_layout = block_layout;
    m->owns = true;
    m->transposed = nullptr;
    m->expanded = nullptr;
    cudaGetDevice(&m->device);
    const int64_t major = format == SDB_FMT_CSC ? cols : rows;
    sdb_status st = dev_alloc(reinterpret_cast<void**>(&m->indptr), size_t(major + 1) * sizeof(int64_t), s);
    if (st == SDB_STATUS_SUCCESS)
        st = dev_alloc(reinterpret_cast<void**>(&m->indices), size_t(nnz) * sizeof(int32_t), s);
    if (st == SDB_STATUS_SUCCESS)
        st = dev_alloc(&m->values, size_t(nnz) * size_t(block * block) * dtype_size(dtype), s);
    if (st != SDB_STATUS_SUCCESS) {
        free_handle(m);
        return st;
    }
    *out = m;
    return SDB_STATUS_SUCCESS;
}

void free_handle(sdb_mat* m) {
    if (!m) return;
    // release on the device the handle lives on, whatever device is current on this thread
    int cur = -1;
    const bool switched = cudaGetDevice(&cur) == cudaSuccess && cur != m->device && m->device >= 0 &&
                          cudaSetDevice(m->device) == cudaSuccess;
    struct Restore {
        bool on;
        int dev;
        ~Restore() {
            if (on) cudaSetDevice(dev);
        }
    } restore{switched, cur};
    if (m->transposed) free_handle(m->transposed);
    if (m->expanded) free_handle(m->expanded);
    if (m->owns) {
        // Stream-ordered release on this thread's stream: host entry points have
        // synchronised by now; *_dev callers keep handles alive until their own
        // stream has drained (documented in sdb200.h).
        Context* ctx = nullptr;
        cudaStream_t s = get_context(&ctx) == SDB_STATUS_SUCCESS ? ctx->stream : nullptr;
        if (m->indptr) cudaFreeAsync(m->indptr, s);
        if (m->indices) cudaFreeAsync(m->indices, s);
        if (m->values) cudaFreeAsync(m->values, s);
    }
    if (m->pos || m->slab_rc || m->slab_val || m->vt_cache || m->long_rows[0] || m->long_rows[1]) {
        Context* ctx = nullptr;
        cudaStream_t fs = get_context(&ctx) == SDB_STATUS_SUCCESS ? ctx->stream : nullptr;
        if (m->pos) cudaFreeAsync(m->pos, fs);
        if (m->slab_rc) cudaFreeAsync(m->slab_rc, fs);
        if (m->slab_val) cudaFreeAsync(m->slab_val, fs);
        drop_spmv_tiles(m, fs);
    }
    m->magic = 0;
    free(m);
}

// upload one host index array (int32 or int64) into a device int64 / int32 array
static sdb_status upload_index(Context* ctx, const void* h, int bits, int64_t n, int64_t* d64, int32_t* d32) {
    if (n <= 0) return SDB_STATUS_SUCCESS;
    cudaStream_t s = ctx->stream;
    if (d64) {
        if (bits == 64) return h2d(ctx, d64, h, size_t(n) * 8);
        DevBuf tmp;
        SDB_TRY(tmp.alloc(size_t(n) * 4, s));
        SDB_TRY(h2d(ctx, tmp.p, h, size_t(n) * 4));
        return widen_i32_to_i64(s, tmp.as<int32_t>(), d64, n);
    }
    if (bits == 32) return h2d(ctx, d32, h, size_t(n) * 4);
    DevBuf tmp;
    SDB_TRY(tmp.alloc(size_t(n) * 8, s));
    SDB_TRY(h2d(ctx, tmp.p, h, size_t(n) * 8));
    return narrow_i64_to_i32(s, tmp.as<int64_t>(), d32, n);
}

static int64_t host_index_at(const void* p, int bits, int64_t i) {
    return bits == 32 ? int64_t(static_cast<const int32_t*>(p)[i]) : static_cast<const int64_t*>(p)[i];
}

static sdb_status create_compressed(sdb_mat** out, int format, int64_t rows, int64_t cols, int64_t block,
                                    int block_layout, const void* indptr, const void* indices,
                                    int index_bits, const void* values, int dtype) {
    SDB_REQUIRE(out != nullptr, SDB_STATUS_INVALID_VALUE, "create: null output handle");
    *out = nullptr;
    SDB_REQUIRE(rows >= 0 && cols >= 0, SDB_STATUS_INVALID_VALUE, "create: negative dimension");
    SDB_REQUIRE(rows < (int64_t(1) << 31) - 1 && cols < (int64_t(1) << 31) - 1, SDB_STATUS_NOT_SUPPORTED,
                "create: dimensions must fit int32 column indices");
    SDB_REQUIRE(index_bits == 32 || index_bits == 64, SDB_STATUS_INVALID_VALUE,
                "create: index_bits must be 32 or 64");
    SDB_REQUIRE(dtype_size(dtype) != 0, SDB_STATUS_NOT_SUPPORTED, "create: unknown dtype %d", dtype);
    SDB_REQUIRE(indptr != nullptr, SDB_STATUS_INVALID_VALUE, "create: null indptr");
    const int64_t major = format == SDB_FMT_CSC ? cols : rows;
    const int64_t first = host_index_at(indptr, index_bits, 0);
    const int64_t nnz = host_index_at(indptr, index_bits, major) - first;
    SDB_REQUIRE(first == 0, SDB_STATUS_INVALID_VALUE, "create: indptr[0] must be 0 (zero-based, unsliced)");
    SDB_REQUIRE(nnz >= 0, SDB_STATUS_INVALID_VALUE, "create: negative nnz");
    SDB_REQUIRE(nnz == 0 || (indices != nullptr && values != nullptr), SDB_STATUS_INVALID_VALUE,
                "create: null indices/values with nnz > 0");
    Context* ctx;
    SDB_TRY(get_context(&ctx));
    sdb_mat* m;
    SDB_TRY(new_handle(&m, format, dtype, rows, cols, nnz, block, block_layout, ctx->stream));
    sdb_status st = upload_index(ctx, indptr, index_bits, major + 1, m->indptr, nullptr);
    if (st == SDB_STATUS_SUCCESS) st = upload_index(ctx, indices, index_bits, nnz, nullptr, m->indices);
    if (st == SDB_STATUS_SUCCESS)
        st = h2d(ctx, m->values, values, size_t(nnz) * size_t(block * block) * dtype_size(dtype));
    // never trust the host arrays: one device pass checks offsets and indices (and notes whether every
    // line is strictly ascending, which the streaming SpMM and the triangular products want to know)
    if (st == SDB_STATUS_SUCCESS && nnz > 0) st = validate_compressed(ctx, m);
    if (st == SDB_STATUS_SUCCESS) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) st = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    }
    if (st != SDB_STATUS_SUCCESS) {
        free_handle(m);
        return st;
    }
    *out = m;
    return SDB_STATUS_SUCCESS;
}

}  // namespace sdb

using namespace sdb;

extern "C" {

sdb_status sdb_create_csr(sdb_mat** out, int64_t rows, int64_t cols, const void* indptr,
                          const void* indices, int index_bits, const void* values, int dtype) {
    return create_compressed(out, SDB_FMT_CSR, rows, cols, 1, SDB_LAYOUT_ROW_MAJOR, indptr, indices,
                             index_bits, values, dtype);
}

sdb_status sdb_create_csc(sdb_mat** out, int64_t rows, int64_t cols, const void* indptr,
                          const void* indices, int index_bits, const void* values, int dtype) {
    return create_compressed(out, SDB_FMT_CSC, rows, cols, 1, SDB_LAYOUT_ROW_MAJOR, indptr, indices,
                             index_bits, values, dtype);
}

sdb_status sdb_create_bsr(sdb_mat** out, int64_t block_rows, int64_t block_cols, int64_t block_size,
                          int block_layout, const void* indptr, const void* indices, int index_bits,
                          const void* values, int dtype) {
    SDB_REQUIRE(block_size >= 1, SDB_STATUS_INVALID_VALUE, "create_bsr: block_size must be >= 1");
    SDB_REQUIRE(block_layout == SDB_LAYOUT_ROW_MAJOR || block_layout == SDB_LAYOUT_COL_MAJOR,
                SDB_STATUS_INVALID_VALUE, "create_bsr: bad block layout %d", block_layout);
    SDB_REQUIRE(block_rows * block_size < (int64_t(1) << 31) - 1 && block_cols * block_size < (int64_t(1) << 31) - 1,
                SDB_STATUS_NOT_SUPPORTED, "create_bsr: expanded dimensions must fit int32");
    return create_compressed(out, SDB_FMT_BSR, block_rows, block_cols, block_size, block_layout, indptr,
                             indices, index_bits, values, dtype);
}

sdb_status sdb_create_csr_dev(sdb_mat** out, int64_t rows, int64_t cols, int64_t nnz,
                              const int64_t* d_indptr, const int32_t* d_indices, const void* d_values,
                              int dtype) {
    SDB_REQUIRE(out != nullptr, SDB_STATUS_INVALID_VALUE, "create_csr_dev: null output handle");
    *out = nullptr;
    SDB_REQUIRE(rows >= 0 && cols >= 0 && nnz >= 0, SDB_STATUS_INVALID_VALUE, "create_csr_dev: negative size");
    SDB_REQUIRE(dtype_size(dtype) != 0, SDB_STATUS_NOT_SUPPORTED, "create_csr_dev: unknown dtype %d", dtype);
    SDB_REQUIRE(d_indptr != nullptr && (nnz == 0 || (d_indices && d_values)), SDB_STATUS_INVALID_VALUE,
                "create_csr_dev: null device array");
    sdb_mat* m = static_cast<sdb_mat*>(calloc(1, sizeof(sdb_mat)));
    SDB_REQUIRE(m != nullptr, SDB_STATUS_ALLOC_FAILED, "out of host memory for a handle");
    m->magic = kMagic;
    m->format = SDB_FMT_CSR;
    m->dtype = dtype;
    m->rows = rows;
    m->cols = cols;
    m->nnz = nnz;
    m->block = 1;
    m->block_layout = SDB_LAYOUT_ROW_MAJOR;
    m->indptr = const_cast<int64_t*>(d_indptr);
    m->indices = const_cast<int32_t*>(d_indices);
    m->values = const_cast<void*>(d_values);
    m->owns = false;
    cudaGetDevice(&m->device);
    *out = m;
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_destroy(sdb_mat* m) {
    SDB_REQUIRE(m != nullptr, SDB_STATUS_NOT_INITIALIZED, "destroy: null handle");
    SDB_REQUIRE(valid(m), SDB_STATUS_INVALID_VALUE, "destroy: not a live sdb_mat handle");
    free_handle(m);
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_get_info(const sdb_mat* m, int* format, int* dtype, int64_t* rows, int64_t* cols,
                        int64_t* nnz, int64_t* block_size, int* block_layout) {
    SDB_REQUIRE(m != nullptr, SDB_STATUS_NOT_INITIALIZED, "get_info: null handle");
    SDB_REQUIRE(valid(m), SDB_STATUS_INVALID_VALUE, "get_info: not a live sdb_mat handle");
    if (format) *format = m->format;
    if (dtype) *dtype = m->dtype;
    if (rows) *rows = m->rows;
    if (cols) *cols = m->cols;
    if (nnz) *nnz = m->nnz;
    if (block_size) *block_size = m->block;
    if (block_layout) *block_layout = m->block_layout;
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_export_dev(const sdb_mat* m, const int64_t** d_indptr, const int32_t** d_indices,
                          const void** d_values) {
    SDB_REQUIRE(m != nullptr, SDB_STATUS_NOT_INITIALIZED, "export_dev: null handle");
    SDB_REQUIRE(valid(m), SDB_STATUS_INVALID_VALUE, "export_dev: not a live sdb_mat handle");
    if (d_indptr) *d_indptr = m->indptr;
    if (d_indices) *d_indices = m->indices;
    if (d_values) *d_values = m->values;
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_export(const sdb_mat* m, void* indptr, int indptr_bits, void* indices, int indices_bits,
                      void* values) {
    SDB_REQUIRE(m != nullptr, SDB_STATUS_NOT_INITIALIZED, "export: null handle");
    SDB_REQUIRE(valid(m), SDB_STATUS_INVALID_VALUE, "export: not a live sdb_mat handle");
    Context* ctx;
    SDB_TRY(get_context(&ctx));
    cudaStream_t s = ctx->stream;
    const int64_t major = major_dim(m);
    if (indptr) {
        SDB_REQUIRE(indptr_bits == 32 || indptr_bits == 64, SDB_STATUS_INVALID_VALUE, "export: bad indptr_bits");
        if (indptr_bits == 64) {
            SDB_TRY(d2h(ctx, indptr, m->indptr, size_t(major + 1) * 8));
        } else {
            SDB_REQUIRE(m->nnz <= INT32_MAX, SDB_STATUS_INVALID_VALUE,
                        "export: nnz %lld does not fit an int32 indptr", (long long)m->nnz);
            DevBuf tmp;
            SDB_TRY(tmp.alloc(size_t(major + 1) * 4, s));
            SDB_TRY(narrow_i64_to_i32(s, m->indptr, tmp.as<int32_t>(), major + 1));
            SDB_TRY(d2h(ctx, indptr, tmp.p, size_t(major + 1) * 4));
        }
    }
    if (indices && m->nnz > 0) {
        SDB_REQUIRE(indices_bits == 32 || indices_bits == 64, SDB_STATUS_INVALID_VALUE, "export: bad indices_bits");
        if (indices_bits == 32) {
            SDB_TRY(d2h(ctx, indices, m->indices, size_t(m->nnz) * 4));
        } else {
            DevBuf tmp;
            SDB_TRY(tmp.alloc(size_t(m->nnz) * 8, s));
            SDB_TRY(widen_i32_to_i64(s, m->indices, tmp.as<int64_t>(), m->nnz));
            SDB_TRY(d2h(ctx, indices, tmp.p, size_t(m->nnz) * 8));
        }
    }
    if (values && m->nnz > 0)
        SDB_TRY(d2h(ctx, values, m->values, size_t(m->nnz) * size_t(m->block * m->block) * dtype_size(m->dtype)));
    SDB_CUDA(cudaStreamSynchronize(s));
    return SDB_STATUS_SUCCESS;
}

}  // extern "C"
