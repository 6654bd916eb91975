/*
 * sdb200.h — C ABI of libsdb200.so, the B200 (sm_100a) sparse-matmul backend that
 * sits where Intel MKL's inspector-executor sparse BLAS sits under sparse_dot_mkl.
 *
 * Every entry point replaces one (family of) MKL routine(s) the reference binds
 * through ctypes in sparse_dot_mkl/_mkl_interface/_cfunctions.py; the reference
 * call site is cited next to each declaration (paths relative to the reference
 * checkout).  Conventions:
 *
 *   - plain C, no C++/torch types; all sizes are explicit-width integers;
 *   - every routine returns an sdb_status with MKL's sparse_status_t numbering
 *     (_mkl_interface/_constants.py:2-10), so the reference's
 *     _check_return_value (_common.py:645-668) works unchanged;
 *   - sdb_last_error() returns a thread-local human-readable message;
 *   - calls are synchronous: when a routine returns, its results are in the
 *     caller's memory (host entry points) or complete on `stream` (the *_dev
 *     entry points are stream-ordered and do not synchronise);
 *   - host input buffers are borrowed only for the duration of the call (they
 *     are copied to HBM); a handle owns its device memory until sdb_destroy;
 *   - nothing here falls back to the CPU: without a usable GPU the compute
 *     entry points return SDB_STATUS_EXECUTION_FAILED.
 */
#ifndef SDB200_H
#define SDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDB_API __attribute__((visibility("default")))

/* ---- status codes: sparse_status_t, _constants.py:2-10 ------------------- */
typedef int sdb_status;
#define SDB_STATUS_SUCCESS          0
#define SDB_STATUS_NOT_INITIALIZED  1
#define SDB_STATUS_ALLOC_FAILED     2
#define SDB_STATUS_INVALID_VALUE    3
#define SDB_STATUS_EXECUTION_FAILED 4
#define SDB_STATUS_INTERNAL_ERROR   5
#define SDB_STATUS_NOT_SUPPORTED    6

/* ---- enums shared with the reference: _constants.py:13-19 ----------------- */
#define SDB_LAYOUT_ROW_MAJOR 101 /* LAYOUT_CODE_C */
#define SDB_LAYOUT_COL_MAJOR 102 /* LAYOUT_CODE_F */
#define SDB_OP_NON_TRANSPOSE       10
#define SDB_OP_TRANSPOSE           11
#define SDB_OP_CONJUGATE_TRANSPOSE 12

/* value types: the s/d/c/z letter of the MKL routine name */
#define SDB_F32  0
#define SDB_F64  1
#define SDB_C64  2
#define SDB_C128 3

/* storage formats of a handle */
#define SDB_FMT_CSR 0
#define SDB_FMT_CSC 1
#define SDB_FMT_BSR 2

/* Opaque device-resident sparse matrix: replaces sparse_matrix_t
 * (_mkl_interface/_structs.py:5-9). */
typedef struct sdb_mat sdb_mat;

/* ======================= handles ============== */

/* mkl_sparse_?_create_csr / _create_csc (_cfunctions.py:525-536, called at
 * _common.py:310-319).  indptr has rows+1 (CSR) / cols+1 (CSC) entries — the
 * reference's rows_start=indptr[:-1], rows_end=indptr[1:] pair collapsed back
 * into scipy's 3-array form.  index_bits is 32 or 64 and describes BOTH host
 * index arrays (scipy keeps them the same width).  Zero-based only. */
SDB_API sdb_status sdb_create_csr(sdb_mat** out, int64_t rows, int64_t cols,
                                  const void* indptr, const void* indices, int index_bits,
                                  const void* values, int dtype);
SDB_API sdb_status sdb_create_csc(sdb_mat** out, int64_t rows, int64_t cols,
                                  const void* indptr, const void* indices, int index_bits,
                                  const void* values, int dtype);

/* mkl_sparse_?_create_bsr (_cfunctions.py:538-551, called at _common.py:368-379).
 * Square blocks; block_layout is SDB_LAYOUT_ROW_MAJOR / _COL_MAJOR for the
 * elements inside one block. */
SDB_API sdb_status sdb_create_bsr(sdb_mat** out, int64_t block_rows, int64_t block_cols,
                                  int64_t block_size, int block_layout,
                                  const void* indptr, const void* indices, int index_bits,
                                  const void* values, int dtype);

/* Zero-copy CSR handle over arrays ALREADY in HBM (borrowed; the caller keeps
 * them alive).  d_indptr is int64[rows+1], d_indices int32[nnz].  This is the
 * analogue of MKL's zero-copy create for device-resident callers (§8f rank 3)
 * and what bench.py uses for the HBM-resident `value`. */
SDB_API sdb_status sdb_create_csr_dev(sdb_mat** out, int64_t rows, int64_t cols, int64_t nnz,
                                      const int64_t* d_indptr, const int32_t* d_indices,
                                      const void* d_values, int dtype);

/* mkl_sparse_destroy (_common.py:671-680).  NULL -> SDB_STATUS_NOT_INITIALIZED,
 * which the reference surfaces as ValueError (tests/test_mkl.py:128-141). */
SDB_API sdb_status sdb_destroy(sdb_mat* m);

/* Shape query; with sdb_export it replaces mkl_sparse_?_export_csr/csc/bsr
 * (_cfunctions.py:553-564, called at _common.py:442-451, 542-553).  For BSR
 * rows/cols are BLOCK counts and nnz counts BLOCKS, as MKL reports them. */
SDB_API sdb_status sdb_get_info(const sdb_mat* m, int* format, int* dtype,
                                int64_t* rows, int64_t* cols, int64_t* nnz,
                                int64_t* block_size, int* block_layout);

/* Copy the three arrays into caller-owned host memory (numpy allocates them, so
 * the reference's extra copy at _common.py:488-491 disappears).  indptr_bits /
 * indices_bits select int32 or int64 on the host side; any pointer may be NULL
 * to skip that array. */
SDB_API sdb_status sdb_export(const sdb_mat* m, void* indptr, int indptr_bits,
                              void* indices, int indices_bits, void* values);

/* The interior DEVICE pointers of a handle — what mkl_sparse_?_export_csr returns on the host
 * (_common.py:442-451: pointers into MKL-owned memory, valid until destroy): int64 row offsets [major + 1],
 * int32 indices [nnz], values [nnz * block^2].  Borrowed, valid until sdb_destroy / sdb_order.  Lets
 * device-resident callers (the multi-GPU SpGEMM exchange over NCCL) move a result without a host round trip. */
SDB_API sdb_status sdb_export_dev(const sdb_mat* m, const int64_t** d_indptr, const int32_t** d_indices,
                                  const void** d_values);

/* mkl_sparse_order (_common.py:683-692): ascending column order inside every
 * row (CSR) / column (CSC) / block row (BSR), values permuted along. */
SDB_API sdb_status sdb_order(sdb_mat* m);

/* Drop everything the library has cached on a handle (transposed / expanded companions, cross
 * positions, the slab-ordered copy of the streaming SpMM, the sortedness flag).  Needed only for handles
 * over BORROWED device arrays (sdb_create_csr_dev) whose owner rewrites the arrays in place: the caches
 * hold copies of the values.  Handles may be shared by several host threads for products; triangular
 * products (sdb_syrk*, sdb_syrkd*), sdb_order and sdb_invalidate must not run concurrently with other
 * calls on the same handle.  No MKL counterpart (MKL's handles borrow and never cache). */
SDB_API sdb_status sdb_invalidate(sdb_mat* m);

/* mkl_sparse_convert_csr (_common.py:695-722): CSC or BSR (or CSR) -> new CSR
 * handle; op must be SDB_OP_NON_TRANSPOSE or SDB_OP_TRANSPOSE. */
SDB_API sdb_status sdb_convert_csr(const sdb_mat* m, int op, sdb_mat** out);

/* ======================= SpMM ====================== */

/* mkl_sparse_{s,d,c,z}_mm (_cfunctions.py:611-625, called at
 * _sparse_dense.py:111-123):   Y := alpha * op(A) * X + beta * Y.
 * alpha/beta point at {re, im} doubles (im ignored for real dtypes) — MKL takes
 * them by value in the matrix dtype.  X is (k x n), Y is (m x n) with
 * (m, k) = shape of op(A); `layout` applies to both; ldx/ldy are leading
 * dimensions in elements.  X and Y are HOST pointers; the copies to and from
 * HBM happen inside the call (pinned fast path when the memory is page-locked).
 * beta == 0 means Y is write-only (never read, NaNs in Y do not propagate). */
SDB_API sdb_status sdb_spmm(int op, const double* alpha, const sdb_mat* A, int layout,
                            const void* X, int64_t n, int64_t ldx,
                            const double* beta, void* Y, int64_t ldy);

/* The reference's create -> mm -> destroy triple (_sparse_dense.py:34-132) as
 * ONE call for the common case op = N, A CSR, X / Y row-major: all arrays are
 * HOST pointers.  With page-locked buffers (sdb_host_alloc or any
 * cudaHostAlloc'ed memory) the uploads, the kernels and the download run as a
 * row-chunk pipeline on three streams; otherwise it is exactly
 * sdb_create_csr + sdb_spmm + sdb_destroy.  After a pipelined call
 * sdb_last_timing reports overlapping spans: [0] start -> last upload landed,
 * [1] sum of kernel times, [2] the whole call. */
SDB_API sdb_status sdb_spmm_csr_host(int64_t rows, int64_t cols, const void* indptr, const void* indices,
                                     int index_bits, const void* values, int dtype, const double* alpha,
                                     const void* X, int64_t n, int64_t ldx, const double* beta, void* Y,
                                     int64_t ldy);

/* Same operation on DEVICE pointers, stream-ordered on `stream` (a
 * cudaStream_t passed as void*; NULL = the library's own stream).
 *
 * Inspector / executor (the analogue of mkl_sparse_optimize, which the reference
 * never calls): when a handle created from HOST arrays is multiplied a second
 * time by a row-major fp32 / fp64 panel whose rows are a multiple of 512 bytes
 * and which is several times larger than the L2 cache, the library builds, once,
 * a column-slab-ordered copy of the stored entries (nnz * (4 + sizeof value)
 * bytes of HBM, released by sdb_order / sdb_destroy) and switches to the
 * L2-tiled streaming kernel (csrc/spmm_slab.cu).  Results are the same up to
 * the order of the floating-point additions inside a row.  SDB_SLAB=1 in the
 * environment disables this; handles over borrowed device arrays
 * (sdb_create_csr_dev) never cache a copy. */
SDB_API sdb_status sdb_spmm_dev(int op, const double* alpha, const sdb_mat* A, int layout,
                                const void* dX, int64_t n, int64_t ldx,
                                const double* beta, void* dY, int64_t ldy, void* stream);

/* Row-sharded SpMM fused with the all-gather of the output panel:
 * this rank owns rows [row0, row0 + A.rows) of the global product and the call
 * leaves them in EACH of the n_peers full-size row-major output panels
 * (peer-mapped device pointers, ld = ldy, own rank included); once `stream` has
 * passed the call, this rank's rows are in every panel.  No separate collective
 * is needed.  beta reads the local panel (dY_peers[self]).  Row-major, op = N.
 * Exchange strategies (environment SDB_ALLGATHER overrides the automatic choice:
 * "ce" up to 4 ranks, "k1" beyond, as measured on one 8-GPU box):
 *   "ce"            the shard is cut into a few row chunks; chunk c's kernel
 *                   writes the local panel and the copy engines push its rows
 *                   to the peers over NVLink (one stream per peer) while chunk
 *                   c + 1's kernel runs, so no SM waits for NVLink;
 *   "stores"        one kernel whose epilogue stores every finished 16-byte
 *                   slice into every peer panel itself;
 *   "k1"            "stores" with the row-gather kernel even where the L2-tiled
 *                   streaming kernel would qualify (its CTAs finish continuously,
 *                   which spreads the peer stores over the whole kernel);
 *   "sm"            row chunks like "ce", but a few copier CTAs push each finished
 *                   chunk: every 16-byte pack is read from the local panel once and
 *                   stored into all peer panels, on SMs the streaming kernel's
 *                   persistent grid leaves free (sdb_set_allgather_sms, default 12). */
SDB_API sdb_status sdb_spmm_dev_allgather(const double* alpha, const sdb_mat* A,
                                          const void* dX, int64_t n, int64_t ldx,
                                          const double* beta, void* const* dY_peers,
                                          int n_peers, int self, int64_t row0, int64_t ldy,
                                          void* stream);

/* Choose the exchange strategy of sdb_spmm_dev_allgather at run time (process-wide; overrides
 * SDB_ALLGATHER): strategy 0 = automatic, 1 = "ce", 2 = "stores", 3 = "k1", 4 = "sm"; chunks = target number of
 * row chunks of the "ce" pipeline (0 = default, about five).  RowShardedSpMM.autotune() times the
 * candidates on the live NVLink topology during warm-up and keeps the fastest.  No reference
 * counterpart (the reference has no multi-device path). */
SDB_API sdb_status sdb_set_allgather(int strategy, int chunks);
/* SMs kept free of SpMM CTAs for the copier CTAs of strategy 4 ("sm"), 1..64. */
SDB_API sdb_status sdb_set_allgather_sms(int sms);

/* ======================= SpGEMM ==================== */

/* mkl_sparse_spmm (_cfunctions.py:376-382, called at _sparse_sparse.py:35-40):
 * C := op(A) * B as a NEW handle in the format of A (CSR x CSR -> CSR,
 * CSC x CSC -> CSC, BSR x BSR -> BSR); structural nonzeros are kept (numeric
 * cancellation does not drop an entry), columns inside a row are NOT sorted
 * until sdb_order is called. */
SDB_API sdb_status sdb_spgemm(int op, const sdb_mat* A, const sdb_mat* B, sdb_mat** C);

/* mkl_sparse_spmm followed by mkl_sparse_order on the result (the reference's
 * reorder_output=True path, _sparse_sparse.py:219-228) as one call: every row of
 * C comes out with ascending columns — the hash bins sort inside shared memory
 * before they write, the bitmap bin emits in order anyway — so the result
 * makes one trip to HBM instead of three. */
SDB_API sdb_status sdb_spgemm_ordered(int op, const sdb_mat* A, const sdb_mat* B, sdb_mat** C);

/* mkl_sparse_?_spmmd (_cfunctions.py:600-609, called at _sparse_sparse.py:94-101):
 * dense C := op(A) * B, OVERWRITING the host array C (no beta). */
SDB_API sdb_status sdb_spgemm_dense(int op, const sdb_mat* A, const sdb_mat* B,
                                    int layout, void* C, int64_t ldc);
SDB_API sdb_status sdb_spgemm_dense_dev(int op, const sdb_mat* A, const sdb_mat* B,
                                        int layout, void* dC, int64_t ldc, void* stream);

/* ======================= SYRK ===================== */

/* mkl_sparse_syrk (_cfunctions.py:456-461, called at _gram_matrix.py:70-74):
 * upper triangle of  A^T*A (op = SDB_OP_TRANSPOSE)  or  A*A^T
 * (op = SDB_OP_NON_TRANSPOSE)  as a new CSR handle. */
SDB_API sdb_status sdb_syrk(int op, const sdb_mat* A, sdb_mat** C);
/* ... with the result ordered (reorder_output=True, _gram_matrix.py:79-80). */
SDB_API sdb_status sdb_syrk_ordered(int op, const sdb_mat* A, sdb_mat** C);

/* mkl_sparse_?_syrkd (_cfunctions.py:639-649, called at _gram_matrix.py:149-157):
 * dense C := alpha * op-product + beta * C on the UPPER triangle of the host
 * array C; the strict lower triangle is left untouched. */
SDB_API sdb_status sdb_syrkd(int op, const sdb_mat* A, const double* alpha, const double* beta,
                             void* C, int layout, int64_t ldc);
/* Same product into a FRESH host array (the reference's out=None case,
 * _gram_matrix.py:136-139 + the lower-triangle clean-up at :168-169): upper
 * triangle = alpha * product, strict lower triangle = 0, nothing is uploaded. */
SDB_API sdb_status sdb_syrkd_new(int op, const sdb_mat* A, const double* alpha, void* C, int layout,
                                 int64_t ldc);
SDB_API sdb_status sdb_syrkd_dev(int op, const sdb_mat* A, const double* alpha,
                                 const double* beta, void* dC, int layout, int64_t ldc,
                                 void* stream);

/* ======================= dense x dense (drop-in completeness) ============== */

/* cblas_{s,d,c,z}gemm (_cfunctions.py:582-598 argtypes, called at _dense_dense.py:55-68) and
 * cblas_{s,d,c,z}syrk (_cfunctions.py:652-668, called at _gram_matrix.py:233-245) for callers that hand
 * dot_product_mkl / gram_matrix_mkl two dense arrays.  CBLAS integer codes: layout 101 / 102, transpose 111 /
 * 112 / 113, uplo 121 (upper; the only one the reference uses).  HOST pointers; alpha / beta as {re, im}
 * doubles; dtype SDB_F32..SDB_C128.  Not part of the sparse hot path: a plain tiled CUDA kernel, not a tuned
 * GEMM.  sdb_syrk_dense writes only the upper triangle and leaves the rest of C as it was. */
SDB_API sdb_status sdb_gemm(int layout, int transa, int transb, int64_t m, int64_t n, int64_t k,
                            const double* alpha, const void* A, int64_t lda, const void* B, int64_t ldb,
                            const double* beta, void* C, int64_t ldc, int dtype);
SDB_API sdb_status sdb_syrk_dense(int layout, int uplo, int trans, int64_t n, int64_t k, const double* alpha,
                                  const void* A, int64_t lda, const double* beta, void* C, int64_t ldc, int dtype);

/* ======================= host-side helpers ================================= */

/* nnz-balanced contiguous row blocks for the multi-GPU path:
 * bounds[0] = 0 <= bounds[1] <= ... <= bounds[parts] = rows.  Pure host code. */
SDB_API sdb_status sdb_partition_rows(const void* indptr, int index_bits, int64_t rows,
                                      int parts, int64_t* bounds);

/* Page-locked host buffers for callers who want the DMA fast path. */
SDB_API sdb_status sdb_host_alloc(void** p, size_t bytes);
SDB_API sdb_status sdb_host_free(void* p);

/* Plain device memory (cudaMalloc, so it can be shared across processes) and
 * synchronous copies, for device-resident callers that do not bring their own
 * allocator.  kind: 1 = host->device, 2 = device->host, 3 = device->device. */
SDB_API sdb_status sdb_dev_alloc(void** p, size_t bytes);
SDB_API sdb_status sdb_dev_free(void* p);
SDB_API sdb_status sdb_memcpy(void* dst, const void* src, size_t bytes, int kind);
/* The same for a panel of `rows` rows of `row_bytes` each with different pitches (bytes between row starts) on
 * the two sides, kind 1 or 2: a column range of a row-major array, or an array with a leading dimension.  Pageable
 * host memory is staged through the page-locked ring by the library's copy threads, as in sdb_memcpy. */
SDB_API sdb_status sdb_memcpy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t row_bytes,
                                 size_t rows, int kind);
SDB_API sdb_status sdb_device_synchronize(void);

/* Peer mapping of an sdb_dev_alloc buffer into another process on the same
 * node (one process per GPU, NVLink/NVSwitch underneath): the owner exports a
 * 64-byte token, every peer opens it and gets a device pointer it can hand to
 * sdb_spmm_dev_allgather.  sdb_ipc_close unmaps (peers only; the owner frees). */
#define SDB_IPC_TOKEN_BYTES 64
SDB_API sdb_status sdb_ipc_export(const void* d_ptr, char token[SDB_IPC_TOKEN_BYTES]);
SDB_API sdb_status sdb_ipc_open(const char token[SDB_IPC_TOKEN_BYTES], void** d_ptr);
SDB_API sdb_status sdb_ipc_close(void* d_ptr);

/* Device selection / introspection (the analogue of mkl_get_max_threads and
 * mkl_get_version_string, _cfunctions.py:738-767). */
SDB_API sdb_status sdb_device_count(int* n);
SDB_API sdb_status sdb_set_device(int device);
SDB_API sdb_status sdb_get_device(int* device);
SDB_API sdb_status sdb_version_string(char* buf, int len);
SDB_API int        sdb_last_error(char* buf, int len);

/* Run-time switches (process-wide), each also readable from the environment at first use:
 *   "bsr_mma"       SDB_BSR_MMA        BSR x dense on the tensor cores (DMMA fp64, 3xTF32 fp32): -1 automatic, 0 off, 1 on
 *   "spgemm_wide"   SDB_SPGEMM_WIDE    wide SpGEMM rows: 0 automatic (bitmap with a shared-memory summary), 1 full-sweep bitmap
 *   "dense_mode"    SDB_DENSE_MODE     dense-output products: 0 automatic, 1 shared-memory tiles, 2 global reductions
 *   "dense_threads" SDB_DENSE_THREADS  threads per CTA of the global-reduction kernel (0 = default)
 *   "dense_ctas"    SDB_DENSE_CTAS     resident CTAs per SM of that kernel (0 = as many as fit)
 *   "slab_keep"     SDB_SLAB_KEEP      streaming SpMM: gathers of X rows carry an L2 evict_last policy (0 / 1)
 *   "spgemm_sorted_cta" SDB_SPGEMM_SORTED_CTA  sorted SpGEMM: 1 keeps rows of 1025..4096 entries in the CTA hash bin (0: bitmap bin)
 *   "spmv_wide"     SDB_SPMV_WIDE      SpMV: 0 automatic (16-byte loads of A, four entries per lane and step), 1 scalar loads
 *   "spmv_tile"     SDB_SPMV_TILE      SpMV with x staged slab by slab in shared memory (inspector + executor, cached on the
 *                                      handle): 0 automatic (from the second product of a large matrix with a vector), 1 never, 2 always
 * Unknown names return SDB_STATUS_INVALID_VALUE.  No reference counterpart (tuning aid for tests and sweeps). */
SDB_API sdb_status sdb_set_option(const char* name, int value);

/* Bandwidth probes measured on the device the numbers are quoted on (bench.py's roofline
 * denominators): kind 0 = HBM read (sequential 16-byte loads over `bytes`, pick a size much larger than
 * L2), 1 = L2 -> SM read (same loop over an L2-resident `bytes`, L1 bypassed), 2 = L2 -> SM gather of
 * whole 512-byte rows at random positions of an L2-resident buffer (the access shape of the SpMM
 * gathers), 3 = the host side of the pageable-memory pipeline: pageable -> page-locked copies in 16 MiB
 * slots by the library's copy threads (no GPU involved).  Result in GB/s (1e9 bytes per second), timed with CUDA events over `iters` launches after
 * one warm-up launch.  No reference counterpart. */
SDB_API sdb_status sdb_probe_bandwidth(int kind, int64_t bytes, int iters, double* gbs);

/* Number of CUDA kernels this library has launched in this process (all
 * threads); bench.py reports the delta over the timed region. */
SDB_API int64_t sdb_kernel_launches(void);

/* Name (with template arguments) of the SpMM kernel the most recent sdb_spmm* call on this thread
 * launched, e.g. "spmm_stream_kernel<float,6,32,2,2>"; empty before the first call.  bench.py labels
 * its roofline with it.  No reference counterpart (tracing aid, like the reference's print_mkl_debug,
 * _mkl_interface/_common.py:97-136). */
SDB_API sdb_status sdb_last_spmm_kernel(char* buf, int len);

/* Device time (ms, CUDA events on the launch stream) the most recent
 * host-pointer entry point on this thread spent in: [0] host->device copies,
 * [1] kernels, [2] device->host copies.  The reference's debug_timer
 * (_common.py:138-155) prints the same phases. */
SDB_API sdb_status sdb_last_timing(double ms[3]);

#ifdef __cplusplus
}
#endif
#endif /* SDB200_H */
