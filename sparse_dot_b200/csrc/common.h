// common.h — shared plumbing of libsdb200: status/error convention, the opaque
// handle, the per-thread execution context (stream, pinned staging ring, phase
// timers) and small RAII helpers.  Host-side C++17; included by every .cu file.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/sdb200.h"

namespace sdb {

// ---------------------------------------------------------------- errors
// The ABI never throws: every internal routine returns an sdb_status and leaves
// a message in a thread-local buffer (sdb_last_error).
void set_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
sdb_status cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define SDB_CUDA(expr)                                                     \
    do {                                                                   \
        cudaError_t _e = (expr);                                           \
        if (_e != cudaSuccess) return ::sdb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define SDB_TRY(expr)                                 \
    do {                                              \
        sdb_status _s = (expr);                       \
        if (_s != SDB_STATUS_SUCCESS) return _s;      \
    } while (0)

#define SDB_REQUIRE(cond, status, ...)     \
    do {                                   \
        if (!(cond)) {                     \
            ::sdb::set_error(__VA_ARGS__); \
            return (status);               \
        }                                  \
    } while (0)

// Every kernel launch goes through this so sdb_kernel_launches() is exact.
extern std::atomic<int64_t> g_launches;
#define SDB_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
    do {                                                                   \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);        \
        ::sdb::g_launches.fetch_add(1, std::memory_order_relaxed);         \
        SDB_CUDA(cudaGetLastError());                                      \
    } while (0)

// Phase tracing for development (SDB_TRACE=1): synchronises `s` and prints the milliseconds since
// the previous trace point on this thread.  A no-op (no sync) when the variable is unset.
void trace(cudaStream_t s, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
// Name of the SpMM kernel the most recent sdb_spmm* call on this thread launched (sdb_last_spmm_kernel).
void note_spmm_kernel(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
extern thread_local char t_spmm_kernel[128];

// ---------------------------------------------------------------- run-time options
// Small integer switches, each initialised from an environment variable on first use and settable at run time
// through sdb_set_option (tests and sweeps flip them inside one process).  -1 / 0 mean "automatic".
enum Option {
    kOptBsrMma = 0,     // "bsr_mma"       SDB_BSR_MMA        BSR x dense on tensor cores: -1 auto, 0 off, 1 on
    kOptSpgemmWide,     // "spgemm_wide"   SDB_SPGEMM_WIDE    wide SpGEMM rows: 0 auto (bitmap with summary), 1 full-sweep bitmap
    kOptDenseMode,      // "dense_mode"    SDB_DENSE_MODE     dense-output products: 0 auto, 1 shared tiles, 2 global reductions
    kOptDenseThreads,   // "dense_threads" SDB_DENSE_THREADS  threads per CTA of the global-reduction kernel (0 = 1024)
    kOptDenseCtas,      // "dense_ctas"    SDB_DENSE_CTAS     resident CTAs per SM of that kernel (0 = what fits)
    kOptSpgemmSortedCta,  // "spgemm_sorted_cta" SDB_SPGEMM_SORTED_CTA  sorted SpGEMM: 1 keeps 1025..4096-entry rows in the CTA hash bin
    kOptSlabKeep,       // "slab_keep"     SDB_SLAB_KEEP      streaming SpMM gathers with an L2 evict_last policy (0 / 1)
    kOptSpmvWide,       // "spmv_wide"     SDB_SPMV_WIDE      SpMV: 0 auto (16-byte loads of A when rows are long enough), 1 scalar loads
    kOptSpmvTile,       // "spmv_tile"     SDB_SPMV_TILE      SpMV with x staged in shared memory: 0 auto (repeated products), 1 never, 2 always
    kOptCount
};
int get_option(Option o);

// ---------------------------------------------------------------- dtypes
inline const char* dtype_cname(int dtype) {
    static const char* const names[4] = {"float", "double", "complex<float>", "complex<double>"};
    return dtype >= 0 && dtype < 4 ? names[dtype] : "?";
}
inline size_t dtype_size(int dtype) {
    switch (dtype) {
        case SDB_F32: return 4;
        case SDB_F64: return 8;
        case SDB_C64: return 8;
        case SDB_C128: return 16;
        default: return 0;
    }
}

// ---------------------------------------------------------------- context
// One per host thread and device: a non-blocking stream, a ring of pinned
// staging chunks for pageable<->HBM copies, and CUDA-event phase timers.
// Page-locked staging slots with one "slot is free again" event each (pipeline.cu).
struct PinnedRing {
    static constexpr int kSlots = 8;
    size_t slot_bytes = 0;
    void* slot[kSlots] = {};
    cudaEvent_t free_ev[kSlots] = {};
    int next = 0;
};

struct Context {
    int device = -1;
    cudaStream_t stream = nullptr;   // compute + H2D
    cudaStream_t d2h_stream = nullptr;
    cudaStream_t h2d_stream = nullptr;  // uploads of the pipelined host entry points
    int sm_count = 0;
    size_t l2_bytes = 0;
    // pinned staging ring
    static constexpr int kChunks = 4;
    static constexpr size_t kChunkBytes = size_t(32) << 20;
    void* chunk[kChunks] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t chunk_free[kChunks] = {nullptr, nullptr, nullptr, nullptr};
    int next_chunk = 0;
    // staging rings of the pipelined host entry point (pageable operands): uploads / downloads
    PinnedRing up_ring, dn_ring;
    // phase timing of the last host-pointer entry point (ms)
    double last_ms[3] = {0, 0, 0};
    // chunk-pipelined all-gather (sdb_spmm_dev_allgather): one copy stream per peer, a ring of events
    static constexpr int kExchangePeers = 8;
    static constexpr int kExchangeEvents = 32;
    cudaStream_t xchg_stream[kExchangePeers] = {};
    cudaEvent_t xchg_event[kExchangeEvents] = {};
    cudaEvent_t xchg_done[kExchangePeers] = {};
    int xchg_next_event = 0;
};

sdb_status get_context(Context** out);

// Stream-ordered device allocation from the device's default memory pool (the
// pool keeps freed blocks, so repeated calls do not hit cudaMalloc).
sdb_status dev_alloc(void** p, size_t bytes, cudaStream_t s);
void dev_free(void* p, cudaStream_t s);

// RAII temp buffer; freed stream-ordered on destruction.
struct DevBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { reset(); }
    sdb_status alloc(size_t bytes, cudaStream_t stream) {
        reset();
        s = stream;
        return dev_alloc(&p, bytes ? bytes : 16, stream);
    }
    void reset() {
        if (p) dev_free(p, s);
        p = nullptr;
    }
    void* release() {
        void* r = p;
        p = nullptr;
        return r;
    }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

// Host <-> HBM copies that DMA straight from/to page-locked memory and stage
// pageable memory through the context's pinned ring.  Both are stream-ordered
// on ctx->stream; h2d returns once the host buffer may be reused, d2h returns
// once the bytes are in `dst`.  2-D variants copy `rows` rows of `row_bytes`
// with independent pitches (dense panels with ld != n).
bool is_pinned(const void* p);  // page-locked (or managed) host memory: DMA without staging
void host_copy(void* dst, const void* src, size_t bytes);  // memcpy split over the library's copy threads
void host_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t row_bytes, size_t rows);
sdb_status ensure_ring(PinnedRing* ring, size_t slot_bytes);
sdb_status h2d(Context* ctx, void* d_dst, const void* h_src, size_t bytes);
sdb_status d2h(Context* ctx, void* h_dst, const void* d_src, size_t bytes);
sdb_status h2d_2d(Context* ctx, void* d_dst, size_t d_pitch, const void* h_src, size_t h_pitch,
                  size_t row_bytes, size_t rows);
sdb_status d2h_2d(Context* ctx, void* h_dst, size_t h_pitch, const void* d_src, size_t d_pitch,
                  size_t row_bytes, size_t rows);

// CUDA-event stopwatch for the H2D / kernel / D2H phases.
struct PhaseTimer {
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t s = nullptr;
    sdb_status init(cudaStream_t stream);
    sdb_status mark(int i);  // 0: start, 1: after H2D, 2: after kernels, 3: after D2H
    void finish(Context* ctx);
    ~PhaseTimer();
};

}  // namespace sdb

// ---------------------------------------------------------------- the handle
// Device-resident sparse matrix.  indptr is always int64 and indices int32 in
// HBM whatever the host handed us (SURVEY §8b "Index width").  For BSR, rows /
// cols / nnz count BLOCKS and `values` holds nnz * block * block elements.
struct sdb_mat {
    uint32_t magic;
    int format;        // SDB_FMT_*
    int dtype;         // SDB_F32 ...
    int64_t rows, cols;
    int64_t nnz;
    int64_t block;     // 1 for CSR/CSC
    int block_layout;  // SDB_LAYOUT_* (BSR only)
    int64_t* indptr;
    int32_t* indices;
    void* values;
    bool owns;         // false for sdb_create_csr_dev (borrowed HBM arrays)
    int device;
    // Lazily built companion holding the transposed compressed form (CSR of
    // A^T); invalidated by sdb_order.  Lets op=T and CSC inputs reuse the
    // gather kernel instead of a scatter/atomic path.
    sdb_mat* transposed;
    // Lazily built CSR expansion of a BSR handle (every stored block becomes
    // block*block explicit entries); lets BSR reuse the CSR kernels.
    sdb_mat* expanded;
    // Optional cross positions (built together with the transposed companion when a triangular
    // product asks for them): for stored entry p = (line i, index k), pos[p] is the position of
    // index i inside line k of the companion.  Valid only while strict_sorted == 1.
    int32_t* pos;
    int strict_sorted;  // 0 unknown, 1 every line strictly ascending (no duplicates), -1 not
    // Optional slab-ordered copy of the stored entries for the L2-tiled SpMM (spmm_slab.cu, built by its
    // inspector on the second multiplication with this handle): inside every group of slab_rpw consecutive
    // rows the entries are ordered by (column / slab_width, row, column); slab_rc[p] = local row << 27 | column,
    // slab_val[p] the value.  Dropped by sdb_order.
    void* slab_rc;
    void* slab_val;
    int slab_rpw;
    int64_t slab_width;
    int spmm_calls;  // multiplications seen so far (inspector policy)
    // Optional tile-ordered copy for the shared-memory SpMV (spmv_tile.cu, built on the second product with a
    // vector): entries grouped by (row block, column slab).  vt_state: 0 not built, 1 built, -1 the inspector
    // could not balance the matrix (stays on the gather kernel).
    // Rows longer than long_threshold entries (SpMV / SpMM give those one CTA each; spmm.cu): built on the first
    // product, dropped with the other caches.  long_state: 0 unknown, 1 built (n_long may be 0).
    int32_t* long_rows[2];  // slot 0: the SpMV kernels' threshold, slot 1: the SpMM kernel's
    int32_t n_long[2];
    int64_t long_threshold[2];
    int long_state[2];
    void* vt_cache;  // spmv_tile.cu's TileCache (host object owning the device arrays)
    int vt_state;
    int spmv_calls;
};

namespace sdb {
constexpr uint32_t kMagic = 0x5db200a5u;
inline bool valid(const sdb_mat* m) { return m != nullptr && m->magic == kMagic; }

// major dimension of the compressed arrays (rows for CSR/BSR, cols for CSC)
inline int64_t major_dim(const sdb_mat* m) { return m->format == SDB_FMT_CSC ? m->cols : m->rows; }
inline int64_t minor_dim(const sdb_mat* m) { return m->format == SDB_FMT_CSC ? m->rows : m->cols; }

sdb_status new_handle(sdb_mat** out, int format, int dtype, int64_t rows, int64_t cols,
                      int64_t nnz, int64_t block, int block_layout, cudaStream_t s);
void free_handle(sdb_mat* m);

// Internal device-level operations shared between translation units -------------
// CSR(A) -> CSR(A^T) as a new owned handle (rows/cols swapped), rows sorted.
sdb_status transpose_compressed(Context* ctx, sdb_mat* a, sdb_mat** out, bool with_pos = false);
// The CSR view of `m` for op(A): returns arrays such that row r lists op(A)[r,:].
struct CsrView {
    int64_t rows, cols, nnz;
    const int64_t* indptr;
    const int32_t* indices;
    const void* values;
    // For entry p = (row i, column k): where column i sits in row k of the TRANSPOSE of this view
    // (nullptr when not requested / not available).  A triangular product L * L^T restricted to
    // col >= row starts its walk of row k of L^T there instead of at the row's first entry.
    const int32_t* pos = nullptr;
    // the handle whose arrays these are (nullptr for ad-hoc views): lets kernels cache per-matrix indexes
    sdb_mat* owner = nullptr;
    // SpMM only: multiply just rows [sub_begin, sub_begin + sub_rows) of the view (sub_rows < 0 = all of them).
    // Used by the chunk-pipelined all-gather; per-matrix indexes are always built from the whole view.
    int64_t sub_begin = 0;
    int64_t sub_rows = -1;
};
sdb_status csr_view(Context* ctx, const sdb_mat* m, bool transpose, CsrView* v, bool want_pos = false);
// shared-memory SpMV (spmv_tile.cu): policy, executor (+ inspector on first use), and release of the cached tiles
bool spmv_tile_wanted(const CsrView& a, int dtype, int64_t incx, int64_t incy);
sdb_status spmv_tile_device(Context* ctx, cudaStream_t s, const CsrView& a, int dtype, const double* alpha,
                            const double* beta, const void* dX, void* dY);
void drop_spmv_tiles(sdb_mat* m, cudaStream_t s);
sdb_status sort_rows(Context* ctx, int dtype, int64_t rows, const int64_t* indptr, int32_t* indices,
                     void* values, int64_t elems_per_entry, int32_t* extra = nullptr);
sdb_status expand_bsr(Context* ctx, const sdb_mat* bsr, sdb_mat** out_csr);
// CSR made of whole b-blocks (rows sorted) -> BSR with row-major blocks.
sdb_status compress_to_bsr(Context* ctx, const sdb_mat* csr, int64_t b, sdb_mat** out);
// True when every row's column indices are non-decreasing (device reduction + sync).
sdb_status rows_sorted(Context* ctx, int64_t rows, const int64_t* indptr, const int32_t* indices,
                       bool* sorted);
// SpGEMM core on CSR views: C = L * R (optionally only entries with col >= row).
sdb_status spgemm_device(Context* ctx, const CsrView& l, const CsrView& r, int dtype, bool upper,
                         sdb_mat** out, bool sort = false);

sdb_status spmm_device(Context* ctx, cudaStream_t s, const CsrView& a, int dtype, bool conj_a, const double* alpha,
                       const double* beta, int layout, const void* dX, int64_t n, int64_t ldx,
                       void* const* dY_peers, int n_peers, int self, int64_t row0, int64_t ldy);
// L2-tiled SpMM (spmm_slab.cu): returns SDB_STATUS_NOT_SUPPORTED when the call does not qualify.
bool spmm_slab_wanted(const CsrView& a, int dtype, int64_t n, int64_t ldx);
// rows one wave of the streaming kernel's persistent grid covers (0 when the call would not use it): row sub-ranges
// handed to it must start at a multiple of this
void spmm_slab_reserve_sms(int sms);  // persistent grid leaves this many SMs free (per host thread; 0 = none)
int64_t spmm_slab_wave_rows(const Context* ctx, const CsrView& a, int dtype, int64_t n, int64_t ldx, bool count_call);
sdb_status spmm_slab_device(Context* ctx, cudaStream_t s, const CsrView& a, int dtype, bool conj_a,
                            const double* alpha, const double* beta, const void* dX, int64_t n, int64_t ldx,
                            void* const* dY_peers, int n_peers, int self, int64_t row0, int64_t ldy);
extern std::mutex g_companion_mutex;  // serialises building / dropping the per-handle caches (format.cu)
sdb_status ensure_strict_flag(Context* ctx, sdb_mat* m);
// range / monotonicity check of freshly uploaded arrays; also sets strict_sorted (format.cu)
sdb_status validate_compressed(Context* ctx, sdb_mat* m);
// Native BSR x dense kernel (spmm_bsr.cu): availability test + launch.
bool spmm_bsr_supported(const sdb_mat* a, int op, int layout, const void* dX, int64_t n, int64_t ldx, const void* dY,
                        int64_t ldy);
// tensor-core variant (spmm_bsr_mma.cu): block sizes 8 / 16 / 32, n a multiple of 8
bool spmm_bsr_mma_supported(const sdb_mat* a, int64_t n);
sdb_status spmm_bsr_mma_device(cudaStream_t s, const sdb_mat* a, const double* alpha, const double* beta,
                               const void* dX, int64_t n, int64_t ldx, void* dY, int64_t ldy, int stages);
sdb_status spmm_bsr_device(cudaStream_t s, const sdb_mat* a, const double* alpha, const double* beta, const void* dX,
                           int64_t n, int64_t ldx, void* dY, int64_t ldy);
}  // namespace sdb
