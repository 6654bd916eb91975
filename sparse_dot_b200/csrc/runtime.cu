// runtime.cu — host runtime of libsdb200: error plumbing, per-thread context,
// stream-ordered memory, pinned staging copies, phase timers and the small
// host-only ABI entry points (version string, device selection, row partitioner).
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

#include "common.h"

namespace sdb {

std::atomic<int64_t> g_launches{0};

static thread_local char t_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

sdb_status cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    const char* base = strrchr(file, '/');
    set_error("CUDA error %d (%s) in %s at %s:%d", int(e), cudaGetErrorString(e), what,
              base ? base + 1 : file, line);
    cudaGetLastError();  // clear the sticky-free error state
    return e == cudaErrorMemoryAllocation ? SDB_STATUS_ALLOC_FAILED : SDB_STATUS_EXECUTION_FAILED;
}

thread_local char t_spmm_kernel[128] = "";

void note_spmm_kernel(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_spmm_kernel, sizeof(t_spmm_kernel), fmt, ap);
    va_end(ap);
}

void trace(cudaStream_t s, const char* fmt, ...) {
    static const bool on = [] {
        const char* e = getenv("SDB_TRACE");
        return e && *e && *e != '0';
    }();
    if (!on) return;
    static thread_local std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    cudaStreamSynchronize(s);
    const auto now = std::chrono::steady_clock::now();
    char msg[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg, sizeof(msg), fmt, ap);
    va_end(ap);
    fprintf(stderr, "[sdb %9.3f ms] %s\n", std::chrono::duration<double, std::milli>(now - last).count(), msg);
    last = std::chrono::steady_clock::now();
}

// ---------------------------------------------------------------- options
namespace {
struct OptionSlot {
    const char* name;
    const char* env;
    int fallback;
    std::atomic<int> value;
    std::atomic<bool> set;
};
OptionSlot g_options[kOptCount] = {
    {"bsr_mma", "SDB_BSR_MMA", -1, {0}, {false}},
    {"spgemm_wide", "SDB_SPGEMM_WIDE", 0, {0}, {false}},
    {"dense_mode", "SDB_DENSE_MODE", 0, {0}, {false}},
    {"dense_threads", "SDB_DENSE_THREADS", 0, {0}, {false}},
    {"dense_ctas", "SDB_DENSE_CTAS", 0, {0}, {false}},
    {"spgemm_sorted_cta", "SDB_SPGEMM_SORTED_CTA", 0, {0}, {false}},
    {"slab_keep", "SDB_SLAB_KEEP", 0, {0}, {false}},
    {"spmv_wide", "SDB_SPMV_WIDE", 0, {0}, {false}},
    {"spmv_tile", "SDB_SPMV_TILE", 0, {0}, {false}},
};
}  // namespace

int get_option(Option o) {
    OptionSlot& s = g_options[o];
    if (!s.set.load(std::memory_order_acquire)) {
        const char* e = getenv(s.env);
        s.value.store(e ? atoi(e) : s.fallback, std::memory_order_relaxed);
        s.set.store(true, std::memory_order_release);
    }
    return s.value.load(std::memory_order_relaxed);
}

// ---------------------------------------------------------------- context
static thread_local Context t_ctx[16];

sdb_status get_context(Context** out) {
    int dev = 0;
    SDB_CUDA(cudaGetDevice(&dev));
    SDB_REQUIRE(dev >= 0 && dev < 16, SDB_STATUS_NOT_SUPPORTED, "device index %d out of range", dev);
    Context* c = &t_ctx[dev];
    if (c->device < 0) {
        SDB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        SDB_CUDA(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
        SDB_CUDA(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
        cudaDeviceProp prop;
        SDB_CUDA(cudaGetDeviceProperties(&prop, dev));
        c->sm_count = prop.multiProcessorCount;
        c->l2_bytes = size_t(prop.l2CacheSize);
        // keep freed blocks in the pool: repeated calls reuse them
        cudaMemPool_t pool;
        SDB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = UINT64_MAX;
        SDB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        c->device = dev;
    }
    *out = c;
    return SDB_STATUS_SUCCESS;
}

static sdb_status ensure_staging(Context* c) {
    if (c->chunk[0]) return SDB_STATUS_SUCCESS;
    for (int i = 0; i < Context::kChunks; ++i) {
        SDB_CUDA(cudaHostAlloc(&c->chunk[i], Context::kChunkBytes, cudaHostAllocDefault));
        SDB_CUDA(cudaEventCreateWithFlags(&c->chunk_free[i], cudaEventDisableTiming));
    }
    return SDB_STATUS_SUCCESS;
}

sdb_status ensure_ring(PinnedRing* ring, size_t slot_bytes) {
    if (ring->slot[0] && ring->slot_bytes >= slot_bytes) return SDB_STATUS_SUCCESS;
    for (int i = 0; i < PinnedRing::kSlots; ++i) {
        if (ring->slot[i]) cudaFreeHost(ring->slot[i]);
        ring->slot[i] = nullptr;
        SDB_CUDA(cudaHostAlloc(&ring->slot[i], slot_bytes, cudaHostAllocDefault));
        if (!ring->free_ev[i]) SDB_CUDA(cudaEventCreateWithFlags(&ring->free_ev[i], cudaEventDisableTiming));
    }
    ring->slot_bytes = slot_bytes;
    ring->next = 0;
    return SDB_STATUS_SUCCESS;
}

sdb_status dev_alloc(void** p, size_t bytes, cudaStream_t s) {
    *p = nullptr;
    cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 16, s);
    if (e != cudaSuccess) {
        *p = nullptr;
        return cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
    }
    return SDB_STATUS_SUCCESS;
}

void dev_free(void* p, cudaStream_t s) {
    if (p) cudaFreeAsync(p, s);
}

bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// ---------------------------------------------------------------- host copy pool
// Pageable host memory cannot be DMA'd; it is staged through page-locked ring buffers, and one core
// moves only ~10 GB/s where PCIe 5 wants ~55.  A small persistent pool of worker threads (created on
// first use, SDB_COPY_THREADS overrides the count) splits every large copy; the calling thread
// takes pieces too.  Several threads may submit copies at the same time (the upload and download
// sides of sdb_spmm_csr_host do).
namespace {

// One piece of a staged copy.  The destination is either a page-locked slot the DMA engine reads next or the
// caller's array after a download: in both cases nothing on this core re-reads it soon, so it is written with
// non-temporal stores (no read-for-ownership of the destination lines: 2 bytes of memory traffic per byte
// copied instead of 3).  glibc's memcpy only does that above a threshold far larger than a piece.
void stream_copy(char* dst, const char* src, size_t bytes) {
#if defined(__SSE2__)
    // head up to 16-byte alignment of the destination
    size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
    if (head > bytes) head = bytes;
    memcpy(dst, src, head);
    dst += head;
    src += head;
    bytes -= head;
    size_t blocks = bytes / 64;
    for (size_t i = 0; i < blocks; ++i) {
        __builtin_prefetch(src + 512);
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 32));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 48), d);
        src += 64;
        dst += 64;
    }
    _mm_sfence();
    memcpy(dst, src, bytes - blocks * 64);
#else
    memcpy(dst, src, bytes);
#endif
}

bool use_stream_copy() {
    static const bool v = [] {
        const char* e = getenv("SDB_COPY_NT");
        return !(e && e[0] == '0');
    }();
    return v;
}

class CopyPool {
  public:
    static CopyPool& get() {
        static CopyPool* pool = new CopyPool();  // never destroyed: workers may outlive static destruction
        return *pool;
    }
    void copy(void* dst, const void* src, size_t bytes) {
        static const size_t kPiece = [] {  // smallest piece handed to a thread (SDB_COPY_PIECE_KB, default 1 MiB)
            const char* e = getenv("SDB_COPY_PIECE_KB");
            const int kb = e ? atoi(e) : 1024;
            return size_t(kb >= 64 && kb <= (1 << 20) ? kb : 1024) << 10;
        }();
        if (bytes < 2 * kPiece || workers_ == 0) {
            if (use_stream_copy() && bytes >= (size_t(256) << 10)) stream_copy(static_cast<char*>(dst), static_cast<const char*>(src), bytes);
            else memcpy(dst, src, bytes);
            return;
        }
        Job job;
        job.dst = static_cast<char*>(dst);
        job.src = static_cast<const char*>(src);
        job.bytes = bytes;
        // about one piece per thread, multiples of 4 KiB
        size_t piece = bytes / size_t(workers_ + 1);
        piece = std::max(kPiece, (piece + 4095) & ~size_t(4095));
        job.piece = piece;
        job.n_pieces = (bytes + piece - 1) / piece;
        {
            std::lock_guard<std::mutex> lk(m_);
            jobs_.push_back(&job);
        }
        cv_.notify_all();
        run_pieces(&job);  // the submitter works too
        std::unique_lock<std::mutex> lk(m_);
        job.done_cv.wait(lk, [&] { return job.finished == job.n_pieces; });
    }
    // `rows` rows of `row_bytes` each, row r from src + r * spitch to dst + r * dpitch: pieces are row ranges
    void copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t row_bytes, size_t rows) {
        if (rows == 0 || row_bytes == 0) return;
        if (dpitch == row_bytes && spitch == row_bytes) return copy(dst, src, row_bytes * rows);
        Job job;
        job.dst = static_cast<char*>(dst);
        job.src = static_cast<const char*>(src);
        job.bytes = rows;  // counted in rows for a 2-D job
        job.row_bytes = row_bytes;
        job.dpitch = dpitch;
        job.spitch = spitch;
        // about one piece per thread, at least ~256 KiB of payload each
        const size_t min_rows = std::max<size_t>(1, (size_t(256) << 10) / row_bytes);
        job.piece = std::max(min_rows, (rows + size_t(workers_)) / size_t(workers_ + 1));
        job.n_pieces = (rows + job.piece - 1) / job.piece;
        if (job.n_pieces > 1 && workers_ > 0) {
            {
                std::lock_guard<std::mutex> lk(m_);
                jobs_.push_back(&job);
            }
            cv_.notify_all();
        } else {
            job.n_pieces = 1;
            job.piece = rows;
            job.local = true;
        }
        run_pieces(&job);
        std::unique_lock<std::mutex> lk(m_);
        job.done_cv.wait(lk, [&] { return job.finished == job.n_pieces; });
    }

  private:
    struct Job {
        char* dst;
        const char* src;
        size_t bytes, piece, n_pieces;
        size_t row_bytes = 0, dpitch = 0, spitch = 0;  // row_bytes != 0: a 2-D job, `bytes` and `piece` count rows
        bool local = false;                            // never queued: the submitter runs its only piece
        size_t next = 0;      // guarded by m_
        size_t finished = 0;  // guarded by m_
        std::condition_variable done_cv;
    };
    CopyPool() {
        unsigned hw = std::thread::hardware_concurrency();
        // one process per GPU (torchrun exports LOCAL_WORLD_SIZE): the ranks of a node share its cores
        if (const char* lw = getenv("LOCAL_WORLD_SIZE")) {
            const int ranks = atoi(lw);
            if (ranks > 1) hw = std::max(1u, hw / unsigned(ranks));
        }
        int n = hw > 1 ? int(hw) - 1 : 0;
        if (n > 15) n = 15;
        if (const char* e = getenv("SDB_COPY_THREADS")) n = std::max(0, std::min(63, atoi(e) - 1));
        workers_ = n;
        for (int i = 0; i < n; ++i) std::thread([this] { worker(); }).detach();
    }
    // Hand out the next piece of `job` (m_ held).  A job leaves the queue when its last piece is handed
    // out, so nobody can pick it up afterwards; it may be destroyed by its submitter as soon as the last
    // `finished` increment has been published.
    size_t claim(Job* job) {
        const size_t idx = job->next++;
        if (job->next == job->n_pieces && !job->local) {
            for (size_t i = 0; i < jobs_.size(); ++i)
                if (jobs_[i] == job) {
                    jobs_.erase(jobs_.begin() + long(i));
                    break;
                }
        }
        return idx;
    }
    void copy_piece(Job* job, size_t idx) {
        const size_t off = idx * job->piece;
        const size_t len = std::min(job->piece, job->bytes - off);
        if (job->row_bytes) {
            const bool nt = use_stream_copy() && job->row_bytes >= 4096;
            for (size_t r = off; r < off + len; ++r) {
                if (nt) stream_copy(job->dst + r * job->dpitch, job->src + r * job->spitch, job->row_bytes);
                else memcpy(job->dst + r * job->dpitch, job->src + r * job->spitch, job->row_bytes);
            }
        } else if (use_stream_copy()) {
            stream_copy(job->dst + off, job->src + off, len);
        } else {
            memcpy(job->dst + off, job->src + off, len);
        }
        std::lock_guard<std::mutex> lk(m_);
        if (++job->finished == job->n_pieces) job->done_cv.notify_all();  // `job` must not be touched after this
    }
    // the submitter's share: pieces of its own job only (the job is alive: it lives on this thread's stack)
    void run_pieces(Job* job) {
        while (true) {
            size_t idx;
            {
                std::lock_guard<std::mutex> lk(m_);
                if (job->next >= job->n_pieces) return;
                idx = claim(job);
            }
            copy_piece(job, idx);
        }
    }
    void worker() {
        while (true) {
            Job* job;
            size_t idx;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return !jobs_.empty(); });
                job = jobs_.front();  // still has at least one piece, or it would have left the queue
                idx = claim(job);
            }
            copy_piece(job, idx);
        }
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::vector<Job*> jobs_;
    int workers_ = 0;
};

}  // namespace

void host_copy(void* dst, const void* src, size_t bytes) { CopyPool::get().copy(dst, src, bytes); }
void host_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t row_bytes, size_t rows) {
    CopyPool::get().copy_2d(dst, dpitch, src, spitch, row_bytes, rows);
}

static void fast_memcpy(void* dst, const void* src, size_t bytes) { host_copy(dst, src, bytes); }

sdb_status h2d(Context* ctx, void* d_dst, const void* h_src, size_t bytes) {
    if (bytes == 0) return SDB_STATUS_SUCCESS;
    if (is_pinned(h_src)) {
        SDB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return SDB_STATUS_SUCCESS;
    }
    SDB_TRY(ensure_staging(ctx));
    size_t off = 0;
    while (off < bytes) {
        int k = ctx->next_chunk;
        ctx->next_chunk = (k + 1) % Context::kChunks;
        SDB_CUDA(cudaEventSynchronize(ctx->chunk_free[k]));
        size_t len = bytes - off < Context::kChunkBytes ? bytes - off : Context::kChunkBytes;
        fast_memcpy(ctx->chunk[k], (const char*)h_src + off, len);
        SDB_CUDA(cudaMemcpyAsync((char*)d_dst + off, ctx->chunk[k], len, cudaMemcpyHostToDevice,
                                 ctx->stream));
        SDB_CUDA(cudaEventRecord(ctx->chunk_free[k], ctx->stream));
        off += len;
    }
    return SDB_STATUS_SUCCESS;
}

sdb_status d2h(Context* ctx, void* h_dst, const void* d_src, size_t bytes) {
    if (bytes == 0) return SDB_STATUS_SUCCESS;
    if (is_pinned(h_dst)) {
        SDB_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        SDB_CUDA(cudaStreamSynchronize(ctx->stream));
        return SDB_STATUS_SUCCESS;
    }
    SDB_TRY(ensure_staging(ctx));
    // software pipeline: DMA chunk i+1 while the host copies chunk i out of the ring
    struct Pending { int k; size_t off, len; };
    Pending pend[Context::kChunks];
    int head = 0, tail = 0, inflight = 0;
    size_t off = 0;
    while (off < bytes || inflight > 0) {
        while (off < bytes && inflight < Context::kChunks - 1) {
            int k = ctx->next_chunk;
            ctx->next_chunk = (k + 1) % Context::kChunks;
            SDB_CUDA(cudaEventSynchronize(ctx->chunk_free[k]));
            size_t len = bytes - off < Context::kChunkBytes ? bytes - off : Context::kChunkBytes;
            SDB_CUDA(cudaMemcpyAsync(ctx->chunk[k], (const char*)d_src + off, len,
                                     cudaMemcpyDeviceToHost, ctx->stream));
            SDB_CUDA(cudaEventRecord(ctx->chunk_free[k], ctx->stream));
            pend[tail] = {k, off, len};
            tail = (tail + 1) % Context::kChunks;
            ++inflight;
            off += len;
        }
        Pending p = pend[head];
        head = (head + 1) % Context::kChunks;
        --inflight;
        SDB_CUDA(cudaEventSynchronize(ctx->chunk_free[p.k]));
        fast_memcpy((char*)h_dst + p.off, ctx->chunk[p.k], p.len);
    }
    return SDB_STATUS_SUCCESS;
}

sdb_status h2d_2d(Context* ctx, void* d_dst, size_t d_pitch, const void* h_src, size_t h_pitch,
                  size_t row_bytes, size_t rows) {
    if (rows == 0 || row_bytes == 0) return SDB_STATUS_SUCCESS;
    if (d_pitch == row_bytes && h_pitch == row_bytes) return h2d(ctx, d_dst, h_src, row_bytes * rows);
    if (is_pinned(h_src) || row_bytes > Context::kChunkBytes) {
        SDB_CUDA(cudaMemcpy2DAsync(d_dst, d_pitch, h_src, h_pitch, row_bytes, rows,
                                   cudaMemcpyHostToDevice, ctx->stream));
        if (!is_pinned(h_src)) SDB_CUDA(cudaStreamSynchronize(ctx->stream));
        return SDB_STATUS_SUCCESS;
    }
    // strided pageable panel: whole rows are packed into the page-locked chunks by the copy pool, the DMA engine
    // spreads them to the device pitch
    SDB_TRY(ensure_staging(ctx));
    const size_t per = Context::kChunkBytes / row_bytes;
    for (size_t r = 0; r < rows; r += per) {
        const size_t nr = std::min(per, rows - r);
        int k = ctx->next_chunk;
        ctx->next_chunk = (k + 1) % Context::kChunks;
        SDB_CUDA(cudaEventSynchronize(ctx->chunk_free[k]));
        host_copy_2d(ctx->chunk[k], row_bytes, (const char*)h_src + r * h_pitch, h_pitch, row_bytes, nr);
        SDB_CUDA(cudaMemcpy2DAsync((char*)d_dst + r * d_pitch, d_pitch, ctx->chunk[k], row_bytes, row_bytes, nr,
                                   cudaMemcpyHostToDevice, ctx->stream));
        SDB_CUDA(cudaEventRecord(ctx->chunk_free[k], ctx->stream));
    }
    return SDB_STATUS_SUCCESS;
}

sdb_status d2h_2d(Context* ctx, void* h_dst, size_t h_pitch, const void* d_src, size_t d_pitch,
                  size_t row_bytes, size_t rows) {
    if (rows == 0 || row_bytes == 0) return SDB_STATUS_SUCCESS;
    if (d_pitch == row_bytes && h_pitch == row_bytes) return d2h(ctx, h_dst, d_src, row_bytes * rows);
    if (is_pinned(h_dst) || row_bytes > Context::kChunkBytes) {
        SDB_CUDA(cudaMemcpy2DAsync(h_dst, h_pitch, d_src, d_pitch, row_bytes, rows,
                                   cudaMemcpyDeviceToHost, ctx->stream));
        SDB_CUDA(cudaStreamSynchronize(ctx->stream));
        return SDB_STATUS_SUCCESS;
    }
    // strided pageable destination: the DMA engine packs whole rows into the page-locked chunks, the copy pool
    // spreads each chunk to the caller's pitch while the next one is in flight
    SDB_TRY(ensure_staging(ctx));
    const size_t per = Context::kChunkBytes / row_bytes;
    struct Pending { int k; size_t r, nr; };
    Pending pend[Context::kChunks];
    int head = 0, tail = 0, inflight = 0;
    size_t r = 0;
    while (r < rows || inflight > 0) {
        while (r < rows && inflight < Context::kChunks - 1) {
            const size_t nr = std::min(per, rows - r);
            int k = ctx->next_chunk;
            ctx->next_chunk = (k + 1) % Context::kChunks;
            SDB_CUDA(cudaEventSynchronize(ctx->chunk_free[k]));
            SDB_CUDA(cudaMemcpy2DAsync(ctx->chunk[k], row_bytes, (const char*)d_src + r * d_pitch, d_pitch, row_bytes, nr,
                                       cudaMemcpyDeviceToHost, ctx->stream));
            SDB_CUDA(cudaEventRecord(ctx->chunk_free[k], ctx->stream));
            pend[tail] = {k, r, nr};
            tail = (tail + 1) % Context::kChunks;
            ++inflight;
            r += nr;
        }
        Pending p = pend[head];
        head = (head + 1) % Context::kChunks;
        --inflight;
        SDB_CUDA(cudaEventSynchronize(ctx->chunk_free[p.k]));
        host_copy_2d((char*)h_dst + p.r * h_pitch, h_pitch, ctx->chunk[p.k], row_bytes, row_bytes, p.nr);
    }
    return SDB_STATUS_SUCCESS;
}

// ---------------------------------------------------------------- timers
sdb_status PhaseTimer::init(cudaStream_t stream) {
    s = stream;
    for (auto& e : ev) SDB_CUDA(cudaEventCreate(&e));
    return SDB_STATUS_SUCCESS;
}
sdb_status PhaseTimer::mark(int i) {
    SDB_CUDA(cudaEventRecord(ev[i], s));
    return SDB_STATUS_SUCCESS;
}
void PhaseTimer::finish(Context* ctx) {
    if (!ev[3]) return;
    if (cudaEventSynchronize(ev[3]) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    for (int i = 0; i < 3; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) != cudaSuccess) {
            cudaGetLastError();
            ms = 0.f;
        }
        ctx->last_ms[i] = ms;
    }
}
PhaseTimer::~PhaseTimer() {
    for (auto& e : ev)
        if (e) cudaEventDestroy(e);
}

}  // namespace sdb

// =============================================================== ABI: misc
using namespace sdb;

extern "C" {

int sdb_last_error(char* buf, int len) {
    if (buf && len > 0) {
        strncpy(buf, t_err, size_t(len) - 1);
        buf[len - 1] = '\0';
    }
    return int(strlen(t_err));
}

sdb_status sdb_set_option(const char* name, int value) {
    SDB_REQUIRE(name != nullptr, SDB_STATUS_INVALID_VALUE, "sdb_set_option: null name");
    for (auto& s : g_options)
        if (strcmp(s.name, name) == 0) {
            s.value.store(value, std::memory_order_relaxed);
            s.set.store(true, std::memory_order_release);
            return SDB_STATUS_SUCCESS;
        }
    set_error("sdb_set_option: unknown option '%s'", name);
    return SDB_STATUS_INVALID_VALUE;
}

int64_t sdb_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

sdb_status sdb_last_spmm_kernel(char* buf, int len) {
    SDB_REQUIRE(buf != nullptr && len > 0, SDB_STATUS_INVALID_VALUE, "sdb_last_spmm_kernel: null output");
    snprintf(buf, size_t(len), "%s", sdb::t_spmm_kernel);
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_last_timing(double ms[3]) {
    SDB_REQUIRE(ms != nullptr, SDB_STATUS_INVALID_VALUE, "sdb_last_timing: null output");
    Context* ctx;
    SDB_TRY(get_context(&ctx));
    for (int i = 0; i < 3; ++i) ms[i] = ctx->last_ms[i];
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_device_count(int* n) {
    SDB_REQUIRE(n != nullptr, SDB_STATUS_INVALID_VALUE, "sdb_device_count: null output");
    *n = 0;
    SDB_CUDA(cudaGetDeviceCount(n));
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_set_device(int device) {
    SDB_CUDA(cudaSetDevice(device));
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_get_device(int* device) {
    SDB_REQUIRE(device != nullptr, SDB_STATUS_INVALID_VALUE, "sdb_get_device: null output");
    SDB_CUDA(cudaGetDevice(device));
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_version_string(char* buf, int len) {
    SDB_REQUIRE(buf != nullptr && len > 0, SDB_STATUS_INVALID_VALUE, "sdb_version_string: bad buffer");
    int rt = 0, drv = 0, ndev = 0, dev = 0;
    cudaRuntimeGetVersion(&rt);
    cudaDriverGetVersion(&drv);
    char gpu[160] = "no CUDA device visible";
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 && cudaGetDevice(&dev) == cudaSuccess) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) == cudaSuccess)
            snprintf(gpu, sizeof(gpu), "%s sm_%d%d, %d SMs, %.0f GiB, L2 %d MiB (device %d of %d)", p.name,
                     p.major, p.minor, p.multiProcessorCount, double(p.totalGlobalMem) / double(1 << 30),
                     p.l2CacheSize >> 20, dev, ndev);
    } else {
        cudaGetLastError();
    }
    snprintf(buf, size_t(len), "sparse_dot_b200 libsdb200 0.1.0 (sm_100a) | CUDA runtime %d.%d, driver %d.%d | %s",
             rt / 1000, (rt % 1000) / 10, drv / 1000, (drv % 1000) / 10, gpu);
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_host_alloc(void** p, size_t bytes) {
    SDB_REQUIRE(p != nullptr, SDB_STATUS_INVALID_VALUE, "sdb_host_alloc: null output");
    *p = nullptr;
    SDB_CUDA(cudaHostAlloc(p, bytes ? bytes : 16, cudaHostAllocDefault));
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_host_free(void* p) {
    if (p) SDB_CUDA(cudaFreeHost(p));
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_dev_alloc(void** p, size_t bytes) {
    SDB_REQUIRE(p != nullptr, SDB_STATUS_INVALID_VALUE, "sdb_dev_alloc: null output");
    *p = nullptr;
    SDB_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_dev_free(void* p) {
    if (p) SDB_CUDA(cudaFree(p));
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_memcpy(void* dst, const void* src, size_t bytes, int kind) {
    SDB_REQUIRE(kind >= 1 && kind <= 3, SDB_STATUS_INVALID_VALUE, "sdb_memcpy: kind must be 1, 2 or 3");
    if (bytes == 0) return SDB_STATUS_SUCCESS;
    SDB_REQUIRE(dst && src, SDB_STATUS_INVALID_VALUE, "sdb_memcpy: null pointer");
    Context* ctx;
    SDB_TRY(get_context(&ctx));
    if (kind == 1) {
        SDB_TRY(h2d(ctx, dst, src, bytes));
    } else if (kind == 2) {
        SDB_TRY(d2h(ctx, dst, src, bytes));
    } else {
        SDB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    SDB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_memcpy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t row_bytes, size_t rows,
                         int kind) {
    SDB_REQUIRE(kind == 1 || kind == 2, SDB_STATUS_INVALID_VALUE, "sdb_memcpy_2d: kind must be 1 (to device) or 2 (to host)");
    if (rows == 0 || row_bytes == 0) return SDB_STATUS_SUCCESS;
    SDB_REQUIRE(dst && src, SDB_STATUS_INVALID_VALUE, "sdb_memcpy_2d: null pointer");
    SDB_REQUIRE(dpitch >= row_bytes && spitch >= row_bytes, SDB_STATUS_INVALID_VALUE,
                "sdb_memcpy_2d: a pitch is smaller than the row");
    Context* ctx;
    SDB_TRY(get_context(&ctx));
    if (kind == 1) SDB_TRY(h2d_2d(ctx, dst, dpitch, src, spitch, row_bytes, rows));
    else SDB_TRY(d2h_2d(ctx, dst, dpitch, src, spitch, row_bytes, rows));
    SDB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_device_synchronize(void) {
    SDB_CUDA(cudaDeviceSynchronize());
    return SDB_STATUS_SUCCESS;
}

static_assert(sizeof(cudaIpcMemHandle_t) == SDB_IPC_TOKEN_BYTES, "IPC token size");

sdb_status sdb_ipc_export(const void* d_ptr, char* token) {
    SDB_REQUIRE(d_ptr && token, SDB_STATUS_INVALID_VALUE, "sdb_ipc_export: null argument");
    cudaIpcMemHandle_t h;
    SDB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    memcpy(token, &h, sizeof(h));
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_ipc_open(const char* token, void** d_ptr) {
    SDB_REQUIRE(d_ptr && token, SDB_STATUS_INVALID_VALUE, "sdb_ipc_open: null argument");
    *d_ptr = nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, token, sizeof(h));
    SDB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SDB_STATUS_SUCCESS;
}

sdb_status sdb_ipc_close(void* d_ptr) {
    if (d_ptr) SDB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return SDB_STATUS_SUCCESS;
}

// nnz-balanced contiguous row blocks (SURVEY §8e "Partitioning"): bounds[p] is
// the first row whose cumulative nnz reaches p/parts of the total.
sdb_status sdb_partition_rows(const void* indptr, int index_bits, int64_t rows, int parts,
                              int64_t* bounds) {
    SDB_REQUIRE(indptr && bounds && parts > 0 && rows >= 0, SDB_STATUS_INVALID_VALUE,
                "sdb_partition_rows: bad arguments");
    SDB_REQUIRE(index_bits == 32 || index_bits == 64, SDB_STATUS_INVALID_VALUE,
                "sdb_partition_rows: index_bits must be 32 or 64");
    auto at = [&](int64_t i) -> int64_t {
        return index_bits == 32 ? int64_t(static_cast<const int32_t*>(indptr)[i])
                                : static_cast<const int64_t*>(indptr)[i];
    };
    const int64_t base = at(0), total = at(rows) - base;
    bounds[0] = 0;
    for (int p = 1; p < parts; ++p) {
        if (total == 0) {  // no nonzeros: split rows evenly
            bounds[p] = rows * p / parts;
            continue;
        }
        // smallest row r with indptr[r] - base >= total * p / parts
        const int64_t want = base + (total / parts) * p + (total % parts) * p / parts;
        int64_t lo = bounds[p - 1], hi = rows;
        while (lo < hi) {
            int64_t mid = lo + (hi - lo) / 2;
            if (at(mid) >= want) hi = mid;
            else lo = mid + 1;
        }
        bounds[p] = lo;
    }
    bounds[parts] = rows;
    return SDB_STATUS_SUCCESS;
}

}  // extern "C"
