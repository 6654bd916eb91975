// pipeline.cu — sdb_spmm_csr_host: the reference's create -> mm -> destroy triple
// (_sparse_dense.py:34-132: _create_mkl_sparse, mkl_sparse_?_mm, _destroy_mkl_handle)
// as ONE call on host arrays, for the common case CSR x row-major dense.
//
// On a GPU this call is PCIe-bound (configs[1]: 1.43 GB up, 0.51 GB down around a
// ~4.5 ms kernel), so the whole job is to keep both DMA directions busy:
//   upload stream   X, then per row chunk: column indices, values, Y rows (if beta != 0)
//   compute stream  the SpMM kernel of chunk c as soon as its upload has landed
//   download stream Y rows of chunk c as soon as its kernel has finished
// With page-locked host buffers all three overlap (full-duplex PCIe); the time
// is ~ upload bytes / link rate + one chunk's kernel + download.  Pageable
// buffers (ordinary numpy arrays — what the reference's callers pass) run the
// same pipeline through two rings of page-locked staging slots: the calling
// thread copies each upload piece into a slot with the library's copy threads
// (runtime.cu, host_copy) and DMAs it from there, and a helper thread DMAs the
// finished Y rows into its own ring and copies them out, so host copies, both
// DMA directions and the kernels all overlap.
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "common.h"
#include "prims.h"

using namespace sdb;

namespace {

struct EventPool {
    std::vector<cudaEvent_t> ev;
    ~EventPool() {
        for (auto e : ev) cudaEventDestroy(e);
    }
    sdb_status get(cudaEvent_t* out, bool timing = false) {
        cudaEvent_t e;
        SDB_CUDA(cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming));
        ev.push_back(e);
        *out = e;
        return SDB_STATUS_SUCCESS;
    }
};

// Declared AFTER the device buffers of a pipelined call, so it is destroyed BEFORE them: whatever way
// the call leaves (including an early error return), all three streams have drained before the
// buffers they use are handed back to the pool.
struct StreamJoin {
    cudaStream_t a, b, c;
    ~StreamJoin() {
        cudaStreamSynchronize(a);
        cudaStreamSynchronize(b);
        cudaStreamSynchronize(c);
        cudaGetLastError();
    }
};

int64_t index_at(const void* p, int bits, int64_t i) {
    return bits == 32 ? int64_t(static_cast<const int32_t*>(p)[i]) : static_cast<const int64_t*>(p)[i];
}

// staging slot size (SDB_STAGE_SLOT_MB, default 16 MiB): large enough to amortise the hand-off to the copy
// threads, small enough that eight of them pipeline well behind a 48 MiB row chunk (measured on a 16-core host,
// profiles/r2_logs/host_copy_sweep*.log: 16 MiB slots, 1 MiB pieces, non-temporal stores -> 35.7 ms per configs[1]
// call from pageable arrays; 8 MiB 40-45 ms, 4 MiB 43 ms, cached stores 71 ms)
size_t slot_bytes() {
    static const size_t v = [] {
        const char* e = getenv("SDB_STAGE_SLOT_MB");
        const int mb = e ? atoi(e) : 16;
        return size_t(mb >= 1 && mb <= 256 ? mb : 16) << 20;
    }();
    return v;
}

// Host -> HBM on `s`: straight DMA from page-locked memory, else through the ring in slot-sized pieces
// (the host copy of piece i + 1 overlaps the DMA of piece i).
struct Uploader {
    cudaStream_t s;
    PinnedRing* ring;
    sdb_status put(void* d_dst, const void* h_src, size_t bytes, bool pinned) {
        if (bytes == 0) return SDB_STATUS_SUCCESS;
        if (pinned) {
            SDB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, s));
            return SDB_STATUS_SUCCESS;
        }
        for (size_t off = 0; off < bytes; off += ring->slot_bytes) {
            const int k = ring->next;
            ring->next = (k + 1) % PinnedRing::kSlots;
            SDB_CUDA(cudaEventSynchronize(ring->free_ev[k]));
            const size_t len = std::min(ring->slot_bytes, bytes - off);
            host_copy(ring->slot[k], static_cast<const char*>(h_src) + off, len);
            SDB_CUDA(cudaMemcpyAsync(static_cast<char*>(d_dst) + off, ring->slot[k], len, cudaMemcpyHostToDevice, s));
            SDB_CUDA(cudaEventRecord(ring->free_ev[k], s));
        }
        return SDB_STATUS_SUCCESS;
    }
};

// HBM -> pageable host memory on a helper thread: jobs (device range, host range, "kernel done" event)
// are cut into slot-sized pieces; up to kSlots - 1 DMAs are in flight on `s` while the thread copies the
// oldest finished slot out to the caller's array.
class Downloader {
  public:
    struct Job {
        cudaEvent_t ready;
        const char* d_src;
        char* h_dst;
        size_t bytes;
    };
    Downloader(int device, cudaStream_t s, PinnedRing* ring) : device_(device), s_(s), ring_(ring) {}
    ~Downloader() { finish(); }
    void start() { th_ = std::thread([this] { run(); }); }
    void push(const Job& j) {
        {
            std::lock_guard<std::mutex> lk(m_);
            q_.push_back(j);
        }
        cv_.notify_one();
    }
    // no more jobs: wait until everything has been copied out; returns the thread's status
    sdb_status finish() {
        if (th_.joinable()) {
            {
                std::lock_guard<std::mutex> lk(m_);
                closed_ = true;
            }
            cv_.notify_one();
            th_.join();
        }
        if (status_ != SDB_STATUS_SUCCESS) set_error("%s", err_);
        return status_;
    }

  private:
    struct Piece {
        int slot;
        char* h_dst;
        size_t len;
    };
    void fail(cudaError_t e, const char* what) {
        status_ = e == cudaErrorMemoryAllocation ? SDB_STATUS_ALLOC_FAILED : SDB_STATUS_EXECUTION_FAILED;
        snprintf(err_, sizeof(err_), "CUDA error %d (%s) in the download thread: %s", int(e), cudaGetErrorString(e),
                 what);
    }
    void run() {
        cudaError_t e = cudaSetDevice(device_);
        if (e != cudaSuccess) return fail(e, "cudaSetDevice");
        std::deque<Piece> inflight;
        Job cur{};
        size_t cur_off = 0;
        bool have = false;
        while (true) {
            // issue DMAs while slots are free and work is queued
            while (int(inflight.size()) < PinnedRing::kSlots - 1) {
                if (!have) {
                    std::unique_lock<std::mutex> lk(m_);
                    if (q_.empty()) {
                        if (!inflight.empty()) break;  // something to copy out meanwhile
                        cv_.wait(lk, [&] { return !q_.empty() || closed_; });
                        if (q_.empty()) return;  // closed and drained
                    }
                    cur = q_.front();
                    q_.pop_front();
                    cur_off = 0;
                    have = true;
                    lk.unlock();
                    if ((e = cudaStreamWaitEvent(s_, cur.ready, 0)) != cudaSuccess) return fail(e, "cudaStreamWaitEvent");
                }
                const int k = ring_->next;
                ring_->next = (k + 1) % PinnedRing::kSlots;
                const size_t len = std::min(ring_->slot_bytes, cur.bytes - cur_off);
                if ((e = cudaMemcpyAsync(ring_->slot[k], cur.d_src + cur_off, len, cudaMemcpyDeviceToHost, s_)) !=
                    cudaSuccess)
                    return fail(e, "cudaMemcpyAsync");
                if ((e = cudaEventRecord(ring_->free_ev[k], s_)) != cudaSuccess) return fail(e, "cudaEventRecord");
                inflight.push_back({k, cur.h_dst + cur_off, len});
                cur_off += len;
                if (cur_off >= cur.bytes) have = false;
            }
            if (inflight.empty()) continue;
            const Piece p = inflight.front();
            inflight.pop_front();
            if ((e = cudaEventSynchronize(ring_->free_ev[p.slot])) != cudaSuccess) return fail(e, "cudaEventSynchronize");
            host_copy(p.h_dst, ring_->slot[p.slot], p.len);
        }
    }
    int device_;
    cudaStream_t s_;
    PinnedRing* ring_;
    std::thread th_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Job> q_;
    bool closed_ = false;
    sdb_status status_ = SDB_STATUS_SUCCESS;
    char err_[256] = "";
};

}  // namespace

extern "C" sdb_status sdb_spmm_csr_host(int64_t rows, int64_t cols, const void* indptr, const void* indices,
                                        int index_bits, const void* values, int dtype, const double* alpha,
                                        const void* X, int64_t n, int64_t ldx, const double* beta, void* Y,
                                        int64_t ldy) {
    SDB_REQUIRE(rows >= 0 && cols >= 0 && n >= 0, SDB_STATUS_INVALID_VALUE, "spmm_csr_host: negative size");
    SDB_REQUIRE(index_bits == 32 || index_bits == 64, SDB_STATUS_INVALID_VALUE, "spmm_csr_host: bad index_bits");
    SDB_REQUIRE(indptr && alpha && beta && X && Y, SDB_STATUS_INVALID_VALUE, "spmm_csr_host: null argument");
    const size_t es = dtype_size(dtype);
    SDB_REQUIRE(es != 0, SDB_STATUS_NOT_SUPPORTED, "spmm_csr_host: unknown dtype %d", dtype);
    SDB_REQUIRE(ldx >= n && ldy >= n, SDB_STATUS_INVALID_VALUE, "spmm_csr_host: leading dimension smaller than n");
    const int64_t nnz = rows > 0 ? index_at(indptr, index_bits, rows) : 0;
    SDB_REQUIRE(index_at(indptr, index_bits, 0) == 0 && nnz >= 0, SDB_STATUS_INVALID_VALUE,
                "spmm_csr_host: indptr must start at 0");
    const bool beta_zero = beta[0] == 0.0 && beta[1] == 0.0;
    const bool packed = ldx == n && ldy == n;
    const bool pipelined = index_bits == 32 && packed && nnz > 0 && rows > 0 && n > 0 &&
                           size_t(nnz) * (4 + es) + size_t(rows + cols) * size_t(n) * es > (size_t(32) << 20);
    if (!pipelined) {
        // small operands (or 64-bit host indices / padded panels): the plain triple
        sdb_mat* a = nullptr;
        SDB_TRY(sdb_create_csr(&a, rows, cols, indptr, indices, index_bits, values, dtype));
        sdb_status st = sdb_spmm(SDB_OP_NON_TRANSPOSE, alpha, a, SDB_LAYOUT_ROW_MAJOR, X, n, ldx, beta, Y, ldy);
        sdb_destroy(a);
        return st;
    }

    Context* ctx;
    SDB_TRY(get_context(&ctx));
    cudaStream_t s0 = ctx->stream, s_up = ctx->h2d_stream, s_dn = ctx->d2h_stream;
    EventPool pool;
    DevBuf d_ptr32, d_ptr, d_idx, d_val, d_x, d_y;
    SDB_TRY(d_ptr32.alloc(size_t(rows + 1) * 4, s0));
    SDB_TRY(d_ptr.alloc(size_t(rows + 1) * 8, s0));
    SDB_TRY(d_idx.alloc(size_t(nnz) * 4, s0));
    SDB_TRY(d_val.alloc(size_t(nnz) * es, s0));
    SDB_TRY(d_x.alloc(size_t(cols) * size_t(n) * es, s0));
    SDB_TRY(d_y.alloc(size_t(rows) * size_t(n) * es, s0));
    StreamJoin join{s_up, s_dn, s0};
    // page-locked arrays are DMA'd in place, pageable ones go through the staging rings
    const bool pin_x = is_pinned(X), pin_y = is_pinned(Y), pin_i = is_pinned(indices), pin_v = is_pinned(values);
    if (!(pin_x && pin_y && pin_i && pin_v)) SDB_TRY(ensure_ring(&ctx->up_ring, slot_bytes()));
    if (!pin_y) SDB_TRY(ensure_ring(&ctx->dn_ring, slot_bytes()));
    Uploader up{s_up, &ctx->up_ring};
    Downloader down(ctx->device, s_dn, &ctx->dn_ring);  // destroyed (joined) before the streams are drained
    if (!pin_y) down.start();
    cudaEvent_t e_start, e_alloc, e_end, e_last_up;
    SDB_TRY(pool.get(&e_start, true));
    SDB_TRY(pool.get(&e_alloc));
    SDB_TRY(pool.get(&e_end, true));
    SDB_TRY(pool.get(&e_last_up, true));
    SDB_CUDA(cudaEventRecord(e_start, s0));
    // row offsets are tiny: upload + widen on the compute stream up front
    SDB_CUDA(cudaMemcpyAsync(d_ptr32.p, indptr, size_t(rows + 1) * 4, cudaMemcpyHostToDevice, s0));
    SDB_TRY(widen_i32_to_i64(s0, d_ptr32.as<int32_t>(), d_ptr.as<int64_t>(), rows + 1));
    SDB_CUDA(cudaEventRecord(e_alloc, s0));
    SDB_CUDA(cudaStreamWaitEvent(s_up, e_alloc, 0));
    SDB_CUDA(cudaStreamWaitEvent(s_dn, e_alloc, 0));

    // X first: every chunk needs all of it
    SDB_TRY(up.put(d_x.p, X, size_t(cols) * size_t(n) * es, pin_x));

    // row chunks of ~kChunkBytes of upload each
    constexpr size_t kChunkBytes = size_t(48) << 20;
    const int32_t* hp = static_cast<const int32_t*>(indptr);
    const size_t row_bytes = size_t(n) * es * (beta_zero ? 0 : 1);
    int64_t r0 = 0;
    double kernel_ms = 0.0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kernel_events;
    while (r0 < rows) {
        // grow the chunk until its upload reaches the target
        int64_t lo = r0 + 1, hi = rows;
        auto bytes_to = [&](int64_t r1) {
            return size_t(hp[r1] - hp[r0]) * (4 + es) + size_t(r1 - r0) * row_bytes;
        };
        while (lo < hi) {
            const int64_t mid = lo + (hi - lo) / 2;
            if (bytes_to(mid) >= kChunkBytes) hi = mid;
            else lo = mid + 1;
        }
        const int64_t r1 = lo;
        const int64_t p0 = hp[r0], p1 = hp[r1];
        if (p1 > p0) {
            SDB_TRY(up.put(d_idx.as<int32_t>() + p0, static_cast<const int32_t*>(indices) + p0, size_t(p1 - p0) * 4,
                           pin_i));
            SDB_TRY(up.put(static_cast<char*>(d_val.p) + size_t(p0) * es,
                           static_cast<const char*>(values) + size_t(p0) * es, size_t(p1 - p0) * es, pin_v));
        }
        const size_t y_off = size_t(r0) * size_t(n) * es, y_len = size_t(r1 - r0) * size_t(n) * es;
        if (!beta_zero)
            SDB_TRY(up.put(static_cast<char*>(d_y.p) + y_off, static_cast<const char*>(Y) + y_off, y_len, pin_y));
        cudaEvent_t e_up, e_k0, e_k1;
        SDB_TRY(pool.get(&e_up));
        SDB_TRY(pool.get(&e_k0, true));
        SDB_TRY(pool.get(&e_k1, true));
        SDB_CUDA(cudaEventRecord(e_up, s_up));
        SDB_CUDA(cudaStreamWaitEvent(s0, e_up, 0));
        CsrView v;
        v.rows = r1 - r0;
        v.cols = cols;
        v.nnz = p1 - p0;
        v.indptr = d_ptr.as<int64_t>() + r0;  // absolute offsets into the full index/value arrays
        v.indices = d_idx.as<int32_t>();
        v.values = d_val.p;
        void* yp[1] = {static_cast<char*>(d_y.p) + y_off};
        SDB_CUDA(cudaEventRecord(e_k0, s0));
        SDB_TRY(spmm_device(ctx, s0, v, dtype, false, alpha, beta, SDB_LAYOUT_ROW_MAJOR, d_x.p, n, n, yp, 1, 0, 0, n));
        SDB_CUDA(cudaEventRecord(e_k1, s0));
        kernel_events.emplace_back(e_k0, e_k1);
        if (pin_y) {
            SDB_CUDA(cudaStreamWaitEvent(s_dn, e_k1, 0));
            SDB_CUDA(cudaMemcpyAsync(static_cast<char*>(Y) + y_off, static_cast<char*>(d_y.p) + y_off, y_len,
                                     cudaMemcpyDeviceToHost, s_dn));
        } else if (y_len > 0) {
            down.push({e_k1, static_cast<const char*>(d_y.p) + y_off, static_cast<char*>(Y) + y_off, y_len});
        }
        r0 = r1;
    }
    SDB_CUDA(cudaEventRecord(e_last_up, s_up));
    SDB_TRY(down.finish());  // pageable Y: every row has been copied out to the caller's array
    // join: frees below are ordered on s0 after everything else
    cudaEvent_t e_dn;
    SDB_TRY(pool.get(&e_dn));
    SDB_CUDA(cudaEventRecord(e_dn, s_dn));
    SDB_CUDA(cudaStreamWaitEvent(s0, e_dn, 0));
    SDB_CUDA(cudaStreamWaitEvent(s0, e_last_up, 0));
    SDB_CUDA(cudaEventRecord(e_end, s0));
    SDB_CUDA(cudaStreamSynchronize(s0));
    float up_ms = 0.f, total_ms = 0.f;
    cudaEventElapsedTime(&up_ms, e_start, e_last_up);
    cudaEventElapsedTime(&total_ms, e_start, e_end);
    for (auto& ke : kernel_events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ke.first, ke.second) == cudaSuccess) kernel_ms += ms;
    }
    cudaGetLastError();
    // overlapping spans: [0] start -> last upload landed, [1] sum of kernel times, [2] whole call
    ctx->last_ms[0] = up_ms;
    ctx->last_ms[1] = kernel_ms;
    ctx->last_ms[2] = total_ms;
    return SDB_STATUS_SUCCESS;
}
