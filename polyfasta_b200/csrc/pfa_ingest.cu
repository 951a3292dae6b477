// Upload of an alignment: host (or device) text -> site-major bit-planes in HBM.  Replaces what readfasta
// (PolyFastA.py:227-250) leaves in memory for the scans.
//
// Small inputs and inputs already on the device: chunks of columns go through K1 (pfa_encode_kernel) one after another,
// the host copies double-buffered on a copy stream.
//
// Large host inputs ("hybrid ingest"): the PCIe link carries 1 byte per base as text but only 1/4 byte per base once
// packed, and the host cores can pack faster than the link can ship text.  So the column chunks of the alignment are
// handed out from one counter to two lanes that run concurrently:
//   raw lane    (the calling thread): cudaMemcpy2DAsync of the text chunk -> K1                      (1 B/base on the link)
//   packed lane (a driver thread + a pool of host threads): the pool packs the chunk 4 bases per byte into pinned memory
//               (pfa_pack.cpp, AVX-512 / AVX2), one contiguous copy -> pfa_encode_packed_kernel       (0.25 B/base)
// Whichever lane is free takes the next chunk, so the split adapts to the box (cores, memory bandwidth, link).  A chunk
// in which the packer meets anything but A/C/G/T (gaps, N, IUPAC codes, ...) is "dirty" and goes to the raw lane, whose
// kernel builds the validity plane and the exception list; when the first chunks are dirty the packed lane stops.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "pfa_common.cuh"
#include "pfa_host.h"

namespace {

// persistent pool: run(f) executes f on every thread and returns when all are done
class PackPool {
public:
    explicit PackPool(int threads) {
        for (int t = 0; t < threads; ++t)
            th_.emplace_back([this] {
                uint64_t seen = 0;
                for (;;) {
                    std::unique_lock<std::mutex> lk(m_);
                    start_.wait(lk, [&] { return stop_ || gen_ != seen; });
                    if (stop_) return;
                    seen = gen_;
                    auto job = job_;
                    lk.unlock();
                    job();
                    lk.lock();
                    if (--pending_ == 0) done_.notify_one();
                }
            });
    }
    ~PackPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        start_.notify_all();
        for (auto& t : th_) t.join();
    }
    void run(const std::function<void()>& f) {
        std::unique_lock<std::mutex> lk(m_);
        job_ = f;
        pending_ = (int)th_.size();
        ++gen_;
        start_.notify_all();
        done_.wait(lk, [&] { return pending_ == 0; });
    }

private:
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable start_, done_;
    std::function<void()> job_;
    uint64_t gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

constexpr int kSlots = 3;     // packed chunks in flight per packed lane (pinned + device staging each)
constexpr int kRawBufs = 2;   // text chunks in flight on the raw lane (device staging each); a third one only shifts chunks away from the packer (measured)
constexpr int kMaxLanes = 2;  // packed lanes: while one waits at its end-of-chunk barrier the other keeps the cores busy

struct Hybrid {
    pfa_aln* a;
    const uint8_t* text;  // first column of the shard
    const int64_t* row_off = nullptr;  // optional: row r starts at text + row_off[r] (rows of any stride)
    const int32_t *wrap_w = nullptr, *wrap_gap = nullptr;  // optional: rows wrapped over lines (mapped files)
    int64_t col_begin = 0;  // first column of the shard (wrapped rows are addressed by column, not by pointer)
    // columns [c0, c0 + cols) of the shard's row r: a pointer into the source, or tmp after gathering the line pieces
    const uint8_t* span(int64_t r, int64_t c0, int64_t cols, uint8_t* tmp) const {
        if (wrap_w && wrap_w[r] > 0) {
            pfa_gather_wrapped(text - col_begin + row_off[r], wrap_w[r], wrap_gap[r], col_begin + c0, cols, tmp);
            return tmp;
        }
        return text + (row_off ? row_off[r] : r * ld) + c0;
    }
    int64_t ld, n, ns, chunk, nchunks, ldt, ldp, ldv;
    unsigned long long* d_count;
    int64_t cap;
    int* d_inv;
    uint8_t* stage_raw[kRawBufs] = {};
    uint8_t* stage_packed[kSlots * kMaxLanes] = {};
    int lanes = 1;
    PackPool* pools[kMaxLanes] = {};
    uint8_t* pinned = nullptr;
    std::atomic<int64_t> next{0};
    std::atomic<bool> give_up{false};
    std::mutex m;
    std::vector<int64_t> dirty;
    int64_t n_packed = 0, n_gappy = 0, n_raw = 0, bytes_raw = 0, bytes_packed = 0;
    int rc_packed = PFA_OK;
    std::string err_packed;
    int threads = 1;
    bool raw_takes_chunks = true;     // false: the raw lane only gets the dirty chunks (pageable source, or PFA_INGEST_HYBRID=2)
    bool bounce = false;              // the source is pageable: text chunks reach the copy engine through pinned bounce buffers
    uint8_t* bounce_buf[kRawBufs] = {};
    PackPool* pool = nullptr;
    std::atomic<int> lanes_up{0};  // packed lanes that are running
    double t_pack = 0, t_wait_slot = 0, t_wait_raw = 0, t_pool_start = 0, t0 = 0, t_lane_end = 0;  // PFA_INGEST_TRACE
};

int raw_chunk(Hybrid& h, int64_t c, int& issued) {
    pfa_ctx* ctx = h.a->ctx;
    const int b = issued % kRawBufs;
    const double tw = now_ms();
    if (issued >= kRawBufs) PFA_CUDA(ctx, cudaEventSynchronize(ctx->ev_encoded[b]));  // throttle: the buffer is free again
    h.t_wait_raw += now_ms() - tw;
    const int64_t c0 = c * h.chunk, cols = std::min(h.chunk, h.ns - c0);
    if (h.bounce) {
        // pageable source: the pool copies the chunk's rows into a pinned buffer (the copy engine cannot read pageable
        // memory asynchronously, and pinning the whole source in place costs more than the upload itself)
        uint8_t* bb = h.bounce_buf[b];
        std::atomic<int64_t> row_next(0);
        h.pool->run([&] {
            for (;;) {
                const int64_t r0 = row_next.fetch_add(32);
                if (r0 >= h.n) break;
                const int64_t r1 = std::min(h.n, r0 + 32);
                for (int64_t r = r0; r < r1; ++r) {
                    const uint8_t* src = h.span(r, c0, cols, bb + r * h.ldt);
                    if (src != bb + r * h.ldt) memcpy(bb + r * h.ldt, src, (size_t)cols);
                }
            }
        });
        PFA_CUDA(ctx, cudaMemcpyAsync(h.stage_raw[b], bb, (size_t)(h.n * h.ldt), cudaMemcpyHostToDevice, ctx->copy_stream));
    } else {
        PFA_CUDA(ctx, cudaMemcpy2DAsync(h.stage_raw[b], (size_t)h.ldt, h.text + c0, (size_t)h.ld, (size_t)cols, (size_t)h.n,
                                        cudaMemcpyHostToDevice, ctx->copy_stream));
    }
    PFA_CUDA(ctx, cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
    PFA_CUDA(ctx, cudaStreamWaitEvent(ctx->enc_stream, ctx->ev_copied[b], 0));
    int rc = pfa_encode_chunk(h.a, h.stage_raw[b], h.ldt, cols, c0, h.d_count, h.cap, h.d_inv, ctx->enc_stream);
    if (rc) return rc;
    PFA_CUDA(ctx, cudaEventRecord(ctx->ev_encoded[b], ctx->enc_stream));
    ++issued;
    ++h.n_raw;
    h.bytes_raw += h.n * cols;
    return PFA_OK;
}

void packed_lane(Hybrid* hp, int lane_id) {
    Hybrid& h = *hp;
    pfa_ctx* ctx = h.a->ctx;
    bool counted_up = false;
    auto fail = [&](cudaError_t e, const char* what) {
        if (!counted_up) ++h.lanes_up;
        counted_up = true;
        std::lock_guard<std::mutex> lk(h.m);
        h.rc_packed = PFA_ERR_CUDA;
        h.err_packed = std::string(what) + " failed: " + cudaGetErrorString(e);
        h.give_up = true;
    };
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(e, "cudaSetDevice");
    PackPool& pool = *h.pools[lane_id];
    ++h.lanes_up;
    counted_up = true;
    const int slot0 = lane_id * kSlots;
    int64_t my_packed = 0, my_gappy = 0, my_bytes = 0;
    double my_pack = 0, my_wait = 0;
    const size_t code_bytes = (size_t)(h.n * h.ldp), slot_bytes = code_bytes + (size_t)(h.n * h.ldv);
    const bool with_validity = pfa_pack3_fast();  // AVX-512 VBMI: gaps, N and ? are packed too (validity bitmap)
    bool used[kSlots] = {};
    int slot = 0;
    int64_t n_dirty = 0;
    while (!h.give_up.load()) {
        const int64_t c = h.next.fetch_add(1);
        if (c >= h.nchunks) break;
        const int64_t c0 = c * h.chunk, cols = std::min(h.chunk, h.ns - c0);
        const double tw = now_ms();
        if (used[slot] && (e = cudaEventSynchronize(ctx->ev_slot[slot0 + slot])) != cudaSuccess) return fail(e, "cudaEventSynchronize");
        const double tq = now_ms();
        my_wait += tq - tw;
        uint8_t* dst = h.pinned + (slot0 + slot) * slot_bytes;
        std::atomic<int64_t> row_next(0);
        std::atomic<int> flags(0);  // bit 0: the chunk holds '-', 'N' or '?'; bit 1: it holds any other symbol (dirty)
        uint8_t* vdst = dst + code_bytes;
        pool.run([&] {
            int mine = 0;
            std::vector<uint8_t> tmp(h.wrap_w ? (size_t)cols : 0);  // wrapped rows: the line pieces are gathered here first
            for (;;) {
                const int64_t r0 = row_next.fetch_add(16);
                if (r0 >= h.n || (flags.load(std::memory_order_relaxed) & 2)) break;
                const int64_t r1 = std::min(h.n, r0 + 16);
                for (int64_t r = r0; r < r1; ++r) {
                    const uint8_t* src = h.span(r, c0, cols, tmp.data());
                    mine |= with_validity ? pfa_pack3_row(src, cols, dst + r * h.ldp, vdst + r * h.ldv)
                                          : 2 * pfa_pack2_row(src, cols, dst + r * h.ldp);
                }
                if (mine) flags.fetch_or(mine, std::memory_order_relaxed);
            }
        });
        my_pack += now_ms() - tq;
        const int fl = flags.load();
        if (fl & 2) {
            std::lock_guard<std::mutex> lk(h.m);
            h.dirty.push_back(c);
            if (++n_dirty >= 2 && n_dirty > my_packed) h.give_up = true;  // symbols the packer does not know: leave the rest to the raw lane
            continue;
        }
        // the copy goes on the SAME stream as the raw lane's copies: on a stream of its own it is starved by the copy engine
        // until the raw lane has nothing left (measured); in one queue it waits for at most the two raw chunks in flight
        const bool gappy = (fl & 1) != 0;
        const size_t copy_bytes = gappy ? slot_bytes : code_bytes;
        if ((e = cudaMemcpyAsync(h.stage_packed[slot0 + slot], dst, copy_bytes, cudaMemcpyHostToDevice, ctx->copy_stream)) != cudaSuccess)
            return fail(e, "cudaMemcpyAsync");
        if ((e = cudaEventRecord(ctx->ev_slot_copied[slot0 + slot], ctx->copy_stream)) != cudaSuccess) return fail(e, "cudaEventRecord");
        if ((e = cudaStreamWaitEvent(ctx->pack_stream, ctx->ev_slot_copied[slot0 + slot], 0)) != cudaSuccess) return fail(e, "cudaStreamWaitEvent");
        const int rc = pfa_encode_packed_chunk(h.a, h.stage_packed[slot0 + slot], h.ldp, gappy ? h.stage_packed[slot0 + slot] + code_bytes : nullptr, h.ldv,
                                               with_validity ? 1 : 0, cols, c0, h.d_inv, ctx->pack_stream);
        if (rc) {
            std::lock_guard<std::mutex> lk(h.m);
            h.rc_packed = rc;
            h.err_packed = ctx->err;
            h.give_up = true;
            return;
        }
        if ((e = cudaEventRecord(ctx->ev_slot[slot0 + slot], ctx->pack_stream)) != cudaSuccess) return fail(e, "cudaEventRecord");
        used[slot] = true;
        ++my_packed;
        my_gappy += gappy ? 1 : 0;
        my_bytes += (int64_t)copy_bytes;
        slot = (slot + 1) % kSlots;
    }
    std::lock_guard<std::mutex> lk(h.m);
    h.n_packed += my_packed;
    h.n_gappy += my_gappy;
    h.bytes_packed += my_bytes;
    h.t_pack += my_pack;
    h.t_wait_slot += my_wait;
    h.t_lane_end = std::max(h.t_lane_end, now_ms());
}

int host_threads(const pfa_ctx* ctx) {
    if (ctx->host_threads > 0) return ctx->host_threads;
    if (const char* s = getenv("PFA_HOST_THREADS")) {
        const int t = atoi(s);
        if (t > 0) return t;
    }
    return (int)std::max(1u, std::thread::hardware_concurrency());
}

// all chunks of one attempt through the two lanes; the planes are complete when ctx->stream reaches the joins at the end
int upload_hybrid(pfa_aln* a, const uint8_t* text, int64_t ld, const int64_t* row_off, const int32_t* wrap_w, const int32_t* wrap_gap,
                  unsigned long long* d_count, int64_t cap, int* d_inv, int threads, bool raw_takes_chunks, bool bounce) {
    pfa_ctx* ctx = a->ctx;
    Hybrid h;
    h.a = a;
    h.text = text;
    h.ld = ld;
    h.row_off = row_off;
    h.wrap_w = wrap_w;
    h.wrap_gap = wrap_gap;
    h.col_begin = a->col_begin;
    h.n = a->n;
    h.ns = a->ns;
    int64_t chunk_mb = 64;
    if (const char* s = getenv("PFA_INGEST_CHUNK_MB")) chunk_mb = std::max(1, atoi(s));
    h.chunk = std::max<int64_t>(256, ((chunk_mb << 20) / h.n) & ~255ll);
    h.nchunks = (h.ns + h.chunk - 1) / h.chunk;
    h.ldt = pfa_round_up(std::min(h.chunk, h.ns), 256);
    h.ldp = h.ldt / 4;
    h.ldv = h.ldt / 8;
    h.d_count = d_count;
    h.cap = cap;
    h.d_inv = d_inv;
    h.threads = threads;
    h.raw_takes_chunks = raw_takes_chunks && !bounce;
    h.bounce = bounce;
    const size_t slot_bytes = (size_t)(h.n * (h.ldp + h.ldv));
    h.lanes = threads >= 8 ? kMaxLanes : 1;
    if (const char* s = getenv("PFA_PACK_LANES")) h.lanes = std::max(1, std::min(kMaxLanes, atoi(s)));
    const size_t nslots = (size_t)kSlots * h.lanes;
    if (ctx->pack_pinned_bytes < nslots * slot_bytes) {
        if (ctx->pack_pinned) cudaFreeHost(ctx->pack_pinned);
        ctx->pack_pinned = nullptr;
        ctx->pack_pinned_bytes = 0;
        PFA_CUDA(ctx, cudaHostAlloc(&ctx->pack_pinned, nslots * slot_bytes, cudaHostAllocDefault));
        ctx->pack_pinned_bytes = nslots * slot_bytes;
    }
    h.pinned = static_cast<uint8_t*>(ctx->pack_pinned);
    auto release = [&]() {
        for (auto& p : h.stage_raw) pfa_dfree(ctx, p);
        for (auto& p : h.stage_packed) pfa_dfree(ctx, p);
    };
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < kRawBufs && e == cudaSuccess; ++i) e = pfa_dmalloc(ctx, &h.stage_raw[i], (size_t)(h.n * h.ldt));
    for (size_t i = 0; i < nslots && e == cudaSuccess; ++i) e = pfa_dmalloc(ctx, &h.stage_packed[i], slot_bytes);
    // the lanes' streams may touch the planes / staging buffers only after ctx->stream has allocated and cleared them
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_ready, ctx->stream);
    for (cudaStream_t st : {ctx->copy_stream, ctx->enc_stream, ctx->pack_stream})
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ctx->ev_ready, 0);
    if (e != cudaSuccess) {
        release();
        return pfa_fail(ctx, PFA_ERR_CUDA, "hybrid ingest setup failed: %s", cudaGetErrorString(e));
    }
    h.t0 = now_ms();
    std::vector<std::unique_ptr<PackPool>> pools;
    for (int l = 0; l < h.lanes; ++l) {
        pools.emplace_back(new PackPool(std::max(1, (threads + (h.lanes - 1 - l)) / h.lanes)));
        h.pools[l] = pools.back().get();
    }
    h.pool = h.pools[0];  // after the lanes have finished: bounce copies of the raw lane
    h.t_pool_start = now_ms() - h.t0;
    std::vector<std::thread> lane;
    for (int l = 0; l < h.lanes; ++l) lane.emplace_back(packed_lane, &h, l);
    int rc = PFA_OK, issued = 0;
    while (h.lanes_up.load() < h.lanes) std::this_thread::yield();  // ~0.2 ms: all lanes start taking chunks together
    while (h.raw_takes_chunks) {
        const int64_t c = h.next.fetch_add(1);
        if (c >= h.nchunks) break;
        if ((rc = raw_chunk(h, c, issued)) != PFA_OK) break;
    }
    if (rc) h.give_up = true;
    for (auto& t : lane) t.join();
    if (!rc && h.rc_packed) rc = pfa_fail(ctx, h.rc_packed, "packed lane: %s", h.err_packed.c_str());
    if (!rc && h.bounce && (h.give_up.load() || !h.dirty.empty())) {
        const size_t need = kRawBufs * (size_t)(h.n * h.ldt);
        if (ctx->raw_pinned_bytes < need) {
            if (ctx->raw_pinned) cudaFreeHost(ctx->raw_pinned);
            ctx->raw_pinned = nullptr;
            ctx->raw_pinned_bytes = 0;
            if ((e = cudaHostAlloc(&ctx->raw_pinned, need, cudaHostAllocDefault)) != cudaSuccess)
                rc = pfa_fail(ctx, PFA_ERR_CUDA, "pinned bounce buffers: %s", cudaGetErrorString(e));
            else
                ctx->raw_pinned_bytes = need;
        }
        for (int i = 0; i < kRawBufs; ++i) h.bounce_buf[i] = static_cast<uint8_t*>(ctx->raw_pinned) + (size_t)i * (size_t)(h.n * h.ldt);
    }
    if (!rc && h.give_up.load())  // the packed lane stopped early: chunks it never took are still on the counter
        for (;;) {
            const int64_t c = h.next.fetch_add(1);
            if (c >= h.nchunks) break;
            if ((rc = raw_chunk(h, c, issued)) != PFA_OK) break;
        }
    for (size_t i = 0; !rc && i < h.dirty.size(); ++i) rc = raw_chunk(h, h.dirty[i], issued);
    // join: ctx->stream continues after both lanes
    if (!rc) {
        e = cudaEventRecord(ctx->ev_join[0], ctx->enc_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join[0], 0);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_join[1], ctx->pack_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join[1], 0);
        if (e != cudaSuccess) rc = pfa_fail(ctx, PFA_ERR_CUDA, "hybrid ingest join failed: %s", cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(ctx->enc_stream);
        cudaStreamSynchronize(ctx->pack_stream);
        cudaStreamSynchronize(ctx->copy_stream);
    }
    release();  // stream-ordered: freed on ctx->stream after the joins
    if (getenv("PFA_INGEST_TRACE"))
        fprintf(stderr, "[pfa ingest] %lld chunks of %lld cols: raw %lld packed %lld dirty %zu | host issue %.1f ms; raw lane blocked %.1f ms; "
                        "packed lane: pool start %.2f ms, packing %.1f ms, waiting for a slot %.1f ms, done at %.1f ms\n",
                (long long)h.nchunks, (long long)h.chunk, (long long)h.n_raw, (long long)h.n_packed, h.dirty.size(), now_ms() - h.t0,
                h.t_wait_raw, h.t_pool_start, h.t_pack, h.t_wait_slot, h.t_lane_end - h.t0);
    ctx->ingest_stats[0] = h.n_raw;
    ctx->ingest_stats[1] = h.n_packed;
    ctx->ingest_stats[2] = (int64_t)h.dirty.size();
    ctx->ingest_stats[3] = threads;
    ctx->ingest_stats[4] = h.bytes_raw;
    ctx->ingest_stats[5] = h.bytes_packed;
    ctx->ingest_stats[6] = h.n_gappy;
    return rc;
}

}  // namespace

int pfa_aln_from_text(pfa_ctx* ctx, const uint8_t* text, bool dev, int64_t n, int64_t L, int64_t ld, int64_t col_begin,
                      int64_t col_end, pfa_aln** out, const int64_t* row_off, const int32_t* wrap_w, const int32_t* wrap_gap) {
    if (!ctx || !out || (dev && row_off) || (wrap_w && (!row_off || !wrap_gap))) return PFA_ERR_ARG;
    *out = nullptr;
    if (n < 0 || L < 0 || col_begin < 0 || col_end < col_begin || col_end > L || (n > 0 && L > 0 && (!text || (ld < L && !row_off))))
        return pfa_fail(ctx, PFA_ERR_ARG, "bad alignment shape n=%lld L=%lld ld=%lld cols=[%lld,%lld)", (long long)n,
                        (long long)L, (long long)ld, (long long)col_begin, (long long)col_end);
    if (n >= (1ll << 24)) return pfa_fail(ctx, PFA_ERR_ARG, "more than 2^24-1 sequences are not supported");
    if (col_end - col_begin >= (1ll << 32)) return pfa_fail(ctx, PFA_ERR_ARG, "a shard holds at most 2^32-1 sites");
    PFA_CUDA(ctx, cudaSetDevice(ctx->device));
    pfa_aln* a = nullptr;
    int rc = pfa_aln_alloc(ctx, n, L, col_begin, col_end, &a);
    if (rc) return rc;
    const int64_t ns = a->ns;
    unsigned long long* d_count = nullptr;
    int* d_inv = nullptr;
    uint8_t* stage[2] = {nullptr, nullptr};
    cudaStream_t cs = ctx->copy_stream;
    cudaEvent_t* ev_copied = ctx->ev_copied;
    cudaEvent_t* ev_encoded = ctx->ev_encoded;
    bool registered = false;
    int64_t cap = 0;
    auto cleanup = [&]() {
        pfa_dfree(ctx, d_count);
        pfa_dfree(ctx, d_inv);
        pfa_dfree(ctx, stage[0]);
        pfa_dfree(ctx, stage[1]);
        if (registered) cudaHostUnregister(const_cast<uint8_t*>(text));
    };
#define UP(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            cleanup();                                                                                    \
            pfa_aln_free(a);                                                                              \
            return pfa_fail(ctx, PFA_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));          \
        }                                                                                                 \
    } while (0)
    if (ns > 0 && n > 0) {
        UP(pfa_dmalloc(ctx, &d_count, sizeof(unsigned long long)));
        UP(pfa_dmalloc(ctx, &d_inv, sizeof(int)));
        // hybrid ingest for large host inputs (PFA_INGEST_HYBRID overrides the size test: 0 plain, 1 hybrid, 2 hybrid with the
        // raw lane restricted to dirty chunks)
        const int threads = host_threads(ctx);
        bool hybrid = !dev && ((threads > 1 && n * ns >= (64ll << 20)) || row_off);
        bool raw_takes_chunks = true;
        if (const char* s = getenv("PFA_INGEST_HYBRID")) {
            hybrid = !dev && (atoi(s) != 0 || row_off);  // rows of any stride exist only on the hybrid path
            raw_takes_chunks = atoi(s) != 2;
        }
        bool is_pinned = true;
        if (!dev) {
            cudaPointerAttributes attr;
            is_pinned = cudaPointerGetAttributes(&attr, text) == cudaSuccess && attr.type == cudaMemoryTypeHost;
            cudaGetLastError();
        }
        // columns per chunk of the plain path: ~256 MB of text, a multiple of 256 columns
        int64_t chunk = ((256ll << 20) / n) & ~255ll;
        if (chunk < 256) chunk = 256;
        if (chunk > ns) chunk = ns;
        const int64_t ldt = pfa_round_up(chunk, 256);
        if (!dev && !hybrid) {
            const int nbuf = chunk < ns ? 2 : 1;
            for (int i = 0; i < nbuf; ++i) UP(pfa_dmalloc(ctx, &stage[i], (size_t)(n * ldt)));
            // the copy stream may touch the staging buffers only after their (stream-ordered) allocation
            UP(cudaEventRecord(ctx->ev_ready, ctx->stream));
            UP(cudaStreamWaitEvent(cs, ctx->ev_ready, 0));
            // pin mid-sized pageable inputs in place so that the 2-D copies run asynchronously (large ones take the hybrid
            // path, whose packer reads pageable memory directly)
            const size_t span = (size_t)((n - 1) * ld + L);
            if (!is_pinned && span >= (32u << 20)) {
                registered = cudaHostRegister(const_cast<uint8_t*>(text), span, cudaHostRegisterReadOnly) == cudaSuccess ||
                             cudaHostRegister(const_cast<uint8_t*>(text), span, cudaHostRegisterDefault) == cudaSuccess;
                cudaGetLastError();
            }
        }
        cap = std::max<int64_t>(1 << 16, n * ns / 256);
        for (int attempt = 0; attempt < 2; ++attempt) {
            pfa_dfree(ctx, a->exc_keys);
            a->exc_keys = nullptr;
            UP(pfa_dmalloc(ctx, &a->exc_keys, sizeof(unsigned long long) * (size_t)cap));
            UP(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), ctx->stream));
            UP(cudaMemsetAsync(d_inv, 0, sizeof(int), ctx->stream));
            if (hybrid) {
                rc = upload_hybrid(a, text + col_begin, ld, row_off, wrap_w, wrap_gap, d_count, cap, d_inv, std::max(1, threads), raw_takes_chunks,
                                   !is_pinned || row_off != nullptr);
                if (rc) {
                    cleanup();
                    pfa_aln_free(a);
                    return rc;
                }
            } else {
                int64_t ci = 0;
                for (int64_t c = 0; c < ns; c += chunk, ++ci) {
                    const int64_t cols = std::min(chunk, ns - c);
                    if (dev) {
                        rc = pfa_encode_chunk(a, text + col_begin + c, ld, cols, c, d_count, cap, d_inv);
                    } else {
                        const int b = (int)(ci & 1);
                        if (ci >= 2) UP(cudaStreamWaitEvent(cs, ev_encoded[b], 0));
                        UP(cudaMemcpy2DAsync(stage[b], (size_t)ldt, text + col_begin + c, (size_t)ld, (size_t)cols, (size_t)n,
                                             cudaMemcpyHostToDevice, cs));
                        UP(cudaEventRecord(ev_copied[b], cs));
                        UP(cudaStreamWaitEvent(ctx->stream, ev_copied[b], 0));
                        rc = pfa_encode_chunk(a, stage[b], ldt, cols, c, d_count, cap, d_inv);
                        if (!rc) UP(cudaEventRecord(ev_encoded[b], ctx->stream));
                    }
                    if (rc) {
                        cleanup();
                        pfa_aln_free(a);
                        return rc;
                    }
                }
                ctx->ingest_stats[0] = ci;
                ctx->ingest_stats[1] = ctx->ingest_stats[2] = ctx->ingest_stats[5] = ctx->ingest_stats[6] = 0;
                ctx->ingest_stats[3] = 1;
                ctx->ingest_stats[4] = dev ? 0 : n * ns;
            }
            unsigned long long count = 0;
            UP(cudaMemcpyAsync(&count, d_count, sizeof(count), cudaMemcpyDeviceToHost, ctx->stream));
            UP(cudaMemcpyAsync(&a->has_invalid, d_inv, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            UP(cudaStreamSynchronize(ctx->stream));
            if ((int64_t)count <= cap) {
                if (count >= (1ull << 31)) {
                    cleanup();
                    pfa_aln_free(a);
                    return pfa_fail(ctx, PFA_ERR_ARG, "more than 2^31 non-ACGT/-/N/? symbols in one shard");
                }
                rc = pfa_finish_exceptions(a, (int64_t)count);
                if (rc) {
                    cleanup();
                    pfa_aln_free(a);
                    return rc;
                }
                break;
            }
            cap = (int64_t)count;  // the exception list overflowed: encode once more with the exact size
        }
    }
#undef UP
    cleanup();
    rc = pfa_aln_default_pop(a);
    if (rc) {
        pfa_aln_free(a);
        return rc;
    }
    *out = a;
    return PFA_OK;
}

extern "C" {

int pfa_ctx_set_host_threads(pfa_ctx* ctx, int threads) {
    if (!ctx || threads < 0) return PFA_ERR_ARG;
    ctx->host_threads = threads;
    return PFA_OK;
}

int pfa_ctx_ingest_stats(const pfa_ctx* ctx, int64_t out[8]) {
    if (!ctx || !out) return PFA_ERR_ARG;
    for (int i = 0; i < 8; ++i) out[i] = ctx->ingest_stats[i];
    return PFA_OK;
}

}  // extern "C"
