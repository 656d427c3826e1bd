// staging.cu — pageable host memory <-> device at PCIe rate.
//
// The reference receives polars / Arrow buffers: ordinary (pageable) allocations.  A plain cudaMemcpyAsync from
// pageable memory is staged by the driver through one internal bounce buffer on ONE thread (~10-25 GB/s); the
// hot path's end-to-end rate is the host link, so the engine does the staging itself:
//   ring of page-locked slots (cudaMallocHost) <- T host threads memcpy the caller's columns, slot by slot
//   each filled slot -> cudaMemcpyAsync on a dedicated copy stream (DMA at full PCIe rate) -> event per slot
// so the host-side memcpy of chunk i+1.. overlaps the DMA of chunk i, and the compute stream only waits on the
// final event.  The reverse direction (large prediction / coefficient outputs into pageable numpy buffers) runs
// the same ring backwards.  Buffers that are already page-locked (b200ols_host_alloc, cudaHostRegister'd,
// torch pinned tensors) are detected with cudaPointerGetAttributes and copied directly.
#include <emmintrin.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>

#include "engine_ctx.h"

namespace b200 {

// pageable -> pinned slot with non-temporal stores: the slot is read next by the DMA engine, not by this core, so the
// destination lines need neither a read-for-ownership nor a place in the cache (one third less host DRAM traffic
// than memcpy, which matters because the copy threads share the memory controllers with the DMA reads)
static void copy_streaming(char *dst, const char *src, size_t n) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) != 0) {
        std::memcpy(dst, src, n);
        return;
    }
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 32));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 48), d);
    }
    _mm_sfence();
    if (i < n) std::memcpy(dst + i, src + i, n - i);
}

struct Stager {
    int device = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t done_ev = nullptr, gate_ev = nullptr;
    char *ring = nullptr;
    size_t slot_bytes = 0;
    int nslots = 0;
    std::vector<cudaEvent_t> ev;
    std::vector<std::atomic<int64_t>> *seq = nullptr;  // per slot: generation that may use it next
    int64_t chunk_base = 0;                            // global index of the first chunk of the current job
    struct Chunk { char *host; char *dev; size_t bytes; };
    std::vector<Chunk> chunks;
    bool d2h = false;
    bool streaming = true;  // non-temporal stores into the ring (B200OLS_STAGE_NT=0: plain memcpy)
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    size_t finished = 0;
    uint64_t job = 0;
    bool stop = false;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::vector<std::thread> workers;

    void work() {  // runs chunks of the current job until none is left
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= chunks.size()) return;
            const Chunk &ch = chunks[i];
            const int64_t gid = chunk_base + static_cast<int64_t>(i);
            const int slot = static_cast<int>(gid % nslots);
            const int64_t gen = gid / nslots;
            while ((*seq)[slot].load(std::memory_order_acquire) != gen) std::this_thread::yield();
            char *buf = ring + static_cast<size_t>(slot) * slot_bytes;
            cudaError_t e = cudaSuccess;
            if (gen > 0) e = cudaEventSynchronize(ev[slot]);  // the DMA that last used this slot has drained
            if (!d2h) {
                if (streaming) copy_streaming(buf, ch.host, ch.bytes); else std::memcpy(buf, ch.host, ch.bytes);
                if (e == cudaSuccess) e = cudaMemcpyAsync(ch.dev, buf, ch.bytes, cudaMemcpyHostToDevice, copy_stream);
                if (e == cudaSuccess) e = cudaEventRecord(ev[slot], copy_stream);
            } else {
                if (e == cudaSuccess) e = cudaMemcpyAsync(buf, ch.dev, ch.bytes, cudaMemcpyDeviceToHost, copy_stream);
                if (e == cudaSuccess) e = cudaEventRecord(ev[slot], copy_stream);
                if (e == cudaSuccess) e = cudaEventSynchronize(ev[slot]);
                if (e == cudaSuccess) std::memcpy(ch.host, buf, ch.bytes);
            }
            if (e != cudaSuccess) failed.store(static_cast<int>(e));
            (*seq)[slot].store(gen + 1, std::memory_order_release);
        }
    }

    void worker_main() {
        cudaSetDevice(device);
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_job.wait(lk, [&] { return stop || job != seen; });
                if (stop) return;
                seen = job;
            }
            work();
            {
                std::lock_guard<std::mutex> lk(mu);
                if (++finished == workers.size()) cv_done.notify_one();
            }
        }
    }

    // enqueue (H2D) or complete (D2H) every chunk; the calling thread works too
    cudaError_t run(bool to_host) {
        if (chunks.empty()) return cudaSuccess;
        d2h = to_host;
        next.store(0);
        failed.store(0);
        {
            std::lock_guard<std::mutex> lk(mu);
            finished = 0;
            ++job;
        }
        cv_job.notify_all();
        work();
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return finished == workers.size(); });
        }
        chunk_base += static_cast<int64_t>(chunks.size());
        chunks.clear();
        return static_cast<cudaError_t>(failed.load());
    }
};

}  // namespace b200

using b200::Stager;

static int stager_get(b200ols_ctx *c, Stager **out) {
    if (c->stager) {
        *out = c->stager;
        return 0;
    }
    Stager *s = new Stager();
    s->device = c->device;
    int threads = static_cast<int>(std::thread::hardware_concurrency());
    threads = std::max(1, std::min(8, threads / 2));
    if (const char *v = std::getenv("B200OLS_STAGE_THREADS")) threads = std::max(1, std::min(64, std::atoi(v)));
    size_t slot_kb = 4096;
    if (const char *v = std::getenv("B200OLS_STAGE_SLOT_KB")) slot_kb = static_cast<size_t>(std::max(64, std::min(65536, std::atoi(v))));
    s->slot_bytes = slot_kb << 10;
    s->nslots = std::max(2 * threads, 8);
    if (const char *v = std::getenv("B200OLS_STAGE_SLOTS")) s->nslots = std::max(threads + 1, std::min(256, std::atoi(v)));
    if (const char *v = std::getenv("B200OLS_STAGE_NT")) s->streaming = std::atoi(v) != 0;
    cudaError_t e = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->done_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->gate_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void **>(&s->ring), s->slot_bytes * s->nslots);
    s->ev.resize(s->nslots, nullptr);
    for (int i = 0; i < s->nslots && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        const int rc = fail(B200OLS_ERR_CUDA, "staging ring setup failed: %s", cudaGetErrorString(e));
        stager_destroy(s);
        return rc;
    }
    s->seq = new std::vector<std::atomic<int64_t>>(s->nslots);
    for (auto &a : *s->seq) a.store(0);
    for (int t = 0; t + 1 < threads; ++t) s->workers.emplace_back([s] { s->worker_main(); });
    c->stager = s;
    *out = s;
    return 0;
}

void stager_destroy(Stager *s) {
    if (!s) return;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->stop = true;
    }
    s->cv_job.notify_all();
    for (auto &t : s->workers) t.join();
    if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
    for (cudaEvent_t e : s->ev)
        if (e) cudaEventDestroy(e);
    if (s->done_ev) cudaEventDestroy(s->done_ev);
    if (s->gate_ev) cudaEventDestroy(s->gate_ev);
    if (s->ring) cudaFreeHost(s->ring);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    delete s->seq;
    delete s;
}

bool host_ptr_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

static constexpr size_t STAGE_MIN_BYTES = 1u << 20;  // below this the driver's own bounce buffer is as good

static void add_chunks(Stager *s, char *host, char *dev, size_t bytes) {
    for (size_t o = 0; o < bytes; o += s->slot_bytes) s->chunks.push_back({host + o, dev + o, std::min(s->slot_bytes, bytes - o)});
}

// Host -> device copies of this call.  Pinned sources go straight onto the compute stream; pageable ones run through
// the ring on the copy stream and the compute stream is made to wait for the last of them.  Returns once every
// copy is ENQUEUED (the pageable sources have been read completely, so the caller may reuse them).
int stage_h2d(b200ols_ctx *c, const StageSeg *segs, int n) {
    Stager *s = nullptr;
    bool any = false;
    for (int i = 0; i < n; ++i) {
        if (segs[i].bytes == 0) continue;
        char *host = static_cast<char *>(const_cast<void *>(segs[i].host));
        char *dev = static_cast<char *>(segs[i].dev);
        if (segs[i].bytes < STAGE_MIN_BYTES || host_ptr_is_pinned(host)) {
            CU(cudaMemcpyAsync(dev, host, segs[i].bytes, cudaMemcpyHostToDevice, c->stream));
            continue;
        }
        if (!s) TRY(stager_get(c, &s));
        add_chunks(s, host, dev, segs[i].bytes);
        any = true;
    }
    if (!any) return 0;
    // the device destinations may still be read by work queued earlier on the compute stream
    CU(cudaEventRecord(s->gate_ev, c->stream));
    CU(cudaStreamWaitEvent(s->copy_stream, s->gate_ev, 0));
    const cudaError_t e = s->run(false);
    if (e != cudaSuccess) return fail(B200OLS_ERR_CUDA, "staged host->device copy failed: %s", cudaGetErrorString(e));
    CU(cudaEventRecord(s->done_ev, s->copy_stream));
    CU(cudaStreamWaitEvent(c->stream, s->done_ev, 0));
    return 0;
}

// Device -> host copies of this call's results; returns when the data is in host memory (synchronises the stream).
int stage_d2h(b200ols_ctx *c, const StageSeg *segs, int n) {
    Stager *s = nullptr;
    bool any = false;
    for (int i = 0; i < n; ++i) {
        if (segs[i].bytes == 0) continue;
        char *host = static_cast<char *>(const_cast<void *>(segs[i].host));
        char *dev = static_cast<char *>(segs[i].dev);
        if (segs[i].bytes < STAGE_MIN_BYTES || host_ptr_is_pinned(host)) {
            CU(cudaMemcpyAsync(host, dev, segs[i].bytes, cudaMemcpyDeviceToHost, c->stream));
            continue;
        }
        if (!s) TRY(stager_get(c, &s));
        add_chunks(s, host, dev, segs[i].bytes);
        any = true;
    }
    if (any) {
        CU(cudaEventRecord(s->gate_ev, c->stream));
        CU(cudaStreamWaitEvent(s->copy_stream, s->gate_ev, 0));
        const cudaError_t e = s->run(true);
        if (e != cudaSuccess) return fail(B200OLS_ERR_CUDA, "staged device->host copy failed: %s", cudaGetErrorString(e));
    }
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}
