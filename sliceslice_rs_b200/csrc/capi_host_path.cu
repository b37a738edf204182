// capi_host_path.cu -- search_in(&[u8]) on a HOST slice: chunked upload overlapped with the scan, with a
// memcpy worker pool that stages pageable memory through pinned buffers.
#include "capi_internal.h"

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <thread>

// ---------------------------------------------------------------------------------------------
// host-resident haystack: chunked upload overlapped with the scan (PCIe-bound by construction)

// A pageable host slice (what a caller's &[u8] normally is) reaches the GPU at the driver's
// single-threaded staging rate (~11 GB/s measured) when handed to cudaMemcpyAsync directly.  For large
// pageable haystacks the library stages each chunk itself: a small pool of worker threads memcpy()s
// slices of the chunk into a pinned ring buffer in parallel, and the DMA engine copies that buffer
// while the workers already fill the next one.  SS_B200_HOST_THREADS=0 turns the pool off.
namespace {

class CopyPool {
public:
    static CopyPool &get()
    {
        static CopyPool *p = new CopyPool(); // never destroyed: its detached workers outlive static teardown
        return *p;
    }
    int threads() const { return (int)workers_.size(); }
    // dst[0..len) = src[0..len), split over the workers and the calling thread; returns when done
    void copy(uint8_t *dst, const uint8_t *src, size_t len)
    {
        const size_t parts = workers_.size() + 1;
        const size_t slice = ((len + parts - 1) / parts + 4095) & ~(size_t)4095;
        Job job;
        size_t off = slice < len ? slice : len; // the caller copies the first slice itself
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (; off < len; off += slice) {
                const size_t n = len - off < slice ? len - off : slice;
                tasks_.push_back(Task{dst + off, src + off, n, &job});
                job.pending++;
            }
        }
        cv_.notify_all();
        memcpy(dst, src, slice < len ? slice : len);
        std::unique_lock<std::mutex> lk(mu_);
        job.cv.wait(lk, [&] { return job.pending == 0; });
    }

private:
    struct Job {
        size_t pending = 0;
        std::condition_variable cv;
    };
    struct Task {
        uint8_t *dst;
        const uint8_t *src;
        size_t n;
        Job *job;
    };
    CopyPool()
    {
        const char *v = getenv("SS_B200_HOST_THREADS");
        int n = v ? atoi(v) : -1;
        if (n < 0) {
            const unsigned hc = std::thread::hardware_concurrency();
            n = hc > 2 ? (int)(hc - 1 < 7 ? hc - 1 : 7) : 0; // 7 workers + the caller by default
        }
        for (int i = 0; i < n; i++)
            workers_.emplace_back([this] { run(); });
        for (auto &t : workers_)
            t.detach(); // process-lifetime pool
    }
    void run()
    {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return !tasks_.empty(); });
                t = tasks_.back();
                tasks_.pop_back();
            }
            memcpy(t.dst, t.src, t.n);
            std::lock_guard<std::mutex> lk(mu_);
            if (--t.job->pending == 0)
                t.job->cv.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<Task> tasks_;
    std::vector<std::thread> workers_;
};

bool host_pointer_is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

} // namespace

static size_t host_chunk_bytes()
{
    const char *v = getenv("SS_B200_HOST_CHUNK_MIB");
    size_t mib = v ? (size_t)atoll(v) : 0;
    if (mib == 0)
        mib = 64;
    return mib << 20;
}

extern "C" int ss_b200_find_in_host(const ss_b200_searcher *s, const uint8_t *host, size_t len, size_t *offset)
{
    if (!s || !offset || (len && !host))
        return SS_B200_E_ARG;
    const size_t k = s->needle.size();
    if (k == 0) {
        *offset = 0;
        return SS_B200_OK;
    }
    if (len < k) {
        *offset = SS_B200_NPOS;
        return SS_B200_OK;
    }
    if (k > 0xFFFFFFFFull)
        return SS_B200_E_ARG;
    SsDeviceInfo dev;
    int rc = ss_capi_device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    SsThreadCtx *c = nullptr;
    rc = ss_capi_get_ctx(&c);
    if (rc != SS_B200_OK)
        return rc;

    if (len <= SS_SMALL_HOST_MAX) {
        // short slice (the reference's short-haystack regime, src/x86.rs:363-375): no DMA, no events --
        // copy it into this thread's mapped pinned buffer and let the scan read it in place over PCIe;
        // one launch and the mapped result word are all that is left of the call
        if (!c->small_host) {
            SS_CUDA(cudaHostAlloc((void **)&c->small_host, SS_SMALL_HOST_MAX + 32, cudaHostAllocMapped));
            SS_CUDA(cudaHostGetDevicePointer((void **)&c->small_dev, c->small_host, 0));
        }
        memcpy(c->small_host, host, len);
        memset(c->small_host + len, 0, 32 - (len & 15)); // the scan reads whole 16-byte chunks
        return ss_capi_find_device_sync(s, c->small_dev, len, offset, 1); // direct loads, never the staged ring
    }

    const size_t halo = k - 1;
    size_t chunk = host_chunk_bytes();
    // large pageable slice: stage through pinned buffers with the copy pool (smaller chunks keep the
    // pinned ring modest and the pipeline busy)
    const bool staged = len >= ((size_t)8 << 20) && CopyPool::get().threads() > 0 && !host_pointer_is_pinned(host);
    if (staged && !getenv("SS_B200_HOST_CHUNK_MIB"))
        chunk = (size_t)32 << 20;
    if (chunk > len)
        chunk = (len + 15) & ~(size_t)15;
    const size_t end_total = len - k + 1;
    const size_t n_chunks = (end_total + chunk - 1) / chunk;
    const size_t need = chunk + halo + 32;
    if (c->dbuf_cap < need) {
        for (int b = 0; b < SsThreadCtx::NBUF; b++) {
            if (c->dbuf[b])
                cudaFree(c->dbuf[b]);
            c->dbuf[b] = nullptr;
            SS_CUDA(cudaMalloc(&c->dbuf[b], need));
            SS_CUDA(cudaMemset(c->dbuf[b], 0, need)); // chunk tails are read as whole 16-byte words
            if (!c->copied[b]) {
                SS_CUDA(cudaEventCreateWithFlags(&c->copied[b], cudaEventDisableTiming));
                SS_CUDA(cudaEventCreateWithFlags(&c->scanned[b], cudaEventDisableTiming));
            }
        }
        c->dbuf_cap = need;
    }
    if (staged && c->stage_cap < need) {
        for (int b = 0; b < SsThreadCtx::NBUF; b++) {
            if (c->stage[b])
                cudaFreeHost(c->stage[b]);
            c->stage[b] = nullptr;
            SS_CUDA(cudaHostAlloc((void **)&c->stage[b], need, cudaHostAllocDefault));
        }
        c->stage_cap = need;
    }
    if (c->chunk_results_cap < n_chunks) {
        if (c->chunk_results)
            cudaFreeHost(c->chunk_results);
        c->chunk_results = nullptr;
        SS_CUDA(cudaHostAlloc((void **)&c->chunk_results, n_chunks * sizeof(unsigned long long), cudaHostAllocMapped));
        SS_CUDA(cudaHostGetDevicePointer((void **)&c->chunk_results_dev, c->chunk_results, 0));
        c->chunk_results_cap = n_chunks;
    }
    for (size_t i = 0; i < n_chunks; i++)
        c->chunk_results[i] = ~0ull; // "not produced yet"

    ScanArgs proto;
    rc = ss_capi_build_args(s, c->dbuf[0], k, 0, (size_t)-1, dev.device, proto); // needle fields; geometry redone per chunk
    if (rc != SS_B200_OK)
        return rc;

    size_t submitted = 0;
    unsigned long long best = SS_NONE_U64;
    for (size_t i = 0; i < n_chunks; i++) {
        // the reference returns at the first match (src/lib.rs:242-244): stop feeding once an
        // already-finished chunk has reported one
        bool hit = false;
        for (size_t j = 0; j < submitted; j++) {
            const unsigned long long v = ((volatile unsigned long long *)c->chunk_results)[j];
            if (v != ~0ull && v != SS_NONE_U64) {
                hit = true;
                break;
            }
        }
        if (hit)
            break;
        const int b = (int)(i % SsThreadCtx::NBUF);
        const size_t off = i * chunk;
        size_t bytes = chunk + halo;
        if (off + bytes > len)
            bytes = len - off;
        if (i >= (size_t)SsThreadCtx::NBUF)
            SS_CUDA(cudaStreamWaitEvent(c->copy_stream, c->scanned[b], 0));
        const uint8_t *src = host + off;
        if (staged) {
            // the pinned buffer is free once its previous DMA has finished; fill it in parallel while the
            // DMA engine is still busy with the previous chunk
            if (i >= (size_t)SsThreadCtx::NBUF)
                SS_CUDA(cudaEventSynchronize(c->copied[b]));
            CopyPool::get().copy(c->stage[b], src, bytes);
            src = c->stage[b];
        }
        SS_CUDA(cudaMemcpyAsync(c->dbuf[b], src, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        SS_CUDA(cudaEventRecord(c->copied[b], c->copy_stream));
        SS_CUDA(cudaStreamWaitEvent(c->stream, c->copied[b], 0));
        ScanArgs a = proto;
        a.hay = c->dbuf[b];
        a.n = bytes;
        a.base = off;
        ss_host_scan_geometry(a, chunk);
        a.ws = c->ws;
        a.out = c->chunk_results_dev + i;
        SS_CUDA(ss_host_launch_scan(a, ss_capi_tuning(), dev, c->stream));
        SS_CUDA(cudaEventRecord(c->scanned[b], c->stream));
        submitted++;
    }
    SS_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t j = 0; j < submitted; j++) {
        const unsigned long long v = c->chunk_results[j];
        if (v != ~0ull && v < best)
            best = v;
    }
    *offset = (best == SS_NONE_U64) ? SS_B200_NPOS : (size_t)best;
    return SS_B200_OK;
}

extern "C" int ss_b200_search_in_host(const ss_b200_searcher *s, const uint8_t *host, size_t len, uint8_t *found)
{
    if (!found)
        return SS_B200_E_ARG;
    size_t off = SS_B200_NPOS;
    int rc = ss_b200_find_in_host(s, host, len, &off);
    if (rc == SS_B200_OK)
        *found = (off != SS_B200_NPOS) ? 1 : 0;
    return rc;
}

