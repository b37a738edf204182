// host_engine.cu -- search_in(&[u8]) on a HOST slice (the literal drop-in call, reference src/x86.rs:523).
//
// One engine serves ss_b200_find_in_host (one device: the calling thread's lane) and
// ss_b200_find_in_host_multi (capi_ctx.cu: one lane per device of a context).  The slice is cut into
// chunks of start positions, handed out in ascending order to whichever lane has room, so the devices
// sweep the slice as one band, all PCIe links are busy from the first chunk on, and a faster link simply
// takes more chunks.  Per lane the chunks run through a ring of
// three device buffers: the H2D copy of chunk j+1 (copy stream) overlaps the scan of chunk j (scan
// stream), each scan writes its result into a mapped pinned word.  The host never runs more than the
// ring depth ahead of the results it has seen, so a match stops the feeding within a few chunks -- the
// reference's early return (src/lib.rs:242-244) -- and the answer is the minimum over everything that
// was submitted (chunks are submitted in ascending order, so nothing left of the winner is missing).
//
// Three data paths:
//   DMA ring   (default for long slices) chunked cudaMemcpyAsync into HBM, scanned there at the HBM rate
//   in place   pinned (page-locked, device-mapped) input is read by the scan kernel straight over PCIe:
//              no staging buffers, no copy/scan hand-off
//   pageable   an ordinary &[u8]: a pool of memcpy workers fills a pinned ring in parallel (the driver's
//              own single-threaded staging reaches ~11 GB/s) and the DMA engine drains it
// A slice of up to 32 KiB (the reference's short-haystack regime, src/x86.rs:363-375) is copied into the
// lane's mapped pinned buffer and scanned in place: one launch, no DMA, no events.
#include "capi_internal.h"

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace {

// memcpy workers for pageable input.  Process-lifetime, grown on demand, never destroyed (detached
// workers must outlive static teardown).
class CopyPool {
public:
    static CopyPool &get()
    {
        static CopyPool *p = new CopyPool();
        return *p;
    }
    // make sure at least `n` workers exist; returns the worker count
    int ensure(int n)
    {
        std::lock_guard<std::mutex> lk(mu_);
        while ((int)n_workers_ < n) {
            std::thread([this] { run(); }).detach();
            n_workers_++;
        }
        return (int)n_workers_;
    }
    // dst[0..len) = src[0..len), split over `parts` - 1 workers and the calling thread; returns when done
    void copy(uint8_t *dst, const uint8_t *src, size_t len, int parts)
    {
        if (parts < 1)
            parts = 1;
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
    size_t n_workers_ = 0;
};

// Is `p` page-locked memory the device can address?  *dev_view receives the device-side pointer.
bool host_pointer_is_pinned(const void *p, const uint8_t **dev_view)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (at.type != cudaMemoryTypeHost && at.type != cudaMemoryTypeManaged)
        return false;
    *dev_view = (const uint8_t *)at.devicePointer; // nullptr when the registration is not mapped
    return true;
}

size_t round_up_pow2(size_t v)
{
    size_t p = 1;
    while (p < v)
        p <<= 1;
    return p;
}

// Chunk size: the caller's setting, else sized from the slice -- an eighth of a lane's share, between 4
// and 64 MiB, so that a lane's ring (3 buffers) is never larger than 3/8 of what it has to move and short
// slices do not allocate 3 x 64 MiB.
size_t pick_chunk(size_t len, int n_lanes, bool staged, const SsHostPathTuning &t)
{
    size_t chunk;
    if (t.chunk_mib > 0) {
        chunk = (size_t)t.chunk_mib << 20;
    } else {
        const size_t share = (len + n_lanes - 1) / n_lanes;
        chunk = round_up_pow2((share + 7) / 8);
        const size_t lo = (size_t)4 << 20, hi = (size_t)64 << 20;
        (void)staged;
        chunk = chunk < lo ? lo : (chunk > hi ? hi : chunk);
    }
    if (chunk > len)
        chunk = (len + 15) & ~(size_t)15;
    return chunk;
}

int ensure_ring(SsLane *c, size_t need, bool staged)
{
    if (c->dbuf_cap < need) {
        for (int b = 0; b < SsLane::NBUF; b++) {
            if (c->dbuf[b])
                cudaFree(c->dbuf[b]);
            c->dbuf[b] = nullptr;
        }
        c->dbuf_cap = 0;
        for (int b = 0; b < SsLane::NBUF; b++) {
            SS_CUDA(cudaMalloc(&c->dbuf[b], need));
            // chunk tails are read as whole 16-byte words: give the bytes behind a copy a defined value.
            // On the copy stream, so the memset is ordered before the first copy into the buffer.
            SS_CUDA(cudaMemsetAsync(c->dbuf[b], 0, need, c->copy_stream));
        }
        c->dbuf_cap = need;
    }
    for (int b = 0; b < SsLane::NBUF; b++) {
        if (!c->copied[b])
            SS_CUDA(cudaEventCreateWithFlags(&c->copied[b], cudaEventDisableTiming));
        if (!c->scanned[b])
            SS_CUDA(cudaEventCreateWithFlags(&c->scanned[b], cudaEventDisableTiming));
    }
    if (staged && c->stage_cap < need) {
        for (int b = 0; b < SsLane::NBUF; b++) {
            if (c->stage[b])
                cudaFreeHost(c->stage[b]);
            c->stage[b] = nullptr;
        }
        c->stage_cap = 0;
        for (int b = 0; b < SsLane::NBUF; b++)
            SS_CUDA(cudaHostAlloc((void **)&c->stage[b], need, cudaHostAllocPortable));
        c->stage_cap = need;
    }
    return SS_B200_OK;
}

int ensure_chunk_results(SsLane *c, size_t n)
{
    if (c->chunk_results_cap >= n)
        return SS_B200_OK;
    if (c->chunk_results)
        cudaFreeHost(c->chunk_results);
    c->chunk_results = c->chunk_results_dev = nullptr;
    c->chunk_results_cap = 0;
    const size_t cap = n < 64 ? 64 : n;
    SS_CUDA(cudaHostAlloc((void **)&c->chunk_results, cap * sizeof(unsigned long long),
                          cudaHostAllocMapped | cudaHostAllocPortable));
    SS_CUDA(cudaHostGetDevicePointer((void **)&c->chunk_results_dev, c->chunk_results, 0));
    c->chunk_results_cap = cap;
    return SS_B200_OK;
}

// in-place auto mode: pinned slices of up to this many bytes PER LANE are read over PCIe by the scan itself.
// Measured (profiles/r02_host_path_n1.json, _n8.json): one launch per lane beats the copy/scan ring's ramp
// below ~2 MiB per device (1 MiB on one GPU: 26.4 vs 23.5 GB/s; 16 MiB over eight: 131 vs 93 GB/s); from
// 16 MiB per device on, the DMA ring wins (49.7 vs 47.8, 55.3 vs 50.2 GB/s).
constexpr size_t SS_INPLACE_AUTO_MAX_PER_LANE = (size_t)2 << 20;

} // namespace

int ss_host_engine_find(SsLane *const *lanes, int n_lanes_in, const ss_b200_searcher *s, const uint8_t *host,
                        size_t len, size_t *offset, SsHostStats *stats)
{
    if (!s || !offset || !lanes || n_lanes_in < 1 || (len && !host))
        return SS_B200_E_ARG;
    int n_lanes = n_lanes_in;
    SsHostStats st_local;
    SsHostStats &st = stats ? *stats : st_local;
    st = SsHostStats();
    const size_t k = s->needle.size();
    if (k == 0) { // N0 => true, even for an empty haystack (src/x86.rs:470,500)
        *offset = 0;
        return SS_B200_OK;
    }
    if (len < k) { // src/x86.rs:357-359; k == 1: src/lib.rs:131-133
        *offset = SS_B200_NPOS;
        return SS_B200_OK;
    }
    if (k > 0xFFFFFFFFull)
        return SS_B200_E_ARG;

    if (len <= SS_SMALL_HOST_MAX) {
        // short slice: copy it into the lane's mapped pinned buffer and let the scan read it in place over
        // PCIe; one launch and the mapped result word are all that is left of the call
        SsLane *c = lanes[0];
        if (!c->small_host) {
            SS_CUDA(cudaHostAlloc((void **)&c->small_host, SS_SMALL_HOST_MAX + 32,
                                  cudaHostAllocMapped | cudaHostAllocPortable));
            SS_CUDA(cudaHostGetDevicePointer((void **)&c->small_dev, c->small_host, 0));
        }
        memcpy(c->small_host, host, len);
        memset(c->small_host + len, 0, 32 - (len & 15)); // the scan reads whole 16-byte chunks
        st.mode = 3;
        st.chunks = 1;
        {
            // through the lane's resident kernel when there is one to be had: no launch (service.cu)
            int on = 0;
            unsigned idle_us = 0;
            ss_capi_service_tuning(&on, &idle_us);
            if (on && ss_service_eligible(s, len)) {
                SsDeviceGuard guard(c->device);
                bool used = false;
                int rc = ss_service_find(c, s, c->small_dev, len, idle_us, offset, &used, true);
                if (rc != SS_B200_OK || used)
                    return rc;
            }
        }
        return ss_capi_find_on_lane(c, s, c->small_dev, len, offset, 1); // direct loads, never the staged ring
    }

    const SsHostPathTuning ht = ss_capi_host_tuning();
    const uint8_t *dev_view = nullptr;
    const bool pinned = host_pointer_is_pinned(host, &dev_view);
    int mode = ht.mode;
    if (mode >= 2 && !(pinned && dev_view))
        mode = 1; // nothing to read in place: the device cannot address this memory
    if (mode == 0)
        mode = (pinned && dev_view && len <= SS_INPLACE_AUTO_MAX_PER_LANE * (size_t)n_lanes) ? 2 : 1;
    const bool inplace = mode >= 2;

    int pool_threads = 0;
    bool staged = false;
    if (!inplace && !pinned && len >= ((size_t)8 << 20) && ht.copy_threads != 0) {
        int want = ht.copy_threads;
        if (want < 0) {
            const unsigned hc = std::thread::hardware_concurrency();
            // workers next to the calling thread.  The staging memcpy is what bounds pageable input, and it
            // scales with threads well past 8 on these hosts (2 GiB pageable, one box: 3 workers 27 GB/s,
            // 7: 36, 11: 44, 15: 45 with 64 MiB chunks; profiles/r02_pageable_threads.txt)
            want = hc > 2 ? (int)(hc - 1 < 15u ? hc - 1 : 15u) : 0;
        }
        if (want > 0) {
            pool_threads = CopyPool::get().ensure(want);
            pool_threads = pool_threads < want ? pool_threads : want;
            staged = true;
        }
    }
    if (staged) {
        // Pageable input is bound by the staging memcpy (the host's DRAM read + write), not by a PCIe link:
        // one device drains the pinned ring as fast as the pool fills it, and more lanes only cut the slice
        // into smaller chunks (measured on 8 x B200, 1 GiB pageable: 49 GB/s over one device, 42 / 39 / 34
        // over 2 / 4 / 8; profiles/r02_host_path_n8.json).
        n_lanes = 1;
    }

    const size_t halo = k - 1;
    size_t chunk = pick_chunk(len, n_lanes, staged, ht);
    if (inplace && ht.chunk_mib == 0) {
        // no ring to keep small: one launch per lane covers up to 64 MiB (the kernel stops early by itself)
        const size_t share = (((len + n_lanes - 1) / n_lanes) + 15) & ~(size_t)15;
        chunk = share < ((size_t)64 << 20) ? share : (size_t)64 << 20;
    }
    const size_t end_total = len - k + 1;
    const size_t n_chunks = (end_total + chunk - 1) / chunk;
    const int used_lanes = (size_t)n_lanes < n_chunks ? n_lanes : (int)n_chunks;
    st.mode = inplace ? 2 : 1;
    st.staged = staged ? 1 : 0;
    st.chunk_bytes = chunk;

    // per-lane setup
    std::vector<SsDeviceInfo> devs(used_lanes);
    std::vector<ScanArgs> protos(used_lanes);
    for (int l = 0; l < used_lanes; l++) {
        SsLane *c = lanes[l];
        SsDeviceGuard guard(c->device);
        int rc = ss_capi_device_info(devs[l]);
        if (rc != SS_B200_OK)
            return rc;
        if (!inplace) {
            rc = ensure_ring(c, chunk + halo + 32, staged);
            if (rc != SS_B200_OK)
                return rc;
        }
        // chunks go to whichever lane has room (below), so any lane may end up with any number of them
        rc = ensure_chunk_results(c, n_chunks);
        if (rc != SS_B200_OK)
            return rc;
        for (size_t j = 0; j < n_chunks; j++)
            c->chunk_results[j] = SS_RESULT_PENDING;
        // needle fields once per lane; the geometry is redone per chunk
        rc = ss_capi_build_args(s, inplace ? (const void *)dev_view : (const void *)c->dbuf[0], k, 0, (size_t)-1,
                                devs[l].device, protos[l]);
        if (rc != SS_B200_OK)
            return rc;
    }
    SsScanTuning tuning = ss_capi_tuning();
    if (inplace)
        tuning.variant = (ht.mode == 3) ? 2 : 1;

    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    int cur_dev = prev_dev;
    auto make_current = [&](int dev) -> cudaError_t {
        if (dev == cur_dev)
            return cudaSuccess;
        cur_dev = dev;
        return cudaSetDevice(dev);
    };

    int rc = SS_B200_OK;
    size_t submitted = 0;
    bool hit = false;
    // per lane: chunks issued so far, and how many of their results (in order) the host has looked at
    std::vector<size_t> issued(used_lanes, 0), seen(used_lanes, 0);
    auto consume = [&](int l) { // look at every result of lane l that has arrived; true on a match
        volatile unsigned long long *r = lanes[l]->chunk_results;
        while (seen[l] < issued[l]) {
            const unsigned long long v = r[seen[l]];
            if (v == SS_RESULT_PENDING)
                break;
            seen[l]++;
            if (v != SS_NONE_U64)
                return true;
        }
        return false;
    };
    int rr = 0; // where the search for a free lane starts: round robin among equally loaded lanes
    for (size_t i = 0; i < n_chunks && rc == SS_B200_OK && !hit; i++) {
        // Which lane takes chunk i?  The one with the fewest chunks in flight -- NOT i % lanes: the PCIe
        // links of a box are not equally fast when all are busy (measured on 8 x B200: four GPUs behind one
        // shared upstream get 23 GB/s each, the other four 35 GB/s; profiles/r02_host_path_n8.json), and a
        // static deal runs every link at the pace of the slowest.  A lane holds at most NBUF chunks (its ring
        // has NBUF buffers): with every lane full, wait for whichever result arrives first.
        int l = -1;
        unsigned spins = 0;
        while (!hit && rc == SS_B200_OK) {
            // any finished chunk (of any lane) with a match ends the feeding: the reference would have
            // returned already (src/lib.rs:242-244)
            for (int l2 = 0; l2 < used_lanes && !hit; l2++)
                hit = consume(l2);
            if (hit)
                break;
            size_t best_load = (size_t)SsLane::NBUF;
            for (int t = 0; t < used_lanes; t++) {
                const int cand = (rr + t) % used_lanes;
                const size_t load = issued[cand] - seen[cand];
                if (load < best_load) {
                    best_load = load;
                    l = cand;
                }
            }
            if (l >= 0)
                break;
            if ((++spins & 0x3FFF) == 0) { // every lane is full: keep an eye on the streams while spinning
                for (int l2 = 0; l2 < used_lanes; l2++) {
                    cudaError_t eq = cudaStreamQuery(lanes[l2]->stream);
                    if (eq != cudaSuccess && eq != cudaErrorNotReady)
                        rc = ss_capi_cuda_fail(eq, "host-slice scan");
                }
            }
        }
        if (hit || rc != SS_B200_OK)
            break;
        rr = (l + 1) % used_lanes;
        const size_t j = issued[l]; // index of the chunk within its lane
        SsLane *c = lanes[l];
        const int b = (int)(j % SsLane::NBUF);
        cudaError_t e = make_current(c->device);
        if (e != cudaSuccess) {
            rc = ss_capi_cuda_fail(e, "cudaSetDevice");
            break;
        }
        const size_t off = i * chunk;
        size_t bytes = chunk + halo;
        if (off + bytes > len)
            bytes = len - off;
        ScanArgs a = protos[l];
        a.base = off;
        a.ws = c->ws;
        a.out = c->chunk_results_dev + j;
        if (inplace) {
            a.hay = dev_view + off;
            a.n = bytes;
            ss_host_scan_geometry(a, chunk);
            e = ss_host_launch_scan(a, tuning, devs[l], c->stream);
        } else {
            if (j >= (size_t)SsLane::NBUF)
                e = cudaStreamWaitEvent(c->copy_stream, c->scanned[b], 0);
            const uint8_t *src = host + off;
            if (e == cudaSuccess && staged) {
                // the pinned buffer is free once its previous DMA has finished; fill it in parallel while
                // the DMA engines are still busy with earlier chunks
                if (j >= (size_t)SsLane::NBUF)
                    e = cudaEventSynchronize(c->copied[b]);
                if (e == cudaSuccess) {
                    CopyPool::get().copy(c->stage[b], src, bytes, pool_threads + 1);
                    src = c->stage[b];
                }
            }
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(c->dbuf[b], src, bytes, cudaMemcpyHostToDevice, c->copy_stream);
            if (e == cudaSuccess)
                e = cudaEventRecord(c->copied[b], c->copy_stream);
            if (e == cudaSuccess)
                e = cudaStreamWaitEvent(c->stream, c->copied[b], 0);
            if (e == cudaSuccess) {
                a.hay = c->dbuf[b];
                a.n = bytes;
                ss_host_scan_geometry(a, chunk);
                e = ss_host_launch_scan(a, tuning, devs[l], c->stream);
            }
            if (e == cudaSuccess)
                e = cudaEventRecord(c->scanned[b], c->stream);
            st.h2d_bytes += bytes;
        }
        if (e != cudaSuccess) {
            rc = ss_capi_cuda_fail(e, "host-slice chunk");
            break;
        }
        issued[l]++;
        submitted++;
    }
    st.chunks = submitted;
    // drain: every submitted chunk must have reported (also on the error path: the ring is reused)
    unsigned long long best = SS_NONE_U64;
    for (int l = 0; l < used_lanes; l++) {
        SsLane *c = lanes[l];
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess && rc == SS_B200_OK)
            rc = ss_capi_cuda_fail(e, "cudaStreamSynchronize(host-slice)");
        for (size_t j = 0; j < issued[l]; j++) {
            const unsigned long long v = c->chunk_results[j];
            if (v == SS_RESULT_PENDING) {
                if (rc == SS_B200_OK) {
                    ss_capi_set_error("host-slice chunk retired without publishing a result");
                    rc = SS_B200_E_CUDA;
                }
            } else if (v < best) {
                best = v;
            }
        }
    }
    if (cur_dev != prev_dev && prev_dev >= 0)
        cudaSetDevice(prev_dev);
    if (rc != SS_B200_OK)
        return rc;
    *offset = (best == SS_NONE_U64) ? SS_B200_NPOS : (size_t)best;
    return SS_B200_OK;
}

extern "C" int ss_b200_find_in_host(const ss_b200_searcher *s, const uint8_t *host, size_t len, size_t *offset)
{
    if (!s || !offset || (len && !host))
        return SS_B200_E_ARG;
    if (s->needle.empty()) { // N0: no device needed (src/x86.rs:470,500)
        *offset = 0;
        return SS_B200_OK;
    }
    if (len < s->needle.size()) {
        *offset = SS_B200_NPOS;
        return SS_B200_OK;
    }
    SsLane *c = nullptr;
    int rc = ss_capi_get_lane(&c);
    if (rc != SS_B200_OK)
        return rc;
    return ss_host_engine_find(&c, 1, s, host, len, offset, nullptr);
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

// Measured host -> device copy bandwidth of the calling thread's device: `bytes` of pinned host memory
// copied `reps` times with cudaMemcpyAsync (the e2e path's PCIe ceiling, reported next to it by bench.py).
extern "C" int ss_b200_measure_h2d(size_t bytes, int reps, double *gb_per_s)
{
    if (!gb_per_s || bytes == 0 || reps < 1)
        return SS_B200_E_ARG;
    SsLane *c = nullptr;
    int rc = ss_capi_get_lane(&c);
    if (rc != SS_B200_OK)
        return rc;
    void *h = nullptr, *d = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
    if (e == cudaSuccess)
        e = cudaMalloc(&d, bytes);
    if (e == cudaSuccess) {
        memset(h, 0x5A, bytes);
        e = cudaEventCreate(&e0);
    }
    if (e == cudaSuccess)
        e = cudaEventCreate(&e1);
    float best_ms = 0.f;
    for (int r = 0; r < reps + 1 && e == cudaSuccess; r++) { // first pass warms up
        cudaEventRecord(e0, c->copy_stream);
        e = cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->copy_stream);
        cudaEventRecord(e1, c->copy_stream);
        if (e == cudaSuccess)
            e = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (e == cudaSuccess)
            e = cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && (best_ms == 0.f || ms < best_ms))
            best_ms = ms;
    }
    if (e0)
        cudaEventDestroy(e0);
    if (e1)
        cudaEventDestroy(e1);
    if (d)
        cudaFree(d);
    if (h)
        cudaFreeHost(h);
    if (e != cudaSuccess)
        return ss_capi_cuda_fail(e, "ss_b200_measure_h2d");
    *gb_per_s = (double)bytes / (best_ms * 1e-3) / 1e9;
    return SS_B200_OK;
}
