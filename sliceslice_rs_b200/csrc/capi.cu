// capi.cu -- the extern "C" boundary declared in include/sliceslice_b200.h.
//
// Host runtime around the sm_100a kernels: opaque searcher / haystack handles,
// thread-local stream + self-resetting workspace + mapped result slot for the
// synchronous calls, the chunked host->device streaming path, and the
// stream-ordered entry used for roofline measurement and the multi-GPU shards.
// No CPU search path exists in this library: without a device every search
// returns SS_B200_E_CUDA.
#include "capi_internal.h"

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

// ---------------------------------------------------------------------------------------------
// errors

static thread_local std::string t_last_error;

static int cuda_fail(cudaError_t e, const char *what)
{
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    t_last_error = buf;
    return e == cudaErrorMemoryAllocation ? SS_B200_E_NOMEM : SS_B200_E_CUDA;
}
int ss_capi_cuda_fail(cudaError_t e, const char *what) { return cuda_fail(e, what); }
void ss_capi_set_error(const char *msg) { t_last_error = msg; }

extern "C" const char *ss_b200_strerror(int status)
{
    switch (status) {
    case SS_B200_OK: return "ok";
    case SS_B200_E_POSITION: return "position is not a valid index for the needle";
    case SS_B200_E_EMPTY_NEEDLE: return "needle is empty";
    case SS_B200_E_ARG: return "invalid argument";
    case SS_B200_E_CUDA: return "CUDA error or no usable device";
    case SS_B200_E_NOMEM: return "out of memory";
    case SS_B200_E_NCCL: return "NCCL error or libnccl not loadable";
    default: return "unknown status";
    }
}
extern "C" const char *ss_b200_last_error(void) { return t_last_error.c_str(); }
extern "C" int ss_b200_abi_version(void) { return SS_B200_ABI_VERSION; }

// ---------------------------------------------------------------------------------------------
// process-wide tuning + per-device facts

// The setters may be called while other threads search: every field is an atomic and each search works
// on one snapshot (ss_capi_tuning).
namespace {
struct AtomicTuning {
    std::atomic<int> variant{0}, ctas_per_sm{0}, unroll{0}, tile_kib{0}, stages{0}, extra_anchors{-1}, pdl{1};
    std::atomic<int> host_mode{0}, host_chunk_mib{0}, host_copy_threads{-1};
    std::atomic<int> service_on{1}, service_idle_us{100};
} g_tuning;
std::mutex g_dev_mutex;
std::map<int, SsDeviceInfo> g_devs;
} // namespace

static int device_info(SsDeviceInfo &out)
{
    int dev = -1;
    SS_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_dev_mutex);
    auto it = g_devs.find(dev);
    if (it == g_devs.end()) {
        SsDeviceInfo d;
        d.device = dev;
        SS_CUDA(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
        SS_CUDA(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        SS_CUDA(cudaDeviceGetAttribute(&d.smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        it = g_devs.emplace(dev, d).first;
    }
    out = it->second;
    return SS_B200_OK;
}

int ss_capi_device_info(SsDeviceInfo &out) { return device_info(out); }
SsScanTuning ss_capi_tuning()
{
    SsScanTuning t;
    t.variant = g_tuning.variant.load(std::memory_order_relaxed);
    t.ctas_per_sm = g_tuning.ctas_per_sm.load(std::memory_order_relaxed);
    t.unroll = g_tuning.unroll.load(std::memory_order_relaxed);
    t.tile_kib = g_tuning.tile_kib.load(std::memory_order_relaxed);
    t.stages = g_tuning.stages.load(std::memory_order_relaxed);
    t.extra_anchors = g_tuning.extra_anchors.load(std::memory_order_relaxed);
    t.pdl = g_tuning.pdl.load(std::memory_order_relaxed);
    return t;
}
SsHostPathTuning ss_capi_host_tuning()
{
    SsHostPathTuning t;
    t.mode = g_tuning.host_mode.load(std::memory_order_relaxed);
    t.chunk_mib = g_tuning.host_chunk_mib.load(std::memory_order_relaxed);
    t.copy_threads = g_tuning.host_copy_threads.load(std::memory_order_relaxed);
    return t;
}

extern "C" int ss_b200_set_scan_variant(int variant)
{
    if (variant < 0 || variant > 2)
        return SS_B200_E_ARG;
    g_tuning.variant = variant;
    return SS_B200_OK;
}
extern "C" int ss_b200_set_scan_tuning(int ctas_per_sm, int unroll, int tile_kib, int stages)
{
    // 0 = auto everywhere; anything else must be a value the launcher knows (scan_long.cu)
    if (ctas_per_sm < 0 || ctas_per_sm > 32 || (unroll != 0 && unroll != 1 && unroll != 4) ||
        (tile_kib != 0 && tile_kib != 16 && tile_kib != 32) || stages < 0 || stages > 8 || stages == 1)
        return SS_B200_E_ARG;
    g_tuning.ctas_per_sm = ctas_per_sm;
    g_tuning.unroll = unroll;
    g_tuning.tile_kib = tile_kib;
    g_tuning.stages = stages;
    return SS_B200_OK;
}
extern "C" int ss_b200_set_extra_anchors(int n)
{
    if (n < -1 || n > 1)
        return SS_B200_E_ARG;
    g_tuning.extra_anchors = n;
    return SS_B200_OK;
}
extern "C" int ss_b200_set_launch_pdl(int on)
{
    if (on != 0 && on != 1)
        return SS_B200_E_ARG;
    g_tuning.pdl = on;
    return SS_B200_OK;
}
extern "C" int ss_b200_set_host_path(int mode, int chunk_mib, int copy_threads)
{
    if (mode < 0 || mode > 3 || chunk_mib < 0 || chunk_mib > 4096 || copy_threads < -1 || copy_threads > 256)
        return SS_B200_E_ARG;
    g_tuning.host_mode = mode;
    g_tuning.host_chunk_mib = chunk_mib;
    g_tuning.host_copy_threads = copy_threads;
    return SS_B200_OK;
}
void ss_capi_service_tuning(int *on, unsigned *idle_us)
{
    *on = g_tuning.service_on.load(std::memory_order_relaxed);
    *idle_us = (unsigned)g_tuning.service_idle_us.load(std::memory_order_relaxed);
}
extern "C" int ss_b200_set_sync_service(int on, int idle_us)
{
    if ((on != 0 && on != 1) || idle_us < 0 || idle_us > 1000000)
        return SS_B200_E_ARG;
    g_tuning.service_on = on;
    if (idle_us > 0)
        g_tuning.service_idle_us = idle_us;
    return SS_B200_OK;
}
extern "C" uint64_t ss_b200_launch_count(void) { return ss_host_launch_count(); }

// ---------------------------------------------------------------------------------------------
// handles

static int make_searcher(const uint8_t *needle, size_t len, size_t position, bool have_position, bool strict,
                         ss_b200_searcher **out)
{
    if (!out || (len && !needle))
        return SS_B200_E_ARG;
    *out = nullptr;
    if (!have_position)
        position = len - 1; // wrapping_sub(1), src/x86.rs:285,457
    if (strict) {
        // Avx2Searcher::with_position, src/x86.rs:297-305
        if (len == 0)
            return SS_B200_E_EMPTY_NEEDLE; // "position < size" cannot hold
        if (position >= len)
            return SS_B200_E_POSITION;
    } else {
        // DynamicAvx2Searcher::with_position, src/x86.rs:468-493
        if (len == 1 && position != 0)
            return SS_B200_E_POSITION; // assert_eq!(position, 0) :473
        if (len >= 2 && position >= len)
            return SS_B200_E_POSITION; // :300
        if (len == 0)
            position = 0; // N0: position ignored (:470)
    }
    ss_b200_searcher *s = new (std::nothrow) ss_b200_searcher();
    if (!s)
        return SS_B200_E_NOMEM;
    s->needle.assign(needle, needle + len);
    s->position = position;
    s->strict = strict;
    *out = s;
    return SS_B200_OK;
}

extern "C" int ss_b200_searcher_new(const uint8_t *needle, size_t len, ss_b200_searcher **out)
{
    return make_searcher(needle, len, 0, false, false, out);
}
extern "C" int ss_b200_searcher_with_position(const uint8_t *needle, size_t len, size_t position,
                                              ss_b200_searcher **out)
{
    return make_searcher(needle, len, position, true, false, out);
}
extern "C" int ss_b200_searcher_new_strict(const uint8_t *needle, size_t len, ss_b200_searcher **out)
{
    return make_searcher(needle, len, 0, false, true, out);
}
extern "C" int ss_b200_searcher_with_position_strict(const uint8_t *needle, size_t len, size_t position,
                                                     ss_b200_searcher **out)
{
    return make_searcher(needle, len, position, true, true, out);
}
extern "C" void ss_b200_searcher_free(ss_b200_searcher *s)
{
    if (!s)
        return;
    for (auto &kv : s->dev_needle)
        cudaFree(kv.second);
    delete s;
}
extern "C" size_t ss_b200_searcher_needle_len(const ss_b200_searcher *s) { return s ? s->needle.size() : 0; }
extern "C" size_t ss_b200_searcher_position(const ss_b200_searcher *s) { return s ? s->position : 0; }

extern "C" int ss_b200_haystack_upload(const uint8_t *host, size_t len, ss_b200_haystack **out)
{
    if (!out || (len && !host))
        return SS_B200_E_ARG;
    *out = nullptr;
    int dev = -1;
    SS_CUDA(cudaGetDevice(&dev));
    uint8_t *d = nullptr;
    const size_t alloc = ((len + 15) & ~(size_t)15) + 16;
    SS_CUDA(cudaMalloc(&d, alloc));
    // the kernels read whole 16-byte chunks: give the padding behind the haystack a defined value
    cudaError_t e = cudaMemset(d + len, 0, alloc - len);
    if (e == cudaSuccess && len)
        e = cudaMemcpy(d, host, len, cudaMemcpyHostToDevice);
    // Both calls go to the legacy default stream and may return before the bytes have landed (memset is
    // asynchronous, a pageable copy of <= 64 KiB returns once it is staged); the searches run on the
    // library's own non-blocking streams, which do not wait for the legacy stream.  Finish here.
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) {
        cudaFree(d);
        return cuda_fail(e, "cudaMemcpy(haystack)");
    }
    ss_b200_haystack *h = new (std::nothrow) ss_b200_haystack();
    if (!h) {
        cudaFree(d);
        return SS_B200_E_NOMEM;
    }
    h->dptr = d;
    h->len = len;
    h->owned = true;
    h->plain_device_memory = true;
    h->device = dev;
    *out = h;
    return SS_B200_OK;
}
extern "C" int ss_b200_haystack_from_device(const void *dptr, size_t len, ss_b200_haystack **out)
{
    if (!out || (len && !dptr))
        return SS_B200_E_ARG;
    ss_b200_haystack *h = new (std::nothrow) ss_b200_haystack();
    if (!h)
        return SS_B200_E_NOMEM;
    h->dptr = (const uint8_t *)dptr;
    h->len = len;
    h->owned = false;
    if (dptr) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, dptr) == cudaSuccess) {
            h->plain_device_memory = at.type == cudaMemoryTypeDevice;
            h->device = at.device;
        } else {
            cudaGetLastError();
        }
    }
    *out = h;
    return SS_B200_OK;
}
extern "C" void ss_b200_haystack_free(ss_b200_haystack *h)
{
    if (!h)
        return;
    if (h->owned && h->dptr)
        cudaFree((void *)h->dptr);
    delete h;
}
extern "C" size_t ss_b200_haystack_len(const ss_b200_haystack *h) { return h ? h->len : 0; }
extern "C" const void *ss_b200_haystack_device_ptr(const ss_b200_haystack *h) { return h ? h->dptr : nullptr; }

// ---------------------------------------------------------------------------------------------
// lanes: per-device resources of the synchronous calls

int SsLane::init(int dev)
{
    SsDeviceGuard guard(dev);
    SS_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    SS_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    SS_CUDA(cudaMalloc(&ws, 64));
    SS_CUDA(cudaMemsetAsync(ws, 0, 64, stream));
    SS_CUDA(cudaHostAlloc((void **)&slot, sizeof(SsHostSlot), cudaHostAllocMapped | cudaHostAllocPortable));
    slot->value = 0;
    slot->pad = 0;
    SS_CUDA(cudaHostGetDevicePointer((void **)&slot_dev, (void *)slot, 0));
    SS_CUDA(cudaStreamSynchronize(stream));
    device = dev;
    return SS_B200_OK;
}

void SsLane::release()
{
    if (device < 0 && !stream && !ws)
        return;
    // at process teardown the runtime may already be gone: every call below then fails harmlessly
    SsDeviceGuard guard(device);
    if (service) {
        ss_service_release(service); // retires the resident kernel first
        service = nullptr;
    }
    if (stream)
        cudaStreamSynchronize(stream);
    if (copy_stream)
        cudaStreamSynchronize(copy_stream);
    for (int b = 0; b < NBUF; b++) {
        if (dbuf[b])
            cudaFree(dbuf[b]);
        if (stage[b])
            cudaFreeHost(stage[b]);
        if (copied[b])
            cudaEventDestroy(copied[b]);
        if (scanned[b])
            cudaEventDestroy(scanned[b]);
        dbuf[b] = stage[b] = nullptr;
        copied[b] = scanned[b] = nullptr;
    }
    dbuf_cap = stage_cap = 0;
    if (small_host)
        cudaFreeHost(small_host);
    small_host = small_dev = nullptr;
    if (chunk_results)
        cudaFreeHost(chunk_results);
    chunk_results = chunk_results_dev = nullptr;
    chunk_results_cap = 0;
    if (slot)
        cudaFreeHost((void *)slot);
    slot = slot_dev = nullptr;
    if (ws)
        cudaFree(ws);
    ws = nullptr;
    if (stream)
        cudaStreamDestroy(stream);
    if (copy_stream)
        cudaStreamDestroy(copy_stream);
    stream = copy_stream = nullptr;
    device = -1;
    cudaGetLastError();
}

size_t SsLane::pinned_bytes() const
{
    return stage_cap * NBUF + (small_host ? SS_SMALL_HOST_MAX + 32 : 0) + chunk_results_cap * 8 +
           (slot ? sizeof(SsHostSlot) : 0);
}

// The calling thread's lanes, one per device it has searched on.  The map's destructor runs at thread
// exit and releases the CUDA resources (a host that churns threads leaks nothing);
// ss_b200_thread_release() does the same on demand.
static thread_local std::map<int, SsLane> t_lanes;

int ss_capi_get_lane(SsLane **out)
{
    int dev = -1;
    SS_CUDA(cudaGetDevice(&dev));
    SsLane &c = t_lanes[dev];
    if (c.device < 0) {
        int rc = c.init(dev);
        if (rc != SS_B200_OK) {
            c.release();
            t_lanes.erase(dev);
            return rc;
        }
    }
    *out = &c;
    return SS_B200_OK;
}

extern "C" int ss_b200_thread_release(void)
{
    t_lanes.clear();
    return SS_B200_OK;
}

extern "C" int ss_b200_thread_footprint(size_t *device_bytes, size_t *pinned_bytes)
{
    size_t d = 0, p = 0;
    for (auto &kv : t_lanes) {
        d += kv.second.device_bytes();
        p += kv.second.pinned_bytes();
    }
    if (device_bytes)
        *device_bytes = d;
    if (pinned_bytes)
        *pinned_bytes = p;
    return SS_B200_OK;
}

int ss_capi_wait_slot(volatile unsigned long long *word, unsigned long long pending, cudaStream_t stream)
{
    // spin on the mapped result word; fall back to the stream status every so often
    unsigned spins = 0;
    while (*word == pending) {
        if ((++spins & 0x3FF) == 0) {
            cudaError_t e = cudaStreamQuery(stream);
            if (e == cudaSuccess)
                break; // everything on the stream retired: the mapped write is visible now
            if (e != cudaErrorNotReady)
                return cuda_fail(e, "scan kernel");
        }
    }
    if (*word == pending) {
        SS_CUDA(cudaStreamSynchronize(stream));
        if (*word == pending) {
            t_last_error = "scan kernel retired without publishing a result";
            return SS_B200_E_CUDA;
        }
    }
    return SS_B200_OK;
}

// one 8-byte store in stream order; the slot may be device memory or a mapped pinned host word
__global__ void store_u64_kernel(unsigned long long *dst, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
}

static int needle_on_device(const ss_b200_searcher *s, int dev, const uint8_t **out)
{
    std::lock_guard<std::mutex> lk(s->mu);
    auto it = s->dev_needle.find(dev);
    if (it == s->dev_needle.end()) {
        uint8_t *d = nullptr;
        SS_CUDA(cudaMalloc(&d, s->needle.size() + 16));
        cudaError_t e = cudaMemcpy(d, s->needle.data(), s->needle.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(0); // a small pageable copy may still be in flight (see haystack_upload)
        if (e != cudaSuccess) {
            cudaFree(d);
            return cuda_fail(e, "cudaMemcpy(needle)");
        }
        it = s->dev_needle.emplace(dev, d).first;
    }
    *out = it->second;
    return SS_B200_OK;
}

// Build the kernel arguments for one scan (k >= 1, len >= k).
int ss_capi_build_args(const ss_b200_searcher *s, const void *dptr, size_t len, uint64_t base, size_t start_limit,
                       int dev, ScanArgs &a)
{
    memset(&a, 0, sizeof a);
    const size_t k = s->needle.size();
    a.hay = (const uint8_t *)dptr;
    a.n = len;
    a.base = base;
    a.k = (uint32_t)k;
    a.pos = (uint32_t)s->position;
    const uint8_t f = s->needle[0], l = s->needle[s->position];
    a.f4 = 0x01010101u * f;
    a.l4 = 0x01010101u * l;
    memcpy(a.needle_inline, s->needle.data(), k < SS_INLINE_NEEDLE_MAX ? k : SS_INLINE_NEEDLE_MAX);
    for (size_t j = 1; j < 17 && j < k; j++)
        a.needle4[j] = 0x01010101u * s->needle[j];
    if (k > SS_INLINE_NEEDLE_MAX) {
        int rc = needle_on_device(s, dev, &a.needle_g);
        if (rc != SS_B200_OK)
            return rc;
    }
    ss_host_scan_geometry(a, start_limit);
    return SS_B200_OK;
}

int ss_capi_find_async(const ss_b200_searcher *s, const void *dptr, size_t len, uint64_t base_offset,
                       size_t start_limit, void *workspace, uint64_t *d_result, void *stream, const SsStopSpec *stop)
{
    if (!s || !d_result || !workspace || (len && !dptr))
        return SS_B200_E_ARG;
    const size_t k = s->needle.size();
    if (k > 0xFFFFFFFFull)
        return SS_B200_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    // trivial outcomes are still delivered in stream order through the result slot
    if (k == 0 || len < k || start_limit == 0) {
        // N0 => true at 0 (src/x86.rs:500); n < k => false (src/x86.rs:357-359, src/lib.rs:131-133)
        // (a store kernel, not a memcpy: d_result may be a mapped host word, and the call must stay legal
        // under stream capture)
        const unsigned long long v = (k == 0) ? (unsigned long long)base_offset : SS_NONE_U64;
        store_u64_kernel<<<1, 1, 0, st>>>((unsigned long long *)d_result, v);
        ss_host_count_launch(1);
        SS_CUDA(cudaGetLastError());
        return SS_B200_OK;
    }
    SsDeviceInfo dev;
    int rc = device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    ScanArgs a;
    rc = ss_capi_build_args(s, dptr, len, base_offset, start_limit, dev.device, a);
    if (rc != SS_B200_OK)
        return rc;
    a.ws = (SsWorkspace *)workspace;
    a.out = (unsigned long long *)d_result;
    if (stop)
        ss_capi_apply_stop(a, *stop);
    SS_CUDA(ss_host_launch_scan(a, ss_capi_tuning(), dev, st));
    return SS_B200_OK;
}

void ss_capi_apply_stop(ScanArgs &a, const SsStopSpec &stop)
{
    a.stop_word = stop.stop_word;
    a.stop_seq = stop.seq;
    a.n_stop_peers = stop.n_peers < SS_MAX_PEERS ? stop.n_peers : SS_MAX_PEERS;
    for (uint32_t p = 0; p < a.n_stop_peers; p++)
        a.stop_peer[p] = stop.peers[p];
}

extern "C" int ss_b200_find_in_device_async(const ss_b200_searcher *s, const void *dptr, size_t len,
                                            uint64_t base_offset, size_t start_limit, void *workspace,
                                            uint64_t *d_result, void *stream)
{
    return ss_capi_find_async(s, dptr, len, base_offset, start_limit, workspace, d_result, stream, nullptr);
}

// Many-haystack mode: one needle against a device-resident set of haystacks in one pass.
__global__ void fill_flags_kernel(uint8_t *flags, unsigned long long n, uint8_t v)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        flags[i] = v;
}

int ss_capi_search_many(const ss_b200_searcher *s, const void *d_blob, const uint64_t *d_offsets, size_t n_haystacks,
                        size_t blob_len, uint8_t *d_flags, void *workspace, const uint32_t *d_hint, size_t n_gran,
                        void *stream)
{
    if (!s || !d_offsets || !d_flags || !workspace || (blob_len && !d_blob))
        return SS_B200_E_ARG;
    if (n_haystacks == 0)
        return SS_B200_OK;
    const size_t k = s->needle.size();
    if (k > 0xFFFFFFFFull)
        return SS_B200_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    SsDeviceInfo dev;
    int rc = device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    // k == 0: every haystack matches (N0, src/x86.rs:500); otherwise start from "not found"
    const unsigned blocks = (unsigned)((n_haystacks + 255) / 256 < (size_t)dev.sm_count * 8 ? (n_haystacks + 255) / 256
                                                                                             : (size_t)dev.sm_count * 8);
    fill_flags_kernel<<<blocks, 256, 0, st>>>(d_flags, n_haystacks, k == 0 ? 1 : 0);
    ss_host_count_launch(1);
    SS_CUDA(cudaGetLastError());
    if (k == 0 || blob_len < k)
        return SS_B200_OK;
    ScanArgs a;
    rc = ss_capi_build_args(s, d_blob, blob_len, 0, (size_t)-1, dev.device, a);
    if (rc != SS_B200_OK)
        return rc;
    a.ws = (SsWorkspace *)workspace;
    a.out = (unsigned long long *)((uint8_t *)workspace + 16); // scratch result slot, unused by callers
    a.seg_off = (const unsigned long long *)d_offsets;
    a.seg_flags = d_flags;
    a.n_seg = n_haystacks;
    a.seg_hint = d_hint;
    a.n_gran = n_gran;
    SS_CUDA(ss_host_launch_scan(a, ss_capi_tuning(), dev, st));
    return SS_B200_OK;
}

extern "C" int ss_b200_search_many_async(const ss_b200_searcher *s, const void *d_blob, const uint64_t *d_offsets,
                                         size_t n_haystacks, size_t blob_len, uint8_t *d_flags, void *workspace,
                                         void *stream)
{
    return ss_capi_search_many(s, d_blob, d_offsets, n_haystacks, blob_len, d_flags, workspace, nullptr, 0, stream);
}

// Count mode: number of occurrences (overlapping ones included) of the needle in device memory.
extern "C" int ss_b200_count_in_device_async(const ss_b200_searcher *s, const void *dptr, size_t len,
                                             size_t start_limit, void *workspace, uint64_t *d_count, void *stream)
{
    if (!s || !d_count || !workspace || (len && !dptr))
        return SS_B200_E_ARG;
    const size_t k = s->needle.size();
    if (k == 0 || k > 0xFFFFFFFFull)
        return SS_B200_E_ARG; // the empty needle has no meaningful occurrence count
    cudaStream_t st = (cudaStream_t)stream;
    SS_CUDA(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
    if (len < k || start_limit == 0)
        return SS_B200_OK;
    SsDeviceInfo dev;
    int rc = device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    ScanArgs a;
    rc = ss_capi_build_args(s, dptr, len, 0, start_limit, dev.device, a);
    if (rc != SS_B200_OK)
        return rc;
    a.ws = (SsWorkspace *)workspace;
    a.out = (unsigned long long *)((uint8_t *)workspace + 16); // scratch result slot
    a.count = (unsigned long long *)d_count;
    SS_CUDA(ss_host_launch_scan(a, ss_capi_tuning(), dev, st));
    return SS_B200_OK;
}

// One synchronous scan of device-visible memory on a given lane (its device is made current for the call).
int ss_capi_find_on_lane(SsLane *c, const ss_b200_searcher *s, const void *dptr, size_t len, size_t *offset,
                         int force_variant, int plain_device)
{
    SsDeviceGuard guard(c->device);
    if (plain_device == c->device && force_variant == 0 && ss_service_eligible(s, len)) {
        // short device-resident haystack: through the resident kernel, no launch (service.cu)
        int on = 0;
        unsigned idle_us = 0;
        ss_capi_service_tuning(&on, &idle_us);
        if (on) {
            bool used = false;
            int rc = ss_service_find(c, s, dptr, len, idle_us, offset, &used);
            if (rc != SS_B200_OK || used)
                return rc;
        }
    }
    SsDeviceInfo dev;
    int rc = device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    ScanArgs a;
    rc = ss_capi_build_args(s, dptr, len, 0, (size_t)-1, dev.device, a);
    if (rc != SS_B200_OK)
        return rc;
    a.ws = c->ws;
    a.out = (unsigned long long *)&c->slot_dev->value;
    c->slot->value = SS_RESULT_PENDING;
    SsScanTuning tuning = ss_capi_tuning();
    if (force_variant)
        tuning.variant = force_variant;
    SS_CUDA(ss_host_launch_scan(a, tuning, dev, c->stream));
    rc = ss_capi_wait_slot(&c->slot->value, SS_RESULT_PENDING, c->stream);
    if (rc != SS_B200_OK)
        return rc;
    const unsigned long long v = c->slot->value;
    *offset = (v == SS_NONE_U64) ? SS_B200_NPOS : (size_t)v;
    return SS_B200_OK;
}

// One synchronous scan of device memory through the calling thread's lane.
int ss_capi_find_device_sync(const ss_b200_searcher *s, const void *dptr, size_t len, size_t *offset,
                             int force_variant, int plain_device)
{
    const size_t k = s->needle.size();
    if (k == 0) { // DynamicAvx2Searcher::N0 => true, even for an empty haystack (src/x86.rs:470,500)
        *offset = 0;
        return SS_B200_OK;
    }
    if (len < k) { // src/x86.rs:357-359 (haystack == needle is false); k==1: src/lib.rs:131-133
        *offset = SS_B200_NPOS;
        return SS_B200_OK;
    }
    if (k > 0xFFFFFFFFull)
        return SS_B200_E_ARG;
    SsLane *c = nullptr;
    int rc = ss_capi_get_lane(&c);
    if (rc != SS_B200_OK)
        return rc;
    return ss_capi_find_on_lane(c, s, dptr, len, offset, force_variant, plain_device);
}

extern "C" int ss_b200_find_in(const ss_b200_searcher *s, const ss_b200_haystack *h, size_t *offset)
{
    if (!s || !h || !offset)
        return SS_B200_E_ARG;
    return ss_capi_find_device_sync(s, h->dptr, h->len, offset, 0, h->plain_device_memory ? h->device : -1);
}

extern "C" int ss_b200_search_in(const ss_b200_searcher *s, const ss_b200_haystack *h, uint8_t *found)
{
    if (!s || !h || !found)
        return SS_B200_E_ARG;
    size_t off = SS_B200_NPOS;
    int rc = ss_capi_find_device_sync(s, h->dptr, h->len, &off, 0, h->plain_device_memory ? h->device : -1);
    if (rc == SS_B200_OK)
        *found = (off != SS_B200_NPOS) ? 1 : 0;
    return rc;
}

// ---------------------------------------------------------------------------------------------
// generators

extern "C" int ss_b200_fill_random(void *d_dst, size_t len, uint64_t global_start, uint64_t seed, void *stream)
{
    if (len && !d_dst)
        return SS_B200_E_ARG;
    SsDeviceInfo dev;
    int rc = device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    SS_CUDA(ss_host_fill_random(d_dst, len, global_start, seed, dev.sm_count, (cudaStream_t)stream));
    return SS_B200_OK;
}

extern "C" int ss_b200_fill_tiled(void *d_dst, size_t len, uint64_t global_start, const void *d_src, size_t src_len,
                                  void *stream)
{
    if ((len && !d_dst) || !d_src || src_len == 0)
        return SS_B200_E_ARG;
    SsDeviceInfo dev;
    int rc = device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    SS_CUDA(ss_host_fill_tiled(d_dst, len, global_start, d_src, src_len, dev.sm_count, (cudaStream_t)stream));
    return SS_B200_OK;
}
