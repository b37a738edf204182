// capi.cu -- the extern "C" boundary declared in include/sliceslice_b200.h.
//
// Host runtime around the sm_100a kernels: opaque searcher / haystack handles,
// thread-local stream + self-resetting workspace + mapped result slot for the
// synchronous calls, the chunked host->device streaming path, and the
// stream-ordered entry used for roofline measurement and the multi-GPU shards.
// No CPU search path exists in this library: without a device every search
// returns SS_B200_E_CUDA.
#include "../../include/sliceslice_b200.h"
#include "ss_host.h"

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
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
#define SS_CUDA(call)                                                                                                \
    do {                                                                                                             \
        cudaError_t e__ = (call);                                                                                    \
        if (e__ != cudaSuccess)                                                                                      \
            return cuda_fail(e__, #call);                                                                            \
    } while (0)

int ss_capi_cuda_fail(cudaError_t e, const char *what) { return cuda_fail(e, what); }

extern "C" const char *ss_b200_strerror(int status)
{
    switch (status) {
    case SS_B200_OK: return "ok";
    case SS_B200_E_POSITION: return "position is not a valid index for the needle";
    case SS_B200_E_EMPTY_NEEDLE: return "needle is empty";
    case SS_B200_E_ARG: return "invalid argument";
    case SS_B200_E_CUDA: return "CUDA error or no usable device";
    case SS_B200_E_NOMEM: return "out of memory";
    default: return "unknown status";
    }
}
extern "C" const char *ss_b200_last_error(void) { return t_last_error.c_str(); }
extern "C" int ss_b200_abi_version(void) { return SS_B200_ABI_VERSION; }

// ---------------------------------------------------------------------------------------------
// process-wide tuning + per-device facts

static SsScanTuning g_tuning;
static std::mutex g_dev_mutex;
static std::map<int, SsDeviceInfo> g_devs;

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
        const char *v = getenv("SS_B200_LONG_VARIANT");
        d.auto_long_variant = (v && atoi(v) == 1) ? 1 : 2;
        it = g_devs.emplace(dev, d).first;
    }
    out = it->second;
    return SS_B200_OK;
}

int ss_capi_device_info(SsDeviceInfo &out) { return device_info(out); }

extern "C" int ss_b200_set_scan_variant(int variant)
{
    if (variant < 0 || variant > 2)
        return SS_B200_E_ARG;
    g_tuning.variant = variant;
    return SS_B200_OK;
}
extern "C" int ss_b200_set_scan_tuning(int ctas_per_sm, int unroll, int tile_kib, int stages)
{
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
extern "C" uint64_t ss_b200_launch_count(void) { return ss_host_launch_count(); }

// ---------------------------------------------------------------------------------------------
// handles

struct ss_b200_searcher {
    std::vector<uint8_t> needle;
    size_t position = 0;
    bool strict = false; // Avx2Searcher flavour: one-byte needles take the two-anchor path too
    // device copies of long needles, one per device that has searched with this handle
    mutable std::mutex mu;
    mutable std::map<int, uint8_t *> dev_needle;
};

struct ss_b200_haystack {
    const uint8_t *dptr = nullptr;
    size_t len = 0;
    bool owned = false;
    int device = -1;
};

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
// per-thread, per-device context for the synchronous calls

struct HostSlot {
    volatile unsigned long long value;
    volatile unsigned long long pad;
};

struct ThreadCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    SsWorkspace *ws = nullptr;
    HostSlot *slot = nullptr;   // pinned + mapped
    HostSlot *slot_dev = nullptr; // device view of the same memory
    // host-path staging (lazily sized)
    static const int NBUF = 3;
    uint8_t *dbuf[NBUF] = {nullptr, nullptr, nullptr};
    size_t dbuf_cap = 0;
    cudaEvent_t copied[NBUF] = {nullptr, nullptr, nullptr};
    cudaEvent_t scanned[NBUF] = {nullptr, nullptr, nullptr};
    uint8_t *stage[NBUF] = {nullptr, nullptr, nullptr}; // pinned staging for pageable host haystacks
    size_t stage_cap = 0;
    unsigned long long *chunk_results = nullptr; // pinned + mapped, one per in-flight chunk
    unsigned long long *chunk_results_dev = nullptr;
    size_t chunk_results_cap = 0;
};

static thread_local std::map<int, ThreadCtx> t_ctx;

static int get_ctx(ThreadCtx **out)
{
    int dev = -1;
    SS_CUDA(cudaGetDevice(&dev));
    ThreadCtx &c = t_ctx[dev];
    if (c.device < 0) {
        SS_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        SS_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        SS_CUDA(cudaMalloc(&c.ws, sizeof(SsWorkspace)));
        SS_CUDA(cudaMemset(c.ws, 0, sizeof(SsWorkspace)));
        SS_CUDA(cudaHostAlloc((void **)&c.slot, sizeof(HostSlot), cudaHostAllocMapped));
        c.slot->value = 0;
        c.slot->pad = 0;
        SS_CUDA(cudaHostGetDevicePointer((void **)&c.slot_dev, (void *)c.slot, 0));
        SS_CUDA(cudaDeviceSynchronize());
        c.device = dev;
    }
    *out = &c;
    return SS_B200_OK;
}

static int needle_on_device(const ss_b200_searcher *s, int dev, const uint8_t **out)
{
    std::lock_guard<std::mutex> lk(s->mu);
    auto it = s->dev_needle.find(dev);
    if (it == s->dev_needle.end()) {
        uint8_t *d = nullptr;
        SS_CUDA(cudaMalloc(&d, s->needle.size() + 16));
        cudaError_t e = cudaMemcpy(d, s->needle.data(), s->needle.size(), cudaMemcpyHostToDevice);
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
static int build_args(const ss_b200_searcher *s, const void *dptr, size_t len, uint64_t base, size_t start_limit,
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
    if (k > SS_INLINE_NEEDLE_MAX) {
        int rc = needle_on_device(s, dev, &a.needle_g);
        if (rc != SS_B200_OK)
            return rc;
    }
    ss_host_scan_geometry(a, start_limit);
    return SS_B200_OK;
}

extern "C" int ss_b200_find_in_device_async(const ss_b200_searcher *s, const void *dptr, size_t len,
                                            uint64_t base_offset, size_t start_limit, void *workspace,
                                            uint64_t *d_result, void *stream)
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
        const unsigned long long v = (k == 0) ? (unsigned long long)base_offset : SS_NONE_U64;
        SS_CUDA(cudaMemcpyAsync(d_result, &v, sizeof v, cudaMemcpyHostToDevice, st));
        return SS_B200_OK;
    }
    SsDeviceInfo dev;
    int rc = device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    ScanArgs a;
    rc = build_args(s, dptr, len, base_offset, start_limit, dev.device, a);
    if (rc != SS_B200_OK)
        return rc;
    a.ws = (SsWorkspace *)workspace;
    a.out = (unsigned long long *)d_result;
    SS_CUDA(ss_host_launch_scan(a, g_tuning, dev, st));
    return SS_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// Peer mailbox exchange: the MIN over ranks of the first offsets without a collective call.
// Mailbox layout (per rank, plain cudaMalloc memory shared through CUDA IPC):
//   u64 slot[SS_MAILBOX_DEPTH][world]; slot[seq % 4][r] = result of rank r for search `seq`
// Search `seq` on rank r: the scan's last CTA stores its result into slot[seq%4][r] of EVERY rank's
// mailbox (scan_finish); mailbox_min_kernel, enqueued right behind the scan, waits until the `world`
// slots of its own mailbox are filled, writes their minimum and empties them again.  A rank can run
// at most one search ahead of the slowest rank (its gather needs everybody's scan), so four slot
// rows never collide.

__global__ void mailbox_fill_kernel(unsigned long long *mb, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        mb[i] = SS_MAILBOX_EMPTY;
}

__global__ void mailbox_post_kernel(ScanArgs a, unsigned long long value)
{
    if (threadIdx.x < a.n_peers)
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a.peer_slot[threadIdx.x]), "l"(value) : "memory");
}

__global__ void mailbox_min_kernel(unsigned long long *row, int world, unsigned long long *out)
{
    const int lane = threadIdx.x;
    unsigned long long v = SS_NONE_U64;
    if (lane < world) {
        unsigned ns = 32;
        for (;;) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(row + lane) : "memory");
            if (v != SS_MAILBOX_EMPTY)
                break;
            __nanosleep(ns);
            if (ns < 1024)
                ns *= 2;
        }
        row[lane] = SS_MAILBOX_EMPTY; // ready for search seq + 4
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long w = __shfl_xor_sync(0xFFFFFFFFu, v, o);
        v = w < v ? w : v;
    }
    if (lane == 0)
        *out = v;
}

extern "C" int ss_b200_mailbox_create(int world, void **d_mailbox)
{
    if (!d_mailbox || world < 1 || world > SS_MAX_PEERS)
        return SS_B200_E_ARG;
    unsigned long long *mb = nullptr;
    const int n = SS_MAILBOX_DEPTH * world;
    SS_CUDA(cudaMalloc((void **)&mb, (size_t)n * 8));
    mailbox_fill_kernel<<<1, 64>>>(mb, n);
    ss_host_count_launch(1);
    SS_CUDA(cudaGetLastError());
    SS_CUDA(cudaDeviceSynchronize());
    *d_mailbox = mb;
    return SS_B200_OK;
}
extern "C" int ss_b200_mailbox_free(void *d_mailbox)
{
    if (d_mailbox)
        SS_CUDA(cudaFree(d_mailbox));
    return SS_B200_OK;
}
extern "C" int ss_b200_ipc_export(const void *dptr, uint8_t handle_out[64])
{
    if (!dptr || !handle_out)
        return SS_B200_E_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    cudaIpcMemHandle_t h;
    SS_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(dptr)));
    memcpy(handle_out, &h, 64);
    return SS_B200_OK;
}
extern "C" int ss_b200_ipc_open(const uint8_t handle[64], void **dptr_out)
{
    if (!handle || !dptr_out)
        return SS_B200_E_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    SS_CUDA(cudaIpcOpenMemHandle(dptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return SS_B200_OK;
}
extern "C" int ss_b200_ipc_close(void *dptr)
{
    if (dptr)
        SS_CUDA(cudaIpcCloseMemHandle(dptr));
    return SS_B200_OK;
}

extern "C" int ss_b200_find_in_device_exchange_async(const ss_b200_searcher *s, const void *dptr, size_t len,
                                                     uint64_t base_offset, size_t start_limit, void *workspace,
                                                     void *const *mailboxes, int world, int rank, uint64_t seq,
                                                     uint64_t *d_result, void *stream)
{
    if (!s || !d_result || !workspace || !mailboxes || (len && !dptr) || world < 1 || world > SS_MAX_PEERS ||
        rank < 0 || rank >= world)
        return SS_B200_E_ARG;
    const size_t k = s->needle.size();
    if (k > 0xFFFFFFFFull)
        return SS_B200_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    SsDeviceInfo dev;
    int rc = device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    const size_t row = (size_t)(seq % SS_MAILBOX_DEPTH) * (size_t)world;
    ScanArgs a;
    if (k == 0 || len < k || start_limit == 0) {
        // trivial local outcome (N0 => found at base; n < k => none): still posted to every rank
        memset(&a, 0, sizeof a);
        a.n_peers = (uint32_t)world;
        for (int p = 0; p < world; p++)
            a.peer_slot[p] = (unsigned long long *)mailboxes[p] + row + rank;
        mailbox_post_kernel<<<1, 32, 0, st>>>(a, k == 0 ? (unsigned long long)base_offset : SS_NONE_U64);
        ss_host_count_launch(1);
        SS_CUDA(cudaGetLastError());
    } else {
        rc = build_args(s, dptr, len, base_offset, start_limit, dev.device, a);
        if (rc != SS_B200_OK)
            return rc;
        a.ws = (SsWorkspace *)workspace;
        a.out = (unsigned long long *)((uint8_t *)workspace + 16); // local copy of this rank's own result
            a.n_peers = (uint32_t)world;
        for (int p = 0; p < world; p++)
            a.peer_slot[p] = (unsigned long long *)mailboxes[p] + row + rank;
        SS_CUDA(ss_host_launch_scan(a, g_tuning, dev, st));
    }
    mailbox_min_kernel<<<1, 32, 0, st>>>((unsigned long long *)mailboxes[rank] + row, world,
                                         (unsigned long long *)d_result);
    ss_host_count_launch(1);
    SS_CUDA(cudaGetLastError());
    return SS_B200_OK;
}

// Many-haystack mode: one needle against a device-resident set of haystacks in one pass.
__global__ void fill_flags_kernel(uint8_t *flags, unsigned long long n, uint8_t v)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        flags[i] = v;
}

extern "C" int ss_b200_search_many_async(const ss_b200_searcher *s, const void *d_blob, const uint64_t *d_offsets,
                                         size_t n_haystacks, size_t blob_len, uint8_t *d_flags, void *workspace,
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
    rc = build_args(s, d_blob, blob_len, 0, (size_t)-1, dev.device, a);
    if (rc != SS_B200_OK)
        return rc;
    a.ws = (SsWorkspace *)workspace;
    a.out = (unsigned long long *)((uint8_t *)workspace + 16); // scratch result slot, unused by callers
    a.seg_off = (const unsigned long long *)d_offsets;
    a.seg_flags = d_flags;
    a.n_seg = n_haystacks;
    SS_CUDA(ss_host_launch_scan(a, g_tuning, dev, st));
    return SS_B200_OK;
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
    rc = build_args(s, dptr, len, 0, start_limit, dev.device, a);
    if (rc != SS_B200_OK)
        return rc;
    a.ws = (SsWorkspace *)workspace;
    a.out = (unsigned long long *)((uint8_t *)workspace + 16); // scratch result slot
    a.count = (unsigned long long *)d_count;
    SS_CUDA(ss_host_launch_scan(a, g_tuning, dev, st));
    return SS_B200_OK;
}

// One synchronous scan of device memory through the thread's context.
static int find_device_sync(const ss_b200_searcher *s, const void *dptr, size_t len, size_t *offset)
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
    SsDeviceInfo dev;
    int rc = device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    ThreadCtx *c = nullptr;
    rc = get_ctx(&c);
    if (rc != SS_B200_OK)
        return rc;
    ScanArgs a;
    rc = build_args(s, dptr, len, 0, (size_t)-1, dev.device, a);
    if (rc != SS_B200_OK)
        return rc;
    a.ws = c->ws;
    a.out = (unsigned long long *)&c->slot_dev->value;
    c->slot->value = SS_RESULT_PENDING;
    SS_CUDA(ss_host_launch_scan(a, g_tuning, dev, c->stream));
    // spin on the mapped result word; fall back to the stream status every so often
    unsigned spins = 0;
    while (c->slot->value == SS_RESULT_PENDING) {
        if ((++spins & 0x3FF) == 0) {
            cudaError_t e = cudaStreamQuery(c->stream);
            if (e == cudaSuccess)
                break; // kernel retired: the mapped write is visible now
            if (e != cudaErrorNotReady)
                return cuda_fail(e, "scan kernel");
        }
    }
    if (c->slot->value == SS_RESULT_PENDING) {
        SS_CUDA(cudaStreamSynchronize(c->stream));
        if (c->slot->value == SS_RESULT_PENDING) {
            t_last_error = "scan kernel retired without publishing a result";
            return SS_B200_E_CUDA;
        }
    }
    const unsigned long long v = c->slot->value;
    *offset = (v == SS_NONE_U64) ? SS_B200_NPOS : (size_t)v;
    return SS_B200_OK;
}

extern "C" int ss_b200_find_in(const ss_b200_searcher *s, const ss_b200_haystack *h, size_t *offset)
{
    if (!s || !h || !offset)
        return SS_B200_E_ARG;
    return find_device_sync(s, h->dptr, h->len, offset);
}

extern "C" int ss_b200_search_in(const ss_b200_searcher *s, const ss_b200_haystack *h, uint8_t *found)
{
    if (!s || !h || !found)
        return SS_B200_E_ARG;
    size_t off = SS_B200_NPOS;
    int rc = find_device_sync(s, h->dptr, h->len, &off);
    if (rc == SS_B200_OK)
        *found = (off != SS_B200_NPOS) ? 1 : 0;
    return rc;
}

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
    int rc = device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    ThreadCtx *c = nullptr;
    rc = get_ctx(&c);
    if (rc != SS_B200_OK)
        return rc;

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
        for (int b = 0; b < ThreadCtx::NBUF; b++) {
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
        for (int b = 0; b < ThreadCtx::NBUF; b++) {
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
    rc = build_args(s, c->dbuf[0], k, 0, (size_t)-1, dev.device, proto); // needle fields; geometry redone per chunk
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
        const int b = (int)(i % ThreadCtx::NBUF);
        const size_t off = i * chunk;
        size_t bytes = chunk + halo;
        if (off + bytes > len)
            bytes = len - off;
        if (i >= (size_t)ThreadCtx::NBUF)
            SS_CUDA(cudaStreamWaitEvent(c->copy_stream, c->scanned[b], 0));
        const uint8_t *src = host + off;
        if (staged) {
            // the pinned buffer is free once its previous DMA has finished; fill it in parallel while the
            // DMA engine is still busy with the previous chunk
            if (i >= (size_t)ThreadCtx::NBUF)
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
            SS_CUDA(ss_host_launch_scan(a, g_tuning, dev, c->stream));
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
