// capi_internal.h -- declarations shared by the translation units that implement the C ABI
// (capi.cu, host_engine.cu, capi_ctx.cu, capi_exchange.cu, batch.cu, hist.cu, hayset.cu).  Not installed.
#pragma once
#include "../../include/sliceslice_b200.h"
#include "ss_host.h"

#include <map>
#include <mutex>
#include <string>
#include <vector>

// errors: every CUDA failure becomes SS_B200_E_CUDA / _NOMEM plus a thread-local detail string
int ss_capi_cuda_fail(cudaError_t e, const char *what);
void ss_capi_set_error(const char *msg);
#define SS_CUDA(call)                                                                                                \
    do {                                                                                                             \
        cudaError_t e__ = (call);                                                                                    \
        if (e__ != cudaSuccess)                                                                                      \
            return ss_capi_cuda_fail(e__, #call);                                                                    \
    } while (0)

// facts about the CURRENT device (cached per device)
int ss_capi_device_info(SsDeviceInfo &out);
// snapshot of the process-wide tuning (the setters store atomically; every search reads one snapshot)
SsScanTuning ss_capi_tuning();

// process-wide knobs of the host-slice path (ss_b200_set_host_path)
struct SsHostPathTuning {
    int mode = 0;          // 0 auto, 1 DMA ring (chunked cudaMemcpyAsync + scan), 2 in place (pinned input read over PCIe)
    int chunk_mib = 0;     // 0 auto (sized from the slice), else chunk size in MiB
    int copy_threads = -1; // memcpy workers staging pageable input; -1 auto, 0 = let the driver stage
};
SsHostPathTuning ss_capi_host_tuning();

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
    bool plain_device_memory = false; // cudaMalloc-style memory (not managed, not mapped host): checked once
    int device = -1;
};

// mapped pinned result word of the synchronous calls
struct SsHostSlot {
    volatile unsigned long long value;
    volatile unsigned long long pad;
};

// Everything ONE device needs to serve synchronous calls: scan / copy streams, the self-resetting
// workspace, a mapped result slot, and (allocated on first use, sized from the slices actually searched)
// the staging ring of the host-slice path.  The calling thread's context is one lane per device it has
// used (released at thread exit or by ss_b200_thread_release); a multi-GPU ss_b200_ctx owns one lane per
// device of the context.
struct SsLane {
    int device = -1;
    cudaStream_t stream = nullptr;      // scans
    cudaStream_t copy_stream = nullptr; // host -> device chunk copies
    SsWorkspace *ws = nullptr;          // 64 bytes of device memory: workspace + scratch result slot
    SsHostSlot *slot = nullptr;         // pinned + mapped
    SsHostSlot *slot_dev = nullptr;     // device view of the same memory
    static const int NBUF = 3;
    uint8_t *dbuf[NBUF] = {nullptr, nullptr, nullptr};
    size_t dbuf_cap = 0;
    cudaEvent_t copied[NBUF] = {nullptr, nullptr, nullptr};
    cudaEvent_t scanned[NBUF] = {nullptr, nullptr, nullptr};
    uint8_t *stage[NBUF] = {nullptr, nullptr, nullptr}; // pinned staging for pageable host haystacks
    size_t stage_cap = 0;
    uint8_t *small_host = nullptr; // pinned + mapped copy of a short host slice (read in place by the kernel)
    uint8_t *small_dev = nullptr;  // device view of the same memory
    unsigned long long *chunk_results = nullptr; // pinned + mapped, one per chunk of a host-slice search
    unsigned long long *chunk_results_dev = nullptr;
    size_t chunk_results_cap = 0;
    void *service = nullptr; // resident scan kernel of the synchronous short-haystack calls (service.cu)

    SsLane() = default;
    SsLane(const SsLane &) = delete;
    SsLane &operator=(const SsLane &) = delete;
    ~SsLane() { release(); }
    int init(int dev);  // creates streams, workspace and result slot on `dev` (leaves the current device unchanged)
    void release();     // frees everything; safe at any time (errors during process teardown are ignored)
    size_t device_bytes() const { return dbuf_cap * NBUF + (ws ? 64 : 0); }
    size_t pinned_bytes() const;
};
int ss_capi_get_lane(SsLane **out); // the calling thread's lane for the current device

// RAII: make `dev` current, restore the previous device on scope exit
struct SsDeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit SsDeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev)
            switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~SsDeviceGuard()
    {
        if (switched)
            cudaSetDevice(prev);
    }
};

// one synchronous scan of device-visible memory through the calling thread's lane
// (force_variant: 0 = the process-wide tuning, 1 / 2 = that scan variant for this call)
// plain_device: ordinal of the device whose plain (cudaMalloc-style) memory dptr is, or -1 when that is
// not known -- the resident service kernel only serves memory of the lane's own device
int ss_capi_find_device_sync(const ss_b200_searcher *s, const void *dptr, size_t len, size_t *offset,
                             int force_variant, int plain_device = -1);
int ss_capi_find_on_lane(SsLane *c, const ss_b200_searcher *s, const void *dptr, size_t len, size_t *offset,
                         int force_variant, int plain_device = -1);
// service.cu: the resident kernel of the synchronous short-haystack calls
bool ss_service_eligible(const ss_b200_searcher *s, size_t len);
int ss_service_find(SsLane *lane, const ss_b200_searcher *s, const void *dptr, size_t len, unsigned idle_us,
                    size_t *offset, bool *used, bool mapped_host = false);
void ss_service_release(void *service);
void ss_capi_service_tuning(int *on, unsigned *idle_us);
// spin on a mapped result word until the kernel behind it has written it (or its stream reports an error)
int ss_capi_wait_slot(volatile unsigned long long *word, unsigned long long pending, cudaStream_t stream);
// host slices up to this size are searched in place from a mapped pinned copy (host_engine.cu)
#define SS_SMALL_HOST_MAX (32u << 10)

// cross-GPU early exit of a sharded first-match search (ScanArgs::stop_word): this shard polls
// `stop_word`, and on its first match stores `seq` into the stop words of the shards to its right
struct SsStopSpec {
    const unsigned long long *stop_word = nullptr;
    unsigned long long seq = 0;
    unsigned long long *peers[SS_MAX_PEERS] = {};
    uint32_t n_peers = 0;
};
void ss_capi_apply_stop(ScanArgs &a, const SsStopSpec &stop);
// ss_b200_find_in_device_async with an optional stop spec
int ss_capi_find_async(const ss_b200_searcher *s, const void *dptr, size_t len, uint64_t base_offset,
                       size_t start_limit, void *workspace, uint64_t *d_result, void *stream, const SsStopSpec *stop);

// kernel arguments for one scan of (dptr, len) with this searcher (k >= 1, len >= k)
int ss_capi_build_args(const ss_b200_searcher *s, const void *dptr, size_t len, uint64_t base, size_t start_limit,
                       int dev, ScanArgs &a);

// The host-slice engine (host_engine.cu): one host slice striped in chunks over `n_lanes` devices
// (chunk i -> lane i % n_lanes), copy/scan overlapped per lane, bounded run-ahead so that a match stops
// the feeding (src/lib.rs:242-244), leftmost offset over all chunks.  n_lanes == 1 is ss_b200_find_in_host.
struct SsHostStats {
    unsigned long long h2d_bytes = 0;  // bytes handed to cudaMemcpyAsync (0 for the in-place mode)
    unsigned long long chunks = 0;     // chunks submitted
    int mode = 0;                      // 1 DMA ring, 2 in place, 3 short slice through the mapped copy
    int staged = 0;                    // pageable input staged through the pinned ring
    unsigned long long chunk_bytes = 0;
};
int ss_host_engine_find(SsLane *const *lanes, int n_lanes, const ss_b200_searcher *s, const uint8_t *host, size_t len,
                        size_t *offset, SsHostStats *stats); // (pageable input: staged through ONE lane's pinned ring)

// many-haystack scan shared by ss_b200_search_many_async (no hints) and ss_b200_hayset_search_async
int ss_capi_search_many(const ss_b200_searcher *s, const void *d_blob, const uint64_t *d_offsets, size_t n_haystacks,
                        size_t blob_len, uint8_t *d_flags, void *workspace, const uint32_t *d_hint, size_t n_gran,
                        void *stream);
