// capi_internal.h -- declarations shared by the translation units that implement the C ABI
// (capi.cu, capi_host_path.cu, capi_exchange.cu, batch.cu).  Not installed.
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

int ss_capi_device_info(SsDeviceInfo &out);
const SsScanTuning &ss_capi_tuning();

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

// mapped pinned result word of the synchronous calls
struct SsHostSlot {
    volatile unsigned long long value;
    volatile unsigned long long pad;
};

// per-thread, per-device context of the synchronous calls: stream, self-resetting workspace, mapped
// result slot, and the staging buffers of the host-slice path
struct SsThreadCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    SsWorkspace *ws = nullptr;
    SsHostSlot *slot = nullptr;     // pinned + mapped
    SsHostSlot *slot_dev = nullptr; // device view of the same memory
    static const int NBUF = 3;
    uint8_t *dbuf[NBUF] = {nullptr, nullptr, nullptr};
    size_t dbuf_cap = 0;
    cudaEvent_t copied[NBUF] = {nullptr, nullptr, nullptr};
    cudaEvent_t scanned[NBUF] = {nullptr, nullptr, nullptr};
    uint8_t *stage[NBUF] = {nullptr, nullptr, nullptr}; // pinned staging for pageable host haystacks
    size_t stage_cap = 0;
    uint8_t *small_host = nullptr; // pinned + mapped copy of a short host slice (read in place by the kernel)
    uint8_t *small_dev = nullptr;  // device view of the same memory
    unsigned long long *chunk_results = nullptr; // pinned + mapped, one per in-flight chunk
    unsigned long long *chunk_results_dev = nullptr;
    size_t chunk_results_cap = 0;
};
int ss_capi_get_ctx(SsThreadCtx **out);
// one synchronous scan of device-visible memory through the calling thread's context
// (force_variant: 0 = the process-wide tuning, 1 / 2 = that scan variant for this call)
int ss_capi_find_device_sync(const ss_b200_searcher *s, const void *dptr, size_t len, size_t *offset,
                             int force_variant);
// host slices up to this size are searched in place from a mapped pinned copy (capi_host_path.cu)
#define SS_SMALL_HOST_MAX (32u << 10)

// kernel arguments for one scan of (dptr, len) with this searcher (k >= 1, len >= k)
int ss_capi_build_args(const ss_b200_searcher *s, const void *dptr, size_t len, uint64_t base, size_t start_limit,
                       int dev, ScanArgs &a);

// many-haystack scan shared by ss_b200_search_many_async (no hints) and ss_b200_hayset_search_async
int ss_capi_search_many(const ss_b200_searcher *s, const void *d_blob, const uint64_t *d_offsets, size_t n_haystacks,
                        size_t blob_len, uint8_t *d_flags, void *workspace, const uint32_t *d_hint, size_t n_gran,
                        void *stream);
