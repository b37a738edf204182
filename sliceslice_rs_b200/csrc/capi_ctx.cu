// capi_ctx.cu -- multi-GPU behind the C ABI: one process, every GPU of the box.
//
// The reference has no multi-device notion (single-threaded CPU code); its FFI precedent binds plain C
// functions from one process (bench/sse4-strstr/build.rs:8-23, bench/sse4-strstr/src/lib.rs:4-15,
// wrapper.h:7).  A Rust host following that pattern reaches all GPUs through the entry points below:
//
//   ss_b200_ctx_create          one lane (streams, workspace, mapped result slot, staging ring) per device,
//                               peer access between all pairs
//   ss_b200_sharded_*           one haystack as contiguous shards of start positions, shard d on device d,
//                               each with a right halo
//   ss_b200_search_sharded      every device scans its shard; the first offsets are MIN-reduced by the
//                               exchange of the context: mapped host words (default), peer mailboxes
//                               fused into the scan epilogue (capi_exchange.cu), or ncclAllReduce(min)
//   ss_b200_find_in_host_multi  search_in(&[u8]) with ONE host slice striped over all PCIe links
//                               (host_engine.cu)
//   ss_b200_ctx_hayset_*        many-haystack mode: the set partitioned over the devices by bytes, one
//                               flag per haystack; every haystack lives on one device, so the OR over
//                               devices is a gather of disjoint slices
// NCCL is loaded with dlopen on demand (libnccl.so.2: the copy already in the process when the host is
// PyTorch, the system one otherwise); nothing links against it, and a missing library is SS_B200_E_NCCL.
#include "capi_internal.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <memory>
#include <new>

// entry points of the other translation units used here
extern "C" int ss_b200_find_in_device_async(const ss_b200_searcher *, const void *, size_t, uint64_t, size_t, void *,
                                            uint64_t *, void *);
extern "C" int ss_b200_find_in_device_exchange_async(const ss_b200_searcher *, const void *, size_t, uint64_t, size_t,
                                                     void *, void *const *, int, int, uint64_t, uint64_t *, void *);
extern "C" int ss_b200_mailbox_create(int, void **);
extern "C" int ss_b200_mailbox_free(void *);

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    bool load(std::string &err)
    {
        if (lib)
            return true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib)
                break;
        }
        if (!lib) {
            err = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
            return false;
        }
        bool ok = true;
        auto sym = [&](const char *n) {
            void *p = dlsym(lib, n);
            if (!p) {
                ok = false;
                err = std::string("libnccl lacks ") + n;
            }
            return p;
        };
        GetVersion = (decltype(GetVersion))sym("ncclGetVersion");
        CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        if (!ok) {
            dlclose(lib);
            lib = nullptr;
        }
        return ok;
    }
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

} // namespace

struct ss_b200_ctx {
    std::vector<int> devices;
    std::vector<std::unique_ptr<SsLane>> lanes;
    std::vector<SsLane *> lane_ptrs;
    std::mutex mu; // a context is single-threaded by contract; the mutex makes misuse safe, not fast
    int exchange = SS_B200_EXCHANGE_HOST;
    bool peer_access = false;                    // every pair of devices can address each other's memory
    std::vector<void *> mailbox;                 // peer exchange: one mailbox per device
    std::vector<unsigned long long *> red;       // nccl exchange: one device word per device
    std::vector<ncclComm_t> comms;
    uint64_t seq = 0;
    SsHostStats last_host;
};

struct ss_b200_sharded {
    struct Shard {
        const uint8_t *dptr = nullptr;
        size_t start = 0, owned = 0, span = 0;
        bool own_mem = false;
    };
    const ss_b200_ctx *ctx = nullptr;
    size_t total = 0;
    std::vector<Shard> shards;
};

struct ss_b200_ctx_hayset {
    struct Part {
        size_t lo = 0, hi = 0; // haystack index range held by this device
        uint8_t *blob = nullptr;
        uint64_t *offsets = nullptr;
        uint8_t *flags = nullptr;
        void *workspace = nullptr;
        ss_b200_hayset *set = nullptr;
        size_t blob_len = 0;
    };
    const ss_b200_ctx *ctx = nullptr; // identity only: the set may outlive its context (freeing uses `devices`)
    std::vector<int> devices;
    size_t n = 0;
    std::vector<Part> parts;
};

static int nccl_fail(ncclResult_t r, const char *what)
{
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error");
    ss_capi_set_error(buf);
    return SS_B200_E_NCCL;
}
#define SS_NCCL(call)                                                                                                \
    do {                                                                                                             \
        ncclResult_t r__ = (call);                                                                                   \
        if (r__ != ncclSuccess)                                                                                      \
            return nccl_fail(r__, #call);                                                                            \
    } while (0)

// ---------------------------------------------------------------------------------------------
// context

extern "C" int ss_b200_ctx_create(int ndev, const int *devices, ss_b200_ctx **out)
{
    if (!out)
        return SS_B200_E_ARG;
    *out = nullptr;
    int visible = 0;
    SS_CUDA(cudaGetDeviceCount(&visible));
    if (ndev <= 0)
        ndev = visible;
    if (ndev < 1 || ndev > SS_MAX_PEERS || (devices == nullptr && ndev > visible))
        return SS_B200_E_ARG;
    std::unique_ptr<ss_b200_ctx> c(new (std::nothrow) ss_b200_ctx());
    if (!c)
        return SS_B200_E_NOMEM;
    for (int i = 0; i < ndev; i++) {
        const int d = devices ? devices[i] : i;
        if (d < 0 || d >= visible)
            return SS_B200_E_ARG;
        for (int j = 0; j < i; j++)
            if (c->devices[j] == d)
                return SS_B200_E_ARG;
        c->devices.push_back(d);
    }
    int prev = -1;
    cudaGetDevice(&prev);
    for (int i = 0; i < ndev; i++) {
        c->lanes.emplace_back(new SsLane());
        int rc = c->lanes.back()->init(c->devices[i]);
        if (rc != SS_B200_OK)
            return rc;
        c->lane_ptrs.push_back(c->lanes.back().get());
    }
    // peer access between all pairs (the fused exchange stores straight into peer HBM)
    bool all = true;
    for (int i = 0; i < ndev && all; i++)
        for (int j = 0; j < ndev && all; j++) {
            if (i == j)
                continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, c->devices[i], c->devices[j]) != cudaSuccess || !can)
                all = false;
        }
    if (all && ndev > 1) {
        for (int i = 0; i < ndev; i++) {
            cudaSetDevice(c->devices[i]);
            for (int j = 0; j < ndev; j++) {
                if (i == j)
                    continue;
                cudaError_t e = cudaDeviceEnablePeerAccess(c->devices[j], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled)
                    cudaGetLastError();
                else if (e != cudaSuccess) {
                    cudaGetLastError();
                    all = false;
                }
            }
        }
    }
    c->peer_access = all || ndev == 1;
    if (prev >= 0)
        cudaSetDevice(prev);
    *out = c.release();
    return SS_B200_OK;
}

static void ctx_drop_exchange(ss_b200_ctx *c)
{
    for (size_t i = 0; i < c->mailbox.size(); i++) {
        SsDeviceGuard g(c->devices[i]);
        ss_b200_mailbox_free(c->mailbox[i]);
    }
    c->mailbox.clear();
    for (size_t i = 0; i < c->comms.size(); i++)
        if (c->comms[i] && g_nccl.CommDestroy)
            g_nccl.CommDestroy(c->comms[i]);
    c->comms.clear();
    for (size_t i = 0; i < c->red.size(); i++) {
        SsDeviceGuard g(c->devices[i]);
        cudaFree(c->red[i]);
    }
    c->red.clear();
}

extern "C" void ss_b200_ctx_free(ss_b200_ctx *c)
{
    if (!c)
        return;
    ctx_drop_exchange(c);
    delete c; // lanes release their CUDA resources
}

extern "C" int ss_b200_ctx_device_count(const ss_b200_ctx *c) { return c ? (int)c->devices.size() : 0; }
extern "C" int ss_b200_ctx_device(const ss_b200_ctx *c, int i)
{
    return (c && i >= 0 && i < (int)c->devices.size()) ? c->devices[i] : -1;
}

extern "C" int ss_b200_ctx_set_exchange(ss_b200_ctx *c, int kind)
{
    if (!c || kind < SS_B200_EXCHANGE_HOST || kind > SS_B200_EXCHANGE_NCCL)
        return SS_B200_E_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    const int n = (int)c->devices.size();
    if (kind == SS_B200_EXCHANGE_PEER && c->mailbox.empty()) {
        if (!c->peer_access) {
            ss_capi_set_error("peer exchange needs peer access between every pair of devices of the context");
            return SS_B200_E_CUDA;
        }
        for (int i = 0; i < n; i++) {
            SsDeviceGuard g(c->devices[i]);
            void *mb = nullptr;
            int rc = ss_b200_mailbox_create(n, &mb);
            if (rc != SS_B200_OK) {
                ctx_drop_exchange(c);
                return rc;
            }
            c->mailbox.push_back(mb);
        }
    }
    if (kind == SS_B200_EXCHANGE_NCCL && c->comms.empty()) {
        std::lock_guard<std::mutex> lk2(g_nccl_mu);
        std::string err;
        if (!g_nccl.load(err)) {
            ss_capi_set_error(err.c_str());
            return SS_B200_E_NCCL;
        }
        c->comms.assign(n, nullptr);
        ncclResult_t r = g_nccl.CommInitAll(c->comms.data(), n, c->devices.data());
        if (r != ncclSuccess) {
            c->comms.clear();
            return nccl_fail(r, "ncclCommInitAll");
        }
        for (int i = 0; i < n; i++) {
            SsDeviceGuard g(c->devices[i]);
            unsigned long long *p = nullptr;
            cudaError_t e = cudaMalloc((void **)&p, 16);
            if (e != cudaSuccess) {
                ctx_drop_exchange(c);
                return ss_capi_cuda_fail(e, "cudaMalloc(nccl word)");
            }
            c->red.push_back(p);
        }
    }
    c->exchange = kind;
    return SS_B200_OK;
}

extern "C" int ss_b200_ctx_nccl_version(int *version)
{
    if (!version)
        return SS_B200_E_ARG;
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    std::string err;
    if (!g_nccl.load(err)) {
        ss_capi_set_error(err.c_str());
        return SS_B200_E_NCCL;
    }
    SS_NCCL(g_nccl.GetVersion(version));
    return SS_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// sharded haystack

extern "C" void ss_b200_sharded_free(ss_b200_sharded *sh)
{
    if (!sh)
        return;
    for (auto &s : sh->shards)
        if (s.own_mem && s.dptr)
            cudaFree((void *)s.dptr);
    delete sh;
}

extern "C" size_t ss_b200_sharded_len(const ss_b200_sharded *sh) { return sh ? sh->total : 0; }

extern "C" int ss_b200_sharded_shard(const ss_b200_sharded *sh, int i, const void **dptr, size_t *start, size_t *owned,
                                     size_t *span)
{
    if (!sh || i < 0 || i >= (int)sh->shards.size())
        return SS_B200_E_ARG;
    const auto &s = sh->shards[i];
    if (dptr)
        *dptr = s.dptr;
    if (start)
        *start = s.start;
    if (owned)
        *owned = s.owned;
    if (span)
        *span = s.span;
    return SS_B200_OK;
}

// Contiguous shards of start positions: shard d owns [d * per, (d+1) * per) with per = ceil(len / ndev)
// rounded up to 16 bytes, and holds `halo` more bytes so that a needle of up to halo + 1 bytes that
// starts in the shard can be verified without the neighbour (right halo only).
extern "C" int ss_b200_sharded_upload(const ss_b200_ctx *c, const uint8_t *host, size_t len, size_t halo,
                                      ss_b200_sharded **out)
{
    if (!c || !out || (len && !host))
        return SS_B200_E_ARG;
    *out = nullptr;
    std::unique_ptr<ss_b200_sharded, void (*)(ss_b200_sharded *)> sh(new (std::nothrow) ss_b200_sharded(),
                                                                    ss_b200_sharded_free);
    if (!sh)
        return SS_B200_E_NOMEM;
    const int n = (int)c->devices.size();
    sh->ctx = c;
    sh->total = len;
    size_t per = (len + n - 1) / n;
    per = (per + 15) & ~(size_t)15;
    sh->shards.resize(n);
    for (int d = 0; d < n; d++) {
        auto &s = sh->shards[d];
        s.start = (size_t)d * per < len ? (size_t)d * per : len;
        s.owned = len - s.start < per ? len - s.start : per;
        s.span = len - s.start < s.owned + halo ? len - s.start : s.owned + halo;
        if (s.span == 0)
            continue;
        SsDeviceGuard g(c->devices[d]);
        uint8_t *p = nullptr;
        const size_t alloc = ((s.span + 15) & ~(size_t)15) + 16;
        SS_CUDA(cudaMalloc(&p, alloc));
        s.dptr = p;
        s.own_mem = true;
        SsLane *lane = c->lane_ptrs[d];
        SS_CUDA(cudaMemsetAsync(p + s.span, 0, alloc - s.span, lane->copy_stream));
        SS_CUDA(cudaMemcpyAsync(p, host + s.start, s.span, cudaMemcpyHostToDevice, lane->copy_stream));
    }
    for (int d = 0; d < n; d++)
        SS_CUDA(cudaStreamSynchronize(c->lane_ptrs[d]->copy_stream));
    *out = sh.release();
    return SS_B200_OK;
}

// Borrow shards that already live in device memory (e.g. generated there): shard d on device d of the
// context holds spans[d] bytes starting at global byte sum(owned[0..d)) and owns the first owned[d]
// start positions of them.
extern "C" int ss_b200_sharded_from_device(const ss_b200_ctx *c, const void *const *dptrs, const size_t *owned,
                                           const size_t *spans, ss_b200_sharded **out)
{
    if (!c || !out || !dptrs || !owned || !spans)
        return SS_B200_E_ARG;
    *out = nullptr;
    const int n = (int)c->devices.size();
    ss_b200_sharded *sh = new (std::nothrow) ss_b200_sharded();
    if (!sh)
        return SS_B200_E_NOMEM;
    sh->ctx = c;
    sh->shards.resize(n);
    size_t start = 0;
    for (int d = 0; d < n; d++) {
        if (spans[d] < owned[d] || (spans[d] && !dptrs[d])) {
            delete sh;
            return SS_B200_E_ARG;
        }
        auto &s = sh->shards[d];
        s.dptr = (const uint8_t *)dptrs[d];
        s.start = start;
        s.owned = owned[d];
        s.span = spans[d];
        start += owned[d];
    }
    // total length = end of the last byte held by anybody
    size_t total = 0;
    for (int d = 0; d < n; d++)
        if (sh->shards[d].span && sh->shards[d].start + sh->shards[d].span > total)
            total = sh->shards[d].start + sh->shards[d].span;
    sh->total = total;
    *out = sh;
    return SS_B200_OK;
}

extern "C" int ss_b200_find_sharded(ss_b200_ctx *c, const ss_b200_searcher *s, const ss_b200_sharded *sh,
                                    size_t *offset)
{
    if (!c || !s || !sh || !offset || sh->ctx != c)
        return SS_B200_E_ARG;
    const size_t k = s->needle.size();
    if (k == 0) { // N0 (src/x86.rs:470,500)
        *offset = 0;
        return SS_B200_OK;
    }
    if (sh->total < k) { // src/x86.rs:357-359
        *offset = SS_B200_NPOS;
        return SS_B200_OK;
    }
    const int n = (int)c->devices.size();
    // every start position must be verifiable inside its shard: a shard that is followed by more bytes
    // needs a halo of k - 1
    for (int d = 0; d < n; d++) {
        const auto &p = sh->shards[d];
        const bool more_behind = p.start + p.span < sh->total;
        if (p.owned && more_behind && p.span - p.owned < k - 1) {
            ss_capi_set_error("needle longer than the halo of the sharded haystack + 1");
            return SS_B200_E_ARG;
        }
    }
    std::lock_guard<std::mutex> lk(c->mu);
    int prev = -1;
    cudaGetDevice(&prev);
    int rc = SS_B200_OK;
    const int ex = n == 1 ? SS_B200_EXCHANGE_HOST : c->exchange;
    c->seq++;
    for (int d = 0; d < n && rc == SS_B200_OK; d++) {
        SsLane *lane = c->lane_ptrs[d];
        const auto &p = sh->shards[d];
        cudaError_t e = cudaSetDevice(c->devices[d]);
        if (e != cudaSuccess) {
            rc = ss_capi_cuda_fail(e, "cudaSetDevice");
            break;
        }
        lane->slot->value = SS_RESULT_PENDING;
        uint64_t *slot = (uint64_t *)&lane->slot_dev->value;
        // cross-GPU early exit (the peer exchange carries its own stop words in the mailboxes): shard d
        // polls the stop word in its lane's workspace block and, on its first match, stores the search's
        // sequence number into the stop words of the shards to its right
        SsStopSpec stop;
        if (c->peer_access && n > 1) {
            stop.stop_word = (const unsigned long long *)((const uint8_t *)lane->ws + 32);
            stop.seq = c->seq;
            for (int r = d + 1; r < n; r++)
                stop.peers[stop.n_peers++] = (unsigned long long *)((uint8_t *)c->lane_ptrs[r]->ws + 32);
        }
        const SsStopSpec *sp = stop.stop_word ? &stop : nullptr;
        if (ex == SS_B200_EXCHANGE_PEER) {
            rc = ss_b200_find_in_device_exchange_async(s, p.dptr, p.span, p.start, p.owned, lane->ws,
                                                       c->mailbox.data(), n, d, c->seq, slot, lane->stream);
        } else if (ex == SS_B200_EXCHANGE_NCCL) {
            rc = ss_capi_find_async(s, p.dptr, p.span, p.start, p.owned, lane->ws, (uint64_t *)c->red[d],
                                    lane->stream, sp);
        } else {
            rc = ss_capi_find_async(s, p.dptr, p.span, p.start, p.owned, lane->ws, slot, lane->stream, sp);
        }
    }
    if (rc == SS_B200_OK && ex == SS_B200_EXCHANGE_NCCL) {
        // one 8-byte ncclAllReduce(min) over the first offsets (NONE = INT64_MAX): found AND leftmost
        ncclResult_t r = g_nccl.GroupStart();
        for (int d = 0; d < n && r == ncclSuccess; d++)
            r = g_nccl.AllReduce(c->red[d], c->red[d], 1, ncclUint64, ncclMin, c->comms[d], c->lane_ptrs[d]->stream);
        ncclResult_t r2 = g_nccl.GroupEnd();
        if (r != ncclSuccess || r2 != ncclSuccess)
            rc = nccl_fail(r != ncclSuccess ? r : r2, "ncclAllReduce(min)");
        for (int d = 0; d < n && rc == SS_B200_OK; d++) {
            cudaError_t e = cudaMemcpyAsync((void *)&c->lane_ptrs[d]->slot->value, c->red[d], 8, cudaMemcpyDeviceToHost,
                                            c->lane_ptrs[d]->stream);
            if (e != cudaSuccess)
                rc = ss_capi_cuda_fail(e, "cudaMemcpyAsync(result)");
        }
    }
    unsigned long long best = SS_NONE_U64;
    for (int d = 0; d < n; d++) {
        SsLane *lane = c->lane_ptrs[d];
        if (rc == SS_B200_OK)
            rc = ss_capi_wait_slot(&lane->slot->value, SS_RESULT_PENDING, lane->stream);
        else
            cudaStreamSynchronize(lane->stream);
        const unsigned long long v = lane->slot->value;
        if (rc == SS_B200_OK && v < best)
            best = v;
    }
    if (prev >= 0)
        cudaSetDevice(prev);
    if (rc != SS_B200_OK)
        return rc;
    *offset = best == SS_NONE_U64 ? SS_B200_NPOS : (size_t)best;
    return SS_B200_OK;
}

extern "C" int ss_b200_search_sharded(ss_b200_ctx *c, const ss_b200_searcher *s, const ss_b200_sharded *sh,
                                      uint8_t *found, size_t *global_offset)
{
    if (!found)
        return SS_B200_E_ARG;
    size_t off = SS_B200_NPOS;
    int rc = ss_b200_find_sharded(c, s, sh, &off);
    if (rc != SS_B200_OK)
        return rc;
    *found = off != SS_B200_NPOS;
    if (global_offset)
        *global_offset = off;
    return SS_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// host slice striped over all devices

extern "C" int ss_b200_find_in_host_multi(ss_b200_ctx *c, const ss_b200_searcher *s, const uint8_t *host, size_t len,
                                          size_t *offset)
{
    if (!c || !s || !offset || (len && !host))
        return SS_B200_E_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    return ss_host_engine_find(c->lane_ptrs.data(), (int)c->lane_ptrs.size(), s, host, len, offset, &c->last_host);
}

extern "C" int ss_b200_search_in_host_multi(ss_b200_ctx *c, const ss_b200_searcher *s, const uint8_t *host, size_t len,
                                            uint8_t *found)
{
    if (!found)
        return SS_B200_E_ARG;
    size_t off = SS_B200_NPOS;
    int rc = ss_b200_find_in_host_multi(c, s, host, len, &off);
    if (rc == SS_B200_OK)
        *found = off != SS_B200_NPOS;
    return rc;
}

extern "C" int ss_b200_ctx_last_host_stats(const ss_b200_ctx *c, uint64_t *h2d_bytes, uint64_t *chunks,
                                           uint64_t *chunk_bytes, int *mode)
{
    if (!c)
        return SS_B200_E_ARG;
    if (h2d_bytes)
        *h2d_bytes = c->last_host.h2d_bytes;
    if (chunks)
        *chunks = c->last_host.chunks;
    if (chunk_bytes)
        *chunk_bytes = c->last_host.chunk_bytes;
    if (mode)
        *mode = c->last_host.mode + (c->last_host.staged ? 10 : 0);
    return SS_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// many-haystack mode over the devices of a context

extern "C" void ss_b200_ctx_hayset_free(ss_b200_ctx_hayset *hs)
{
    if (!hs)
        return;
    for (size_t d = 0; d < hs->parts.size(); d++) {
        auto &p = hs->parts[d];
        SsDeviceGuard g(hs->devices[d]);
        ss_b200_hayset_free(p.set);
        cudaFree(p.blob);
        cudaFree(p.offsets);
        cudaFree(p.flags);
        cudaFree(p.workspace);
    }
    delete hs;
}

extern "C" size_t ss_b200_ctx_hayset_len(const ss_b200_ctx_hayset *hs) { return hs ? hs->n : 0; }

// Partition: contiguous index ranges with balanced bytes (range d ends at the first haystack whose end
// passes d+1 n-ths of the blob), needles replicated (they travel in the kernel arguments).
extern "C" int ss_b200_ctx_hayset_upload(const ss_b200_ctx *c, const uint8_t *blob, const uint64_t *offsets, size_t n,
                                         ss_b200_ctx_hayset **out)
{
    if (!c || !out || !offsets || offsets[0] != 0 || (offsets[n] && !blob))
        return SS_B200_E_ARG;
    *out = nullptr;
    for (size_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i])
            return SS_B200_E_ARG;
    std::unique_ptr<ss_b200_ctx_hayset, void (*)(ss_b200_ctx_hayset *)> hs(new (std::nothrow) ss_b200_ctx_hayset(),
                                                                          ss_b200_ctx_hayset_free);
    if (!hs)
        return SS_B200_E_NOMEM;
    const int nd = (int)c->devices.size();
    hs->ctx = c;
    hs->devices = c->devices;
    hs->n = n;
    hs->parts.resize(nd);
    const uint64_t total = offsets[n];
    size_t i = 0;
    for (int d = 0; d < nd; d++) {
        auto &p = hs->parts[d];
        p.lo = i;
        if (d == nd - 1) {
            i = n;
        } else {
            const unsigned __int128 target = (unsigned __int128)total * (d + 1) / nd;
            while (i < n && offsets[i + 1] <= (uint64_t)target)
                i++;
        }
        p.hi = i;
    }
    std::vector<uint64_t> rebased;
    for (int d = 0; d < nd; d++) {
        auto &p = hs->parts[d];
        const size_t cnt = p.hi - p.lo;
        if (cnt == 0)
            continue;
        SsDeviceGuard g(c->devices[d]);
        SsLane *lane = c->lane_ptrs[d];
        const uint64_t b0 = offsets[p.lo];
        p.blob_len = (size_t)(offsets[p.hi] - b0);
        rebased.resize(cnt + 1);
        for (size_t j = 0; j <= cnt; j++)
            rebased[j] = offsets[p.lo + j] - b0;
        const size_t alloc = ((p.blob_len + 15) & ~(size_t)15) + 16;
        SS_CUDA(cudaMalloc(&p.blob, alloc));
        SS_CUDA(cudaMalloc(&p.offsets, (cnt + 1) * sizeof(uint64_t)));
        SS_CUDA(cudaMalloc(&p.flags, cnt));
        SS_CUDA(cudaMalloc(&p.workspace, 32));
        SS_CUDA(cudaMemsetAsync(p.workspace, 0, 32, lane->stream));
        SS_CUDA(cudaMemsetAsync(p.blob + p.blob_len, 0, alloc - p.blob_len, lane->stream));
        if (p.blob_len)
            SS_CUDA(cudaMemcpyAsync(p.blob, blob + b0, p.blob_len, cudaMemcpyHostToDevice, lane->stream));
        SS_CUDA(cudaMemcpyAsync(p.offsets, rebased.data(), (cnt + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice,
                                lane->stream));
        SS_CUDA(cudaStreamSynchronize(lane->stream)); // `rebased` is reused for the next device
        int rc = ss_b200_hayset_create(p.blob, p.offsets, cnt, p.blob_len, lane->stream, &p.set);
        if (rc != SS_B200_OK)
            return rc;
        SS_CUDA(cudaStreamSynchronize(lane->stream));
    }
    *out = hs.release();
    return SS_B200_OK;
}

extern "C" int ss_b200_ctx_hayset_part(const ss_b200_ctx_hayset *hs, int i, size_t *lo, size_t *hi)
{
    if (!hs || i < 0 || i >= (int)hs->parts.size())
        return SS_B200_E_ARG;
    if (lo)
        *lo = hs->parts[i].lo;
    if (hi)
        *hi = hs->parts[i].hi;
    return SS_B200_OK;
}

// flags[h] = search_in(haystack h) (src/x86.rs:523) for every haystack of the set; `flags` is host memory.
extern "C" int ss_b200_ctx_hayset_search(ss_b200_ctx *c, const ss_b200_searcher *s, const ss_b200_ctx_hayset *hs,
                                         uint8_t *flags)
{
    if (!c || !s || !hs || hs->ctx != c || (hs->n && !flags))
        return SS_B200_E_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    int prev = -1;
    cudaGetDevice(&prev);
    int rc = SS_B200_OK;
    const int nd = (int)c->devices.size();
    for (int d = 0; d < nd && rc == SS_B200_OK; d++) {
        const auto &p = hs->parts[d];
        if (p.hi == p.lo)
            continue;
        cudaError_t e = cudaSetDevice(c->devices[d]);
        if (e != cudaSuccess) {
            rc = ss_capi_cuda_fail(e, "cudaSetDevice");
            break;
        }
        SsLane *lane = c->lane_ptrs[d];
        rc = ss_b200_hayset_search_async(s, p.set, p.flags, p.workspace, lane->stream);
        if (rc == SS_B200_OK) {
            e = cudaMemcpyAsync(flags + p.lo, p.flags, p.hi - p.lo, cudaMemcpyDeviceToHost, lane->stream);
            if (e != cudaSuccess)
                rc = ss_capi_cuda_fail(e, "cudaMemcpyAsync(flags)");
        }
    }
    for (int d = 0; d < nd; d++) {
        cudaError_t e = cudaStreamSynchronize(c->lane_ptrs[d]->stream);
        if (e != cudaSuccess && rc == SS_B200_OK)
            rc = ss_capi_cuda_fail(e, "cudaStreamSynchronize(flags)");
    }
    if (prev >= 0)
        cudaSetDevice(prev);
    return rc;
}
