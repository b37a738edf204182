// service.cu -- a resident scan kernel for SYNCHRONOUS searches over short device-resident haystacks.
//
// The reference's callers loop `searcher.search_in(haystack)` over thousands of needles
// (bench/benches/i386.rs:252-256); on the CPU such a call costs ~8 us.  A synchronous GPU search that
// launches a kernel per call pays the launch and the completion signal -- 7-8 us on these (virtualised)
// boxes before any byte is scanned (tests/cpp/latency_floor.cu).  This file removes the launch from the
// call: one persistent grid per calling thread stays resident while calls keep coming, and a call is
//     host: write a 128-byte request into mapped pinned memory      (no driver call)
//     GPU : CTA 0 polls that memory over PCIe, publishes the request to the other CTAs through L2,
//           every CTA scans its share of the haystack (same filter + register verify as the launched
//           kernels), the last CTA to finish stores the result word into mapped pinned memory
//     host: spins on the result word
// i.e. one PCIe round trip plus the scan.  The kernel retires by itself after `idle_us` without a request
// (Dekker-style handshake on an `alive` word, so a request can never fall between a retiring kernel and
// the next launch); every CTA also carries a watchdog, so a lost host cannot leave the GPU spinning.
//
// Served: first-match searches (ss_b200_find_in / ss_b200_search_in) over device memory of up to
// SS_SERVICE_MAX_BYTES with needles of up to 64 bytes, and short HOST slices (ss_b200_find_in_host, up to
// 32 KiB, needles of up to 17 bytes) out of the lane's mapped pinned copy; everything else takes the
// launched kernels.
// Loads of the haystack bypass L1 (ld.global.cg): the grid outlives any number of host-side writes to
// the haystack between two calls.
#include "capi_internal.h"

#include <atomic>
#include <cstring>

#define SS_SERVICE_THREADS 256
#define SS_SERVICE_MAX_BYTES ((size_t)4 << 20)
#define SS_SERVICE_MAX_NEEDLE 64u
// Residency: a grid is two CTAs of 256 threads per SM (the 857 KB text of the reference's benches is then one
// pass of 16-byte loads; measured against 1 and 0.5 CTAs per SM and 512-thread CTAs, profiles/
// r02_service_variants.txt).  The kernel is capped at 64 registers (__launch_bounds__(256, 4)), so two grids
// -- four CTAs per SM -- are the most that can be co-resident; a third calling thread launches kernels.
#define SS_SERVICE_MAX_PER_DEVICE 2
#ifndef SS_SERVICE_GRID_NUM // CTAs of the resident grid = SMs * NUM / DEN
#define SS_SERVICE_GRID_NUM 2
#define SS_SERVICE_GRID_DEN 1
#endif

// request, 128 bytes of mapped pinned host memory = two 64-byte lines, each closed by the request number
struct SsServiceDesc {
    uint32_t seq0;
    uint32_t k;
    uint32_t pos;
    uint32_t cmd; // 1 search device memory, 2 retire now, 3 search a haystack in mapped HOST memory
    uint64_t hay;
    uint64_t n;
    uint8_t needle_lo[32];
    uint8_t needle_hi[32];
    uint32_t pad[7];
    uint32_t seq1;
};
static_assert(sizeof(SsServiceDesc) == 128, "one coalesced 128-byte read per poll");

// device-side control block
struct SsServiceCtl {
    uint32_t desc[32];       // the request as published by CTA 0
    unsigned int go;         // number of the request every CTA should run now
    unsigned int exit_epoch; // == the kernel's epoch: retire
    unsigned int served;     // last request completed (read by the next incarnation)
    unsigned int pad;
    SsWorkspace ws;          // key / ticket of the grid-wide reduction
};

// host-visible status, mapped pinned
struct SsServiceStatus {
    volatile unsigned long long result; // SS_RESULT_PENDING until the last CTA stores the answer
    volatile unsigned int alive;        // 1 while the resident kernel will still pick up requests
    volatile unsigned int pad;
};

namespace {

__device__ __forceinline__ uint4 ld_cg16(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// haystack chunk: device memory through L2 (never L1: the grid outlives host-side writes); mapped host memory
// with a volatile load, which goes to the host every time (nothing of it may linger in a GPU cache between
// two requests -- the host rewrites that buffer before each call)
template <bool SYS>
__device__ __forceinline__ uint4 ld_hay16(const uint4 *p)
{
    if (!SYS)
        return ld_cg16(p);
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_sys_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_gpu_u32(const unsigned int *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const unsigned int *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// one request: every thread of the grid takes chunks c, c + T, c + 2T, ... (coalesced 16-byte loads)
template <int WS, bool BSZ, bool K1, bool SYS>
__device__ __forceinline__ void service_scan(const ScanArgs &a)
{
    const uint4 *chunks = reinterpret_cast<const uint4 *>(a.hay - a.head);
    FilterConsts fc;
    fc.f4 = a.f4;
    fc.l4 = a.l4;
    fc.bs = a.bs;
    fc.e4[0] = fc.e4[1] = 0;
    fc.xbs = 0;
    const unsigned long long last = a.last_chunk;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long q = a.q;
    for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < a.n_chunks; c += stride) {
        // the reference's early return (src/lib.rs:242-244): nothing right of a match matters
        const unsigned long long key = ld_relaxed_u64(&a.ws->key);
        if (key != 0 && (long long)(c * 16ull) - (long long)a.head > (long long)~key)
            break;
        uint4 av[1], nx[1], lo[1], hi[1];
        av[0] = ld_hay16<SYS>(chunks + (c < last ? c : last));
        nx[0] = ld_hay16<SYS>(chunks + (c + 1 < last ? c + 1 : last));
        if (K1 || q == 0) {
            lo[0] = av[0];
            hi[0] = nx[0];
        } else {
            lo[0] = ld_hay16<SYS>(chunks + (c + q < last ? c + q : last));
            hi[0] = ld_hay16<SYS>(chunks + (c + q + 1 < last ? c + q + 1 : last));
        }
        uint32_t fl[1];
        fl[0] = chunk_flag_x<WS, BSZ, K1, 0>(av[0], nx[0], lo[0], hi[0], fc);
        if (fl[0])
            step_hits<WS, BSZ, K1, 1>(a, av, nx, lo, hi, fl, c, false);
    }
}

__global__ void __launch_bounds__(SS_SERVICE_THREADS, 4)
    service_kernel(SsServiceCtl *ctl, const uint32_t *host_desc, SsServiceStatus *status, unsigned long long idle_ns,
                   unsigned int epoch)
{
    __shared__ ScanArgs sa;
    __shared__ uint32_t s_desc[32];
    __shared__ unsigned int s_seq;
    const int lane = threadIdx.x & 31;
    const bool leader_cta = blockIdx.x == 0;
    unsigned int last = ld_gpu_u32(&ctl->served); // requests up to here are history
    unsigned long long t_idle = now_ns();
    const unsigned long long watchdog_ns = 8ull * idle_ns + 200000000ull; // a CTA never spins longer than this

    for (;;) {
        // ---- wait for the next request (warp 0; every decision below is warp-uniform) ----
        if (threadIdx.x < 32) {
            unsigned int got = 0;
            if (leader_cta) {
                for (;;) {
                    const uint32_t v = ld_sys_u32(host_desc + lane); // one 128-byte read over PCIe
                    uint32_t seq0 = __shfl_sync(0xFFFFFFFFu, v, 0), seq1 = __shfl_sync(0xFFFFFFFFu, v, 31);
                    if (seq0 == seq1 && seq0 != last && seq0 != 0u) {
                        const uint32_t cmd = __shfl_sync(0xFFFFFFFFu, v, 3);
                        if (cmd == 2u) { // retire on request
                            if (lane == 0) {
                                ctl->served = seq0;
                                asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(&status->alive), "r"(0u) : "memory");
                                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(&status->result), "l"(SS_NONE_U64)
                                             : "memory");
                                st_release_u32(&ctl->exit_epoch, epoch);
                            }
                            got = 0xFFFFFFFFu;
                            break;
                        }
                        // publish: the request and a fresh reduction state, then (release) its number
                        ctl->desc[lane] = v;
                        if (lane == 0) {
                            ctl->ws.key = 0ull;
                            ctl->ws.done = 0u;
                        }
                        __threadfence();
                        __syncwarp();
                        if (lane == 0)
                            st_release_u32(&ctl->go, seq0);
                        s_desc[lane] = v;
                        got = seq0;
                        break;
                    }
                    const unsigned long long now = __shfl_sync(0xFFFFFFFFu, now_ns(), 0);
                    if (now - t_idle > idle_ns) {
                        // retire: say so FIRST, then look once more -- the host writes its request and then
                        // reads `alive`, so either we see the request here or the host sees alive == 0
                        if (lane == 0)
                            asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(&status->alive), "r"(0u) : "memory");
                        __threadfence_system();
                        __syncwarp();
                        const uint32_t v2 = ld_sys_u32(host_desc + lane);
                        seq0 = __shfl_sync(0xFFFFFFFFu, v2, 0);
                        seq1 = __shfl_sync(0xFFFFFFFFu, v2, 31);
                        if (seq0 == seq1 && seq0 != last && seq0 != 0u) {
                            if (lane == 0)
                                asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(&status->alive), "r"(1u)
                                             : "memory");
                            t_idle = now;
                            continue; // picked up at the top of the loop
                        }
                        if (lane == 0)
                            st_release_u32(&ctl->exit_epoch, epoch);
                        got = 0xFFFFFFFFu;
                        break;
                    }
                }
            } else {
                const unsigned long long t0 = __shfl_sync(0xFFFFFFFFu, now_ns(), 0);
                for (;;) {
                    // every lane polls (one broadcast transaction), so every lane has acquired what it reads next
                    const unsigned int g = __shfl_sync(0xFFFFFFFFu, ld_acquire_u32(&ctl->go), 0);
                    if (g != last) {
                        (void)ld_acquire_u32(&ctl->go);
                        s_desc[lane] = ld_gpu_u32(ctl->desc + lane); // from L2, never L1: rewritten per request
                        got = g;
                        break;
                    }
                    const unsigned int ex = __shfl_sync(0xFFFFFFFFu, ld_gpu_u32(&ctl->exit_epoch), 0);
                    const unsigned long long now = __shfl_sync(0xFFFFFFFFu, now_ns(), 0);
                    if (ex == epoch || now - t0 > watchdog_ns) {
                        got = 0xFFFFFFFFu;
                        break;
                    }
                }
            }
            if (lane == 0)
                s_seq = got;
        }
        __syncthreads();
        const unsigned int seq = s_seq;
        if (seq == 0xFFFFFFFFu)
            return;

        // ---- kernel arguments of this search, as ss_capi_build_args + ss_host_scan_geometry make them ----
        if (threadIdx.x < 32) {
            const uint8_t *nb = reinterpret_cast<const uint8_t *>(s_desc + 8); // needle bytes 0..63
            if (lane == 0) {
                const uint32_t k = s_desc[1], pos = s_desc[2];
                const unsigned long long hay = (unsigned long long)s_desc[4] | ((unsigned long long)s_desc[5] << 32);
                const unsigned long long n = (unsigned long long)s_desc[6] | ((unsigned long long)s_desc[7] << 32);
                sa.hay = reinterpret_cast<const uint8_t *>(hay);
                sa.n = n;
                sa.end = n - k + 1;
                sa.base = 0;
                sa.head = (uint32_t)(hay & 15ull);
                sa.n_chunks = (sa.head + sa.end + 15) / 16;
                sa.last_chunk = (sa.head + n - 1) / 16;
                sa.needle_g = nullptr;
                sa.ws = &ctl->ws;
                sa.out = nullptr;
                sa.k = k;
                sa.pos = pos;
                sa.q = pos / 16u;
                sa.f4 = 0x01010101u * nb[0];
                sa.l4 = 0x01010101u * nb[pos];
                sa.bs = 8u * (pos % 4u);
                sa.xk = 0;
                sa.n_peers = 0;
                sa.n_stop_peers = 0;
                sa.stop_word = nullptr;
                sa.seg_off = nullptr;
                sa.seg_flags = nullptr;
                sa.seg_hint = nullptr;
                sa.count = nullptr;
                sa.filter_is_exact = 0;
            }
            sa.needle_inline[lane] = nb[lane];
            sa.needle_inline[lane + 32] = nb[lane + 32];
            if (lane >= 1 && lane <= 16)
                sa.needle4[lane] = 0x01010101u * nb[lane];
        }
        cta_best_reset();
        __syncthreads();

        // ---- scan ----
        {
            const uint32_t r = sa.k == 1u ? 0u : sa.pos % 16u;
            const uint32_t cls = sa.k == 1u ? 8u : 2u * (r / 4u) + ((r % 4u) ? 1u : 0u);
            if (s_desc[3] == 3u) { // launch-uniform: the haystack is mapped host memory (needles of up to 17 bytes)
                switch (cls) {
                case 0: service_scan<0, true, false, true>(sa); break;
                case 1: service_scan<0, false, false, true>(sa); break;
                case 2: service_scan<1, true, false, true>(sa); break;
                case 3: service_scan<1, false, false, true>(sa); break;
                case 4: service_scan<2, true, false, true>(sa); break;
                case 5: service_scan<2, false, false, true>(sa); break;
                case 6: service_scan<3, true, false, true>(sa); break;
                case 7: service_scan<3, false, false, true>(sa); break;
                default: service_scan<0, true, true, true>(sa); break;
                }
            } else {
                switch (cls) {
                case 0: service_scan<0, true, false, false>(sa); break;
                case 1: service_scan<0, false, false, false>(sa); break;
                case 2: service_scan<1, true, false, false>(sa); break;
                case 3: service_scan<1, false, false, false>(sa); break;
                case 4: service_scan<2, true, false, false>(sa); break;
                case 5: service_scan<2, false, false, false>(sa); break;
                case 6: service_scan<3, true, false, false>(sa); break;
                case 7: service_scan<3, false, false, false>(sa); break;
                default: service_scan<0, true, true, false>(sa); break;
                }
            }
        }

        // ---- grid-wide result: the last CTA to finish answers the host ----
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int prev;
            asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(&ctl->ws.done) : "memory");
            if (prev == gridDim.x - 1) {
                const unsigned long long key = ld_relaxed_u64(&ctl->ws.key);
                const unsigned long long res = key ? ~key : SS_NONE_U64;
                ctl->served = seq;
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(&status->result), "l"(res) : "memory");
            }
        }
        last = seq;
        t_idle = now_ns();
    }
}

std::atomic<int> g_services_per_device[64];

} // namespace

struct SsService {
    int device = -1;
    cudaStream_t stream = nullptr;
    SsServiceDesc *desc = nullptr;     // mapped pinned
    uint32_t *desc_dev = nullptr;
    SsServiceStatus *status = nullptr; // mapped pinned
    SsServiceStatus *status_dev = nullptr;
    SsServiceCtl *ctl = nullptr;       // device
    unsigned int seq = 0, epoch = 0;
    int grid = 0;
    bool have_slot = false;
};

static void service_free(SsService *sv)
{
    if (!sv)
        return;
    if (sv->stream)
        cudaStreamDestroy(sv->stream);
    if (sv->desc)
        cudaFreeHost(sv->desc);
    if (sv->status)
        cudaFreeHost((void *)sv->status);
    if (sv->ctl)
        cudaFree(sv->ctl);
    if (sv->have_slot && sv->device >= 0 && sv->device < 64)
        g_services_per_device[sv->device].fetch_sub(1);
    delete sv;
    cudaGetLastError();
}

static int service_launch(SsService *sv, unsigned idle_us)
{
    sv->status->alive = 1;
    sv->epoch++;
    if (sv->epoch == 0)
        sv->epoch = 1;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    service_kernel<<<sv->grid, SS_SERVICE_THREADS, 0, sv->stream>>>(sv->ctl, sv->desc_dev, sv->status_dev,
                                                                   (unsigned long long)idle_us * 1000ull, sv->epoch);
    ss_host_count_launch(1);
    SS_CUDA(cudaGetLastError());
    return SS_B200_OK;
}

// Post one request and wait for its result word.
static int service_roundtrip(SsService *sv, uint32_t cmd, const ss_b200_searcher *s, const void *dptr, size_t len,
                             unsigned idle_us, unsigned long long *result)
{
    sv->seq++;
    if (sv->seq == 0u || sv->seq == 0xFFFFFFFFu)
        sv->seq = 1;
    SsServiceDesc *d = sv->desc;
    sv->status->result = SS_RESULT_PENDING;
    d->cmd = cmd;
    if (s) {
        const size_t k = s->needle.size();
        d->k = (uint32_t)k;
        d->pos = (uint32_t)s->position;
        d->hay = (uint64_t)(uintptr_t)dptr;
        d->n = (uint64_t)len;
        memcpy(d->needle_lo, s->needle.data(), k < 32 ? k : 32);
        if (k > 32)
            memcpy(d->needle_hi, s->needle.data() + 32, k - 32);
    }
    std::atomic_thread_fence(std::memory_order_release); // the request body before its number
    reinterpret_cast<volatile uint32_t &>(d->seq1) = sv->seq;
    reinterpret_cast<volatile uint32_t &>(d->seq0) = sv->seq;
    std::atomic_thread_fence(std::memory_order_seq_cst); // ... and the number before reading `alive`
    if (sv->status->alive == 0) {
        if (cmd == 2u) { // nothing resident: nothing to retire
            *result = SS_NONE_U64;
            return SS_B200_OK;
        }
        int rc = service_launch(sv, idle_us);
        if (rc != SS_B200_OK)
            return rc;
    }
    unsigned spins = 0;
    while (sv->status->result == SS_RESULT_PENDING) {
        if ((++spins & 0xFFF) == 0) {
            cudaError_t e = cudaStreamQuery(sv->stream);
            if (e == cudaSuccess) {
                // no kernel resident any more (it retired between our `alive` read and now, or never saw
                // the request): if the answer is not there, start another one -- it reads the pending request
                if (sv->status->result != SS_RESULT_PENDING)
                    break;
                if (cmd == 2u) {
                    *result = SS_NONE_U64;
                    return SS_B200_OK;
                }
                int rc = service_launch(sv, idle_us);
                if (rc != SS_B200_OK)
                    return rc;
            } else if (e != cudaErrorNotReady) {
                return ss_capi_cuda_fail(e, "service kernel");
            }
        }
    }
    *result = sv->status->result;
    return SS_B200_OK;
}

void ss_service_release(void *p)
{
    SsService *sv = (SsService *)p;
    if (!sv)
        return;
    SsDeviceGuard guard(sv->device);
    if (sv->stream && sv->status && sv->status->alive) {
        unsigned long long r = 0;
        service_roundtrip(sv, 2u, nullptr, nullptr, 0, 0, &r); // retire now
    }
    if (sv->stream)
        cudaStreamSynchronize(sv->stream);
    service_free(sv);
}

// Can this synchronous search go through the resident kernel?  (The caller knows whether dptr is plain
// device memory: haystack handles check it once at creation.)
bool ss_service_eligible(const ss_b200_searcher *s, size_t len)
{
    const size_t k = s->needle.size();
    return k != 0 && k <= SS_SERVICE_MAX_NEEDLE && len >= k && len <= SS_SERVICE_MAX_BYTES;
}

// One synchronous first-match search through the lane's resident kernel (created on first use).
// SS_B200_E_ARG with *used = false when no service slot is free: the caller launches a kernel instead.
int ss_service_find(SsLane *lane, const ss_b200_searcher *s, const void *dptr, size_t len, unsigned idle_us,
                    size_t *offset, bool *used, bool mapped_host)
{
    *used = false;
    if (mapped_host && s->needle.size() > 17)
        return SS_B200_OK; // the tail of a longer needle is compared with cached loads: launch instead
    SsService *sv = (SsService *)lane->service;
    if (!sv) {
        if (lane->device < 0 || lane->device >= 64)
            return SS_B200_OK;
        if (g_services_per_device[lane->device].fetch_add(1) >= SS_SERVICE_MAX_PER_DEVICE) {
            g_services_per_device[lane->device].fetch_sub(1);
            return SS_B200_OK; // enough resident grids on this device already
        }
        sv = new (std::nothrow) SsService();
        if (!sv) {
            g_services_per_device[lane->device].fetch_sub(1);
            return SS_B200_E_NOMEM;
        }
        sv->device = lane->device;
        sv->have_slot = true;
        SsDeviceInfo dev;
        int rc = ss_capi_device_info(dev);
        cudaError_t e = cudaSuccess;
        if (rc == SS_B200_OK) {
            sv->grid = dev.sm_count * SS_SERVICE_GRID_NUM / SS_SERVICE_GRID_DEN;
            e = cudaStreamCreateWithFlags(&sv->stream, cudaStreamNonBlocking);
        }
        if (rc == SS_B200_OK && e == cudaSuccess)
            e = cudaHostAlloc((void **)&sv->desc, sizeof(SsServiceDesc), cudaHostAllocMapped);
        if (rc == SS_B200_OK && e == cudaSuccess)
            e = cudaHostGetDevicePointer((void **)&sv->desc_dev, sv->desc, 0);
        if (rc == SS_B200_OK && e == cudaSuccess)
            e = cudaHostAlloc((void **)&sv->status, sizeof(SsServiceStatus), cudaHostAllocMapped);
        if (rc == SS_B200_OK && e == cudaSuccess)
            e = cudaHostGetDevicePointer((void **)&sv->status_dev, (void *)sv->status, 0);
        if (rc == SS_B200_OK && e == cudaSuccess)
            e = cudaMalloc((void **)&sv->ctl, sizeof(SsServiceCtl));
        if (rc == SS_B200_OK && e == cudaSuccess)
            e = cudaMemsetAsync(sv->ctl, 0, sizeof(SsServiceCtl), sv->stream);
        if (rc == SS_B200_OK && e == cudaSuccess)
            e = cudaStreamSynchronize(sv->stream);
        if (rc != SS_B200_OK || e != cudaSuccess) {
            service_free(sv);
            return rc != SS_B200_OK ? rc : ss_capi_cuda_fail(e, "service setup");
        }
        memset(sv->desc, 0, sizeof(SsServiceDesc));
        sv->status->result = 0;
        sv->status->alive = 0;
        lane->service = sv;
    }
    unsigned long long r = 0;
    int rc = service_roundtrip(sv, mapped_host ? 3u : 1u, s, dptr, len, idle_us, &r);
    if (rc != SS_B200_OK)
        return rc;
    *offset = r == SS_NONE_U64 ? SS_B200_NPOS : (size_t)r;
    *used = true;
    return SS_B200_OK;
}
