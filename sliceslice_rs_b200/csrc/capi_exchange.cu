// capi_exchange.cu -- sharded search with the exchange fused into the scan epilogue (peer mailboxes
// over NVLink, CUDA IPC); see include/sliceslice_b200.h "Sharded search with the exchange fused".
#include "capi_internal.h"

#include <cstring>

// ---------------------------------------------------------------------------------------------
// Peer mailbox exchange: the MIN over ranks of the first offsets without a collective call.
// Mailbox layout (per rank, plain cudaMalloc memory shared through CUDA IPC):
//   u64 slot[SS_MAILBOX_DEPTH][world]; slot[seq % 4][r] = result of rank r for search `seq`
//   u64 stop;  the sequence number of a search for which a rank to the LEFT has already matched: this
//              rank's scan polls it and stops (cross-GPU early exit, ScanArgs::stop_word)
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
    const int n = SS_MAILBOX_DEPTH * world + 2; // + the stop word (all-ones: no search has that number)
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
    int rc = ss_capi_device_info(dev);
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
        rc = ss_capi_build_args(s, dptr, len, base_offset, start_limit, dev.device, a);
        if (rc != SS_B200_OK)
            return rc;
        a.ws = (SsWorkspace *)workspace;
        a.out = (unsigned long long *)((uint8_t *)workspace + 16); // local copy of this rank's own result
        a.n_peers = (uint32_t)world;
        for (int p = 0; p < world; p++)
            a.peer_slot[p] = (unsigned long long *)mailboxes[p] + row + rank;
        // cross-GPU early exit: poll the own stop word, tell the ranks to the right about the first match
        SsStopSpec stop;
        const size_t stop_index = (size_t)SS_MAILBOX_DEPTH * (size_t)world;
        stop.stop_word = (const unsigned long long *)mailboxes[rank] + stop_index;
        stop.seq = seq;
        for (int p = rank + 1; p < world; p++)
            stop.peers[stop.n_peers++] = (unsigned long long *)mailboxes[p] + stop_index;
        ss_capi_apply_stop(a, stop);
        SS_CUDA(ss_host_launch_scan(a, ss_capi_tuning(), dev, st));
    }
    mailbox_min_kernel<<<1, 32, 0, st>>>((unsigned long long *)mailboxes[rank] + row, world,
                                         (unsigned long long *)d_result);
    ss_host_count_launch(1);
    SS_CUDA(cudaGetLastError());
    return SS_B200_OK;
}

