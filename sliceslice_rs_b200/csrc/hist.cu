// hist.cu -- byte histogram of device memory + second-anchor (`position`) selection.
//
// The reference leaves the second anchor to the caller (`with_position`, src/x86.rs:289-297,
// :461-468; rationale :252-255: the last byte is usually good but can be made a worst case).  Results
// never depend on it (src/lib.rs:375-378), only the candidate rate does.  This file provides the
// data-driven choice: a 256-bin histogram of the haystack (whole range, or evenly spaced 4 KiB blocks
// of it) and a host-side rule that picks the needle index whose byte is rarest under that histogram.
#include "capi_internal.h"

#include <cstring>

namespace {

constexpr int HIST_THREADS = 256;
constexpr int HIST_WARPS = HIST_THREADS / 32;
constexpr unsigned HIST_BLOCK = 4096; // sample granule in bytes
// What sample_bytes == 0 means: 16 MiB in evenly spaced 4 KiB granules.  The histogram only ranks needle
// bytes by frequency for the second-anchor choice; a sample of that size does it in ~20 us on any
// haystack, whereas counting every byte costs one shared-memory atomic per byte (1.8 TB/s measured, a
// quarter of what the scan itself reads).
constexpr size_t SS_HIST_DEFAULT_SAMPLE = (size_t)16 << 20;

__device__ __forceinline__ void hist_word(uint32_t *h, uint32_t w)
{
    atomicAdd(&h[w & 0xFF], 1u);
    atomicAdd(&h[(w >> 8) & 0xFF], 1u);
    atomicAdd(&h[(w >> 16) & 0xFF], 1u);
    atomicAdd(&h[w >> 24], 1u);
}

// Granule g of n_granules covers bytes [start(g), start(g) + HIST_BLOCK) clipped to len, with
// start(g) = g * stride_granules * HIST_BLOCK (stride 1 = every byte).  One private histogram per
// warp in shared memory (text is dominated by a few bytes: per-warp copies cut the atomic conflicts),
// folded into the global 64-bit bins once per CTA.
__global__ void __launch_bounds__(HIST_THREADS) byte_hist_kernel(const uint8_t *__restrict__ p, unsigned long long len,
                                                                 unsigned long long n_granules,
                                                                 unsigned long long stride_granules,
                                                                 unsigned long long *__restrict__ hist)
{
    __shared__ uint32_t sh[HIST_WARPS][256];
    for (int i = threadIdx.x; i < HIST_WARPS * 256; i += HIST_THREADS)
        (&sh[0][0])[i] = 0;
    __syncthreads();
    uint32_t *h = sh[threadIdx.x >> 5];
    for (unsigned long long g = blockIdx.x; g < n_granules; g += gridDim.x) {
        const unsigned long long b0 = g * stride_granules * HIST_BLOCK;
        const unsigned long long b1 = (len - b0 < HIST_BLOCK) ? len : b0 + HIST_BLOCK;
        // 16 bytes per thread per trip; unaligned edges byte by byte
        const uintptr_t a0 = reinterpret_cast<uintptr_t>(p + b0);
        unsigned long long lo = b0 + ((16 - (a0 & 15)) & 15);
        if (lo > b1)
            lo = b1;
        const unsigned long long hi = lo + ((b1 - lo) & ~15ull);
        for (unsigned long long i = b0 + threadIdx.x; i < lo; i += HIST_THREADS)
            atomicAdd(&h[p[i]], 1u);
        for (unsigned long long i = lo + 16ull * threadIdx.x; i < hi; i += 16ull * HIST_THREADS) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p + i));
            hist_word(h, v.x);
            hist_word(h, v.y);
            hist_word(h, v.z);
            hist_word(h, v.w);
        }
        for (unsigned long long i = hi + threadIdx.x; i < b1; i += HIST_THREADS)
            atomicAdd(&h[p[i]], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 256; b += HIST_THREADS) {
        unsigned long long s = 0;
#pragma unroll
        for (int w = 0; w < HIST_WARPS; w++)
            s += sh[w][b];
        if (s)
            atomicAdd(&hist[b], s);
    }
}

// Background weights used when the caller has no histogram: a coarse picture of "text or binary"
// (space and the common English letters heavy, NUL heavy, everything unusual light).  Letter order:
// the classic frequency ranking "etaoinshrdlcumwfgypbvkjxqz".
struct DefaultWeights {
    uint64_t w[256];
    DefaultWeights()
    {
        for (int b = 0; b < 256; b++)
            w[b] = 4;
        for (int b = 0x80; b < 0x100; b++)
            w[b] = 8; // UTF-8 continuation / high bytes
        w[0x00] = 120;
        w[0xFF] = 40;
        w['\n'] = w['\r'] = w['\t'] = 60;
        for (int b = 0x21; b < 0x7F; b++)
            w[b] = 30; // punctuation (letters and digits overwritten below)
        for (int b = '0'; b <= '9'; b++)
            w[b] = 50;
        const char *order = "etaoinshrdlcumwfgypbvkjxqz";
        for (int r = 0; order[r]; r++) {
            const uint64_t lw = 250 - 8 * (uint64_t)r;
            w[(uint8_t)order[r]] = lw;
            w[(uint8_t)(order[r] - 'a' + 'A')] = lw / 4 + 10;
        }
        w[' '] = 255;
    }
};
const DefaultWeights g_default_weights;

} // namespace

// A second anchor further away than this would push the scan off its TMA-staged variant (the staged
// halo is SS_TMA_HALO_MAX bytes, scan_long.cu); positions beyond it are not considered.
#define SS_RAREST_MAX_POSITION 2032

extern "C" int ss_b200_rarest_position(const uint8_t *needle, size_t len, const uint64_t *hist, size_t *position)
{
    if (!position || (len && !needle))
        return SS_B200_E_ARG;
    *position = 0;
    if (len < 2)
        return SS_B200_OK; // N0: ignored (src/x86.rs:470); N1: must be 0 (:473)
    const uint64_t *w = hist ? hist : g_default_weights.w;
    const size_t last = len - 1 < SS_RAREST_MAX_POSITION ? len - 1 : SS_RAREST_MAX_POSITION;
    // cost of index p = frequency of needle[p], weighted 17/16 once the second anchor leaves the first
    // chunk's register window (p >= 16: two more shared-memory loads per chunk, measured ~5-8 %);
    // equal costs go to the larger p (anchors far apart are closer to independent in natural text,
    // and p = len - 1 is the reference's own default)
    size_t best = 1;
    uint64_t best_cost = ~0ull;
    for (size_t p = 1; p <= last; p++) {
        const uint64_t f = w[needle[p]] < (1ull << 58) ? w[needle[p]] : (1ull << 58); // keep 17 * f in range
        const uint64_t cost = f * (p < 16 ? 16u : 17u);
        if (cost <= best_cost) {
            best_cost = cost;
            best = p;
        }
    }
    *position = best;
    return SS_B200_OK;
}

extern "C" int ss_b200_searcher_new_rarest(const uint8_t *needle, size_t len, const uint64_t *hist,
                                           ss_b200_searcher **out)
{
    size_t position = 0;
    int rc = ss_b200_rarest_position(needle, len, hist, &position);
    if (rc != SS_B200_OK)
        return rc;
    return ss_b200_searcher_with_position(needle, len, position, out);
}

extern "C" int ss_b200_byte_histogram_device_async(const void *dptr, size_t len, size_t sample_bytes,
                                                   uint64_t *d_hist, void *stream)
{
    if (!d_hist || (len && !dptr))
        return SS_B200_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    SS_CUDA(cudaMemsetAsync(d_hist, 0, 256 * sizeof(uint64_t), st));
    if (len == 0)
        return SS_B200_OK;
    SsDeviceInfo dev;
    int rc = ss_capi_device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    const unsigned long long total = ((unsigned long long)len + HIST_BLOCK - 1) / HIST_BLOCK;
    unsigned long long n_granules = total, stride = 1;
    if (sample_bytes == 0)
        sample_bytes = SS_HIST_DEFAULT_SAMPLE; // the default is a sample: see the header
    if (sample_bytes < len) {
        unsigned long long want = ((unsigned long long)sample_bytes + HIST_BLOCK - 1) / HIST_BLOCK;
        stride = total / want; // >= 1 because sample_bytes < len
        n_granules = (total + stride - 1) / stride;
    }
    unsigned long long grid = (unsigned long long)dev.sm_count * 8;
    if (grid > n_granules)
        grid = n_granules;
    byte_hist_kernel<<<(unsigned)grid, HIST_THREADS, 0, st>>>((const uint8_t *)dptr, len, n_granules, stride,
                                                             (unsigned long long *)d_hist);
    ss_host_count_launch(1);
    SS_CUDA(cudaGetLastError());
    return SS_B200_OK;
}

extern "C" int ss_b200_haystack_byte_histogram(const ss_b200_haystack *h, size_t sample_bytes, uint64_t hist[256])
{
    if (!h || !hist)
        return SS_B200_E_ARG;
    SsLane *c = nullptr;
    int rc = ss_capi_get_lane(&c);
    if (rc != SS_B200_OK)
        return rc;
    uint64_t *d_hist = nullptr;
    SS_CUDA(cudaMalloc(&d_hist, 256 * sizeof(uint64_t)));
    rc = ss_b200_byte_histogram_device_async(h->dptr, h->len, sample_bytes, d_hist, c->stream);
    cudaError_t e = cudaSuccess;
    if (rc == SS_B200_OK) {
        e = cudaMemcpyAsync(hist, d_hist, 256 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(c->stream);
    }
    cudaFree(d_hist);
    if (rc != SS_B200_OK)
        return rc;
    if (e != cudaSuccess)
        return ss_capi_cuda_fail(e, "byte histogram");
    return SS_B200_OK;
}
