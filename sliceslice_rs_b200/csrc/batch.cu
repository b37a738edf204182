// batch.cu -- batched modes (K2): many (needle, haystack) pairs per launch.
//
// Workload definitions come from the reference's benches:
//   short sweep  bench/benches/i386.rs:118-131  needle i vs every haystack j >= i of the
//                length-sorted word list (10.5 M pairs of <= 24-byte strings)
//   long sweep   bench/benches/i386.rs:246-257  every needle over one long haystack
// Semantics per pair are DynamicAvx2Searcher::new(needle).search_in(haystack)
// (src/x86.rs:454-459, :498-519): empty needle => true; one byte => memchr; n < k => false;
// otherwise leftmost i with hay[i..i+k] == needle.  The filter is the same two-anchor
// test (first byte, last byte) as the long scan.
//
// The short-haystack regime is dispatch-latency-bound on the CPU (~7.5 ns per search); here
// one thread owns one pair and a warp ballot packs 32 results into one bitmap word.
#include "../../include/sliceslice_b200.h"
#include "ss_host.h"

#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#define SS_PAIR_NONE 0xFFFFFFFFFFFFFFFFull

struct NeedleDesc {
    unsigned long long off; // byte offset into the needle blob
    uint32_t k;             // length
    uint32_t pos;           // second anchor index (k - 1; 0 for k <= 1)
    uint32_t f4, l4;        // splatted anchors
    uint32_t skip;          // 1: result decided on the host (k == 0 or k > n)
    uint32_t pad;
};

struct ss_b200_batch {
    uint8_t *d_nblob = nullptr;
    unsigned long long *d_noff = nullptr;
    uint8_t *d_hblob = nullptr;
    unsigned long long *d_hoff = nullptr;
    NeedleDesc *d_desc = nullptr;
    size_t n_needles = 0, n_hay = 0;
    std::vector<uint8_t> h_nblob;
    std::vector<unsigned long long> h_noff;
    int device = -1;
};

// error plumbing shared with capi.cu
extern "C" const char *ss_b200_last_error(void);
int ss_capi_cuda_fail(cudaError_t e, const char *what);
int ss_capi_device_info(SsDeviceInfo &out);
#define SS_CUDA(call)                                                                                                \
    do {                                                                                                             \
        cudaError_t e__ = (call);                                                                                    \
        if (e__ != cudaSuccess)                                                                                      \
            return ss_capi_cuda_fail(e__, #call);                                                                    \
    } while (0)

namespace {

// One (needle, haystack) pair, scalar.  Returns first offset or SS_PAIR_NONE.
__device__ __forceinline__ unsigned long long pair_find(const uint8_t *__restrict__ nd, unsigned long long k,
                                                        const uint8_t *__restrict__ hs, unsigned long long n)
{
    if (k == 0)
        return 0; // N0 (src/x86.rs:500)
    if (n < k)
        return SS_PAIR_NONE; // src/x86.rs:357-359 / src/lib.rs:131-133
    const uint8_t f = __ldg(nd), l = __ldg(nd + k - 1);
    const unsigned long long end = n - k + 1;
    for (unsigned long long i = 0; i < end; i++) {
        if (__ldg(hs + i) == f && __ldg(hs + i + k - 1) == l) {
            unsigned long long j = 1;
            while (j + 1 < k && __ldg(hs + i + j) == __ldg(nd + j))
                j++;
            if (j + 1 >= k)
                return i;
        }
    }
    return SS_PAIR_NONE;
}

__global__ void __launch_bounds__(256) pairs_kernel(const uint8_t *__restrict__ nblob,
                                                    const unsigned long long *__restrict__ noff,
                                                    const uint8_t *__restrict__ hblob,
                                                    const unsigned long long *__restrict__ hoff,
                                                    const uint32_t *__restrict__ pn, const uint32_t *__restrict__ ph,
                                                    unsigned long long n_pairs, uint32_t *__restrict__ bitmap,
                                                    unsigned long long *__restrict__ offsets)
{
    // whole warps stay in the loop so the ballot is always full
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long rounded = (n_pairs + 31) & ~31ull;
    for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < rounded; p += stride) {
        unsigned long long r = SS_PAIR_NONE;
        if (p < n_pairs) {
            const uint32_t a = pn[p], b = ph[p];
            const unsigned long long no = noff[a], ho = hoff[b];
            r = pair_find(nblob + no, noff[a + 1] - no, hblob + ho, hoff[b + 1] - ho);
            if (offsets)
                offsets[p] = r;
        }
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, r != SS_PAIR_NONE);
        if (bitmap && (threadIdx.x & 31) == 0)
            bitmap[p >> 5] = word;
    }
}

// Triangular rule: pair p <-> (i, j >= i), p = i*W - i(i-1)/2 + (j - i).
__device__ __forceinline__ unsigned long long tri_row_start(unsigned long long i, unsigned long long w)
{
    return i * w - (i * (i - 1)) / 2; // i == 0 -> 0 (0 * anything; (0 * -1)/2 wraps to 0 as well)
}

__global__ void __launch_bounds__(256) triangular_kernel(const uint8_t *__restrict__ nblob,
                                                         const unsigned long long *__restrict__ noff,
                                                         const uint8_t *__restrict__ hblob,
                                                         const unsigned long long *__restrict__ hoff,
                                                         unsigned long long w, unsigned long long n_pairs,
                                                         uint32_t *__restrict__ bitmap,
                                                         unsigned long long *__restrict__ matches)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long rounded = (n_pairs + 31) & ~31ull;
    unsigned int local = 0;
    for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < rounded; p += stride) {
        bool hit = false;
        if (p < n_pairs) {
            // invert the triangular index: i = floor(((2W+1) - sqrt((2W+1)^2 - 8p)) / 2), then fix up
            const double t = 2.0 * (double)w + 1.0;
            long long i = (long long)((t - sqrt(t * t - 8.0 * (double)p)) * 0.5);
            if (i < 0)
                i = 0;
            if ((unsigned long long)i >= w)
                i = (long long)w - 1;
            while (i > 0 && tri_row_start((unsigned long long)i, w) > p)
                i--;
            while ((unsigned long long)i + 1 < w && tri_row_start((unsigned long long)i + 1, w) <= p)
                i++;
            const unsigned long long j = (unsigned long long)i + (p - tri_row_start((unsigned long long)i, w));
            const unsigned long long no = noff[i], ho = hoff[j];
            hit = pair_find(nblob + no, noff[i + 1] - no, hblob + ho, hoff[j + 1] - ho) != SS_PAIR_NONE;
        }
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, hit);
        if ((threadIdx.x & 31) == 0) {
            bitmap[p >> 5] = word;
            local += __popc(word);
        }
    }
    if ((threadIdx.x & 31) == 0 && local)
        atomicAdd(matches, (unsigned long long)local);
}

// ---- every needle over one long haystack -------------------------------------------------
// CTA (segment s, needle group g): the segment's chunks are loaded into registers once and
// tested against every needle of the group; CTAs are numbered segment-major so that early
// segments run first and `best[w]` prunes later segments (the reference's early return).
#define SS_MN_THREADS 256
#define SS_MN_U 2
#define SS_MN_SEG_CHUNKS (SS_MN_THREADS * SS_MN_U)

struct MultiArgs {
    const uint8_t *hay;
    unsigned long long n;
    unsigned long long last_chunk;
    const uint8_t *nblob;
    const NeedleDesc *desc;
    unsigned long long *best; // per needle, first offset (atomicMin), init all-ones
    uint32_t n_needles;
    uint32_t n_groups;
    uint32_t head;
};

template <int R>
__device__ __forceinline__ void multi_test(const MultiArgs &m, const NeedleDesc &d, unsigned long long w,
                                           const uint4 (&av)[SS_MN_U], unsigned long long c0,
                                           const uint4 *__restrict__ chunks, unsigned long long end)
{
    const bool k1 = d.k == 1;
    const unsigned long long q = d.pos >> 4;
    uint4 lo[SS_MN_U], hi[SS_MN_U];
    uint32_t fl[SS_MN_U];
    uint32_t any = 0;
#pragma unroll
    for (int u = 0; u < SS_MN_U; u++) {
        const unsigned long long c = c0 + (unsigned long long)u * SS_MN_THREADS + q;
        if (q == 0)
            lo[u] = av[u];
        else
            lo[u] = ldg16(chunks + (c < m.last_chunk ? c : m.last_chunk));
        if (R > 0)
            hi[u] = ldg16(chunks + (c + 1 < m.last_chunk ? c + 1 : m.last_chunk));
        else
            hi[u] = lo[u];
        fl[u] = k1 ? chunk_flag<0, true>(av[u], lo[u], hi[u], d.f4, d.l4)
                   : chunk_flag<R, false>(av[u], lo[u], hi[u], d.f4, d.l4);
        any |= fl[u];
    }
    if (any == 0)
        return;
    const uint8_t *nd = m.nblob + d.off;
#pragma unroll
    for (int u = 0; u < SS_MN_U; u++) {
        if (!fl[u])
            continue;
        const unsigned long long c = c0 + (unsigned long long)u * SS_MN_THREADS;
        const long long p0 = (long long)(c * 16ull) - (long long)m.head;
        const uint32_t aw[4] = {av[u].x, av[u].y, av[u].z, av[u].w};
        bool done = false;
#pragma unroll
        for (int j = 0; j < 4 && !done; j++) {
            uint32_t x = aw[j] ^ d.f4;
            if (!k1)
                x |= window_word<R>(lo[u], hi[u], j) ^ d.l4;
            uint32_t z = swar_zero_exact(x);
            while (z && !done) {
                const int bit = __ffs((int)z) - 1;
                z &= z - 1;
                const long long i = p0 + 4 * j + (bit >> 3);
                if (i < 0 || (unsigned long long)i >= end)
                    continue;
                bool eq = true;
                for (uint32_t t = 1; t < d.k; t++) {
                    if (__ldg(m.hay + i + t) != __ldg(nd + t)) {
                        eq = false;
                        break;
                    }
                }
                if (eq) {
                    atomicMin(&m.best[w], (unsigned long long)i);
                    done = true;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(SS_MN_THREADS) multi_needle_kernel(const __grid_constant__ MultiArgs m)
{
    const unsigned long long seg = blockIdx.x / m.n_groups;
    const uint32_t grp = blockIdx.x % m.n_groups;
    const uint4 *chunks = reinterpret_cast<const uint4 *>(m.hay - m.head);
    const unsigned long long c0 = seg * SS_MN_SEG_CHUNKS + threadIdx.x;
    uint4 av[SS_MN_U];
#pragma unroll
    for (int u = 0; u < SS_MN_U; u++) {
        const unsigned long long c = c0 + (unsigned long long)u * SS_MN_THREADS;
        av[u] = ldg16(chunks + (c < m.last_chunk ? c : m.last_chunk));
    }
    const long long seg_first = (long long)(seg * SS_MN_SEG_CHUNKS * 16ull) - (long long)m.head;
    for (uint32_t w = grp; w < m.n_needles; w += m.n_groups) {
        const NeedleDesc d = m.desc[w];
        if (d.skip)
            continue;
        const unsigned long long end = m.n - d.k + 1;
        if (seg_first >= (long long)end)
            continue; // no start position of this needle in the segment
        const unsigned long long cur = ld_relaxed_u64(&m.best[w]);
        if (cur != SS_PAIR_NONE && seg_first > (long long)cur)
            continue; // an earlier segment already matched
        switch (d.k == 1 ? 0 : (d.pos & 15)) {
#define SS_CASE(R)                                                                                                   \
    case R:                                                                                                          \
        multi_test<R>(m, d, w, av, c0, chunks, end);                                                                 \
        break;
            SS_CASE(0) SS_CASE(1) SS_CASE(2) SS_CASE(3) SS_CASE(4) SS_CASE(5) SS_CASE(6) SS_CASE(7) SS_CASE(8)
            SS_CASE(9) SS_CASE(10) SS_CASE(11) SS_CASE(12) SS_CASE(13) SS_CASE(14) SS_CASE(15)
#undef SS_CASE
        }
    }
}

template <typename T>
int upload(T **d, const void *h, size_t bytes)
{
    SS_CUDA(cudaMalloc((void **)d, bytes ? bytes + 16 : 16));
    if (bytes)
        SS_CUDA(cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice));
    return SS_B200_OK;
}

} // namespace

extern "C" int ss_b200_batch_create(const uint8_t *needle_blob, const uint64_t *needle_off, size_t n_needles,
                                    const uint8_t *hay_blob, const uint64_t *hay_off, size_t n_haystacks,
                                    ss_b200_batch **out)
{
    if (!out || !needle_off || !hay_off)
        return SS_B200_E_ARG;
    *out = nullptr;
    const size_t nb = (size_t)needle_off[n_needles], hb = (size_t)hay_off[n_haystacks];
    if ((nb && !needle_blob) || (hb && !hay_blob))
        return SS_B200_E_ARG;
    ss_b200_batch *b = new (std::nothrow) ss_b200_batch();
    if (!b)
        return SS_B200_E_NOMEM;
    b->n_needles = n_needles;
    b->n_hay = n_haystacks;
    b->h_nblob.assign(needle_blob, needle_blob + nb);
    b->h_noff.assign(needle_off, needle_off + n_needles + 1);
    int rc = SS_B200_OK;
    cudaError_t e = cudaGetDevice(&b->device);
    if (e != cudaSuccess)
        rc = ss_capi_cuda_fail(e, "cudaGetDevice");
    if (rc == SS_B200_OK)
        rc = upload(&b->d_nblob, needle_blob, nb);
    if (rc == SS_B200_OK)
        rc = upload(&b->d_noff, needle_off, (n_needles + 1) * sizeof(uint64_t));
    if (rc == SS_B200_OK)
        rc = upload(&b->d_hblob, hay_blob, hb);
    if (rc == SS_B200_OK)
        rc = upload(&b->d_hoff, hay_off, (n_haystacks + 1) * sizeof(uint64_t));
    if (rc == SS_B200_OK) {
        cudaError_t e2 = cudaMalloc((void **)&b->d_desc, (n_needles + 1) * sizeof(NeedleDesc));
        if (e2 != cudaSuccess)
            rc = ss_capi_cuda_fail(e2, "cudaMalloc(desc)");
    }
    if (rc != SS_B200_OK) {
        ss_b200_batch_free(b);
        return rc;
    }
    *out = b;
    return SS_B200_OK;
}

extern "C" void ss_b200_batch_free(ss_b200_batch *b)
{
    if (!b)
        return;
    cudaFree(b->d_nblob);
    cudaFree(b->d_noff);
    cudaFree(b->d_hblob);
    cudaFree(b->d_hoff);
    cudaFree(b->d_desc);
    delete b;
}

extern "C" int ss_b200_batch_search_pairs(const ss_b200_batch *b, const uint32_t *pair_needle,
                                          const uint32_t *pair_hay, size_t n_pairs, uint32_t *bitmap,
                                          uint64_t *offsets)
{
    if (!b || (n_pairs && (!pair_needle || !pair_hay)))
        return SS_B200_E_ARG;
    if (n_pairs == 0)
        return SS_B200_OK;
    for (size_t p = 0; p < n_pairs; p++)
        if (pair_needle[p] >= b->n_needles || pair_hay[p] >= b->n_hay)
            return SS_B200_E_ARG;
    SsDeviceInfo dev;
    int rc = ss_capi_device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    uint32_t *d_pn = nullptr, *d_ph = nullptr, *d_bm = nullptr;
    unsigned long long *d_off = nullptr;
    const size_t words = (n_pairs + 31) / 32;
    SS_CUDA(cudaMalloc((void **)&d_pn, n_pairs * 4));
    SS_CUDA(cudaMalloc((void **)&d_ph, n_pairs * 4));
    SS_CUDA(cudaMalloc((void **)&d_bm, words * 4));
    if (offsets)
        SS_CUDA(cudaMalloc((void **)&d_off, n_pairs * 8));
    SS_CUDA(cudaMemcpy(d_pn, pair_needle, n_pairs * 4, cudaMemcpyHostToDevice));
    SS_CUDA(cudaMemcpy(d_ph, pair_hay, n_pairs * 4, cudaMemcpyHostToDevice));
    unsigned long long blocks = (n_pairs + 255) / 256;
    const unsigned long long cap = (unsigned long long)dev.sm_count * 16;
    if (blocks > cap)
        blocks = cap;
    pairs_kernel<<<(unsigned)blocks, 256>>>(b->d_nblob, b->d_noff, b->d_hblob, b->d_hoff, d_pn, d_ph, n_pairs, d_bm,
                                           d_off);
    ss_host_count_launch(1);
    SS_CUDA(cudaGetLastError());
    if (bitmap)
        SS_CUDA(cudaMemcpy(bitmap, d_bm, words * 4, cudaMemcpyDeviceToHost));
    if (offsets)
        SS_CUDA(cudaMemcpy(offsets, d_off, n_pairs * 8, cudaMemcpyDeviceToHost));
    SS_CUDA(cudaDeviceSynchronize());
    cudaFree(d_pn);
    cudaFree(d_ph);
    cudaFree(d_bm);
    cudaFree(d_off);
    return SS_B200_OK;
}

extern "C" int ss_b200_batch_search_triangular(const ss_b200_batch *b, uint32_t *bitmap, uint64_t *matches)
{
    if (!b || !bitmap)
        return SS_B200_E_ARG;
    // the rule pairs word i with word j >= i of ONE list: the haystack set is the list
    const unsigned long long w = b->n_hay;
    if (b->n_needles != b->n_hay)
        return SS_B200_E_ARG;
    if (matches)
        *matches = 0;
    if (w == 0)
        return SS_B200_OK;
    SsDeviceInfo dev;
    int rc = ss_capi_device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    const unsigned long long n_pairs = w * (w + 1) / 2;
    const size_t words = (size_t)((n_pairs + 31) / 32);
    uint32_t *d_bm = nullptr;
    unsigned long long *d_m = nullptr;
    SS_CUDA(cudaMalloc((void **)&d_bm, words * 4));
    SS_CUDA(cudaMalloc((void **)&d_m, 8));
    SS_CUDA(cudaMemset(d_m, 0, 8));
    unsigned long long blocks = (n_pairs + 255) / 256;
    const unsigned long long cap = (unsigned long long)dev.sm_count * 16;
    if (blocks > cap)
        blocks = cap;
    // needle i is taken from the needle set, haystack j from the haystack set (for the
    // reference's workload both hold the same length-sorted word list)
    triangular_kernel<<<(unsigned)blocks, 256>>>(b->d_nblob, b->d_noff, b->d_hblob, b->d_hoff, w, n_pairs, d_bm,
                                                  d_m);
    ss_host_count_launch(1);
    SS_CUDA(cudaGetLastError());
    SS_CUDA(cudaMemcpy(bitmap, d_bm, words * 4, cudaMemcpyDeviceToHost));
    unsigned long long m = 0;
    SS_CUDA(cudaMemcpy(&m, d_m, 8, cudaMemcpyDeviceToHost));
    if (matches)
        *matches = m;
    cudaFree(d_bm);
    cudaFree(d_m);
    return SS_B200_OK;
}

// haystack handle internals (capi.cu)
extern "C" size_t ss_b200_haystack_len(const ss_b200_haystack *h);
extern "C" const void *ss_b200_haystack_device_ptr(const ss_b200_haystack *h);

extern "C" int ss_b200_batch_find_all_in(const ss_b200_batch *b, const ss_b200_haystack *h, uint64_t *offsets)
{
    if (!b || !h || !offsets)
        return SS_B200_E_ARG;
    const size_t nn = b->n_needles;
    if (nn == 0)
        return SS_B200_OK;
    const unsigned long long n = ss_b200_haystack_len(h);
    const uint8_t *hay = (const uint8_t *)ss_b200_haystack_device_ptr(h);
    SsDeviceInfo dev;
    int rc = ss_capi_device_info(dev);
    if (rc != SS_B200_OK)
        return rc;

    // needle descriptors: DynamicAvx2Searcher::new => position = k - 1
    std::vector<NeedleDesc> desc(nn);
    unsigned long long max_end = 0;
    for (size_t w = 0; w < nn; w++) {
        NeedleDesc &d = desc[w];
        memset(&d, 0, sizeof d);
        d.off = b->h_noff[w];
        const unsigned long long k = b->h_noff[w + 1] - b->h_noff[w];
        if (k > 0xFFFFFFFFull)
            return SS_B200_E_ARG;
        d.k = (uint32_t)k;
        d.skip = (k == 0 || k > n) ? 1u : 0u;
        if (!d.skip) {
            d.pos = (uint32_t)(k - 1);
            d.f4 = 0x01010101u * b->h_nblob[d.off];
            d.l4 = 0x01010101u * b->h_nblob[d.off + d.pos];
            if (n - k + 1 > max_end)
                max_end = n - k + 1;
        }
    }
    std::vector<unsigned long long> best(nn, SS_PAIR_NONE);
    if (max_end > 0) {
        unsigned long long *d_best = nullptr;
        SS_CUDA(cudaMemcpy(b->d_desc, desc.data(), nn * sizeof(NeedleDesc), cudaMemcpyHostToDevice));
        SS_CUDA(cudaMalloc((void **)&d_best, nn * 8));
        SS_CUDA(cudaMemset(d_best, 0xFF, nn * 8));
        MultiArgs m;
        memset(&m, 0, sizeof m);
        m.hay = hay;
        m.n = n;
        m.head = (uint32_t)(reinterpret_cast<uintptr_t>(hay) & 15);
        m.last_chunk = (m.head + n - 1) / 16;
        m.nblob = b->d_nblob;
        m.desc = b->d_desc;
        m.best = d_best;
        m.n_needles = (uint32_t)nn;
        const unsigned long long n_chunks = (m.head + max_end + 15) / 16;
        const unsigned long long n_seg = (n_chunks + SS_MN_SEG_CHUNKS - 1) / SS_MN_SEG_CHUNKS;
        // enough needle groups to fill the machine a few times over, at least ~32 needles each
        unsigned long long groups = ((unsigned long long)dev.sm_count * 32 + n_seg - 1) / n_seg;
        if (groups > (nn + 31) / 32)
            groups = (nn + 31) / 32;
        if (groups < 1)
            groups = 1;
        if (n_seg * groups > 0x7FFFFFFFull)
            groups = 0x7FFFFFFFull / n_seg ? 0x7FFFFFFFull / n_seg : 1;
        m.n_groups = (uint32_t)groups;
        multi_needle_kernel<<<(unsigned)(n_seg * groups), SS_MN_THREADS>>>(m);
        ss_host_count_launch(1);
        SS_CUDA(cudaGetLastError());
        SS_CUDA(cudaMemcpy(best.data(), d_best, nn * 8, cudaMemcpyDeviceToHost));
        cudaFree(d_best);
    }
    for (size_t w = 0; w < nn; w++) {
        const unsigned long long k = b->h_noff[w + 1] - b->h_noff[w];
        offsets[w] = (k == 0) ? 0 : best[w]; // N0 => found at 0; k > n stays NONE
    }
    return SS_B200_OK;
}
