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
#include "capi_internal.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#define SS_PAIR_NONE 0xFFFFFFFFFFFFFFFFull

struct MultiNeedle {       // one needle, sorted by class
    unsigned long long off; // byte offset into the needle blob
    uint32_t k, pos;
    uint32_t f4, l4;
    uint32_t w;            // original needle index
    uint32_t pad;
};

struct MultiGroup {
    uint32_t first, count; // range in the sorted needle array
    uint32_t cls;          // 8 = one-byte needles, else 2 * WS + (bs != 0)
    uint32_t pad;
};

struct ss_b200_batch {
    uint8_t *d_nblob = nullptr;
    unsigned long long *d_noff = nullptr;
    uint8_t *d_hblob = nullptr;
    unsigned long long *d_hoff = nullptr;
    size_t n_needles = 0, n_hay = 0;
    std::vector<uint8_t> h_nblob;
    std::vector<unsigned long long> h_noff;
    int device = -1;
    // built once at creation for ss_b200_batch_find_all_in: needles sorted by filter class + groups
    MultiNeedle *d_multi = nullptr;
    MultiGroup *d_groups = nullptr;
    size_t n_multi = 0, n_groups = 0;
    unsigned long long min_k = 0; // shortest non-empty needle
    // per-call scratch, allocated on first use and kept (calls on one handle are serialised by `mu`)
    mutable std::mutex mu;
    mutable unsigned long long *d_best = nullptr; // n_needles first offsets
    mutable uint32_t *d_bitmap = nullptr;          // triangular bitmap
    mutable size_t bitmap_words = 0;
    mutable unsigned long long *d_count = nullptr;
};

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
// CTA (segment s, needle group g).  The segment's chunks (and their successors) are loaded into
// registers once and tested against every needle of the group.  Needles are sorted on the host by
// filter class -- (second-anchor word offset, shift or no shift, one-byte needle), the same template
// parameters as the long scan -- so a group runs one specialised loop with no per-needle dispatch; its
// descriptors and the current per-needle best offsets are staged in shared memory once per CTA.  CTAs
// are numbered segment-major so early segments run first and later ones are pruned by `best[w]` (the
// reference's early return, src/lib.rs:242-244).  Candidates are verified from the register window as
// in the long scan (exact_alive), needle bytes coming from the L1-cached needle blob.
#define SS_MN_THREADS 256
#define SS_MN_U 2
#define SS_MN_SEG_CHUNKS (SS_MN_THREADS * SS_MN_U)
#define SS_MN_GROUP 64

struct MultiArgs {
    const uint8_t *hay;
    unsigned long long n;
    unsigned long long last_chunk;
    const uint8_t *nblob;
    const MultiNeedle *needles;
    const MultiGroup *groups;
    unsigned long long *best; // per original needle index: first offset (atomicMin), init all-ones
    uint32_t n_groups;
    uint32_t head;
};

// exact decode + register refinement + publish for one flagged chunk of one needle
template <int WS, bool BSZ, bool K1>
__device__ __noinline__ void multi_verify(const MultiArgs &m, const MultiNeedle &d, uint4 av, uint4 nx, uint4 lo,
                                          uint4 hi, unsigned long long chunk, unsigned long long end)
{
    FilterConsts fc;
    fc.f4 = d.f4;
    fc.l4 = d.l4;
    fc.bs = 8u * (d.pos & 3u);
    uint32_t z[4];
    const uint8_t *nd = m.nblob + d.off;
    if (!exact_alive<WS, BSZ, K1>(av, nx, lo, hi, fc, d.k, [&](uint32_t j) { return 0x01010101u * (uint32_t)__ldg(nd + j); }, z))
        return;
    const long long p0 = (long long)(chunk * 16ull) - (long long)m.head;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t zz = z[j];
        while (zz) {
            const int bit = __ffs((int)zz) - 1;
            zz &= zz - 1;
            const long long i = p0 + 4 * j + (bit >> 3);
            if (i < 0 || (unsigned long long)i >= end)
                continue;
            bool eq = true;
            for (uint32_t t = 17; t < d.k; t++) { // only needles longer than the register window
                if (__ldg(m.hay + i + t) != __ldg(nd + t)) {
                    eq = false;
                    break;
                }
            }
            if (eq) {
                atomicMin(&m.best[d.w], (unsigned long long)i);
                return; // ascending order within the chunk
            }
        }
    }
}

template <int WS, bool BSZ, bool K1>
__device__ __forceinline__ void multi_group_loop(const MultiArgs &m, const MultiNeedle *sd,
                                                 const unsigned long long *sbest, uint32_t count,
                                                 const uint4 (&av)[SS_MN_U], const uint4 (&nx)[SS_MN_U],
                                                 unsigned long long c0, const uint4 *__restrict__ chunks,
                                                 long long seg_first)
{
    for (uint32_t t = 0; t < count; t++) {
        const MultiNeedle d = sd[t]; // shared-memory broadcast
        if (d.k > m.n)
            continue; // needle longer than the haystack: not found (src/x86.rs:357-359)
        const unsigned long long end = m.n - d.k + 1;
        if (seg_first >= (long long)end)
            continue; // no start position of this needle in the segment
        const unsigned long long cur = sbest[t];
        if (cur != SS_PAIR_NONE && seg_first > (long long)cur)
            continue; // an earlier segment already matched (CTA-uniform)
        FilterConsts fc;
        fc.f4 = d.f4;
        fc.l4 = d.l4;
        fc.bs = 8u * (d.pos & 3u);
        const unsigned long long q = d.pos >> 4;
        uint4 lo[SS_MN_U], hi[SS_MN_U];
        uint32_t fl[SS_MN_U];
        uint32_t any = 0;
#pragma unroll
        for (int u = 0; u < SS_MN_U; u++) {
            if (K1 || q == 0) {
                lo[u] = av[u];
                hi[u] = nx[u];
            } else { // second anchor 16 or more bytes away: rare (needles longer than 16 bytes)
                const unsigned long long c = c0 + (unsigned long long)u * SS_MN_THREADS + q;
                lo[u] = ldg16(chunks + (c < m.last_chunk ? c : m.last_chunk));
                hi[u] = ldg16(chunks + (c + 1 < m.last_chunk ? c + 1 : m.last_chunk));
            }
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < 4; j++)
                acc |= swar_zero_term(filter_word<WS, BSZ, K1, 0>(av[u], nx[u], lo[u], hi[u], j, fc));
            fl[u] = acc & 0x80808080u;
            any |= fl[u];
        }
        if (any) {
#pragma unroll
            for (int u = 0; u < SS_MN_U; u++)
                if (fl[u])
                    multi_verify<WS, BSZ, K1>(m, d, av[u], nx[u], lo[u], hi[u],
                                              c0 + (unsigned long long)u * SS_MN_THREADS, end);
        }
    }
}

__global__ void __launch_bounds__(SS_MN_THREADS) multi_needle_kernel(const __grid_constant__ MultiArgs m)
{
    __shared__ MultiNeedle sd[SS_MN_GROUP];
    __shared__ unsigned long long sbest[SS_MN_GROUP];
    const unsigned long long seg = blockIdx.x / m.n_groups;
    const MultiGroup g = m.groups[blockIdx.x % m.n_groups];
    if (threadIdx.x < g.count) {
        sd[threadIdx.x] = m.needles[g.first + threadIdx.x];
        sbest[threadIdx.x] = ld_relaxed_u64(&m.best[sd[threadIdx.x].w]);
    }
    const uint4 *chunks = reinterpret_cast<const uint4 *>(m.hay - m.head);
    const unsigned long long c0 = seg * SS_MN_SEG_CHUNKS + threadIdx.x;
    uint4 av[SS_MN_U], nx[SS_MN_U];
#pragma unroll
    for (int u = 0; u < SS_MN_U; u++) {
        const unsigned long long c = c0 + (unsigned long long)u * SS_MN_THREADS;
        av[u] = ldg16(chunks + (c < m.last_chunk ? c : m.last_chunk));
        nx[u] = ldg16(chunks + (c + 1 < m.last_chunk ? c + 1 : m.last_chunk));
    }
    const long long seg_first = (long long)(seg * SS_MN_SEG_CHUNKS * 16ull) - (long long)m.head;
    __syncthreads();
    switch (g.cls) {
#define SS_CASE(WS)                                                                                                  \
    case 2 * WS:                                                                                                     \
        multi_group_loop<WS, true, false>(m, sd, sbest, g.count, av, nx, c0, chunks, seg_first);                     \
        break;                                                                                                       \
    case 2 * WS + 1:                                                                                                 \
        multi_group_loop<WS, false, false>(m, sd, sbest, g.count, av, nx, c0, chunks, seg_first);                    \
        break;
        SS_CASE(0) SS_CASE(1) SS_CASE(2) SS_CASE(3)
#undef SS_CASE
    default:
        multi_group_loop<0, true, true>(m, sd, sbest, g.count, av, nx, c0, chunks, seg_first);
        break;
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

// Needle table of ss_b200_batch_find_all_in, built once per batch: every non-empty needle with
// DynamicAvx2Searcher::new's position (k - 1), sorted by filter class, cut into groups of <= 64.
static int build_multi_tables(ss_b200_batch *b)
{
    std::vector<MultiNeedle> nds;
    nds.reserve(b->n_needles);
    b->min_k = 0;
    for (size_t w = 0; w < b->n_needles; w++) {
        const unsigned long long k = b->h_noff[w + 1] - b->h_noff[w];
        if (k == 0)
            continue; // N0: found at 0, decided on the host
        if (k > 0xFFFFFFFFull)
            return SS_B200_E_ARG;
        MultiNeedle d;
        memset(&d, 0, sizeof d);
        d.off = b->h_noff[w];
        d.k = (uint32_t)k;
        d.pos = (uint32_t)(k - 1);
        d.f4 = 0x01010101u * b->h_nblob[d.off];
        d.l4 = 0x01010101u * b->h_nblob[d.off + d.pos];
        d.w = (uint32_t)w;
        nds.push_back(d);
        if (b->min_k == 0 || k < b->min_k)
            b->min_k = k;
    }
    auto cls_of = [](const MultiNeedle &d) -> uint32_t {
        if (d.k == 1)
            return 8u;
        const uint32_t r = d.pos & 15u;
        return 2u * (r >> 2) + ((r & 3u) ? 1u : 0u);
    };
    std::stable_sort(nds.begin(), nds.end(),
                     [&](const MultiNeedle &x, const MultiNeedle &y) { return cls_of(x) < cls_of(y); });
    std::vector<MultiGroup> groups;
    for (size_t i = 0; i < nds.size();) {
        const uint32_t c = cls_of(nds[i]);
        size_t j = i;
        while (j < nds.size() && j - i < SS_MN_GROUP && cls_of(nds[j]) == c)
            j++;
        MultiGroup g;
        g.first = (uint32_t)i;
        g.count = (uint32_t)(j - i);
        g.cls = c;
        g.pad = 0;
        groups.push_back(g);
        i = j;
    }
    b->n_multi = nds.size();
    b->n_groups = groups.size();
    if (nds.empty())
        return SS_B200_OK;
    int rc = upload(&b->d_multi, nds.data(), nds.size() * sizeof(MultiNeedle));
    if (rc == SS_B200_OK)
        rc = upload(&b->d_groups, groups.data(), groups.size() * sizeof(MultiGroup));
    return rc;
}

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
    if (rc == SS_B200_OK)
        rc = build_multi_tables(b);
    if (rc == SS_B200_OK) {
        // small pageable uploads may still be in flight on the legacy stream, which the searches' own
        // non-blocking streams do not wait for
        cudaError_t e2 = cudaStreamSynchronize(0);
        if (e2 != cudaSuccess)
            rc = ss_capi_cuda_fail(e2, "cudaStreamSynchronize(batch upload)");
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
    cudaFree(b->d_multi);
    cudaFree(b->d_groups);
    cudaFree(b->d_best);
    cudaFree(b->d_bitmap);
    cudaFree(b->d_count);
    delete b;
}

// ---------------------------------------------------------------------------------------------
// stream-ordered forms: inputs and outputs in device memory, no host synchronisation, no state shared
// between calls (any number of streams may use one batch handle concurrently)

namespace {
// empty needles are found at offset 0 (N0, src/x86.rs:500): the scan kernels never see them
__global__ void fix_empty_needles_kernel(const unsigned long long *__restrict__ noff, unsigned long long n,
                                         unsigned long long *__restrict__ out)
{
    const unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < n && noff[w + 1] == noff[w])
        out[w] = 0;
}
} // namespace

extern "C" int ss_b200_batch_search_pairs_async(const ss_b200_batch *b, const uint32_t *d_pair_needle,
                                                const uint32_t *d_pair_hay, size_t n_pairs, uint32_t *d_bitmap,
                                                uint64_t *d_offsets, void *stream)
{
    if (!b || (n_pairs && (!d_pair_needle || !d_pair_hay)))
        return SS_B200_E_ARG;
    if (n_pairs == 0)
        return SS_B200_OK;
    SsDeviceInfo dev;
    int rc = ss_capi_device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    unsigned long long blocks = (n_pairs + 255) / 256;
    const unsigned long long cap = (unsigned long long)dev.sm_count * 16;
    if (blocks > cap)
        blocks = cap;
    pairs_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        b->d_nblob, b->d_noff, b->d_hblob, b->d_hoff, d_pair_needle, d_pair_hay, n_pairs, d_bitmap,
        (unsigned long long *)d_offsets);
    ss_host_count_launch(1);
    SS_CUDA(cudaGetLastError());
    return SS_B200_OK;
}

extern "C" int ss_b200_batch_search_triangular_async(const ss_b200_batch *b, uint32_t *d_bitmap, uint64_t *d_matches,
                                                     void *stream)
{
    if (!b || !d_bitmap || !d_matches)
        return SS_B200_E_ARG;
    // the rule pairs word i with word j >= i of ONE list: the haystack set is the list
    const unsigned long long w = b->n_hay;
    if (b->n_needles != b->n_hay)
        return SS_B200_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    SS_CUDA(cudaMemsetAsync(d_matches, 0, 8, st));
    if (w == 0)
        return SS_B200_OK;
    SsDeviceInfo dev;
    int rc = ss_capi_device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    const unsigned long long n_pairs = w * (w + 1) / 2;
    unsigned long long blocks = (n_pairs + 255) / 256;
    const unsigned long long cap = (unsigned long long)dev.sm_count * 16;
    if (blocks > cap)
        blocks = cap;
    // needle i is taken from the needle set, haystack j from the haystack set (for the
    // reference's workload both hold the same length-sorted word list)
    triangular_kernel<<<(unsigned)blocks, 256, 0, st>>>(b->d_nblob, b->d_noff, b->d_hblob, b->d_hoff, w, n_pairs,
                                                         d_bitmap, (unsigned long long *)d_matches);
    ss_host_count_launch(1);
    SS_CUDA(cudaGetLastError());
    return SS_B200_OK;
}

extern "C" int ss_b200_batch_find_all_in_device_async(const ss_b200_batch *b, const void *dptr, size_t len,
                                                      uint64_t *d_offsets, void *stream)
{
    if (!b || !d_offsets || (len && !dptr))
        return SS_B200_E_ARG;
    const size_t nn = b->n_needles;
    if (nn == 0)
        return SS_B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long n = len;
    const uint8_t *hay = (const uint8_t *)dptr;
    SsDeviceInfo dev;
    int rc = ss_capi_device_info(dev);
    if (rc != SS_B200_OK)
        return rc;
    SS_CUDA(cudaMemsetAsync(d_offsets, 0xFF, nn * 8, st)); // UINT64_MAX = not found; k > n stays that way
    if (b->n_groups > 0 && n >= b->min_k) {
        MultiArgs m;
        memset(&m, 0, sizeof m);
        m.hay = hay;
        m.n = n;
        m.head = (uint32_t)(reinterpret_cast<uintptr_t>(hay) & 15);
        m.last_chunk = (m.head + n - 1) / 16;
        m.nblob = b->d_nblob;
        m.needles = b->d_multi;
        m.groups = b->d_groups;
        m.best = (unsigned long long *)d_offsets;
        m.n_groups = (uint32_t)b->n_groups;
        const unsigned long long max_end = n - b->min_k + 1;
        const unsigned long long n_chunks = (m.head + max_end + 15) / 16;
        const unsigned long long n_seg = (n_chunks + SS_MN_SEG_CHUNKS - 1) / SS_MN_SEG_CHUNKS;
        if (n_seg * b->n_groups > 0x7FFFFFFFull)
            return SS_B200_E_ARG; // haystack x needle table too large for one launch
        multi_needle_kernel<<<(unsigned)(n_seg * b->n_groups), SS_MN_THREADS, 0, st>>>(m);
        ss_host_count_launch(1);
        SS_CUDA(cudaGetLastError());
    }
    if (b->n_multi != nn) { // some needles are empty: N0 => found at 0
        fix_empty_needles_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(b->d_noff, nn,
                                                                               (unsigned long long *)d_offsets);
        ss_host_count_launch(1);
        SS_CUDA(cudaGetLastError());
    }
    return SS_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// synchronous forms: host arrays in and out, on the calling thread's own stream

namespace {
struct Scratch { // per-call device scratch, released on every exit path
    void *p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    ~Scratch()
    {
        for (void *q : p)
            cudaFree(q);
    }
};
} // namespace

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
    SsLane *lane = nullptr;
    int rc = ss_capi_get_lane(&lane);
    if (rc != SS_B200_OK)
        return rc;
    cudaStream_t st = lane->stream;
    Scratch sc;
    const size_t words = (n_pairs + 31) / 32;
    SS_CUDA(cudaMalloc(&sc.p[0], n_pairs * 4));
    SS_CUDA(cudaMalloc(&sc.p[1], n_pairs * 4));
    SS_CUDA(cudaMalloc(&sc.p[2], words * 4));
    if (offsets)
        SS_CUDA(cudaMalloc(&sc.p[3], n_pairs * 8));
    uint32_t *d_pn = (uint32_t *)sc.p[0], *d_ph = (uint32_t *)sc.p[1], *d_bm = (uint32_t *)sc.p[2];
    SS_CUDA(cudaMemcpyAsync(d_pn, pair_needle, n_pairs * 4, cudaMemcpyHostToDevice, st));
    SS_CUDA(cudaMemcpyAsync(d_ph, pair_hay, n_pairs * 4, cudaMemcpyHostToDevice, st));
    rc = ss_b200_batch_search_pairs_async(b, d_pn, d_ph, n_pairs, d_bm, (uint64_t *)sc.p[3], st);
    if (rc != SS_B200_OK)
        return rc;
    if (bitmap)
        SS_CUDA(cudaMemcpyAsync(bitmap, d_bm, words * 4, cudaMemcpyDeviceToHost, st));
    if (offsets)
        SS_CUDA(cudaMemcpyAsync(offsets, sc.p[3], n_pairs * 8, cudaMemcpyDeviceToHost, st));
    SS_CUDA(cudaStreamSynchronize(st));
    return SS_B200_OK;
}

extern "C" int ss_b200_batch_search_triangular(const ss_b200_batch *b, uint32_t *bitmap, uint64_t *matches)
{
    if (!b || !bitmap)
        return SS_B200_E_ARG;
    const unsigned long long w = b->n_hay;
    if (b->n_needles != b->n_hay)
        return SS_B200_E_ARG;
    if (matches)
        *matches = 0;
    if (w == 0)
        return SS_B200_OK;
    SsLane *lane = nullptr;
    int rc = ss_capi_get_lane(&lane);
    if (rc != SS_B200_OK)
        return rc;
    cudaStream_t st = lane->stream;
    const unsigned long long n_pairs = w * (w + 1) / 2;
    const size_t words = (size_t)((n_pairs + 31) / 32);
    // the bitmap scratch is kept with the handle (1.3 MB for the reference's word list); calls that share
    // a handle are serialised here -- the _async form has no such state
    std::lock_guard<std::mutex> lk(b->mu);
    if (b->bitmap_words < words) {
        cudaFree(b->d_bitmap);
        b->d_bitmap = nullptr;
        b->bitmap_words = 0;
        SS_CUDA(cudaMalloc((void **)&b->d_bitmap, words * 4));
        b->bitmap_words = words;
    }
    if (!b->d_count)
        SS_CUDA(cudaMalloc((void **)&b->d_count, 8));
    rc = ss_b200_batch_search_triangular_async(b, b->d_bitmap, (uint64_t *)b->d_count, st);
    if (rc != SS_B200_OK)
        return rc;
    SS_CUDA(cudaMemcpyAsync(bitmap, b->d_bitmap, words * 4, cudaMemcpyDeviceToHost, st));
    unsigned long long m = 0;
    SS_CUDA(cudaMemcpyAsync(&m, b->d_count, 8, cudaMemcpyDeviceToHost, st));
    SS_CUDA(cudaStreamSynchronize(st));
    if (matches)
        *matches = m;
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
    SsLane *lane = nullptr;
    int rc = ss_capi_get_lane(&lane);
    if (rc != SS_B200_OK)
        return rc;
    cudaStream_t st = lane->stream;
    std::lock_guard<std::mutex> lk(b->mu);
    if (!b->d_best)
        SS_CUDA(cudaMalloc((void **)&b->d_best, nn * 8));
    rc = ss_b200_batch_find_all_in_device_async(b, ss_b200_haystack_device_ptr(h), ss_b200_haystack_len(h),
                                                (uint64_t *)b->d_best, st);
    if (rc != SS_B200_OK)
        return rc;
    SS_CUDA(cudaMemcpyAsync(offsets, b->d_best, nn * 8, cudaMemcpyDeviceToHost, st));
    SS_CUDA(cudaStreamSynchronize(st));
    return SS_B200_OK;
}
