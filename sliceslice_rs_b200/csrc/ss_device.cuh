// ss_device.cuh -- shared device-side definitions for the sm_100a scan kernels.
//
// Replaces, on the GPU, the instruction mix of the reference's AVX2 `Vector`
// impl (src/x86.rs:202-235: vmovdqu / vpcmpeqb / vpand / vpmovmskb) and the
// block loop of src/lib.rs:199-287.  GPUs have no byte-compare/movemask, so the
// two-anchor filter is expressed as 32-bit SWAR over 16-byte chunks:
//     x = (A ^ splat(first)) | (B ^ splat(last))      B = haystack shifted by `position`
//     any zero byte in x  <=>  candidate              (x - 0x01010101) & ~x & 0x80808080
// The SWAR test is exact as an "any candidate in this word" test (false positives
// only above a true zero byte), so no candidate is ever missed; exact per-byte
// decode and the memcmp verify (src/lib.rs:216-244) run on the hit path, out of the
// same registers (step_alive_mask).  Extra anchors may be folded into the filter where
// candidates are frequent (filter_word, AdaptiveFilter); they never change a result.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SS_INLINE_NEEDLE_MAX 64
#define SS_MAX_PEERS 16
#define SS_MAILBOX_DEPTH 4                         // result slots per (rank) in flight, indexed by seq % 4
#define SS_MAILBOX_EMPTY 0xFFFFFFFFFFFFFFFFull     // never a result: offsets and NONE are <= INT64_MAX
#define SS_HINT_SHIFT 12                           // segment lookup hints: one per 4 KiB of blob
#define SS_HINT_GRANULE (1ull << SS_HINT_SHIFT)
#define SS_NONE_U64 0x7FFFFFFFFFFFFFFFull // SS_B200_DEVICE_NONE
#define SS_RESULT_PENDING 0xFFFFFFFFFFFFFFFFull // host-side marker of a mapped result slot before the kernel wrote it

// 16-byte self-resetting per-stream workspace.  `key` holds ~(smallest local
// offset found so far) so that ZERO means "nothing found": a zero-filled
// allocation is a valid initial state and the last CTA restores it.
struct SsWorkspace {
    unsigned long long key;
    unsigned int done;
    unsigned int pad;
};

struct ScanArgs {
    const uint8_t *hay;          // haystack bytes (any alignment), device memory
    unsigned long long n;        // haystack length
    unsigned long long end;      // number of start positions to test (>= 1, <= n - k + 1)
    unsigned long long base;     // global coordinate of hay[0]
    unsigned long long n_chunks; // 16-byte chunks (from the aligned-down base) holding >= 1 start position
    unsigned long long last_chunk; // index of the chunk holding hay[n-1]; loads are clamped to it
    const uint8_t *needle_g;     // device copy of the needle when k > SS_INLINE_NEEDLE_MAX, else nullptr
    SsWorkspace *ws;
    unsigned long long *out;     // result slot: base + first offset, or SS_NONE_U64
    uint32_t k;    // needle length (>= 1)
    uint32_t pos;  // second anchor index (`position`), < k
    uint32_t q;    // pos / 16
    uint32_t head; // hay - align_down(hay, 16), 0..15
    uint32_t f4;   // needle[0] splatted x4
    uint32_t l4;   // needle[pos] splatted x4
    uint32_t bs;   // 8 * (pos % 4): bit shift of the second-anchor stream inside a word
    uint32_t xk;   // extra-anchor kind the kernel was instantiated with (see filter_word)
    uint32_t e4[2]; // extra anchor bytes splatted x4
    uint32_t xbs;  // xk == 3: 8 * needle offset (1..3) of the unaligned extra anchor
    // 1 when the anchors of the kernel's filter (first, `pos`, and the extra ones of kind xk) cover EVERY
    // needle byte, i.e. a zero byte of the filter word is an occurrence (needles of up to three bytes)
    uint32_t filter_is_exact;
    // peer exchange (n_peers == 0: off): the last CTA also stores the result into slot
    // [seq % 4][rank] of every rank's mailbox (peer memory over NVLink), fused into the scan epilogue
    unsigned long long *peer_slot[SS_MAX_PEERS];
    uint32_t n_peers;
    // cross-GPU early exit of a sharded first-match search (n_stop_peers == 0 and stop_word == nullptr:
    // off).  The first verified match on this GPU stores `stop_seq` into the stop word of every shard to
    // its RIGHT (peer memory): whatever those find is larger than this match, so they may stop -- the
    // reference's early return (src/lib.rs:242-244) across devices.  This GPU polls its own stop word
    // where it polls `key`; a value other than the current search's sequence number is stale and ignored.
    uint32_t n_stop_peers;
    unsigned long long *stop_peer[SS_MAX_PEERS];
    const unsigned long long *stop_word;
    unsigned long long stop_seq;
    // many-haystack mode (nullptr otherwise): hay is the concatenation of n_seg haystacks, haystack h
    // = bytes [seg_off[h], seg_off[h+1]) with seg_off[0] == 0; seg_flags[h] <- 1 when it contains the needle
    const unsigned long long *seg_off;
    uint8_t *seg_flags;
    unsigned long long n_seg;
    // optional lookup hints of a prepared set (ss_b200_hayset): seg_hint[g] = index of the haystack
    // holding blob byte g * SS_HINT_GRANULE, for g < n_gran; nullptr = plain binary search
    const uint32_t *seg_hint;
    unsigned long long n_gran;
    // count mode (nullptr otherwise): incremented once per occurrence
    unsigned long long *count;
    uint8_t needle_inline[SS_INLINE_NEEDLE_MAX]; // first min(k, 64) needle bytes
    // needle bytes 1..16 splatted over a word each (index 0 unused): the register verify of the hit path
    // takes them straight from the constant bank as instruction operands
    uint32_t needle4[17];
};

__device__ __forceinline__ uint4 ldg16(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

#include "ss_filter.cuh"

// Has a shard to the left already matched (see ScanArgs::stop_word)?  Launch-uniform branch on the pointer.
// The word lives in THIS GPU's memory; peers store into it over NVLink, which lands in this GPU's L2 -- so a
// gpu-scope load (what `key` is polled with) sees it.  Callers poll it every few tiles only: the first
// version polled it with a sys-scope load before every tile and cost the sharded scan 3-4 %
// (1.296 vs 1.247 ms per 8 GiB search, profiles/r02_bench_n2.json).
__device__ __forceinline__ bool peer_stop_requested(const ScanArgs &a)
{
    if (a.stop_word == nullptr)
        return false;
    return ld_relaxed_u64(a.stop_word) == a.stop_seq;
}

// Per-warp, per-tile switch between the plain two-anchor filter and the one with extra anchors.
// The plain filter is cheapest when candidates are rare (random bytes); on natural text a common
// anchor pair sends almost every warp step through the verify path and the extra anchors pay for
// themselves.  `s` (0..4) is the share of a tile's steps that took the verify path, in quarters.
//   PLAIN  (period == 0): est <- 3/4 est + 4 s settles at 16 s; two consecutive all-trip tiles (or a
//          sustained trip rate above ~40%) reach 24 and switch the extras on.
//   EXTRAS (period != 0): every `period`-th tile runs the plain filter as a probe.  A probe with fewer
//          than half of its steps tripping returns to PLAIN at once; otherwise the period doubles
//          (4, 8, ... 128 tiles), so on text the probes cost ~1% of the tiles.
// The bookkeeping runs once per tile, outside the step loop.  All fields are warp-uniform.
struct AdaptiveFilter {
    uint32_t est = 0;
    uint32_t period = 0;
    uint32_t left = 0;
    __device__ __forceinline__ bool begin_tile()
    {
        if (period == 0u || left == 0u)
            return false;
        left--;
        return true;
    }
    __device__ __forceinline__ void end_tile(bool used_extras, uint32_t s)
    {
        if (used_extras)
            return;
        if (period == 0u) {
            est = est - (est >> 2) + 4u * s;
            if (est >= 24u)
                period = left = 4u;
        } else if (s >= 2u) {
            period = period < 128u ? period * 2u : 128u;
            left = period;
        } else {
            period = 0u;
            est = 0u;
        }
    }
};

// haystack loads of the slow tail: through L2 only (ld.global.cg), because the resident service kernel
// (service.cu) outlives host-side writes to the haystack and must never see a stale L1 line
__device__ __forceinline__ uint32_t ld_hay_u8(const uint8_t *p)
{
    uint32_t v;
    asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned long long ld_hay_u64(const uint8_t *p)
{
    unsigned long long v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// memcmp of needle[from..k) against h[from..k) from global memory -- only reached by needles longer
// than the 17 bytes the register window covers, after those 17 bytes already matched.
static __device__ __noinline__ bool needle_rest_equal(const ScanArgs &a, const uint8_t *h, uint32_t from)
{
    const uint32_t k = a.k;
    if (k <= SS_INLINE_NEEDLE_MAX) {
        for (uint32_t j = from; j < k; j++)
            if (ld_hay_u8(h + j) != a.needle_inline[j])
                return false;
        return true;
    }
    const uint8_t *nd = a.needle_g;
    uint32_t j = from;
    // byte steps until h + j is 8-byte aligned, then 8 bytes of haystack per step
    for (; j < k && ((reinterpret_cast<uintptr_t>(h + j)) & 7); j++)
        if (ld_hay_u8(h + j) != __ldg(nd + j))
            return false;
    for (; j + 8 <= k; j += 8) {
        unsigned long long hv = ld_hay_u64(h + j);
        unsigned long long nv = 0;
#pragma unroll
        for (int t = 0; t < 8; t++)
            nv |= (unsigned long long)__ldg(nd + j + t) << (8 * t);
        if (hv != nv)
            return false;
    }
    for (; j < k; j++)
        if (ld_hay_u8(h + j) != __ldg(nd + j))
            return false;
    return true;
}

// The CTA's best (smallest) verified offset of a first-match search, in shared memory: every matching
// lane improves it with a shared-memory atomic, and only a lane that lowers it touches the global word
// (a common word matches in every tile at once; thousands of same-address global atomics were what a
// found search paid for).  Reset by the kernels before the first tile.
static __shared__ unsigned long long ss_cta_best;

__device__ __forceinline__ void cta_best_reset()
{
    if (threadIdx.x == 0)
        ss_cta_best = ~0ull;
}

__device__ __forceinline__ uint8_t ld_relaxed_u8(const uint8_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return (uint8_t)v;
}

// Many-haystack mode: mark the haystack that wholly contains the match at blob position i.  Returns the
// first start position behind which a match can no longer fall into that haystack (so the caller can
// skip the other occurrences inside it), or 0 when nothing was flagged.
static __device__ __noinline__ unsigned long long segment_hit(const ScanArgs &a, unsigned long long i)
{
    // h = last segment with seg_off[h] <= i
    unsigned long long lo = 0, hi = a.n_seg; // invariant: seg_off[lo] <= i < seg_off[hi]
    if (a.seg_hint) {
        // the haystacks holding the first byte of this granule and of the next one bracket the answer:
        // the 20-odd dependent probes of a search over the whole table become one or two
        const unsigned long long g = i >> SS_HINT_SHIFT;
        lo = __ldg(a.seg_hint + g);
        if (g + 1 < a.n_gran)
            hi = (unsigned long long)__ldg(a.seg_hint + g + 1) + 1;
    }
    while (hi - lo > 1) {
        const unsigned long long mid = (lo + hi) >> 1;
        if (__ldg(a.seg_off + mid) <= i)
            lo = mid;
        else
            hi = mid;
    }
    const unsigned long long e = __ldg(a.seg_off + lo + 1);
    if (i + a.k > e)
        return 0; // the match straddles the end of the haystack: it belongs to nobody
    a.seg_flags[lo] = 1;
    return e - a.k + 1;
}

// Hit path of one warp step, part 1 (inlined; no memory access): the exact compare of every flagged
// chunk of the lane out of its registers -- the ctz loop + memcmp of src/lib.rs:216-248 for 16 start
// positions at once (exact_alive; needle bytes as constant-bank operands).  The lane already holds the 32
// haystack bytes [16c, 16c+32) of each of its U chunks, which cover needle bytes 0..16 of every start
// position.  Returns one bit per start position whose first min(k, 17) bytes equal the needle's:
// bit 16u + p <-> position p of the lane's chunk u (chunk u lies 32 chunks = 512 bytes behind chunk 0).
// The false candidates of natural text leave after one needle byte and give 0.
template <int WS, bool BSZ, bool K1, int U>
__device__ __forceinline__ unsigned long long step_alive_mask(const ScanArgs &a, const uint4 (&av)[U],
                                                              const uint4 (&nx)[U], const uint4 (&lo)[U],
                                                              const uint4 (&hi)[U], const uint32_t (&fl)[U],
                                                              unsigned long long c_lane)
{
    FilterConsts fc;
    fc.f4 = a.f4;
    fc.l4 = a.l4;
    fc.bs = a.bs;
    unsigned long long m = 0;
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (fl[u] && c_lane + u * 32 < a.n_chunks) {
            uint32_t z[4];
            if (exact_alive<WS, BSZ, K1>(av[u], nx[u], lo[u], hi[u], fc, a.k, [&](uint32_t j) { return a.needle4[j]; }, z))
                m |= (unsigned long long)pack_alive16(z) << (16 * u);
        }
    }
    return m;
}

// Position of bit `bit` of a step_alive_mask whose chunk 0 starts at position p_lane.
__device__ __forceinline__ long long alive_bit_position(long long p_lane, int bit)
{
    return p_lane + (long long)(bit >> 4) * 512 + (bit & 15);
}

// Hit path, part 2: what is left once a lane holds positions whose first min(k, 17) bytes equal the
// needle's -- rare unless the needle really occurs.  One compact loop over the set bits: range check,
// the rest of a long needle from memory, and the action of the mode.
//   first match  leftmost offset, early exit: CTA-level best first, the global word only when this lane
//                lowered it, the stop words of the shards to the right on the GPU's first match
//   count        every occurrence (overlapping ones included), no early exit; returns the number counted
//                (the caller keeps the running total in a register, count_flush adds it once per warp)
//   many         the blob is a concatenation of haystacks; flag the haystack that wholly contains the
//                match (plain lookup; the staged variant places matches of a prepared set itself), no early exit
static __device__ __noinline__ uint32_t hit_tail(const ScanArgs &a, unsigned long long m, long long p_lane)
{
    const bool whole_needle_compared = a.k <= 17u;
    uint32_t occ = 0;
    unsigned long long flagged_until = 0; // many mode: positions below this lie in a haystack flagged just now
#pragma unroll 1
    while (m) {
        const int bit = __ffsll((long long)m) - 1;
        m &= m - 1;
        const long long i = alive_bit_position(p_lane, bit);
        if (i < 0 || (unsigned long long)i >= a.end)
            continue;
        if (a.seg_off != nullptr && (unsigned long long)i < flagged_until)
            continue;
        if (!whole_needle_compared && !needle_rest_equal(a, a.hay + i, 17u))
            continue;
        if (a.count != nullptr) {
            occ++;
            continue;
        }
        if (a.seg_off != nullptr) {
            flagged_until = segment_hit(a, (unsigned long long)i);
            continue;
        }
        // No fence: `key` is only ever touched with atomics and relaxed loads, and the acq_rel ticket of
        // scan_finish (behind a CTA barrier) orders every atomicMax before the final read.
        const unsigned long long old = atomicMin(&ss_cta_best, (unsigned long long)i);
        if ((unsigned long long)i < old) {
            const unsigned long long prev = atomicMax(&a.ws->key, ~(unsigned long long)i);
            if (prev == 0ull) {
                // first match on this GPU: the shards to the right can stop (8-byte NVLink stores)
                for (uint32_t p = 0; p < a.n_stop_peers; p++)
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a.stop_peer[p]), "l"(a.stop_seq)
                                 : "memory");
            }
        }
        return 0; // ascending order: later positions of this lane cannot be smaller
    }
    return occ;
}

// Parts 1 + 2 for the modes that need no warp cooperation.  Count mode takes the survivors of an interior
// step with one popcount (`interior`: every start position of the step is in range -- warp-uniform).
template <int WS, bool BSZ, bool K1, int U>
__device__ __forceinline__ uint32_t step_hits(const ScanArgs &a, const uint4 (&av)[U], const uint4 (&nx)[U],
                                              const uint4 (&lo)[U], const uint4 (&hi)[U], const uint32_t (&fl)[U],
                                              unsigned long long c_lane, bool interior)
{
    const unsigned long long m = step_alive_mask<WS, BSZ, K1, U>(a, av, nx, lo, hi, fl, c_lane);
    if (m == 0)
        return 0;
    if (a.count != nullptr && a.k <= 17u && interior)
        return (uint32_t)__popcll(m); // the register window covered the whole needle: every bit is an occurrence
    return hit_tail(a, m, (long long)(c_lane * 16ull) - (long long)a.head);
}

// Count mode epilogue: add the warp's occurrences to *a.count with one atomic.  Called by whole,
// converged warps before scan_finish.
__device__ __forceinline__ void count_flush(const ScanArgs &a, uint32_t occ)
{
    if (a.count == nullptr)
        return; // launch-uniform
    const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, occ);
    if ((threadIdx.x & 31) == 0 && total)
        atomicAdd(a.count, (unsigned long long)total);
}

// Grid-wide epilogue: the last CTA to finish publishes the result and restores the
// workspace to all-zero so the next launch on this stream needs no memset.
__device__ __forceinline__ void scan_finish(const ScanArgs &a)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        // acq_rel ticket: the release half orders this CTA's atomicMax updates of `key` (made visible to
        // this thread by the barrier above) before the ticket, the acquire half lets the CTA that draws
        // the last ticket read every other CTA's updates.  Cheaper than two sequentially consistent
        // __threadfence()s on the critical path of short scans.
        unsigned int prev;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(&a.ws->done) : "memory");
        if (prev == gridDim.x - 1) {
            const unsigned long long key = atomicExch(&a.ws->key, 0ull);
            a.ws->done = 0;
            const unsigned long long r = key ? (a.base + ~key) : SS_NONE_U64;
            // one 8-byte store: the slot may be device memory or a mapped pinned host word the caller
            // spins on (it holds SS_RESULT_PENDING until this store lands; no fence or second flag needed)
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a.out), "l"(r) : "memory");
            // fused exchange: one 8-byte store per rank, straight into peer HBM (no fence needed:
            // the slot value itself is the arrival signal, see mailbox_min_kernel)
            for (uint32_t p = 0; p < a.n_peers; p++)
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a.peer_slot[p]), "l"(r) : "memory");
        }
    }
}

// ---- mbarrier / TMA (cp.async.bulk) PTX wrappers -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                 "selp.u32 %0, 1, 0, p;\n"
                 "}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
// Requires 16-byte aligned src, dst and size.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar,
                                            uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, "
                 "[%3], %4;" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 lds16(const void *p)
{
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "r"(smem_u32(p)));
    return r;
}
