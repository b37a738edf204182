// scan_long.cuh -- K1, the long-haystack scan (the roofline kernel).
//
// GPU re-expression of vector_search_in / vector_search_in_chunk
// (reference src/lib.rs:253-287, :199-251): the haystack is cut into 16-byte
// chunks counted from the 16-byte-aligned address at or below hay[0]; every lane
// tests the 16 start positions of one chunk per step with the SWAR two-anchor
// filter (ss_device.cuh), a warp vote collapses the per-lane candidate flags,
// and only warps that saw a candidate enter the exact decode + register-resident
// verify (step_alive_mask + hit_tail).  The leftmost match is kept with one atomicMax on
// ~offset; tiles are visited in ascending order and a tile whose first position
// lies beyond the current best is skipped (the reference's early return,
// src/lib.rs:242-244, made parallel).  The same kernels also serve the count
// mode and the many-haystack mode (different action on a verified match).
//
// Two data paths, same arithmetic:
//   scan_ldg_kernel  coalesced 16-byte LDG straight from HBM/L2 (variant 1)
//   scan_tma_kernel  1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) into a
//                    shared-memory ring guarded by mbarriers; consumers read the
//                    tile with 16-byte LDS (variant 2)
#pragma once
#include "ss_device.cuh"

#define SS_LDG_THREADS 256
#define SS_TMA_CONSUMER_WARPS 8
#define SS_TMA_THREADS ((SS_TMA_CONSUMER_WARPS + 1) * 32)
#define SS_TMA_HALO_MAX 2048 // bytes of right halo a TMA stage can carry

// ------------------------------------------------------------------------------------------
// Template parameters shared by both variants (see ss_device.cuh filter_word):
//   WS, BSZ  second-anchor byte shift R = pos % 16 split as word offset WS = R / 4 (static) and bit
//            shift bs = 8 * (R % 4) (launch-uniform runtime value; BSZ <=> bs == 0, no funnel shift)
//   QZ       pos / 16 == 0: the second-anchor window is the lane's own chunk + the next one
//   K1       one-byte needle (memchr path of src/lib.rs:130-136): first anchor only
//   XK       kind of extra anchors this needle offers (0 none); each warp switches them on and off
//            with the verify-path trip rate it observes (AdaptiveFilter)

// Filter flags for the U chunks of one warp step; returns the OR of the flags.
template <int WS, bool BSZ, bool K1, int XK, int U>
__device__ __forceinline__ uint32_t step_flags(const uint4 (&av)[U], const uint4 (&nx)[U], const uint4 (&lo)[U],
                                               const uint4 (&hi)[U], const FilterConsts &fc, bool extras,
                                               uint32_t (&fl)[U])
{
    uint32_t any = 0;
    if (XK != 0 && extras) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            fl[u] = chunk_flag_x<WS, BSZ, K1, XK>(av[u], nx[u], lo[u], hi[u], fc);
            any |= fl[u];
        }
    } else {
#pragma unroll
        for (int u = 0; u < U; u++) {
            fl[u] = chunk_flag_x<WS, BSZ, K1, 0>(av[u], nx[u], lo[u], hi[u], fc);
            any |= fl[u];
        }
    }
    return any;
}

__device__ __forceinline__ FilterConsts load_filter_consts(const ScanArgs &a)
{
    FilterConsts fc;
    fc.f4 = a.f4;
    fc.l4 = a.l4;
    fc.bs = a.bs;
    fc.e4[0] = a.e4[0];
    fc.e4[1] = a.e4[1];
    fc.xbs = a.xbs;
    return fc;
}

// ------------------------------------------------------------------------------------------
// Variant 1: direct LDG.  CTA tile = WARPS * U * 32 chunks; warp w owns a contiguous run of
// U*32 chunks of it; tiles are dealt blocked-cyclically so the grid sweeps the haystack as a
// moving band (good for early exit and DRAM page locality).
//   U   = chunks per lane per step (independent 16-byte loads in flight)
// One warp step: U*32 chunks starting at chunk `cw` (lane l takes cw + l + 32u).
// CLAMP=false is the interior fast path (every load provably in range: plain base+immediate
// addressing); CLAMP=true clamps each chunk index to the last loadable chunk (tail tiles).
// Returns true when the early-exit test says this CTA can stop.
template <int WS, bool BSZ, bool QZ, bool K1, int XK, int U, bool CLAMP>
__device__ __forceinline__ bool ldg_step(const ScanArgs &a, const uint4 *__restrict__ chunks, unsigned long long cw,
                                         int lane, const FilterConsts &fc, AdaptiveFilter &af, uint32_t &occ,
                                         bool look_left)
{
    // early exit: nothing at or right of this warp's first position can beat the current best
    const unsigned long long key = ld_relaxed_u64(&a.ws->key);
    const unsigned long long last = a.last_chunk;
    const unsigned long long c0 = cw + lane;
    const uint4 *p = chunks + c0;
    const uint4 *pq = p + a.q;
    constexpr bool NEED_HI = !(BSZ && WS == 0); // second-anchor window spills into the following chunk

    // The next chunk `nx` is part of the second-anchor window when QZ; otherwise only the extra
    // anchors and the verify path read it, so it is fetched lazily (warp-uniform conditions).
    constexpr bool NX_EAGER = QZ && !K1;
    auto load_nx = [&](int u) -> uint4 {
        if (CLAMP) {
            const unsigned long long c = c0 + u * 32 + 1;
            return ldg16(chunks + (c < last ? c : last));
        }
        return ldg16(p + u * 32 + 1);
    };
    uint4 av[U], nx[U], lo[U], hi[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (CLAMP) {
            const unsigned long long c = c0 + u * 32;
            av[u] = ldg16(chunks + (c < last ? c : last));
        } else {
            av[u] = ldg16(p + u * 32);
        }
        nx[u] = NX_EAGER ? load_nx(u) : av[u];
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (K1 || QZ) {
            lo[u] = av[u];
            hi[u] = nx[u];
        } else if (CLAMP) {
            const unsigned long long c = c0 + u * 32 + a.q;
            lo[u] = ldg16(chunks + (c < last ? c : last));
            hi[u] = NEED_HI ? ldg16(chunks + (c + 1 < last ? c + 1 : last)) : lo[u];
        } else {
            lo[u] = ldg16(pq + u * 32);
            hi[u] = NEED_HI ? ldg16(pq + u * 32 + 1) : lo[u];
        }
    }
    {
        // every observed value of `key` is a valid bound (it only ever improves), so the warp stops as
        // soon as any lane's reading says so; the vote keeps the decision warp-uniform whatever each
        // lane's load returned
        const long long first_pos = (long long)(cw * 16ull) - (long long)a.head;
        if (__any_sync(0xFFFFFFFFu, (key != 0 && first_pos > (long long)~key) || (look_left && peer_stop_requested(a))))
            return true;
    }
    uint32_t fl[U];
    const bool extras = (XK != 0) && af.begin_tile(); // one step per warp per tile in this variant
    if (!NX_EAGER && !K1 && extras) {
#pragma unroll
        for (int u = 0; u < U; u++)
            nx[u] = load_nx(u);
    }
    const uint32_t any = step_flags<WS, BSZ, K1, XK, U>(av, nx, lo, hi, fc, extras, fl);
    const bool slow = __any_sync(0xFFFFFFFFu, any != 0);
    if (XK != 0)
        af.end_tile(extras, slow ? 4u : 0u);
    if (slow) {
        if (!NX_EAGER && !K1 && !extras) {
#pragma unroll
            for (int u = 0; u < U; u++)
                nx[u] = load_nx(u);
        }
        // every start position of the step in range?  (warp-uniform; lets count mode popcount)
        const bool interior = !CLAMP && cw * 16ull >= a.head && (cw + U * 32) * 16ull - a.head <= a.end;
        occ += step_hits<WS, BSZ, K1, U>(a, av, nx, lo, hi, fl, c0, interior);
    }
    return false;
}

template <int WS, bool BSZ, bool QZ, bool K1, int XK, int U>
__global__ void __launch_bounds__(SS_LDG_THREADS) scan_ldg_kernel(const __grid_constant__ ScanArgs a)
{
    constexpr int WARPS = SS_LDG_THREADS / 32;
    constexpr unsigned long long CTA_CHUNKS = (unsigned long long)WARPS * U * 32;
    const uint4 *chunks = reinterpret_cast<const uint4 *>(a.hay - a.head);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned long long n_tiles = (a.n_chunks + CTA_CHUNKS - 1) / CTA_CHUNKS;
    // tiles [0, n_interior): every chunk holds start positions and every load (incl. the next chunk
    // and the second-anchor window at +q, +q+1) stays at or below last_chunk
    unsigned long long n_interior = a.n_chunks / CTA_CHUNKS;
    {
        const unsigned long long reach = a.q + 1;
        const unsigned long long lim = (a.last_chunk >= reach) ? (a.last_chunk - reach + 1) / CTA_CHUNKS : 0;
        if (lim < n_interior)
            n_interior = lim;
    }
    const FilterConsts fc = load_filter_consts(a);
    AdaptiveFilter af;
    uint32_t occ = 0; // count mode: occurrences seen by this thread

    // Programmatic dependent launch (scan_long.cu launches this variant with programmatic stream
    // serialisation): let the next kernel of the stream start launching now, and touch no global memory
    // (haystack, workspace, result) before every earlier kernel of the stream has completed and flushed.
    // Back-to-back short searches then overlap one launch with the previous scan; without the launch
    // attribute both instructions are no-ops.
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    cta_best_reset();
    __syncthreads();

    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long cw = tile * CTA_CHUNKS + (unsigned long long)warp * (U * 32);
        // the stop word of a sharded search is looked at every eighth tile of the CTA
        const bool look_left = ((tile / gridDim.x) & 7ull) == 0ull;
        bool stop;
        if (tile < n_interior) {
            stop = ldg_step<WS, BSZ, QZ, K1, XK, U, false>(a, chunks, cw, lane, fc, af, occ, look_left);
        } else {
            if (cw >= a.n_chunks)
                continue; // this warp's run holds no start position (warp-uniform)
            stop = ldg_step<WS, BSZ, QZ, K1, XK, U, true>(a, chunks, cw, lane, fc, af, occ, look_left);
        }
        if (stop)
            break; // every later tile of this CTA is further right still
    }
    count_flush(a, occ);
    scan_finish(a);
}

// ------------------------------------------------------------------------------------------
// Many-haystack mode over a prepared set, one warp step (staged variant).  The per-match lookup of the
// plain path (segment_hit: four to six DEPENDENT global loads per match) is what a needle that occurs in
// most haystacks pays for.  Here the warp fetches, once per 4 KiB granule of blob (= its run of a tile),
// the 32 haystack boundaries that follow the granule's first haystack into a shared-memory row -- one
// hint load and one coalesced load of offsets -- and every lane then places its own matches with a
// five-probe search of that row: no global load per match, one byte store per flagged haystack.
// A match beyond the row (more than 32 boundaries inside 6 KiB: clusters of tiny haystacks) falls back
// to segment_hit.  Flags are identical to the plain path.  `m`: the lane's step_alive_mask.
template <bool K1>
__device__ __forceinline__ void many_step(const ScanArgs &a, unsigned long long m, unsigned long long c_lane, int lane,
                                          unsigned long long *row, unsigned long long &row_g,
                                          unsigned long long &row_h)
{
    if (!__any_sync(0xFFFFFFFFu, m != 0))
        return; // false candidates only: nothing to look up
    long long p_first = (long long)((c_lane - lane) * 16ull) - (long long)a.head;
    if (p_first < 0)
        p_first = 0;
    const unsigned long long g = (unsigned long long)p_first >> SS_HINT_SHIFT;
    if (g != row_g) { // warp-uniform
        const unsigned long long h_lo = __ldg(a.seg_hint + g); // last haystack starting at or before the granule
        const unsigned long long idx = h_lo + 1 + lane;
        const unsigned long long e = idx <= a.n_seg ? __ldg(a.seg_off + idx) : ~0ull;
        __syncwarp(); // everybody is done with the previous row
        row[lane] = e;
        row_g = g;
        row_h = h_lo;
        __syncwarp();
    }
    if (m == 0)
        return;
    const unsigned long long row_last = row[31];
    const long long p_lane = (long long)(c_lane * 16ull) - (long long)a.head; // position of byte 0 of chunk 0
    unsigned long long flagged_until = 0; // positions below this lie in a haystack this lane has just flagged
#pragma unroll 1
    while (m) {
        const int bit = __ffsll((long long)m) - 1;
        m &= m - 1;
        const long long i = alive_bit_position(p_lane, bit);
        if (i < 0 || (unsigned long long)i >= a.end || (unsigned long long)i < flagged_until)
            continue;
        if (!K1 && a.k > 17u && !needle_rest_equal(a, a.hay + i, 17u))
            continue;
        if (row_last <= (unsigned long long)i) { // beyond the row: the plain lookup
            flagged_until = segment_hit(a, (unsigned long long)i);
            continue;
        }
        uint32_t cnt = 0; // boundaries <= i among row[0..30] (row[31] > i)
#pragma unroll
        for (uint32_t sft = 16; sft; sft >>= 1)
            if (row[cnt + sft - 1] <= (unsigned long long)i)
                cnt += sft;
        const unsigned long long e = row[cnt]; // end of the haystack holding i
        if ((unsigned long long)i + a.k <= e) {
            a.seg_flags[row_h + cnt] = 1;
            flagged_until = e - a.k + 1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Variant 2: TMA-staged.  One producer warp (one elected lane) streams tiles of TILE bytes plus
// a right halo of 16*max(q + (R>0), 1) bytes into a ring of `stages` shared-memory buffers with
// cp.async.bulk; SS_TMA_CONSUMER_WARPS warps wait on the stage's "full" mbarrier, run the same
// SWAR filter out of shared memory, and release the stage through its "empty" mbarrier.
// Dynamic smem layout: [stages][stage_stride] data, then full[stages], empty[stages] mbarriers,
// then one uint32 "valid" word per stage (0 = producer stopped: early exit or end of work).
template <int WS, bool BSZ, bool QZ, bool K1, int XK, int TILE>
__global__ void __launch_bounds__(SS_TMA_THREADS, 2)
    scan_tma_kernel(const __grid_constant__ ScanArgs a, int stages, uint32_t stage_stride, uint32_t halo)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int CW = SS_TMA_CONSUMER_WARPS;
    constexpr int TILE_CHUNKS = TILE / 16;
    constexpr int WARP_CHUNKS = TILE_CHUNKS / CW; // contiguous run per consumer warp
    constexpr int U = QZ ? 4 : 2; // four live 16-byte vectors per chunk when the second anchor is >= 16 away
    constexpr bool NEED_HI = !(BSZ && WS == 0);
    static_assert(WARP_CHUNKS % (32 * U) == 0, "tile must split into whole 32*U-chunk steps per warp");

    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)stages * stage_stride);
    uint64_t *empty = full + stages;
    volatile uint32_t *valid = reinterpret_cast<volatile uint32_t *>(empty + stages);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned long long n_tiles = (a.n_chunks + TILE_CHUNKS - 1) / TILE_CHUNKS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CW);
        }
        mbar_fence_init();
    }
    cta_best_reset();
    __syncthreads();

    if (warp == CW) {
        // ===== producer =====
        if (lane == 0) {
            const uint64_t policy = l2_policy_evict_first();
            const uint8_t *gbase = a.hay - a.head;
            const unsigned long long data_bytes = (a.last_chunk + 1) * 16ull; // loadable bytes from gbase
            int s = 0;
            uint32_t ph = 0;
            unsigned long long tile = blockIdx.x;
            for (;; tile += gridDim.x) {
                mbar_wait(&empty[s], ph ^ 1);
                bool go = tile < n_tiles;
                if (go) {
                    const unsigned long long key = ld_relaxed_u64(&a.ws->key);
                    const long long first_pos = (long long)(tile * (unsigned long long)TILE) - (long long)a.head;
                    // (the stop word of a sharded search is looked at every eighth tile of the CTA)
                    if ((key && first_pos > (long long)~key) || (((tile / gridDim.x) & 7ull) == 0ull && peer_stop_requested(a)))
                        go = false;
                }
                if (!go) {
                    valid[s] = 0;
                    mbar_arrive(&full[s]); // release: consumers see valid == 0 and leave
                    break;
                }
                const unsigned long long off = tile * (unsigned long long)TILE;
                unsigned long long bytes = (unsigned long long)TILE + halo;
                if (off + bytes > data_bytes)
                    bytes = data_bytes - off; // still a multiple of 16
                valid[s] = 1;
                mbar_arrive_expect_tx(&full[s], (uint32_t)bytes);
                tma_load_1d(smem + (size_t)s * stage_stride, gbase + off, (uint32_t)bytes, &full[s], policy);
                if (++s == stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else {
        // ===== consumers =====
        const FilterConsts fc = load_filter_consts(a);
        AdaptiveFilter af;
        uint32_t occ = 0; // count mode: occurrences seen by this thread
        // many-haystack mode over a prepared set: this warp's row of haystack boundaries (many_step)
        __shared__ unsigned long long seg_rows[CW][32];
        unsigned long long row_g = ~0ull, row_h = 0;
        // count mode, needles of up to three bytes: with its anchors the filter word IS the whole compare
        // (k == 1: first byte; k == 2: both bytes; k == 3: the middle byte is the extra anchor), so a zero
        // byte of it is an occurrence.  While a warp sees occurrences in every step it counts them straight
        // from the filter words -- branch-free, no hit path -- instead of flag + vote + verify.
        const bool exact_kind = a.count != nullptr && a.filter_is_exact != 0u && (XK == 0 || XK == 3);
        bool count_hot = false;
        const uint32_t qb = a.q * 16u;
        int s = 0;
        uint32_t ph = 0;
        for (unsigned long long tile = blockIdx.x;; tile += gridDim.x) {
            mbar_wait(&full[s], ph);
            if (valid[s] == 0)
                break;
            const uint8_t *st = smem + (size_t)s * stage_stride;
            const unsigned long long tile_c0 = tile * (unsigned long long)TILE_CHUNKS;
            constexpr int STEPS = WARP_CHUNKS / (32 * U);
            static_assert(STEPS == 1 || STEPS == 2 || STEPS == 4, "trip share is counted in quarters");
            if (exact_kind && count_hot && tile_c0 * 16ull >= a.head &&
                (tile_c0 + TILE_CHUNKS) * 16ull - a.head <= a.end) {
                // every start position of the tile is in range: count from the filter words
                uint32_t tocc = 0;
#pragma unroll 1
                for (int step = 0; step < STEPS; step++) {
                    const uint32_t lc0 = (uint32_t)warp * WARP_CHUNKS + step * (32 * U) + lane;
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const uint4 cav = lds16(st + (lc0 + u * 32) * 16u);
                        const uint4 cnx = K1 ? cav : lds16(st + (lc0 + u * 32) * 16u + 16u);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            tocc += __popc(swar_zero_exact(filter_word<WS, BSZ, K1, XK>(cav, cnx, cav, cnx, j, fc)));
                    }
                }
                occ += tocc;
                // stay while the warp still finds a few occurrences per step; else back to the filter
                count_hot = __reduce_add_sync(0xFFFFFFFFu, tocc) >= 4u * STEPS;
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&empty[s]);
                if (++s == stages) {
                    s = 0;
                    ph ^= 1;
                }
                continue;
            }
            const bool extras = (XK != 0) && af.begin_tile();
            uint32_t trips = 0;
#pragma unroll 1
            for (int step = 0; step < STEPS; step++) {
                const uint32_t lc0 = (uint32_t)warp * WARP_CHUNKS + step * (32 * U) + lane; // chunk within tile
                // `nx` (the next chunk) is fetched lazily unless it is part of the second-anchor window
                constexpr bool NX_EAGER = QZ && !K1;
                uint4 av[U], nx[U], lo[U], hi[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    av[u] = lds16(st + (lc0 + u * 32) * 16u);
                    nx[u] = NX_EAGER ? lds16(st + (lc0 + u * 32) * 16u + 16u) : av[u];
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (K1 || QZ) {
                        lo[u] = av[u];
                        hi[u] = nx[u];
                    } else {
                        const uint32_t b = (lc0 + u * 32) * 16u + qb;
                        lo[u] = lds16(st + b);
                        hi[u] = NEED_HI ? lds16(st + b + 16u) : lo[u];
                    }
                }
                uint32_t fl[U];
                if (!NX_EAGER && !K1 && extras) {
#pragma unroll
                    for (int u = 0; u < U; u++)
                        nx[u] = lds16(st + (lc0 + u * 32) * 16u + 16u);
                }
                const uint32_t any = step_flags<WS, BSZ, K1, XK, U>(av, nx, lo, hi, fc, extras, fl);
                const bool slow = __any_sync(0xFFFFFFFFu, any != 0);
                trips += slow ? 1u : 0u;
                if (slow) {
                    if (!NX_EAGER && !K1 && !extras) {
#pragma unroll
                        for (int u = 0; u < U; u++)
                            nx[u] = lds16(st + (lc0 + u * 32) * 16u + 16u);
                    }
                    const unsigned long long c_lane = tile_c0 + lc0;
                    if (a.seg_hint != nullptr) { // launch-uniform: many-haystack mode over a prepared set
                        const unsigned long long m = step_alive_mask<WS, BSZ, K1, U>(a, av, nx, lo, hi, fl, c_lane);
                        many_step<K1>(a, m, c_lane, lane, seg_rows[warp], row_g, row_h);
                    } else {
                        // every start position of the step in range?  (warp-uniform; lets count mode popcount)
                        const unsigned long long sc0 = c_lane - lane;
                        const bool interior = sc0 * 16ull >= a.head && (sc0 + 32 * U) * 16ull - a.head <= a.end;
                        occ += step_hits<WS, BSZ, K1, U>(a, av, nx, lo, hi, fl, c_lane, interior);
                    }
                }
            }
            if (XK != 0)
                af.end_tile(extras, trips * (4u / STEPS));
            if (exact_kind)
                count_hot = trips == (uint32_t)STEPS; // every step of the tile saw an occurrence candidate
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&empty[s]);
            if (++s == stages) {
                s = 0;
                ph ^= 1;
            }
        }
        count_flush(a, occ);
    }
    scan_finish(a);
}
