// scan_long.cu -- geometry and launcher for K1 (kernels: scan_long.cuh, tables: scan_tables.cuh).
#include "scan_long.cuh"
#include "ss_host.h"

#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

namespace {
std::atomic<uint64_t> g_launches{0};
} // namespace

uint64_t ss_host_launch_count() { return g_launches.load(std::memory_order_relaxed); }
void ss_host_count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Fill the geometry fields of ScanArgs from (hay, n, k, pos, start_limit).
void ss_host_scan_geometry(ScanArgs &a, unsigned long long start_limit)
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(a.hay);
    a.head = (uint32_t)(addr & 15);
    unsigned long long end = a.n - a.k + 1;
    if (start_limit < end)
        end = start_limit;
    a.end = end;
    a.n_chunks = (a.head + end + 15) / 16;
    a.last_chunk = (a.head + a.n - 1) / 16;
    a.q = a.pos / 16;
    a.bs = 8u * (a.pos % 4u);
}

// Choose the extra anchors the filter may switch on (filter_word in ss_device.cuh): word-aligned
// needle offsets 4 and 8 when the needle has them, else one unaligned offset for short needles.
// Offsets equal to 0 or `pos` would repeat an anchor and are skipped.  Needs the first bytes of the
// needle in a.needle_inline.
static void choose_extra_anchors(ScanArgs &a, int allow)
{
    a.xk = 0;
    a.e4[0] = a.e4[1] = 0;
    a.xbs = 0;
    if (allow == 0 || a.k < 3)
        return; // k == 2: the two anchors already are the whole needle
    const bool has4 = a.k > 4 && a.pos != 4;
    const bool has8 = a.k > 8 && a.pos != 8;
    if (has4 && has8) {
        a.xk = 2;
        a.e4[0] = 0x01010101u * a.needle_inline[4];
        a.e4[1] = 0x01010101u * a.needle_inline[8];
    } else if (has4) {
        a.xk = 1;
        a.e4[0] = 0x01010101u * a.needle_inline[4];
    } else {
        // unaligned: the needle offset in 1..3 furthest from both anchors
        uint32_t best = 0;
        for (uint32_t o = 1; o <= 3 && o < a.k; o++)
            if (o != a.pos && (best == 0 || o == 2))
                best = o;
        if (best) {
            a.xk = 3;
            a.e4[0] = 0x01010101u * a.needle_inline[best];
            a.xbs = 8u * best;
        }
    }
}

// cudaOccupancyMaxActiveBlocksPerMultiprocessor per (kernel, dynamic smem), cached
static int ss_tma_occupancy(const void *fn, size_t smem)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, size_t>, int> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair(fn, smem);
    auto it = cache.find(key);
    if (it != cache.end())
        return it->second;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, SS_TMA_THREADS, smem) != cudaSuccess) {
        cudaGetLastError();
        n = 2;
    }
    cache[key] = n;
    return n;
}

cudaError_t ss_host_launch_scan(const ScanArgs &a_in, const SsScanTuning &t, const SsDeviceInfo &dev,
                                cudaStream_t stream)
{
    ScanArgs a = a_in;
    const bool k1 = (a.k == 1);
    const int r = k1 ? 0 : (int)(a.pos % 16);
    const int ws = r / 4;
    const bool bsz = (r % 4) == 0;
    const bool qz = k1 || a.pos < 16;
    const unsigned long long scan_bytes = a.n_chunks * 16ull;
    // right halo of a staged tile: the next chunk (register window of the verify path and the
    // extra anchors) and the second-anchor window at +q, +q+1
    const uint32_t reach = k1 ? 0u : (a.q + (r > 0 ? 1u : 0u));
    const uint32_t halo = k1 ? 0u : 16u * (reach > 1u ? reach : 1u);

    choose_extra_anchors(a, t.extra_anchors);
    const int xk = (int)a.xk;
    // count mode counts needles of up to three bytes straight from the filter words when the anchors of this
    // launch compare every needle byte
    a.filter_is_exact = filter_covers_needle(a.k, a.pos, xk, a.xbs / 8u) ? 1u : 0u;

    int variant = t.variant;
    if (variant == 0)
        variant = (scan_bytes >= (8ull << 20) && halo <= SS_TMA_HALO_MAX) ? dev.auto_long_variant : 1;
    if (variant == 2 && halo > SS_TMA_HALO_MAX)
        variant = 1; // second anchor too far away for a staged tile; LDG path handles any distance

    if (variant == 2) {
        // measured (profiles/r02_tile_choice.txt; both run two CTAs per SM, register-bound): 32 KiB x 3 stages
        // wins or ties from 256 MiB up (6.1 vs 5.7-5.9 TB/s at 256 MiB, 6.7 vs 6.3 at 1 GiB), 16 KiB x 6 stages
        // below (more, smaller tiles spread a short scan better: 4.1 vs 3.7 TB/s at 64 MiB)
        int tile_kib = t.tile_kib;
        if (tile_kib != 16 && tile_kib != 32)
            tile_kib = (scan_bytes >= (128ull << 20)) ? 32 : 16;
        const uint32_t tile = (uint32_t)tile_kib * 1024u;
        // two CTAs per SM either way (register-bound): ~192 KB of loads in flight per SM
        int stages = t.stages > 0 ? t.stages : (tile_kib == 32 ? 3 : 6);
        const uint32_t stage_stride = tile + ((halo + 127u) & ~127u);
        SsTmaFn fn = (tile_kib == 32) ? ss_table_tma_32(ws, bsz, qz, k1, xk) : ss_table_tma_16(ws, bsz, qz, k1, xk);
        size_t smem = (size_t)stages * stage_stride + (size_t)stages * 16 + (size_t)stages * 4 + 16;
        while (smem > (size_t)dev.max_smem_optin && stages > 2) {
            stages--;
            smem = (size_t)stages * stage_stride + (size_t)stages * 16 + (size_t)stages * 4 + 16;
        }
        cudaError_t e = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
        int per_sm = t.ctas_per_sm;
        if (per_sm <= 0) {
            // resident CTAs per SM as the hardware will grant them (shared memory AND registers): a grid
            // larger than one wave would leave CTAs queued behind persistent ones
            per_sm = ss_tma_occupancy((const void *)fn, smem);
            if (per_sm < 1)
                per_sm = 1;
            if (per_sm > 4)
                per_sm = 4;
        }
        const unsigned long long n_tiles = (scan_bytes + tile - 1) / tile;
        unsigned long long grid = (unsigned long long)dev.sm_count * per_sm;
        if (grid > n_tiles)
            grid = n_tiles;
        if (grid < 1)
            grid = 1;
        fn<<<(unsigned)grid, SS_TMA_THREADS, smem, stream>>>(a, stages, stage_stride, halo);
        ss_host_count_launch(1);
        return cudaGetLastError();
    }

    // variant 1: direct LDG
    int u = t.unroll;
    if (u != 1 && u != 4)
        u = (scan_bytes >= (4ull << 20)) ? 4 : 1;
    SsLdgFn fn = (u == 4) ? ss_table_ldg_u4(ws, bsz, qz, k1, xk) : ss_table_ldg_u1(ws, bsz, qz, k1, xk);
    const unsigned long long cta_bytes = (unsigned long long)(SS_LDG_THREADS / 32) * u * 32 * 16;
    const unsigned long long n_tiles = (scan_bytes + cta_bytes - 1) / cta_bytes;
    int per_sm = t.ctas_per_sm > 0 ? t.ctas_per_sm : 6;
    unsigned long long grid = (unsigned long long)dev.sm_count * per_sm;
    if (grid > n_tiles)
        grid = n_tiles;
    if (grid < 1)
        grid = 1;
    // programmatic stream serialisation: this launch may begin while the previous kernel of the stream is
    // still running; the kernel itself waits (griddepcontrol.wait) before its first global access.  Short
    // searches issued back to back are launch-bound, and this hides most of the launch
    // (ss_b200_set_launch_pdl(0) falls back to a plain launch).
    const bool pdl = t.pdl != 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(SS_LDG_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    void *args[] = {(void *)&a};
    const cudaError_t e = cudaLaunchKernelExC(&cfg, (const void *)fn, args);
    ss_host_count_launch(1);
    return e != cudaSuccess ? e : cudaGetLastError();
}
