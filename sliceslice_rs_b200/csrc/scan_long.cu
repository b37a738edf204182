// scan_long.cu -- instantiation table and launcher for K1 (see scan_long.cuh).
#include "scan_long.cuh"
#include "ss_host.h"

#include <atomic>
#include <mutex>

namespace {

using LdgFn = void (*)(const ScanArgs);
using TmaFn = void (*)(const ScanArgs, int, uint32_t);

// index: [R][QZ]
template <int U>
struct LdgTable {
    static LdgFn get(int r, bool qz)
    {
#define SS_ROW(R)                                                                                                    \
    case R:                                                                                                          \
        return qz ? (LdgFn)scan_ldg_kernel<R, true, false, U> : (LdgFn)scan_ldg_kernel<R, false, false, U>;
        switch (r) {
            SS_ROW(0) SS_ROW(1) SS_ROW(2) SS_ROW(3) SS_ROW(4) SS_ROW(5) SS_ROW(6) SS_ROW(7) SS_ROW(8) SS_ROW(9)
            SS_ROW(10) SS_ROW(11) SS_ROW(12) SS_ROW(13) SS_ROW(14) SS_ROW(15)
        }
#undef SS_ROW
        return nullptr;
    }
    static LdgFn k1() { return (LdgFn)scan_ldg_kernel<0, true, true, U>; }
};

template <int TILE>
struct TmaTable {
    static TmaFn get(int r, bool qz)
    {
#define SS_ROW(R)                                                                                                    \
    case R:                                                                                                          \
        return qz ? (TmaFn)scan_tma_kernel<R, true, false, TILE> : (TmaFn)scan_tma_kernel<R, false, false, TILE>;
        switch (r) {
            SS_ROW(0) SS_ROW(1) SS_ROW(2) SS_ROW(3) SS_ROW(4) SS_ROW(5) SS_ROW(6) SS_ROW(7) SS_ROW(8) SS_ROW(9)
            SS_ROW(10) SS_ROW(11) SS_ROW(12) SS_ROW(13) SS_ROW(14) SS_ROW(15)
        }
#undef SS_ROW
        return nullptr;
    }
    static TmaFn k1() { return (TmaFn)scan_tma_kernel<0, true, true, TILE>; }
};

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
}

cudaError_t ss_host_launch_scan(const ScanArgs &a, const SsScanTuning &t, const SsDeviceInfo &dev, cudaStream_t stream)
{
    const bool k1 = (a.k == 1);
    const int r = k1 ? 0 : (int)(a.pos % 16);
    const bool qz = k1 || a.pos < 16;
    const unsigned long long scan_bytes = a.n_chunks * 16ull;
    const uint32_t halo = k1 ? 0u : 16u * (a.q + (r > 0 ? 1u : 0u));

    int variant = t.variant;
    if (variant == 0)
        variant = (scan_bytes >= (8ull << 20) && halo <= SS_TMA_HALO_MAX) ? dev.auto_long_variant : 1;
    if (variant == 2 && halo > SS_TMA_HALO_MAX)
        variant = 1; // second anchor too far away for a staged tile; LDG path handles any distance

    if (variant == 2) {
        const int tile_kib = (t.tile_kib == 32) ? 32 : 16;
        const uint32_t tile = (uint32_t)tile_kib * 1024u;
        int stages = t.stages > 0 ? t.stages : 4;
        const uint32_t stage_stride = tile + ((halo + 127u) & ~127u);
        TmaFn fn = (tile_kib == 32) ? (k1 ? TmaTable<32768>::k1() : TmaTable<32768>::get(r, qz))
                                    : (k1 ? TmaTable<16384>::k1() : TmaTable<16384>::get(r, qz));
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
            per_sm = (int)((size_t)dev.smem_per_sm / (smem + 1024));
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
        fn<<<(unsigned)grid, SS_TMA_THREADS, smem, stream>>>(a, stages, stage_stride);
        ss_host_count_launch(1);
        return cudaGetLastError();
    }

    // variant 1: direct LDG
    int u = t.unroll;
    if (u != 1 && u != 2 && u != 4)
        u = (scan_bytes >= (4ull << 20)) ? 4 : 1;
    LdgFn fn;
    if (u == 4)
        fn = k1 ? LdgTable<4>::k1() : LdgTable<4>::get(r, qz);
    else if (u == 2)
        fn = k1 ? LdgTable<2>::k1() : LdgTable<2>::get(r, qz);
    else
        fn = k1 ? LdgTable<1>::k1() : LdgTable<1>::get(r, qz);
    const unsigned long long cta_bytes = (unsigned long long)(SS_LDG_THREADS / 32) * u * 32 * 16;
    const unsigned long long n_tiles = (scan_bytes + cta_bytes - 1) / cta_bytes;
    int per_sm = t.ctas_per_sm > 0 ? t.ctas_per_sm : 6;
    unsigned long long grid = (unsigned long long)dev.sm_count * per_sm;
    if (grid > n_tiles)
        grid = n_tiles;
    if (grid < 1)
        grid = 1;
    fn<<<(unsigned)grid, SS_LDG_THREADS, 0, stream>>>(a);
    ss_host_count_launch(1);
    return cudaGetLastError();
}
