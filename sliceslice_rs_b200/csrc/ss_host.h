// ss_host.h -- internal host-side declarations shared by the translation units of
// libsliceslice_b200.so (not installed; the public surface is include/sliceslice_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ss_device.cuh"

struct SsScanTuning {
    int variant = 0;     // 0 auto, 1 LDG, 2 TMA
    int ctas_per_sm = 0; // 0 auto
    int unroll = 0;      // LDG: chunks per lane per step (1, 2, 4); 0 auto
    int tile_kib = 0;    // TMA: 16 or 32; 0 auto
    int stages = 0;      // TMA ring depth; 0 auto
    int extra_anchors = -1; // 0 = never use extra anchors, -1 = auto (adaptive, see AdaptiveFilter)
    int pdl = 1;         // short-scan variant launched with programmatic stream serialisation (1) or plainly (0)
};

struct SsDeviceInfo {
    int device = -1;
    int sm_count = 0;
    int max_smem_optin = 0;
    int smem_per_sm = 0;
    int auto_long_variant = 2; // variant picked by "auto" for long haystacks (TMA ring; measured)
};

using SsLdgFn = void (*)(const ScanArgs);
using SsTmaFn = void (*)(const ScanArgs, int, uint32_t, uint32_t);
// kernel tables, one translation unit each (scan_ldg_u1.cu, scan_ldg_u4.cu, scan_tma_16.cu, scan_tma_32.cu)
SsLdgFn ss_table_ldg_u1(int ws, bool bsz, bool qz, bool k1, int xk);
SsLdgFn ss_table_ldg_u4(int ws, bool bsz, bool qz, bool k1, int xk);
SsTmaFn ss_table_tma_16(int ws, bool bsz, bool qz, bool k1, int xk);
SsTmaFn ss_table_tma_32(int ws, bool bsz, bool qz, bool k1, int xk);

uint64_t ss_host_launch_count();
void ss_host_count_launch(uint64_t n);
void ss_host_scan_geometry(ScanArgs &a, unsigned long long start_limit);
cudaError_t ss_host_launch_scan(const ScanArgs &a, const SsScanTuning &t, const SsDeviceInfo &dev, cudaStream_t stream);

// gen.cu
cudaError_t ss_host_fill_random(void *d_dst, size_t len, uint64_t global_start, uint64_t seed, int sm_count,
                                cudaStream_t stream);
cudaError_t ss_host_fill_tiled(void *d_dst, size_t len, uint64_t global_start, const void *d_src, size_t src_len,
                               int sm_count, cudaStream_t stream);
