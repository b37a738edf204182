// scan_tables.cuh -- instantiation tables for K1 (scan_long.cuh).  Each translation unit
// scan_ldg_u*.cu / scan_tma_*.cu instantiates one table (65 kernels) so they compile in parallel.
#pragma once
#include "scan_long.cuh"
#include "ss_host.h"

// Expands to a function `NAME(ws, bsz, qz, k1, xk)` returning the kernel KERNEL<WS,BSZ,QZ,K1,XK,LAST>.
#define SS_TAB_NE(KERNEL, FN, LAST, WS, BSZ, QZ)                                                                     \
    switch (xk) {                                                                                                    \
    case 0: return (FN)KERNEL<WS, BSZ, QZ, false, 0, LAST>;                                                          \
    case 1: return (FN)KERNEL<WS, BSZ, QZ, false, 1, LAST>;                                                          \
    case 2: return (FN)KERNEL<WS, BSZ, QZ, false, 2, LAST>;                                                          \
    default: return (FN)KERNEL<WS, BSZ, QZ, false, 3, LAST>;                                                         \
    }
#define SS_TAB_QZ(KERNEL, FN, LAST, WS, BSZ)                                                                         \
    if (qz) {                                                                                                        \
        SS_TAB_NE(KERNEL, FN, LAST, WS, BSZ, true)                                                                   \
    } else {                                                                                                         \
        SS_TAB_NE(KERNEL, FN, LAST, WS, BSZ, false)                                                                  \
    }
#define SS_TAB_BSZ(KERNEL, FN, LAST, WS)                                                                             \
    if (bsz) {                                                                                                       \
        SS_TAB_QZ(KERNEL, FN, LAST, WS, true)                                                                        \
    } else {                                                                                                         \
        SS_TAB_QZ(KERNEL, FN, LAST, WS, false)                                                                       \
    }
#define SS_DEFINE_TABLE(NAME, KERNEL, FN, LAST)                                                                      \
    FN NAME(int ws, bool bsz, bool qz, bool k1, int xk)                                                              \
    {                                                                                                                \
        if (k1)                                                                                                      \
            return (FN)KERNEL<0, true, true, true, 0, LAST>;                                                         \
        switch (ws) {                                                                                                \
        case 0: SS_TAB_BSZ(KERNEL, FN, LAST, 0)                                                                      \
        case 1: SS_TAB_BSZ(KERNEL, FN, LAST, 1)                                                                      \
        case 2: SS_TAB_BSZ(KERNEL, FN, LAST, 2)                                                                      \
        default: SS_TAB_BSZ(KERNEL, FN, LAST, 3)                                                                     \
        }                                                                                                            \
    }
