// TMA variant, 16 KiB tiles.
#include "scan_tables.cuh"
SS_DEFINE_TABLE(ss_table_tma_16, scan_tma_kernel, SsTmaFn, 16384)
