// TMA variant, 32 KiB tiles.
#include "scan_tables.cuh"
SS_DEFINE_TABLE(ss_table_tma_32, scan_tma_kernel, SsTmaFn, 32768)
