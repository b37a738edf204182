// LDG variant, four chunks per lane per step.
#include "scan_tables.cuh"
SS_DEFINE_TABLE(ss_table_ldg_u4, scan_ldg_kernel, SsLdgFn, 4)
