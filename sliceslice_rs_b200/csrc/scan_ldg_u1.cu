// LDG variant, one chunk per lane per step (short haystacks: more, smaller tiles).
#include "scan_tables.cuh"
SS_DEFINE_TABLE(ss_table_ldg_u1, scan_ldg_kernel, SsLdgFn, 1)
