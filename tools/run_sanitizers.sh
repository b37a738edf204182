#!/usr/bin/env bash
# compute-sanitizer passes over subsets of the GPU parity tests (the ASan analogue of the reference's CI
# job, .github/workflows/check.yml:42-58).  Run on a GPU box from the repository root:
#   gpurun --timeout 1700 -- 'bash tools/run_sanitizers.sh'
# Logs land in gpurun_out/; copy the summaries you want to keep into profiles/.
set -u
mkdir -p gpurun_out
T=tests/test_gpu_parity.py
run() { # tool, -k expression
    timeout 900 compute-sanitizer --tool "$1" --error-exitcode 9 python -m pytest $T -x -q -k "$2" \
        > "gpurun_out/sanitizer_$1.log" 2>&1
    echo "$1 rc=$?"; tail -3 "gpurun_out/sanitizer_$1.log"
}
run memcheck  "kats or memchr or edge or mula or unaligned or async_entry or many_haystack or short_sweep or batched_single or pairs_mode or random_bench or count_mode or histogram or prepared"
run racecheck "mula or unaligned or async_entry or histogram or prepared"
run synccheck "mula or unaligned or async_entry or count_mode or histogram"
run initcheck "kats or batched_single or ipsum_absent or memchr or edge or short_sweep or pairs_mode or histogram or prepared or count_mode"
