#!/usr/bin/env bash
# compute-sanitizer passes over subsets of the GPU parity tests (the ASan analogue of the reference's CI
# job, .github/workflows/check.yml:42-58).  Run on a GPU box from the repository root:
#   gpurun --timeout 1700 -- 'bash tools/run_sanitizers.sh'
# Logs land in gpurun_out/; copy the summaries you want to keep into profiles/.
set -u
mkdir -p gpurun_out
T=tests/test_gpu_parity.py
run() { # tool, -k expression, [test file], [log suffix]
    local f="${3:-$T}" sfx="${4:-}"
    timeout 900 compute-sanitizer --tool "$1" --error-exitcode 9 python -m pytest $f -x -q -k "$2" \
        > "gpurun_out/sanitizer_$1$sfx.log" 2>&1
    echo "$1$sfx rc=$?"; tail -3 "gpurun_out/sanitizer_$1$sfx.log"
}
run memcheck  "kats or memchr or edge or mula or unaligned or async_entry or many_haystack or short_sweep or batched_single or pairs_mode or random_bench or count_mode or histogram or prepared"
run racecheck "mula or unaligned or async_entry or histogram or prepared"
run synccheck "mula or unaligned or async_entry or count_mode or histogram"
# initcheck: only tests whose haystacks are uploads (ss_b200_haystack_upload zeroes the padding) or exact
# byte reads (histogram).  The scans read whole 16-byte chunks, so on a BORROWED buffer whose length is
# not a multiple of 16 the last chunk includes bytes past the end (inside the allocation granule, masked
# out before any compare counts) -- initcheck reports those by design.
run initcheck "kats or batched_single or ipsum_absent or memchr or edge or short_sweep or pairs_mode or histogram"
# round 2: the resident service kernel, the many-haystack boundary rows / count-from-filter-words step, flag
# packing, the peer mailbox at world 1, the stream-ordered batch entries
C=tests/test_gpu_ctx.py
run memcheck  "service or pack_flags or peer_exchange or batch_stream or upload_then_search" $C _r2
run racecheck "service or pack_flags or peer_exchange" $C _r2
run synccheck "service or pack_flags or peer_exchange or many_mode" $C _r2
