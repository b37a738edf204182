#!/bin/bash
# round 2, call O (1 GPU): compute-sanitizer over the many-haystack rows, the count step and the context
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 280 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_modes.py > $O/sanitize_modes_$tool.log 2>&1
  echo "$tool rc=$?" >> $O/steps.log
done
cat $O/steps.log; for tool in memcheck racecheck synccheck; do tail -3 $O/sanitize_modes_$tool.log; done
