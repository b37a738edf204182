#!/bin/bash
# round 2, call Q (1 GPU): host-path tests after the staging change + pageable rate with the new defaults
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ctx.py -x -q -m gpu -k "host or ring or lane or smoke" > $O/pytest_host.log 2>&1
echo "pytest_host rc=$?" >> $O/steps.log
timeout 200 python tools/host_path_timing.py 2 > $O/host_path_timing.txt 2>&1
echo "timing rc=$?" >> $O/steps.log
cat $O/steps.log; tail -3 $O/pytest_host.log; cat $O/host_path_timing.txt
