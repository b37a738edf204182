#!/bin/bash
# round 2, call R (1 GPU): short host slices through the resident kernel -- tests + per-call latency
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out tools/ab/cur
O=gpurun_out
cp sliceslice_rs_b200/libsliceslice_b200.so tools/ab/cur/
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$?" >> $O/steps.log
for sv in 1 0; do
  LD_LIBRARY_PATH=tools/ab/cur:/usr/local/cuda/lib64 timeout 300 tools/ab/bench_latency data/i386.txt data/words.txt $sv > $O/latency_service$sv.txt 2>&1
done
echo "latency rc=$?" >> $O/steps.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke rc=$?" >> $O/steps.log
cat $O/steps.log; tail -3 $O/pytest_gpu.log; grep -h "host slice\|one find_in per word" $O/latency_service1.txt $O/latency_service0.txt; tail -1 $O/smoke.log
