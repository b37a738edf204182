#!/bin/bash
# round 2, call K (1 GPU): final build (service grid 2 CTAs/SM) -- suite, latency, bench, reference arm, smoke
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$?" >> $O/steps.log
for sv in 1 0; do
  LD_LIBRARY_PATH=tools/ab/cur:/usr/local/cuda/lib64 timeout 300 tools/ab/bench_latency data/i386.txt data/words.txt $sv > $O/latency_service$sv.txt 2>&1
done
echo "latency rc=$?" >> $O/steps.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/steps.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err
echo "bench_ref rc=$?" >> $O/steps.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke rc=$?" >> $O/steps.log
cat $O/steps.log; grep -h "long sweep (one find_in per\|absent needle" $O/latency_service1.txt $O/latency_service0.txt; tail -3 $O/pytest_gpu.log; tail -1 $O/smoke.log
