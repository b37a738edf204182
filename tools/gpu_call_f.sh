#!/bin/bash
# round 2, call F (1 GPU): resident service kernel (tests + latency), dynamic host engine, full bench
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
cp sliceslice_rs_b200/libsliceslice_b200.so tools/ab/cur/
timeout 600 python -m pytest tests/test_gpu_ctx.py -x -q -m gpu -k "service or host or ring or lane" --durations=5 > $O/pytest_service.log 2>&1
echo "pytest_service rc=$?" >> $O/steps.log
for sv in 1 0; do
  LD_LIBRARY_PATH=tools/ab/cur:/usr/local/cuda/lib64 timeout 300 tools/ab/bench_latency data/i386.txt data/words.txt $sv > $O/latency_service$sv.txt 2>&1
done
echo "latency rc=$?" >> $O/steps.log
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$?" >> $O/steps.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/steps.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke rc=$?" >> $O/steps.log
cat $O/steps.log; tail -4 $O/pytest_service.log; grep -h "long sweep (one find_in per\|absent needle\|host slice" $O/latency_service1.txt $O/latency_service0.txt; tail -4 $O/pytest_gpu.log; tail -2 $O/smoke.log
