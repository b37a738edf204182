#!/bin/bash
# round 2, call J (1 GPU): A/B of service-kernel variants on one box + final validation of HEAD
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
for rep in 1 2; do
for v in cur svc_scouts256 svc_one512 svc_g2x svc_ghalf; do
  echo "== $v" >> $O/svc_ab.txt
  LD_LIBRARY_PATH=tools/ab/$v:/usr/local/cuda/lib64 timeout 120 tools/ab/bench_latency data/i386.txt data/words.txt 1 2>&1 | grep "one find_in per word\|absent needle" | tail -3 >> $O/svc_ab.txt
done
done
echo "svc_ab rc=$?" >> $O/steps.log
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$?" >> $O/steps.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/steps.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke rc=$?" >> $O/steps.log
cat $O/steps.log; cat $O/svc_ab.txt; tail -3 $O/pytest_gpu.log; tail -1 $O/smoke.log
