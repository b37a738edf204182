#!/bin/bash
# round 2, call D (1 GPU): validate the many-mode boundary rows and the count-from-filter-words step
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
cp sliceslice_rs_b200/libsliceslice_b200.so tools/ab/cur/
timeout 900 python -m pytest tests/test_gpu_ctx.py -x -q -m gpu -k "many_mode or pack or sharded_haystack_set" > $O/pytest_new.log 2>&1
echo "pytest_new rc=$?" >> $O/steps.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/pytest_parity.log 2>&1
echo "pytest_parity rc=$?" >> $O/steps.log
run() { echo "== $*" >> $O/ab2.txt; LD_LIBRARY_PATH=tools/ab/cur:/usr/local/cuda/lib64 timeout 120 tools/ab/bench_scan data/i386.txt "$@" 2>&1 | tail -1 >> $O/ab2.txt; }
run 8 20 the count
run 8 20 e count
run 8 20 th count
run 8 20 segment count
run 8 20 ipsum count
run 8 20 zq count
run 8 60 zq
run 8 60 ipsum
run 1 200 ipsum
run 0.25 400 ipsum
run 0.0625 400 consecteturadipi
timeout 300 python bench.py --mode many --steps 20 > $O/bench_many.json 2> $O/bench_many.err
echo "bench_many rc=$?" >> $O/steps.log
cat $O/steps.log; cat $O/ab2.txt; tail -3 $O/pytest_new.log; tail -3 $O/pytest_parity.log; python -c "
import json;d=json.load(open('$O/bench_many.json'));print(d['value'],d['ms_per_step'],d['present_needle_gbs_per_gpu'])"
