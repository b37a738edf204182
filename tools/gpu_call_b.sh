#!/bin/bash
# round 2, call B (1 GPU): A/B of the scan kernels against the round-1 build on one box, new tests, host-path probe
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
cp sliceslice_rs_b200/libsliceslice_b200.so tools/ab/cur/
ab() { # label args...
  for lib in r1 cur r1 cur; do
    echo "== $lib: $*" >> $O/ab.txt
    LD_LIBRARY_PATH=tools/ab/$lib:/usr/local/cuda/lib64 timeout 120 tools/ab/bench_scan data/i386.txt "$@" 2>&1 | tail -2 >> $O/ab.txt
  done
}
ab 8 60 ipsum
ab 8 60 consecteturadipi
ab 1 200 ipsum
ab 0.25 400 ipsum
ab 8 20 the count
ab 8 20 segment count
ab 8 60 zq
echo "ab rc=$?" >> $O/steps.log
timeout 900 python -m pytest tests/test_gpu_ctx.py -x -q -m gpu --durations=10 > $O/pytest_ctx.log 2>&1
echo "pytest_ctx rc=$?" >> $O/steps.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/pytest_parity.log 2>&1
echo "pytest_parity rc=$?" >> $O/steps.log
timeout 400 python tools/host_path_probe.py --gib 4 --out $O/host_path.json > $O/host_path.log 2>&1
echo "probe rc=$?" >> $O/steps.log
timeout 300 python bench.py --mode many --steps 20 > $O/bench_many.json 2> $O/bench_many.err
echo "bench_many rc=$?" >> $O/steps.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/steps.log
cat $O/steps.log; cat $O/ab.txt; tail -3 $O/pytest_ctx.log; tail -3 $O/pytest_parity.log
