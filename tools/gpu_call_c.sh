#!/bin/bash
# round 2, call C (2 GPUs): tests of the context on 2 devices, tile-size choice for mid-size scans, many/count modes,
# ncu capture of the count-mode scan
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
cp sliceslice_rs_b200/libsliceslice_b200.so tools/ab/cur/
timeout 900 python -m pytest tests/test_gpu_ctx.py tests/test_cpp_host_mirror.py -x -q -m gpu --durations=10 > $O/pytest_ctx.log 2>&1
echo "pytest_ctx rc=$?" >> $O/steps.log
run() { echo "== $*" >> $O/tiles.txt; LD_LIBRARY_PATH=tools/ab/cur:/usr/local/cuda/lib64 timeout 120 tools/ab/bench_scan data/i386.txt "$@" 2>&1 | tail -1 >> $O/tiles.txt; }
for g in 0.0625 0.25 1 2; do
  for nd in ipsum consecteturadipi; do
    run $g 200 $nd find 16 6
    run $g 200 $nd find 16 4
    run $g 200 $nd find 32 3
  done
done
echo "== r1 1 GiB" >> $O/tiles.txt; LD_LIBRARY_PATH=tools/ab/r1:/usr/local/cuda/lib64 tools/ab/bench_scan data/i386.txt 1 200 ipsum | tail -1 >> $O/tiles.txt
echo "== r1 0.25 GiB" >> $O/tiles.txt; LD_LIBRARY_PATH=tools/ab/r1:/usr/local/cuda/lib64 tools/ab/bench_scan data/i386.txt 0.25 200 ipsum | tail -1 >> $O/tiles.txt
echo "tiles rc=$?" >> $O/steps.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/pytest_parity.log 2>&1
echo "pytest_parity rc=$?" >> $O/steps.log
timeout 300 python bench.py --mode many --steps 20 > $O/bench_many.json 2> $O/bench_many.err
echo "bench_many rc=$?" >> $O/steps.log
LD_LIBRARY_PATH=tools/ab/cur:/usr/local/cuda/lib64 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 4 -c 1 -o $O/count_the tools/ab/bench_scan data/i386.txt 8 3 the count > $O/ncu_count.log 2>&1
echo "ncu rc=$?" >> $O/steps.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 40 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench_n2 rc=$?" >> $O/steps.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --impl reference --steps 5 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
echo "bench_ref rc=$?" >> $O/steps.log
cat $O/steps.log; cat $O/tiles.txt; tail -3 $O/pytest_ctx.log; tail -3 $O/pytest_parity.log; tail -c 600 $O/bench_n2.err
