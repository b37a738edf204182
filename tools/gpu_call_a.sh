#!/bin/bash
# round 2, call A (2 GPUs): new boundary code (context, host engine), kernel regressions, host-path probe
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_ctx.py -x -q -m gpu --durations=10 > $O/pytest_ctx.log 2>&1
echo "pytest_ctx rc=$?" >> $O/steps.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cpp_host_mirror.py -x -q -m gpu --durations=10 > $O/pytest_parity.log 2>&1
echo "pytest_parity rc=$?" >> $O/steps.log
timeout 400 python tools/host_path_probe.py --gib 2 --out $O/host_path.json > $O/host_path.log 2>&1
echo "probe rc=$?" >> $O/steps.log
timeout 500 python bench.py --steps 30 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/steps.log
timeout 300 python bench.py --mode many --steps 20 > $O/bench_many.json 2> $O/bench_many.err
echo "bench_many rc=$?" >> $O/steps.log
cat $O/steps.log; tail -5 $O/pytest_ctx.log; tail -5 $O/pytest_parity.log; tail -c 1500 $O/bench_many.json
