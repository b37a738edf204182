#!/bin/bash
# One gpurun call that refreshes the round's GPU evidence: new tests first, then the bench lines
# (ours + reference arm), the whole GPU parity suite, and the ncu launch list.  Every step writes
# under gpurun_out/ as it goes so that a cut-off call still leaves what it finished.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_ctx.py -x -q -m gpu > $O/pytest_new.log 2>&1
echo "pytest_new rc=$?" >> $O/steps.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/steps.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
echo "bench_ref rc=$?" >> $O/steps.log
timeout 900 python -m pytest tests -x -q -m gpu --durations=15 > $O/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$?" >> $O/steps.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-extras --no-cpu --no-e2e > $O/bench_under_ncu.log 2>&1
echo "ncu rc=$?" >> $O/steps.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke rc=$?" >> $O/steps.log
cat $O/steps.log; tail -3 $O/pytest_new.log; tail -5 $O/pytest_gpu.log; head -c 600 $O/bench_n1.json
