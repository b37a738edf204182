#!/bin/bash
# round 2, call G (2 GPUs): whole GPU suite on two devices (final build), ncu evidence, sanitizers
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
cp sliceslice_rs_b200/libsliceslice_b200.so tools/ab/cur/
timeout 1500 python -m pytest tests -x -q -m gpu --durations=6 > $O/pytest_gpu2.log 2>&1
echo "pytest_gpu2 rc=$?" >> $O/steps.log
export CUDA_VISIBLE_DEVICES=0
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-extras --no-cpu --no-e2e --sustained-steps 0 > $O/bench_under_ncu.log 2>&1
echo "ncu_list rc=$?" >> $O/steps.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 2 -c 1 -o $O/r02_find_ipsum python tools/profile_modes.py find ipsum 8 > $O/ncu_find.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 2 -c 1 -o $O/r02_count_the python tools/profile_modes.py count the 8 > $O/ncu_count.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 2 -c 1 -o $O/r02_many_the python tools/profile_modes.py many the 8 > $O/ncu_many.log 2>&1
echo "ncu_full rc=$?" >> $O/steps.log
timeout 1700 bash tools/run_sanitizers.sh > $O/sanitizers.log 2>&1
echo "sanitizers rc=$?" >> $O/steps.log
cat $O/steps.log; tail -8 $O/pytest_gpu2.log; grep "rc=" $O/sanitizers.log
