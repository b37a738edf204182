#!/bin/bash
# round 2, call N (1 GPU): last build -- whole GPU suite + smoke + a short bench line
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
echo "pytest_gpu rc=$?" >> $O/steps.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke rc=$?" >> $O/steps.log
timeout 600 python bench.py --steps 50 --no-cpu > $O/bench_n1_short.json 2> $O/bench_n1_short.err
echo "bench rc=$?" >> $O/steps.log
cat $O/steps.log; tail -3 $O/pytest_gpu.log; tail -1 $O/smoke.log; head -c 300 $O/bench_n1_short.json
