#!/bin/bash
# round 2, call L (2 GPUs): stop-word polling made cheap -- whole suite on two devices + the N = 2 bench line
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu2.log 2>&1
echo "pytest_gpu2 rc=$?" >> $O/steps.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 40 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench_n2 rc=$?" >> $O/steps.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --impl reference --steps 5 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
echo "bench_ref rc=$?" >> $O/steps.log
cat $O/steps.log; tail -3 $O/pytest_gpu2.log; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_n2.json").read().split("\n") if l.startswith("{")][0])
print(d["value"], d["ms_per_step"], d["single_search_latency_ms"], d["found_needle"], d["e2e"]["value"], d["e2e"]["roofline"]["frac"])
PY
