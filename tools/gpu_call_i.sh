#!/bin/bash
# round 2, call I (8 GPUs): both bench arms at N = 8 on the final build (least-loaded deal of the host slice)
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 60 > $O/bench_n8.json 2> $O/bench_n8.err
echo "bench_n8 rc=$?" >> $O/steps.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --impl reference --steps 5 --warmup 1 > $O/bench_ref_n8.json 2> $O/bench_ref_n8.err
echo "bench_ref8 rc=$?" >> $O/steps.log
cat $O/steps.log; head -c 400 $O/bench_n8.json; echo; python - <<'PY'
import json
for f in ("gpurun_out/bench_n8.json",):
    try:
        d=json.loads([l for l in open(f).read().split("\n") if l.startswith("{")][0])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["roofline"])
    except Exception as e: print(f, e)
print(open("gpurun_out/bench_ref_n8.json").read()[:300])
PY
