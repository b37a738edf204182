#!/bin/bash
# round 2, call M (8 GPUs): final build -- context tests on all eight devices, both bench arms at N = 8
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ctx.py tests/test_cpp_host_mirror.py -x -q -m gpu -k "ctx or host_multi or torchrun or peer_exchange" > $O/pytest_ctx8.log 2>&1
echo "pytest_ctx8 rc=$?" >> $O/steps.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 60 > $O/bench_n8.json 2> $O/bench_n8.err
echo "bench_n8 rc=$?" >> $O/steps.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 8 --impl reference --steps 5 --warmup 1 > $O/bench_ref_n8.json 2> $O/bench_ref_n8.err
echo "bench_ref8 rc=$?" >> $O/steps.log
cat $O/steps.log; tail -3 $O/pytest_ctx8.log; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_n8.json").read().split("\n") if l.startswith("{")][0])
print(d["value"], d["ms_per_step"], d["single_search_latency_ms"], d["e2e"]["value"], d["e2e"]["roofline"]["frac"], d["extras"]["many_haystack_mode"]["ms_per_step"])
print(open("gpurun_out/bench_ref_n8.json").read()[:200])
PY
