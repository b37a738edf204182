#!/bin/bash
# round 2, call E (8 GPUs): the context on every GPU of the box, PCIe topology probe, both bench arms at N = 8
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/topo8.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_ctx.py tests/test_cpp_host_mirror.py -x -q -m gpu --durations=8 > $O/pytest_ctx8.log 2>&1
echo "pytest_ctx8 rc=$?" >> $O/steps.log
timeout 600 python tools/host_path_probe.py --gib 1 --quick --out $O/host_path8.json > $O/host_path8.log 2>&1
echo "probe8 rc=$?" >> $O/steps.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 60 > $O/bench_n8.json 2> $O/bench_n8.err
echo "bench_n8 rc=$?" >> $O/steps.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --impl reference --steps 5 --warmup 1 > $O/bench_ref_n8.json 2> $O/bench_ref_n8.err
echo "bench_ref8 rc=$?" >> $O/steps.log
cat $O/steps.log; tail -12 $O/pytest_ctx8.log; head -c 1500 $O/bench_n8.json; echo; head -c 600 $O/bench_ref_n8.json
