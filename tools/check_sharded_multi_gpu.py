#!/usr/bin/env python
"""Functional check of the multi-GPU modes under torchrun (one rank per GPU, NCCL):
sharded single haystack (MIN-allreduce of first offsets) and many-haystack flags (MAX-allreduce),
both compared with plain Python bytes.find on rank 0's view of the same generator data.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_sharded_multi_gpu.py
"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sliceslice_rs_b200 as ss  # noqa: E402
from sliceslice_rs_b200.sharded import ShardedHaystackSet, ShardedSearch, shard_bounds  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    seed = 0x5EEDB20000000001
    # --gib G: G GiB of start positions per rank (BASELINE config 5 is 8 GiB x 8 ranks = 64 GiB)
    gib = float(sys.argv[sys.argv.index("--gib") + 1]) if "--gib" in sys.argv else 0.0
    total = int(gib * (1 << 30)) * world if gib else (256 << 20) + 12345
    needles = []
    for k in (1, 4, 16, 64):
        nd = bytearray(random.Random(1000 + k).randrange(255) for _ in range(k))
        nd[min(1, k - 1)] = 0xFF
        needles.append(bytes(nd))
    ok = True
    for nd, exchange in [(n_, e_) for n_ in needles for e_ in ("nccl", "peer")]:
        k = len(nd)
        start, owned, span = shard_bounds(total, k, world, rank)
        shard = torch.empty(span, dtype=torch.uint8, device="cuda")
        ss.fill_random(shard, start, seed)
        sh = ShardedSearch(shard, start, owned, exchange=exchange)
        s = ss.DynamicB200Searcher.new(nd)
        assert sh.find(s) is None
        per = shard_bounds(total, k, world, 0)[1]
        # plants: straddling the rank 0/1 boundary, at the very end, and early in the last rank
        # SURVEY 8d config 5 plants: the last k bytes of the last rank; straddling a rank boundary (0/1 and
        # the middle one); early in the last rank; inside rank 0 (the MIN must pick it over later ranks)
        plants = sorted({max(0, per - k // 2 - 1) if world > 1 else 777, total - k,
                         min(total - k, (world - 1) * per + 99),
                         max(0, min(total - k, (world // 2) * per - k // 2)), min(total - k, 4242)}, reverse=True)
        ndt = torch.tensor(list(nd), dtype=torch.uint8, device="cuda")
        for plant in plants:  # descending: each new plant is the global leftmost
            lo, hi = max(plant, start), min(plant + k, start + span)
            if lo < hi:
                shard[lo - start:hi - start] = ndt[lo - plant:hi - plant]
            got = sh.find(s)
            got2 = sh.find_many([s, s])
            if got != plant or got2 != [plant, plant]:
                ok = False
                print(f"rank {rank}: {exchange} k={k} plant={plant} got={got} got2={got2}", flush=True)
        # trivial outcomes go through the same exchange: empty needle => found at global offset 0
        if sh.find(ss.DynamicB200Searcher.new(b"")) != 0:
            ok = False
            print(f"rank {rank}: {exchange} empty needle", flush=True)
        if sh.peer is not None:
            sh.peer.close()
    # many-haystack mode
    rng = random.Random(5)
    hays = [bytes(rng.randrange(97, 101) for _ in range(rng.randrange(0, 3000))) for _ in range(4000)]
    hs = ShardedHaystackSet(hays, rank=rank, world=world)
    for nd in (b"abc", b"dddd", b"abcdabcd", b"a"):
        got = hs.search(ss.DynamicB200Searcher.new(nd))
        exp = [h.find(nd) >= 0 for h in hays]
        if got.tolist() != exp:
            ok = False
            print(f"rank {rank}: many-haystack mismatch for {nd!r}", flush=True)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multi-gpu check:", "ok" if int(t.item()) == 1 else "FAILED", f"(world {world})", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
