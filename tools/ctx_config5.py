#!/usr/bin/env python
"""BASELINE config 5 through the one-process context: 8 GiB of the counter-based generator per GPU, sharded by
start position with a k-1 halo, searched with ss_b200_search_sharded through each exchange (mapped host
words, peer mailboxes, NCCL loaded by dlopen).  Absent needle, then the SURVEY 8d plants: the last k bytes of
the last shard, straddling the middle shard boundary, shard 0 and a later shard at once.  Prints one JSON line.

    python tools/ctx_config5.py [--gib 8] [--ndev 0]
"""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import sliceslice_rs_b200 as ss  # noqa: E402

gib = float(sys.argv[sys.argv.index("--gib") + 1]) if "--gib" in sys.argv else 8.0
ndev = int(sys.argv[sys.argv.index("--ndev") + 1]) if "--ndev" in sys.argv else torch.cuda.device_count()
S = int(gib * (1 << 30))
k = 16
seed = 0x5EEDB20000000001
nd = bytearray(random.Random(1000 + k).randrange(255) for _ in range(k))
nd[1] = 0xFF  # absent from the generator's alphabet
nd = bytes(nd)
total = S * ndev
ctx = ss.Context(ndev)
tensors, owned = [], []
for d in range(ndev):
    span = min(S + k - 1, total - d * S)
    with torch.cuda.device(d):
        t = torch.empty(span, dtype=torch.uint8, device=f"cuda:{d}")
        ss.fill_random(t, d * S, seed)
        torch.cuda.synchronize()
    tensors.append(t)
    owned.append(S if d < ndev - 1 else span - k + 1)
sh = ctx.sharded_from_tensors(tensors, owned)
s = ss.DynamicB200Searcher.new(nd)
ndt = torch.tensor(list(nd), dtype=torch.uint8)


def plant(pos):
    saved = []
    for d in range(ndev):
        lo, hi = max(pos, d * S), min(pos + k, d * S + tensors[d].numel())
        if lo < hi:
            view = tensors[d][lo - d * S:hi - d * S]
            saved.append((view, view.clone()))
            view.copy_(ndt[lo - pos:hi - pos].to(view.device))
    torch.cuda.synchronize()
    return saved


def unplant(saved):
    for view, old in saved:
        view.copy_(old)
    torch.cuda.synchronize()


out = {"ndev": ndev, "bytes_total": total, "needle_len": k, "exchanges": {}}
names = {ss.EXCHANGE_HOST: "host_words", ss.EXCHANGE_PEER: "peer_mailboxes", ss.EXCHANGE_NCCL: "nccl_allreduce_min"}
for ex in ([ss.EXCHANGE_HOST] if ndev == 1 else [ss.EXCHANGE_HOST, ss.EXCHANGE_PEER, ss.EXCHANGE_NCCL]):
    ctx.set_exchange(ex)
    row = {}
    assert ctx.find_sharded(s, sh) is None
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        r = ctx.find_sharded(s, sh)
        ts.append(time.perf_counter() - t0)
    assert r is None
    row["absent_ms_per_search"] = round(min(ts) * 1e3, 4)
    row["absent_gbs"] = round(total / min(ts) / 1e9, 1)
    cases = {"last_k_bytes_of_last_shard": ([total - k], total - k),
             "straddling_middle_boundary": ([(ndev // 2) * S - k // 2], (ndev // 2) * S - k // 2),
             "shard0_and_a_later_shard": ([4242, (ndev - 1) * S + S // 2], 4242)}
    for name, (spots, expect) in cases.items():
        saved = [x for p in spots for x in plant(p)]
        t0 = time.perf_counter()
        got = ctx.find_sharded(s, sh)
        dt = time.perf_counter() - t0
        unplant(list(reversed(saved)))
        assert got == expect, (names[ex], name, got, expect)
        row[name] = {"expected": expect, "got": got, "ms": round(dt * 1e3, 4)}
    out["exchanges"][names[ex]] = row
print(json.dumps(out))
