#!/usr/bin/env python
"""Tuning sweep for the long scan (development tool, run under gpurun).

Times ss_b200_find_in_device_async over an i386-tiled (or random) haystack for a set of absent
needles and kernel variants / tunings; prints one table line per combination.

    python tools/sweep.py [--gib 4] [--reps 5] [--random] [--grid "1:0,0,0,0 2:0,0,16,4 ..."]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import sliceslice_rs_b200 as ss  # noqa: E402

TEXT_NEEDLES = ["\xff", "zq", "ipsum", "ipsumdol", "consecteturadipi", "ipsumdolorsitametconsectetur"]


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--gib", type=float, default=4.0)
    p.add_argument("--reps", type=int, default=5)
    p.add_argument("--random", action="store_true")
    p.add_argument("--needles", default="")
    p.add_argument("--grid", default="1:0,4,0,0 1:8,4,0,0 2:0,0,16,4 2:0,0,32,4 2:0,0,16,6 2:0,0,32,3")
    a = p.parse_args()
    n = int(a.gib * (1 << 30))
    hay = torch.empty(n, dtype=torch.uint8, device="cuda")
    if a.random:
        ss.fill_random(hay, 0, 0x5EEDB20000000001)
        import random

        needles = []
        for k in (1, 4, 16, 64):
            nd = bytearray(random.Random(1000 + k).randrange(255) for _ in range(k))
            nd[min(1, k - 1)] = 0xFF
            needles.append(bytes(nd))
    else:
        with open(os.path.join(ROOT, "data", "i386.txt"), "rb") as f:
            src = torch.frombuffer(bytearray(f.read()), dtype=torch.uint8).cuda()
        ss.fill_tiled(hay, 0, src)
        needles = [x.encode("latin-1") for x in (a.needles.split(",") if a.needles else TEXT_NEEDLES)]
    ws = torch.zeros(16, dtype=torch.uint8, device="cuda")
    res = torch.zeros(1, dtype=torch.int64, device="cuda")
    print(f"haystack {a.gib} GiB {'random' if a.random else 'i386-tiled'}; GB/s per (variant:ctas,unroll,tile_kib,stages[,extra_anchors])")
    hdr = "needle".ljust(34) + "".join(g.rjust(18) for g in a.grid.split())
    print(hdr)
    for nd in needles:
        row = (repr(nd)[:30] + f" k={len(nd)}").ljust(34)
        s = ss.DynamicB200Searcher.new(nd)
        for g in a.grid.split():
            v, t = g.split(":")
            ss.set_scan_variant(int(v))
            tt = [int(x) for x in t.split(",")]
            ss.set_scan_tuning(*tt[:4])
            ss.set_extra_anchors(tt[4] if len(tt) > 4 else -1)
            for _ in range(6):
                s.find_in_async(hay, res, ws)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                s.find_in_async(hay, res, ws)
            e1.record()
            torch.cuda.synchronize()
            assert int(res.item()) == ss.DEVICE_NONE
            row += f"{n * a.reps / (e0.elapsed_time(e1) * 1e-3) / 1e9:18.1f}"
        print(row, flush=True)


if __name__ == "__main__":
    main()
