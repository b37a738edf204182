#!/usr/bin/env python
"""Run a few scans of one mode so that ncu can capture one of them:

    ncu --set full --clock-control none -k regex:scan_tma -s 2 -c 1 -o out python tools/profile_modes.py many the 8

mode: find | count | many (prepared set of ~8 KiB haystacks cut from the i386 tiling, as bench.py --mode many)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sliceslice_rs_b200 as ss  # noqa: E402

mode, needle, gib = sys.argv[1], sys.argv[2].encode(), float(sys.argv[3]) if len(sys.argv) > 3 else 8.0
n = int(gib * (1 << 30))
i386 = open(os.path.join(ROOT, "data", "i386.txt"), "rb").read()
src = torch.frombuffer(bytearray(i386), dtype=torch.uint8).cuda()
hay = torch.empty(n, dtype=torch.uint8, device="cuda")
ss.fill_tiled(hay, 0, src)
s = ss.DynamicB200Searcher.new(needle)
ws = torch.zeros(32, dtype=torch.uint8, device="cuda")
res = torch.zeros(1, dtype=torch.int64, device="cuda")
if mode == "many":
    rng = np.random.default_rng(20260101)
    lens = rng.integers(0, 16384, n // 8192 + n // 131072 + 16, dtype=np.int64)
    off = np.zeros(lens.size + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    cut = int(np.searchsorted(off, n, side="right")) - 1
    off = np.append(off[:cut + 1], n) if off[cut] < n else off[:cut + 1]
    hs = ss.HaystackSet.from_device(hay, torch.from_numpy(off).cuda())
    flags = torch.zeros(off.size - 1, dtype=torch.uint8, device="cuda")
for _ in range(4):
    if mode == "find":
        s.find_in_async(hay, res, ws)
    elif mode == "count":
        s.count_in_async(hay, res, ws)
    else:
        s.search_many_async(hs, flags)
torch.cuda.synchronize()
print(mode, needle, int(res.item()) if mode != "many" else int(flags.sum().item()))
