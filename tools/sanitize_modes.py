#!/usr/bin/env python
"""A small run of the many-haystack boundary rows, the count-from-filter-words step and the context's
sharded search, sized for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_modes.py
"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sliceslice_rs_b200 as ss  # noqa: E402

i386 = open(os.path.join(ROOT, "data", "i386.txt"), "rb").read()
ss.set_scan_variant(2)  # the staged kernel on a blob this small
rng = random.Random(3)
lens = []
while sum(lens) < (3 << 20):
    lens += ([rng.choice([0, 1, 2, 5]) for _ in range(60)] if rng.random() < 0.2 else [rng.randrange(0, 16384)])
text = (i386 * 5)[:sum(lens)]
off = np.zeros(len(lens) + 1, np.int64)
np.cumsum(lens, out=off[1:])
for shift in (0, 3):
    blob = torch.frombuffer(bytearray(b"\0" * shift + text), dtype=torch.uint8).cuda()[shift:]
    hs = ss.HaystackSet.from_device(blob, torch.from_numpy(off).cuda())
    hays = [text[int(a):int(b)] for a, b in zip(off[:-1], off[1:])]
    for nd in (b"the", b"e", b"segment", b"the 80386 provides a"):
        got = ss.DynamicB200Searcher.new(nd).search_many_async(hs).cpu().numpy().astype(bool)
        assert got.tolist() == [nd in h for h in hays], nd
    ws = torch.zeros(32, dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    for nd in (b"e", b"th", b"the", b"ing"):
        ss.DynamicB200Searcher.new(nd).count_in_async(blob, cnt, ws)
        assert int(cnt.item()) == sum(1 for i in range(len(text) - len(nd) + 1) if text.startswith(nd, i)), nd
ss.set_scan_variant(0)
ctx = ss.Context()
h = bytearray(i386[:600000])
h[599990:] = b"\x01\x02needle\x03\x04"
sh = ctx.upload_sharded(bytes(h), halo=64)
assert ctx.find_sharded(ss.DynamicB200Searcher.new(b"\x01\x02needle\x03\x04"), sh) == 599990
assert ctx.find_in_host(ss.DynamicB200Searcher.new(b"\x01\x02needle\x03\x04"), bytes(h) * 20) == 599990
print("sanitize_modes ok")
