#!/usr/bin/env python
"""Host wall-clock of the batched modes (development tool): config 2 batched and config 3."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sliceslice_rs_b200 as ss  # noqa: E402

i386 = open(os.path.join(ROOT, "data", "i386.txt"), "rb").read()
words = [w for w in open(os.path.join(ROOT, "data", "words.txt"), "rb").read().split(b"\n") if w]
sw = [words[i] for i in sorted(range(len(words)), key=lambda i: (len(words[i]), i))]
hs = ss.DeviceHaystack.upload(i386)
b = ss.Batch(words, [])
tri = ss.Batch(sw, sw)
for it in range(5):
    t0 = time.perf_counter()
    o = b.find_all_in(hs)
    t1 = time.perf_counter()
    bm, m = tri.search_triangular()
    t2 = time.perf_counter()
    print(f"find_all_in {1e3 * (t1 - t0):.3f} ms (sum {int(o.sum())})   triangular {1e3 * (t2 - t1):.3f} ms ({m} matches)")
