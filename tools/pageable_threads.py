#!/usr/bin/env python
"""Pageable host slice through ss_b200_find_in_host for several sizes of the staging memcpy pool."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import sliceslice_rs_b200 as ss  # noqa: E402

n = 2 << 30
i386 = np.frombuffer(open(os.path.join(ROOT, "data", "i386.txt"), "rb").read(), np.uint8)
buf = np.resize(i386, n)
s = ss.DynamicB200Searcher.new(b"ipsum")
print("cores", len(os.sched_getaffinity(0)))
for chunk in (0, 16, 64):
    for threads in (3, 7, 11, 15):
        ss.set_host_path(0, chunk, threads)
        assert s.find_in(buf) is None
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            s.find_in(buf)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        print(f"chunk_mib {chunk or 'auto(32)'} copy_threads {threads}: {n / best / 1e9:.1f} GB/s")
