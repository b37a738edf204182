#!/usr/bin/env python
"""Throughput of the host-slice entry (ss_b200_find_in_host) for pageable vs pinned host memory."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sliceslice_rs_b200 as ss  # noqa: E402

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
n = int(gib * (1 << 30))
i386 = np.frombuffer(open(os.path.join(ROOT, "data", "i386.txt"), "rb").read(), np.uint8)
pageable = np.resize(i386, n)
pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
pinned.numpy()[:] = pageable
s = ss.DynamicB200Searcher.new(b"ipsum")
for name, buf in (("pageable numpy", pageable), ("pinned torch", pinned)):
    for it in range(4):
        t0 = time.perf_counter()
        r = s.find_in(buf)
        dt = time.perf_counter() - t0
        assert r is None
        if it:
            print(f"{name}: {n / dt / 1e9:.1f} GB/s ({dt * 1e3:.1f} ms)")
