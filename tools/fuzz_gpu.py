#!/usr/bin/env python
"""Randomized differential fuzz of the GPU path against the CPU oracle (development tool; the pytest
suite runs a shorter version).  Every case: random alphabet size, haystack length (log-uniform up to
--max-len), device byte alignment 0..31, needle length up to 300 (often cut from the haystack so that
matches exist), random `position`, random kernel variant.  Compares find_in with the oracle's
DynamicAvx2Searcher restatement and with bytes.find.

    python tools/fuzz_gpu.py [--seconds 120] [--max-len 4194304] [--seed 1]
"""
import argparse
import math
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle  # noqa: E402
import sliceslice_rs_b200 as ss  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--seconds", type=float, default=120)
    p.add_argument("--max-len", type=int, default=4 << 20)
    p.add_argument("--seed", type=int, default=1)
    a = p.parse_args()
    rng = random.Random(a.seed)
    nrng = np.random.default_rng(a.seed)
    pool = torch.empty(a.max_len + 4096, dtype=torch.uint8, device="cuda")
    t_end = time.time() + a.seconds
    cases = found = 0
    by_variant = {1: 0, 2: 0}
    while time.time() < t_end:
        alphabet = rng.choice([1, 2, 2, 3, 4, 16, 64, 256])
        n = int(math.exp(rng.uniform(0, math.log(a.max_len)))) if rng.random() < 0.9 else rng.randrange(0, 64)
        h = nrng.integers(0, alphabet, size=n, dtype=np.uint8) + (0 if alphabet == 256 else 97)
        k = rng.choice([rng.randrange(0, 8), rng.randrange(0, 40), rng.randrange(0, 300)])
        if k and n >= k and rng.random() < 0.6:
            st = rng.randrange(0, n - k + 1)
            nd = bytes(h[st:st + k])
            if rng.random() < 0.3 and k > 1:  # near miss: corrupt one byte
                j = rng.randrange(k)
                nd = nd[:j] + bytes([(nd[j] + 1) % 256]) + nd[j + 1:]
        else:
            nd = bytes(nrng.integers(0, alphabet, size=k, dtype=np.uint8) + (0 if alphabet == 256 else 97))
        pos = 0 if k <= 1 else rng.randrange(k)
        align = rng.randrange(32)
        variant = rng.choice([1, 2])
        ss.set_scan_variant(variant)
        ss.set_extra_anchors(rng.choice([-1, -1, 0]))
        if n:
            pool[align:align + n] = torch.from_numpy(h).cuda()
        if k:
            pool[align + n:align + n + 64] = nd[0]  # poison the bytes behind the slice
        s = ss.DynamicB200Searcher.with_position(nd, pos)
        got = s.find_in(pool[align:align + n])
        s.close()
        exp = oracle.find(h, nd, pos)
        hb = h.tobytes()
        e2 = hb.find(nd)
        if got != exp or exp != (None if e2 < 0 else e2):
            print(f"MISMATCH n={n} k={k} pos={pos} align={align} variant={variant} alphabet={alphabet} got={got} "
                  f"oracle={exp} find={e2} needle={nd[:40]!r}")
            sys.exit(1)
        cases += 1
        found += got is not None
        by_variant[variant] += 1
    print(f"fuzz ok: {cases} cases ({found} with a match), variants {by_variant}, max_len {a.max_len}, seed {a.seed}")


if __name__ == "__main__":
    main()
