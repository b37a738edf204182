import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sliceslice_rs_b200 as ss
rng = np.random.default_rng(7)
for n, align, nd in ((4087216, 8, b"bbab"), (15047586, 24, b"bab"), (4087216, 0, b"bbab"), (1 << 20, 8, b"bbab"), (40 << 20, 8, b"abba")):
    h = rng.integers(0, 2, size=n, dtype=np.uint8) + 97
    exp = h.tobytes().find(nd)
    pool = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
    pool[align:align + n] = torch.from_numpy(h).cuda()
    for variant in (1, 2):
        for ea in (-1, 0):
            for pos in (0, len(nd) - 1):
                ss.set_scan_variant(variant); ss.set_extra_anchors(ea)
                s = ss.DynamicB200Searcher.with_position(nd, pos)
                got = [s.find_in(pool[align:align + n]) for _ in range(30)]
                bad = [g for g in got if g != exp]
                print(f"n={n} align={align} nd={nd} variant={variant} extras={ea} pos={pos} exp={exp} bad={len(bad)}/30 {sorted(set(bad))[:8]}")
