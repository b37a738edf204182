#!/usr/bin/env python
"""Host-slice path probe: where the PCIe ceiling is and which data path gets closest to it.

    python tools/host_path_probe.py [--gib 4] [--ndev 0] [--out gpurun_out/host_path.json]

1. topology: nvidia-smi topo -m, NUMA nodes, cores, memory
2. pinned H2D cudaMemcpyAsync bandwidth per device (alone), and of device subsets copying concurrently
   (one thread per device) -- the ceiling the e2e number is reported against, and what limits N > 1
3. ss_b200_find_in_host_multi over 1, 2, 4, .. ndev devices: DMA ring vs in place (LDG / TMA), per size
4. pageable input (memcpy pool + pinned ring)
Everything goes through the C ABI; prints one JSON document.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sliceslice_rs_b200 as ss  # noqa: E402


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout
    except Exception as e:  # noqa: BLE001
        return f"{type(e).__name__}: {e}"


def concurrent_h2d(devs, nbytes, reps=3):
    """GB/s of each device when all of `devs` copy `nbytes` of their own pinned buffer at once."""
    out = {}
    barrier = threading.Barrier(len(devs))

    def work(d):
        torch.cuda.set_device(d)
        h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h.fill_(7)
        t = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}")
        st = torch.cuda.Stream(device=d)
        with torch.cuda.stream(st):
            t.copy_(h, non_blocking=True)
        st.synchronize()
        barrier.wait()
        t0 = time.perf_counter()
        with torch.cuda.stream(st):
            for _ in range(reps):
                t.copy_(h, non_blocking=True)
        st.synchronize()
        out[d] = nbytes * reps / (time.perf_counter() - t0) / 1e9

    th = [threading.Thread(target=work, args=(d,)) for d in devs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return {"devices": list(devs), "per_device_gbs": [round(out[d], 2) for d in devs],
            "total_gbs": round(sum(out.values()), 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=4.0, help="host slice GiB per device for the big case")
    ap.add_argument("--ndev", type=int, default=0)
    ap.add_argument("--out", default="")
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    ndev = a.ndev or torch.cuda.device_count()
    doc = {"ndev": ndev, "cores": len(os.sched_getaffinity(0))}
    doc["topo"] = sh("nvidia-smi topo -m")
    doc["numa"] = sh("ls -d /sys/devices/system/node/node* 2>/dev/null; cat /sys/devices/system/node/node*/meminfo 2>/dev/null | grep -i memtotal; "
                     "for g in /sys/bus/pci/devices/*; do if [ -e $g/numa_node ] && grep -qi 0x10de $g/vendor 2>/dev/null; then echo $g $(cat $g/numa_node) $(cat $g/current_link_speed 2>/dev/null) x$(cat $g/current_link_width 2>/dev/null); fi; done")
    doc["lscpu"] = sh("lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)|Thread'")
    doc["mem"] = sh("free -g | head -2")

    # 2. raw H2D ceilings
    nb = 1 << 30
    doc["h2d_alone_gbs"] = {}
    for d in range(ndev):
        torch.cuda.set_device(d)
        doc["h2d_alone_gbs"][d] = round(ss.measure_h2d(nb, 3), 2)
    torch.cuda.set_device(0)
    subsets = [list(range(m)) for m in (2, 4, 8) if m <= ndev]
    if ndev >= 8:
        subsets += [[0, 2], [0, 4], [0, 1, 4, 5], [4, 5, 6, 7]]
    doc["h2d_concurrent"] = [concurrent_h2d(sub, nb) for sub in subsets]

    # 3. the host-slice engine
    i386 = np.frombuffer(open(os.path.join(ROOT, "data", "i386.txt"), "rb").read(), np.uint8)
    s = ss.DynamicB200Searcher.new(b"ipsum")
    rows = []
    counts = sorted({1, 2, 4, 8, ndev} & set(range(1, ndev + 1)))
    big = int(a.gib * (1 << 30))
    max_bytes = big * ndev
    t0 = time.perf_counter()
    host = torch.empty(max_bytes, dtype=torch.uint8, pin_memory=True)
    doc["pin_seconds_per_gib"] = round((time.perf_counter() - t0) / (max_bytes / (1 << 30)), 3)
    hv = host.numpy()
    m = i386.size
    for off in range(0, max_bytes, m):  # tile the text (phase is irrelevant for the probe)
        c = min(m, max_bytes - off)
        hv[off:off + c] = i386[:c]
    for nd in counts:
        ctx = ss.Context(nd)
        sizes = [1 << 20, 16 << 20, 256 << 20, big * nd] if not a.quick else [16 << 20, big * nd]
        for size in sizes:
            for mode, name in ((1, "dma"), (2, "inplace_ldg"), (3, "inplace_tma")):
                ss.set_host_path(mode, 0, -1)
                buf = host[:size]
                assert ctx.find_in_host(s, buf) is None
                reps = 3 if size >= (256 << 20) else 20
                t0 = time.perf_counter()
                for _ in range(reps):
                    ctx.find_in_host(s, buf)
                dt = (time.perf_counter() - t0) / reps
                rows.append({"ndev": nd, "bytes": size, "mode": name, "gbs": round(size / dt / 1e9, 2),
                             "ms": round(dt * 1e3, 4), "stats": ctx.last_host_stats()})
        ctx.close()
    ss.set_host_path(0, 0, -1)
    doc["engine"] = rows

    # 4. pageable
    pg_rows = []
    n_pg = min(big, 2 << 30)
    pageable = np.empty(n_pg, np.uint8)
    pageable[:] = hv[:n_pg]
    for nd in counts:
        ctx = ss.Context(nd)
        for threads in (-1, 0):
            if threads == 0 and nd > 1:
                continue
            ss.set_host_path(0, 0, threads)
            assert ctx.find_in_host(s, pageable) is None
            t0 = time.perf_counter()
            for _ in range(3):
                ctx.find_in_host(s, pageable)
            dt = (time.perf_counter() - t0) / 3
            pg_rows.append({"ndev": nd, "bytes": n_pg, "copy_threads": threads, "gbs": round(n_pg / dt / 1e9, 2),
                            "stats": ctx.last_host_stats()})
        ctx.close()
    ss.set_host_path(0, 0, -1)
    doc["pageable"] = pg_rows
    txt = json.dumps(doc, indent=1)
    print(txt)
    if a.out:
        with open(a.out, "w") as f:
            f.write(txt)


if __name__ == "__main__":
    main()
