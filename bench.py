#!/usr/bin/env python
"""bench.py -- haystack GB/s scanned on the i386 long-haystack workload (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun ... bench.py --gpus N ...          (one rank per GPU; driver launches this for N > 1)

Workload (SURVEY.md section 8d, config 2'): data/i386.txt tiled to --gib GiB per GPU (default 8),
needle 'ipsum' built with DynamicAvx2Searcher::new semantics (second anchor = last byte).  The
needle is absent, so every step scans the whole haystack: algorithmic bytes = haystack bytes, 1
byte read per haystack byte, 0 written.  The haystack is 60x larger than L2, so no L2 flush is
needed between steps.  One step = one search = one kernel launch per GPU.

  value      whole-job GB/s with the haystack resident in HBM (ss_b200_find_in_device_async on
             the torch stream, CUDA events, max over ranks).  N > 1: rank r owns the start
             positions of global bytes [r*S, (r+1)*S) plus a k-1 byte right halo, and every step
             ends with the 8-byte NCCL all_reduce(MIN) of the first offsets (weak scaling).
  e2e        the same search through the C-ABI host-slice call on a pinned HOST buffer: chunked
             host->device copies, scans and the result read are all inside the timed region.
             N = 1: ss_b200_find_in_host.  N > 1: rank 0 alone calls ss_b200_find_in_host_multi on ONE
             host slice of N x the per-GPU size, striped by the library over all N GPUs / PCIe links
             (the other ranks wait on a CPU barrier).  e2e.roofline puts it against the pinned
             host->device copy bandwidth measured in the same run.
  roofline   achieved HBM GB/s of the scan kernel (per-launch CUDA events) / measured copy peak.
  parity     before anything is timed the needle is PLANTED (last k bytes of the last rank; straddling
             a rank boundary; in rank 0 and in a later rank at once) and searched through every
             exchange; expected and reported offsets are printed.
  cpu_baseline  the C restatement of DynamicAvx2Searcher (oracle/, the checker) timed on this
             box's host cores over a bounded sample of the same workload (rank 0, N = 1).

--impl reference times that CPU restatement with all host threads (rank 0 only).
The product path never touches oracle/: it is imported only inside cpu_baseline()/reference_arm().
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "haystack GB/s scanned (i386 long-haystack)"
UNIT = "GB/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--gib", type=float, default=8.0, help="haystack GiB per GPU")
    p.add_argument("--total-gib", type=float, default=0.0,
                   help="strong-scaling point: total haystack GiB split evenly over the GPUs (overrides --gib)")
    p.add_argument("--needle", default="ipsum")
    p.add_argument("--variant", type=int, default=0, help="0 auto, 1 LDG, 2 TMA")
    p.add_argument("--tuning", default="", help="ctas_per_sm,unroll,tile_kib,stages")
    p.add_argument("--e2e-steps", type=int, default=5)
    p.add_argument("--e2e-gib", type=float, default=0.0, help="host haystack GiB per GPU (0 = same as --gib)")
    p.add_argument("--cpu-sample-gib", type=float, default=1.0)
    p.add_argument("--exchange", default="nccl", choices=["nccl", "peer"],
                   help="N > 1: how the per-shard first offsets are MIN-reduced (NCCL all_reduce, or stores into "
                        "peer mailboxes fused into the scan epilogue)")
    p.add_argument("--mode", default="single", choices=["single", "many"],
                   help="single = one haystack sharded by start position (the headline); many = the batched "
                        "many-haystack mode: every GPU holds its own set of haystacks, per-haystack flags are "
                        "OR-ed across GPUs as a bit-packed bitmap (all_reduce(SUM) over disjoint bits)")
    p.add_argument("--sustained-steps", type=int, default=300,
                   help="a second, longer timed run for the sustained (power-capped) rate; 0 = skip")
    p.add_argument("--no-extras", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


def load_i386() -> bytes:
    with open(os.path.join(ROOT, "data", "i386.txt"), "rb") as f:
        return f.read()


def peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy read+write)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def committed_read_peak():
    """Read-only stream ceiling measured with tools/read_peak.cu (profiles/r01_read_peak.txt)."""
    try:
        best = 0.0
        with open(os.path.join(ROOT, "profiles", "r01_read_peak.txt")) as f:
            for line in f:
                if line.startswith("read-only stream") and line.rstrip().endswith("GB/s"):
                    best = max(best, float(line.split(":")[1].split()[0]))
        return best or None
    except Exception:
        return None


def committed_traffic():
    """DRAM bytes per launch of the scan kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def workload_config(args, world: int) -> dict:
    """The `config` object, identical for both arms (the reference arm times a bounded sample of this
    workload; what the sample was is said in its cpu_baseline.sample, not here)."""
    k = len(args.needle.encode())
    return {"workload": f"i386 long-haystack: data/i386.txt tiled to {args.gib:g} GiB per GPU, needle "
                        f"{args.needle!r} (absent => full scan), DynamicAvx2Searcher::new semantics",
            "haystack_bytes_per_gpu": int(args.gib * (1 << 30)), "n_gpus": world, "needle_len": k, "position": k - 1,
            "l2": "haystack >> L2 (126 MB): every step streams from memory, no flush needed"}


# ---------------------------------------------------------------------------------------------
# clocks


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                    pw.append(float(parts[3]))
                except ValueError:
                    continue
                for nm, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU legs (the only places that touch oracle/)


def tiled_host(i386: bytes, length: int, global_start: int = 0):
    import numpy as np

    src = np.frombuffer(i386, np.uint8)
    m = src.size
    out = np.empty(length, np.uint8)
    ph = global_start % m
    first = min(length, m - ph)
    out[:first] = src[ph:ph + first]
    pos = first
    while pos < length:
        c = min(m, length - pos)
        out[pos:pos + c] = src[:c]
        pos += c
    return out


def cpu_baseline(i386: bytes, needle: bytes, sample_gib: float):
    """DynamicAvx2Searcher restatement (oracle/sliceslice_oracle.c), 1 thread = the reference's own
    execution model (no threads in src/), plus an all-cores figure for context."""
    import oracle

    n = int(sample_gib * (1 << 30))
    hay = tiled_host(i386, n)
    cores = len(os.sched_getaffinity(0))
    assert oracle.find(hay[: 4 << 20], needle) is None
    times = []
    t_end = time.perf_counter() + 12.0
    while len(times) < 3 or (time.perf_counter() < t_end and len(times) < 50):
        t0 = time.perf_counter()
        r = oracle.find(hay, needle)
        times.append(time.perf_counter() - t0)
        assert r is None
    one = n / statistics.median(times) / 1e9
    mt = []
    for _ in range(5):
        t0 = time.perf_counter()
        r = oracle.find(hay, needle, threads=cores)
        mt.append(time.perf_counter() - t0)
        assert r is None
    # config 1: 'ipsum' over the 857 KB i386.txt itself (cache-resident), one thread
    c1 = []
    for _ in range(2000):
        t0 = time.perf_counter()
        r = oracle.find(i386, needle)
        c1.append(time.perf_counter() - t0)
    c1_us = statistics.median(c1) * 1e6
    # the literal configs 2 and 3 on one host thread, for the `extras` block of the GPU arm
    with open(os.path.join(ROOT, "data", "words.txt"), "rb") as f:
        words = [w for w in f.read().split(b"\n") if w]
    sw = [words[i] for i in sorted(range(len(words)), key=lambda i: (len(words[i]), i))]
    c2, c3 = [], []
    for _ in range(3):
        t0 = time.perf_counter()
        offs = oracle.long_sweep(words, i386)
        c2.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        m, _bm = oracle.short_sweep(sw, want_bitmap=False)
        c3.append(time.perf_counter() - t0)
    assert int(offs.sum()) == 809985317 and m == 39105
    import numpy as np

    g_nd, g_hs, g_pn, g_ph, g_exp = random_grid()
    g_pn_k, g_ph_k = np.tile(g_pn, 1000), np.tile(g_ph, 1000)
    r_ = oracle.pairs(g_nd, g_hs, g_pn, g_ph)
    assert [None if v == oracle.NPOS else int(v) for v in r_] == g_exp
    gt = []
    for _ in range(5):
        t0 = time.perf_counter()
        oracle.pairs(g_nd, g_hs, g_pn_k, g_ph_k)
        gt.append(time.perf_counter() - t0)
    return {"value": round(one, 3), "unit": UNIT, "cores": 1, "kind": "port",
            "random_bench_grid_ns_per_search": round(min(gt) * 1e9 / g_pn_k.size, 2),
            "config1_ipsum_i386_us": round(c1_us, 2), "config1_gbs_cache_resident": round(len(i386) / c1_us / 1e3, 2),
            "config2_literal_ms": round(min(c2) * 1e3, 3), "config3_short_ms": round(min(c3) * 1e3, 3),
            "sample": f"i386.txt tiled to {sample_gib:g} GiB in host DRAM, needle {needle!r} absent, "
                      f"{len(times)} full scans, median; C restatement of DynamicAvx2Searcher (gcc -O3 -mavx2)",
            "all_cores": {"value": round(n / min(mt) / 1e9, 3), "cores": cores}}


def reference_arm(args):
    """The reference's CPU implementation of the path (restated in C; rustc is not in the image) with
    all host threads, on a bounded sample of the same workload.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import oracle

    i386 = load_i386()
    needle = args.needle.encode()
    cores = len(os.sched_getaffinity(0))
    gib = min(args.gib, 4.0)
    n = int(gib * (1 << 30))
    hay = tiled_host(i386, n)
    for _ in range(max(args.warmup, 1)):
        assert oracle.find(hay, needle, threads=cores) is None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.find(hay, needle, threads=cores)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt / 1e9
    # the reference's own execution model is one thread per search_in (no threads in src/): time that
    # too, on the first GiB of the sample, so the line carries both readings
    n1 = min(n, 1 << 30)
    one = []
    for _ in range(3):
        t1 = time.perf_counter()
        assert oracle.find(hay[:n1], needle) is None
        one.append(time.perf_counter() - t1)
    single = n1 / min(one) / 1e9
    sample = (f"i386.txt tiled to {gib:g} GiB in host DRAM (bounded sample of the {args.gib:g} GiB/GPU workload; "
              f"a GB/s rate, so it compares with the full-size figure), needle {args.needle!r} absent, {cores} threads "
              f"over contiguous slices with a k-1 halo; C restatement of DynamicAvx2Searcher (gcc -O3 -mavx2)")
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    emit({
        "impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "sample_bytes": n,
                         "single_thread": {"value": round(single, 3), "cores": 1,
                                           "sample": f"first {n1 / (1 << 30):g} GiB of the same buffer, best of 3"}},
        "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ---------------------------------------------------------------------------------------------
# our arm


def extras_single_gpu(ss, torch, i386: bytes, hay, args):
    """Context numbers outside the headline timed region (rank 0, N = 1): other absent needles on the
    same haystack, and the literal config 2 (4 585 words x 857 KB i386.txt; L2-resident, launch-bound,
    so no HBM fraction is claimed for it)."""
    out = {}
    ws = torch.zeros(16, dtype=torch.uint8, device="cuda")
    res = torch.zeros(1, dtype=torch.int64, device="cuda")
    sweep = {}
    for nd in ("zq", "ipsumdol", "consecteturadipi", "\xff"):
        s = ss.DynamicB200Searcher.new(nd.encode("latin-1"))
        for _ in range(2):
            s.find_in_async(hay, res, ws)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            s.find_in_async(hay, res, ws)
        e1.record()
        torch.cuda.synchronize()
        assert int(res.item()) == ss.DEVICE_NONE
        sweep[repr(nd.encode("latin-1"))] = round(hay.numel() * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
    out["absent_needle_sweep_gbs"] = sweep

    # second anchor chosen from a sampled histogram of the haystack (16 MiB in 4 KiB granules)
    d_hist = torch.zeros(256, dtype=torch.int64, device="cuda")
    ss._check(ss.lib().ss_b200_byte_histogram_device_async(hay.data_ptr(), hay.numel(), 16 << 20, d_hist.data_ptr(),
                                                           torch.cuda.current_stream().cuda_stream))
    hist = d_hist.cpu().numpy().astype("uint64")
    rare = {}
    for nd in (b"consecteturadipi", b"the quick brown fox jumps ov"):
        row = {}
        for label, s in (("new", ss.DynamicB200Searcher.new(nd)),
                         ("rarest", ss.DynamicB200Searcher.with_rarest_position(nd, hist))):
            for _ in range(2):
                s.find_in_async(hay, res, ws)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                s.find_in_async(hay, res, ws)
            e1.record()
            torch.cuda.synchronize()
            assert int(res.item()) == ss.DEVICE_NONE
            row[label] = {"position": s.position,
                          "gbs": round(hay.numel() * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)}
        rare[repr(nd)] = row
    out["rarest_position_gbs"] = rare
    # the histogram kernel itself: every byte of the 8 GiB haystack, and the 16 MiB sample used above
    hrow = {}
    for label, sample in (("every_byte_exact", hay.numel()), ("default_sample_16MiB", 0)):
        for _ in range(2):
            ss._check(ss.lib().ss_b200_byte_histogram_device_async(hay.data_ptr(), hay.numel(), sample,
                                                                   d_hist.data_ptr(),
                                                                   torch.cuda.current_stream().cuda_stream))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ss._check(ss.lib().ss_b200_byte_histogram_device_async(hay.data_ptr(), hay.numel(), sample,
                                                                   d_hist.data_ptr(),
                                                                   torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        hrow[label] = {"ms": round(ms, 4), "bytes_counted": int(d_hist.sum().item()),
                       "gbs": round(int(d_hist.sum().item()) / (ms * 1e-3) / 1e9, 1)}
    out["byte_histogram"] = hrow

    # SURVEY 8f-1 count mode: the same scan without the early return, every occurrence counted; the
    # expected count of the 8 GiB tiling follows from the match positions of one period
    ws32 = torch.zeros(32, dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    counts = {}
    for nd in (b"ipsum", b"segment", b"the"):
        P, m = periodic_matches(i386, nd)
        n = hay.numel()
        expect = int(sum((n - len(nd) - int(p)) // m + 1 for p in P if p <= n - len(nd)))
        s = ss.DynamicB200Searcher.new(nd)
        for _ in range(2):
            s.count_in_async(hay, cnt, ws32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            s.count_in_async(hay, cnt, ws32)
        e1.record()
        torch.cuda.synchronize()
        assert int(cnt.item()) == expect, (nd, int(cnt.item()), expect)
        counts[repr(nd)] = {"occurrences": expect,
                            "gbs": round(n * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)}
    out["count_mode_gbs"] = counts

    with open(os.path.join(ROOT, "data", "words.txt"), "rb") as f:
        words = [w for w in f.read().split(b"\n") if w]
    hs = ss.DeviceHaystack.upload(i386)
    searchers = [ss.DynamicB200Searcher.new(w) for w in words]
    lib_ = ss.lib()
    import ctypes as _C

    def literal_loop():
        # the loop of bench/benches/i386.rs:252-256 with the least host code around each call:
        # ss_b200_find_in per needle, straight through ctypes
        out = _C.c_size_t(0)
        ref = _C.byref(out)
        fn, hh = lib_.ss_b200_find_in, hs._h
        tot = 0
        for s_ in searchers:
            fn(s_._s, hh, ref)
            tot += out.value
        return tot

    sync_ms = {}
    for label, on in (("resident_service_kernel", True), ("one_launch_per_call", False)):
        ss.set_sync_service(on)
        best = None
        for it in range(4):
            t0 = time.perf_counter()
            tot = literal_loop()
            dt = time.perf_counter() - t0
            if it:
                best = dt if best is None else min(best, dt)
        assert tot == 809985317
        sync_ms[label] = round(best * 1e3, 3)
    ss.set_sync_service(True)
    best = sync_ms["resident_service_kernel"] * 1e-3
    # the same loop stream-ordered: one find_in_async per needle (its own searcher, its own launch, no
    # batching API), results in device memory, one synchronisation at the end
    t_i386 = torch.frombuffer(bytearray(i386), dtype=torch.uint8).cuda()
    res_all = torch.zeros(len(words), dtype=torch.int64, device="cuda")
    slots = [res_all[i:i + 1] for i in range(len(words))]
    abest = None
    for it in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s_, slot in zip(searchers, slots):
            s_.find_in_async(t_i386, slot, ws)
        total_off = int(res_all.sum().item())
        dt = time.perf_counter() - t0
        if it:
            abest = dt if abest is None else min(abest, dt)
    assert total_off == 809985317
    batch = ss.Batch(words, [])
    bbest = None
    for it in range(4):
        t0 = time.perf_counter()
        o2 = batch.find_all_in(hs)
        dt = time.perf_counter() - t0
        if it:
            bbest = dt if bbest is None else min(bbest, dt)
    assert int(o2.sum()) == 809985317
    # the same workload end to end from HOST buffers: haystack upload, needle table upload, one launch,
    # offsets back on the host -- everything a caller holding host memory pays
    ebest = None
    for it in range(4):
        t0 = time.perf_counter()
        hs2 = ss.DeviceHaystack.upload(i386)
        b2 = ss.Batch(words, [])
        o3 = b2.find_all_in(hs2)
        dt = time.perf_counter() - t0
        b2.close()
        hs2.close()
        if it:
            ebest = dt if ebest is None else min(ebest, dt)
    assert int(o3.sum()) == 809985317
    cpp_host = compiled_host_literal_loop()
    out["config2_literal"] = {
        "e2e_host_buffers_ms_per_iteration": round(ebest * 1e3, 3),
        "what": "all 4585 words.txt needles over the 857425-byte i386.txt, device-resident, host wall clock",
        # the loop of bench/benches/i386.rs:252-256 is compiled code on both sides: the headline figure is the
        # compiled C++ host's (tests/cpp/bench_latency.cpp); the same loop driven from Python is beside it
        "api_faithful_ms_per_iteration": ((cpp_host or {}).get("resident_service_kernel_ms") or round(best * 1e3, 3)),
        "api_faithful_how": "one synchronous ss_b200_find_in per needle; short device-resident haystacks are served by "
                            "a resident kernel (one PCIe round trip per call, no launch)",
        "api_faithful_compiled_host": cpp_host,
        "api_faithful_python_ctypes_ms": round(best * 1e3, 3),
        "api_faithful_python_ctypes_one_launch_per_call_ms": sync_ms["one_launch_per_call"],
        "api_faithful_async_ms_per_iteration": ((cpp_host or {}).get("stream_ordered_ms") or round(abest * 1e3, 3)),
        "api_faithful_async_python_ctypes_ms": round(abest * 1e3, 3),
        "batched_single_launch_ms_per_iteration": round(bbest * 1e3, 3),
        "examined_bytes": 810016020, "sum_first_offsets": 809985317,
        "readme_i7_6700_ms": 35.181,
        "note": "L2-resident and early-exit: launch-latency-bound, no HBM-fraction claim",
    }
    # config 4: the same buffer refilled with the counter-based generator (alphabet 0..254), needles
    # of length 1/4/16/64 that contain 0xFF (absent by construction), DynamicAvx2Searcher::new positions
    import random as _random

    ss.fill_random(hay, 0, 0x5EEDB20000000001)
    c4 = {}
    for k in (1, 4, 16, 64):
        nd = bytearray(_random.Random(1000 + k).randrange(255) for _ in range(k))
        nd[min(1, k - 1)] = 0xFF
        s = ss.DynamicB200Searcher.new(bytes(nd))
        for _ in range(3):
            s.find_in_async(hay, res, ws)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            s.find_in_async(hay, res, ws)
        e1.record()
        torch.cuda.synchronize()
        assert int(res.item()) == ss.DEVICE_NONE
        c4[f"k={k}"] = round(hay.numel() * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
    out["config4_random_haystack_gbs"] = c4

    sw = [words[i] for i in sorted(range(len(words)), key=lambda i: (len(words[i]), i))]
    tri = ss.Batch(sw, sw)
    tbest = None
    for it in range(4):
        t0 = time.perf_counter()
        bm, matches = tri.search_triangular()
        dt = time.perf_counter() - t0
        if it:
            tbest = dt if tbest is None else min(tbest, dt)
    assert matches == 39105
    out["config3_short_haystack"] = {
        "what": "every word in every not-shorter word (10513405 pairs, bench/benches/i386.rs:118-131), one "
                "batched launch, bitmap copied back to the host, host wall clock",
        "ms_per_iteration": round(tbest * 1e3, 3), "ns_per_pair": round(tbest * 1e9 / 10513405, 4),
        "matches": int(matches), "readme_i7_6700_ms": 79.416,
    }
    # SURVEY 8f-4, bench/benches/random.rs:12-99: prefixes of data/needle in prefixes of data/haystack, sizes
    # {1,5,10,20,50,100,1000}, every needle size against every not-smaller haystack size (28 searches).
    # Pure latency regime: one batched launch for the grid, and the grid replicated 1000x in one launch
    # for a per-search figure; the CPU restatement's time is in cpu_baseline.random_bench_grid_ns_per_search.
    import numpy as np

    nd_blob, hs_blob, pn, ph, exp = random_grid()
    gb = ss.Batch(nd_blob, hs_blob)
    bm, offs = gb.search_pairs(pn, ph)
    assert [None if v == ss.NPOS else int(v) for v in offs] == exp
    t1 = None
    for it in range(30):
        t0 = time.perf_counter()
        gb.search_pairs(pn, ph, want_offsets=False)
        dt = time.perf_counter() - t0
        if it:
            t1 = dt if t1 is None else min(t1, dt)
    pn_k, ph_k = np.tile(pn, 1000), np.tile(ph, 1000)
    tk = None
    for it in range(6):
        t0 = time.perf_counter()
        bmk, _ = gb.search_pairs(pn_k, ph_k, want_offsets=False)
        dt = time.perf_counter() - t0
        if it:
            tk = dt if tk is None else min(tk, dt)
    assert int(np.unpackbits(bmk.view(np.uint8)).sum()) == 1000 * sum(e is not None for e in exp)
    out["random_bench_grid"] = {
        "what": "bench/benches/random.rs:12-99: 28 (needle prefix, haystack prefix) searches, sizes 1..1000 bytes; "
                "host arrays in and out, host wall clock",
        "searches": len(pn), "found": sum(e is not None for e in exp),
        "one_launch_for_the_grid_us": round(t1 * 1e6, 2),
        "grid_x1000_in_one_launch_ns_per_search": round(tk * 1e9 / pn_k.size, 3),
    }
    return out


def compiled_host_literal_loop():
    """The same per-needle loop from a COMPILED host (tests/cpp/bench_latency.cpp through the C++ mirror), the
    way the reference's own bench calls it from Rust: what a call costs without the Python interpreter
    around it.  Built with g++ on the spot; None if that is not possible."""
    try:
        import re

        import sliceslice_rs_b200 as ss

        build_dir = os.path.join(ROOT, "tests", "cpp", "_build")
        os.makedirs(build_dir, exist_ok=True)
        exe = os.path.join(build_dir, "bench_latency")
        lib_dir = os.path.dirname(ss.LIB_PATH)
        cmd = ["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
               os.path.join(ROOT, "tests", "cpp", "bench_latency.cpp"), "-o", exe, ss.LIB_PATH,
               "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{lib_dir}", "-Wl,-rpath,/usr/local/cuda/lib64"]
        subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=120)
        res = {}
        for label, flag in (("resident_service_kernel_ms", "1"), ("one_launch_per_call_ms", "0")):
            r = subprocess.run([exe, os.path.join(ROOT, "data", "i386.txt"), os.path.join(ROOT, "data", "words.txt"), flag],
                               capture_output=True, text=True, timeout=300)
            sync = [float(x) for x in re.findall(r"one find_in per word, device haystack\): ([0-9.]+) ms/iteration", r.stdout)]
            asyn = [float(x) for x in re.findall(r"one find_in_device_async per word, one sync\): ([0-9.]+) ms/iteration", r.stdout)]
            if not sync or "sum 809985317" not in r.stdout:
                return None
            res[label] = round(min(sync), 3)
            res["stream_ordered_ms"] = round(min(asyn), 3) if asyn else None
        res["what"] = "tests/cpp/bench_latency.cpp: one synchronous find_in per needle from C++ (best of 4 iterations)"
        return res
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


def random_grid():
    """The (needle, haystack) prefixes of bench/benches/random.rs:12-99 and what each search must return."""
    import numpy as np

    with open(os.path.join(ROOT, "data", "haystack"), "rb") as f:
        haystack = f.read()
    with open(os.path.join(ROOT, "data", "needle"), "rb") as f:
        needle = f.read()
    sizes = [1, 5, 10, 20, 50, 100, 1000]
    needles = [needle[:s_] for s_ in sizes]
    hays = [haystack[:s_] for s_ in sizes]
    pn = np.array([i for i in range(len(sizes)) for j in range(i, len(sizes))], np.uint32)
    ph = np.array([j for i in range(len(sizes)) for j in range(i, len(sizes))], np.uint32)
    exp = []
    for a_, b_ in zip(pn, ph):
        e = hays[b_].find(needles[a_])
        exp.append(None if e < 0 else e)
    return needles, hays, pn, ph, exp


def periodic_matches(i386: bytes, needle: bytes):
    """Sorted start positions in [0, m) of `needle` in the infinite tiling of i386 (m = len(i386))."""
    import numpy as np

    m, k = len(i386), len(needle)
    ext = i386 + i386[:k - 1]
    out, i = [], ext.find(needle)
    while i >= 0:
        out.append(i)
        i = ext.find(needle, i + 1)
    return np.asarray(out, np.int64), m


def expected_set_flags(i386: bytes, needle: bytes, global_start: int, offsets):
    """flags[h] = needle in tiled[global_start + offsets[h] : global_start + offsets[h+1]] for every
    haystack of a set cut out of the i386 tiling (haystacks shorter than one period), computed from the
    match positions of one period -- the CPU-side expectation for the full-size many-haystack run."""
    import numpy as np

    P, m = periodic_matches(i386, needle)
    k = len(needle)
    s = global_start + offsets[:-1].astype(np.int64)
    e = global_start + offsets[1:].astype(np.int64)
    if P.size == 0:
        return np.zeros(s.size, np.uint8)
    r = s % m
    idx = np.searchsorted(P, r)
    cand = np.where(idx < P.size, P[np.minimum(idx, P.size - 1)], P[0] + m)
    return ((s - r + cand + k) <= e).astype(np.uint8)


def many_mode_run(args, ss, torch, dist, world, rank, local, steps, with_clocks=True):
    """North-star batched mode: the haystack SET is partitioned across the GPUs (each rank holds its own
    haystacks, needles replicated), one pass per needle at the long-scan rate.  Every haystack lives on
    one rank, so the per-rank flags are disjoint: each rank bit-packs its slice into its bit range of one
    global bitmap and the bitmaps are OR-ed with all_reduce(SUM) on int32 words (NCCL has no bitwise OR; a
    sum of disjoint bits is one) -- 1 bit per haystack on the wire.  Weak scaling: --gib per GPU.
    Returns the result dict on rank 0 (None elsewhere)."""
    import numpy as np

    from sliceslice_rs_b200 import sharded

    i386 = load_i386()
    needle = args.needle.encode()
    S = int(args.gib * (1 << 30))
    start = rank * S
    # haystack lengths 0..16383 from a fixed generator (same sequence on every rank), cut from the i386 tiling
    rng = np.random.default_rng(20260101)
    n_h = max(1, S // 8192 + S // 131072 + 16)  # ~6 % more than fit: the cut below always finds S
    lens = rng.integers(0, 16384, n_h, dtype=np.int64)
    off = np.zeros(n_h + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    cut = int(np.searchsorted(off, S, side="right")) - 1  # keep the haystacks that fit, the last takes the rest
    off = off[:cut + 1].copy()
    if off[-1] < S and S - off[-1] < len(i386):
        off = np.append(off, S)
    elif off[-1] < S:
        off = np.append(off, off[-1] + len(i386) - 1)
    n_h = off.size - 1
    total_h = n_h * world
    blob_len = int(off[-1])
    src = torch.frombuffer(bytearray(i386), dtype=torch.uint8).cuda()
    blob = torch.empty(blob_len, dtype=torch.uint8, device="cuda")
    ss.fill_tiled(blob, start, src)
    hset = ss.HaystackSet.from_device(blob, torch.from_numpy(off).cuda())
    searcher = ss.DynamicB200Searcher.new(needle)
    local_flags = torch.zeros(n_h, dtype=torch.uint8, device="cuda")
    n_words = (total_h + 31) // 32
    ring = [torch.zeros(n_words, dtype=torch.int32, device="cuda") for _ in range(4)]
    pending = [None] * len(ring)

    def step(i, s_=None):
        j = i % len(ring)
        if pending[j] is not None:
            pending[j].wait()
            pending[j] = None
        (s_ or searcher).search_many_async(hset, local_flags)
        sharded.pack_flags_async(local_flags, rank * n_h, ring[j], total_h)
        if world > 1:
            pending[j] = sharded.reduce_packed_flags(ring[j], async_op=True)
        return ring[j]

    def drain():
        for j, w in enumerate(pending):
            if w is not None:
                w.wait()
                pending[j] = None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # parity at full size, before timing: the bench needle and present ones, every flag of every rank
    # against the CPU expectation
    checks = {}
    for nd in (needle, b"segment", b"the", b"descriptor table"):
        s2 = ss.DynamicB200Searcher.new(nd)
        words = step(0, s2)
        drain()
        got = sharded.unpack_flags(words.cpu(), total_h)
        exp = np.concatenate([expected_set_flags(i386, nd, r * S, off) for r in range(world)])
        assert np.array_equal(got, exp), f"many-haystack flags differ from the CPU expectation for {nd!r}"
        checks[repr(nd)] = int(exp.sum())
    W = max(args.warmup, 3)
    for i in range(W):
        step(i)
    drain()
    barrier()
    K = steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0 and with_clocks:
        sampler.start()
        time.sleep(0.25)
    launches0 = ss.launch_count()
    barrier()
    e0.record()
    for i in range(K):
        step(i)
    drain()
    e1.record()
    barrier()
    launches = ss.launch_count() - launches0
    clocks = sampler.stop() if (rank == 0 and with_clocks) else None
    tt = torch.tensor([e0.elapsed_time(e1), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, launches = float(mx[0]), int(sm[1])
    else:
        ms_total = float(tt[0])
    # a present needle on the same set: the hit path marks haystacks instead of stopping
    hot = {}
    plain = ss.HaystackSet.from_device(blob, hset.offsets, prepared=False)
    for nd in (b"segment", b"the"):
        s2 = ss.DynamicB200Searcher.new(nd)
        row = {}
        for label, st in (("prepared_set", hset), ("no_hints", plain)):
            for _ in range(2):
                s2.search_many_async(st, local_flags)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                s2.search_many_async(st, local_flags)
            b.record()
            torch.cuda.synchronize()
            row[label] = round(blob_len * 5 / (a.elapsed_time(b) * 1e-3) / 1e9, 1)
        hot[repr(nd)] = row
    if rank != 0:
        return None
    peak, peak_src = peak_hbm()
    value = blob_len * world * K / (ms_total * 1e-3) / 1e9
    return {
        "metric": "haystack GB/s scanned (many-haystack batch)", "value": round(value, 3), "unit": UNIT,
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms_total / K, 5),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"many-haystack batch: {n_h} haystacks per GPU (lengths 0..16383, cut from "
                               f"data/i386.txt tiled to {args.gib:g} GiB per GPU), needle {args.needle!r}, "
                               "flags[h] = search_in(haystack h)",
                   "haystacks_per_gpu": n_h, "blob_bytes_per_gpu": blob_len,
                   "sharding": "haystack set partitioned across GPUs; per step each rank bit-packs its flags into "
                               "its bit range of one global bitmap and the bitmaps are OR-ed with all_reduce(SUM) on "
                               f"int32 words ({n_words * 4} bytes; async, all waited for inside the timed region)",
                   "l2": "blob >> L2 (126 MB)"},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": round(value / world, 2), "peak": peak, "unit": UNIT,
                     "frac": round(value / world / peak, 4), "peak_source": peak_src,
                     "note": "whole step (flag fill + scan + pack + exchange) per GPU, not the kernel alone"},
        "parity": {"flags_equal_cpu_expectation_for": checks},
        "present_needle_gbs_per_gpu": hot,
    }


def many_mode(args, ss, torch, dist, world, rank, local):
    line = many_mode_run(args, ss, torch, dist, world, rank, local, args.steps)
    if line is not None:
        emit(line)


def concurrent_h2d_gbs(torch, devices, nbytes=1 << 30, reps=3):
    """Pinned host->device copy bandwidth with every device of `devices` copying its own buffer at the
    same time (one host thread per device): the PCIe ceiling of the striped host-slice call."""
    import threading

    out = {}
    gate = threading.Barrier(len(devices))

    def work(d):
        torch.cuda.set_device(d)
        h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h.fill_(7)
        t = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}")
        st = torch.cuda.Stream(device=d)
        with torch.cuda.stream(st):
            t.copy_(h, non_blocking=True)
        st.synchronize()
        gate.wait()
        t0 = time.perf_counter()
        with torch.cuda.stream(st):
            for _ in range(reps):
                t.copy_(h, non_blocking=True)
        st.synchronize()
        out[d] = nbytes * reps / (time.perf_counter() - t0) / 1e9

    th = [threading.Thread(target=work, args=(d,)) for d in devices]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return [round(out[d], 2) for d in devices]


_REAL_STDOUT = None


def guard_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print banners there (NCCL's version line comes
    from C code, past sys.stdout): point fd 1 at stderr for the run and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse_args()
    guard_stdout()
    strong = args.total_gib > 0
    if strong:
        args.gib = args.total_gib / max(int(os.environ.get("WORLD_SIZE", "1")), 1)
    if args.impl == "reference":
        reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import sliceslice_rs_b200 as ss

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA path is the only path (no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")  # host-side barriers that keep the GPUs idle
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)

    ss.lib()
    ss.set_scan_variant(args.variant)
    if args.tuning:
        ss.set_scan_tuning(*[int(x) for x in args.tuning.split(",")])
    if args.mode == "many":
        many_mode(args, ss, torch, dist, world, rank, local)
        if world > 1:
            dist.destroy_process_group()
        return

    i386 = load_i386()
    needle = args.needle.encode()
    k = len(needle)
    assert (i386 + i386).find(needle) < 0, "the bench needle must be absent from i386.txt and the tiling seam"
    S = int(args.gib * (1 << 30))  # start positions owned per rank
    total = S * world
    start = rank * S
    span = min(S + k - 1, total - start)  # owned + right halo (the last rank has none)
    owned = S if rank < world - 1 else span - k + 1

    src = torch.frombuffer(bytearray(i386), dtype=torch.uint8).cuda()
    shard = torch.empty(span, dtype=torch.uint8, device="cuda")
    ss.fill_tiled(shard, start, src)
    searcher = ss.DynamicB200Searcher.new(needle)
    ws = torch.zeros(32, dtype=torch.uint8, device="cuda")
    peer = None
    if world > 1:
        from sliceslice_rs_b200.sharded import PeerExchange

        peer = PeerExchange()  # CUDA-IPC mailboxes; used by --exchange peer, the parity plants and the latency rows
    use_peer = peer is not None and args.exchange == "peer"
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    one = torch.zeros(1, dtype=torch.int64, device="cuda")

    def search_once(exchange: str) -> int:
        """One complete sharded search (scan + exchange), synchronous: the global first offset."""
        if exchange == "peer" and peer is not None:
            peer.find_async(searcher, shard, start, owned, ws, one)
        else:
            searcher.find_in_async(shard, one, ws, base_offset=start, start_limit=owned)
            if world > 1:
                dist.all_reduce(one, op=dist.ReduceOp.MIN)
        return int(one.item())

    # ---- parity with PLANTED needles, before anything is timed (src/lib.rs:242-244) ----
    needle_t = torch.tensor(list(needle), dtype=torch.uint8, device="cuda")

    def plant(pos):
        lo, hi = max(pos, start), min(pos + k, start + span)
        if lo >= hi:
            return None
        saved = shard[lo - start:hi - start].clone()
        shard[lo - start:hi - start] = needle_t[lo - pos:hi - pos]
        return (lo - start, saved)

    def unplant(tok):
        if tok is not None:
            shard[tok[0]:tok[0] + tok[1].numel()] = tok[1]

    exchanges = ["nccl", "peer"] if world > 1 else ["single"]
    cases = [("last_k_bytes_of_last_rank", [total - k], total - k),
             ("rank0_and_a_later_position", [4242, (world - 1) * S + S // 2 + 99], 4242)]
    if world > 1:
        cases.insert(1, ("straddling_rank_boundary", [(world // 2) * S - k // 2], (world // 2) * S - k // 2))
    parity = {"absent": {}, "planted": {}}
    for ex in exchanges:
        got = search_once(ex)
        parity["absent"][ex] = {"expected": None, "got": None if got == ss.DEVICE_NONE else got}
        assert got == ss.DEVICE_NONE, f"needle must be absent ({ex}): {got}"
    for name, spots, expect in cases:
        toks = [plant(p) for p in spots]
        torch.cuda.synchronize()
        row = {"planted_at": spots, "expected": expect}
        for ex in exchanges:
            got = search_once(ex)
            row[ex] = got
            assert got == expect, f"planted-needle parity failed: {name} via {ex}: expected {expect}, got {got}"
        parity["planted"][name] = row
        for t in reversed(toks):
            unplant(t)
        torch.cuda.synchronize()
    got = search_once(exchanges[0])
    assert got == ss.DEVICE_NONE, "the plants must be gone again before the timed region"

    # One result slot per step: consecutive searches are independent, so the 8-byte MIN-allreduce of
    # step i runs on NCCL's stream (async_op) while step i+1 already scans; every reduction is
    # waited for before the closing event of the timed region.
    n_slots = max(args.steps, args.warmup, args.sustained_steps, 3)
    results = torch.zeros(n_slots, dtype=torch.int64, device="cuda")
    pending = []

    def step(i, ev_a=None, ev_b=None):
        if ev_a is not None:
            ev_a.record()
        if use_peer:
            # scan + exchange in one call: the scan's last CTA stores into every rank's mailbox
            peer.find_async(searcher, shard, start, owned, ws, results[i:i + 1])
            if ev_b is not None:
                ev_b.record()
            return
        searcher.find_in_async(shard, results[i:i + 1], ws, base_offset=start, start_limit=owned)
        if ev_b is not None:
            ev_b.record()
        if world > 1:
            pending.append(dist.all_reduce(results[i:i + 1], op=dist.ReduceOp.MIN, async_op=True))

    def drain():
        for w in pending:
            w.wait()
        pending.clear()

    def timed_run(K, per_step_events):
        ka = [torch.cuda.Event(enable_timing=True) for _ in range(K)] if per_step_events else [None] * K
        kb = [torch.cuda.Event(enable_timing=True) for _ in range(K)] if per_step_events else [None] * K
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        launches0 = ss.launch_count()
        barrier()
        e0.record()
        for i in range(K):
            step(i, ka[i], kb[i])
        drain()
        e1.record()
        barrier()
        launches = ss.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        assert results[:K].eq(ss.DEVICE_NONE).all().item()
        ms_total = e0.elapsed_time(e1)
        kern = sum(a.elapsed_time(b) for a, b in zip(ka, kb)) / K if per_step_events else 0.0
        tt = torch.tensor([ms_total, kern, float(launches)], dtype=torch.float64, device="cuda")
        if world > 1:
            mx = tt.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tt.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            return float(mx[0]), float(mx[1]), int(sm[2]), clocks
        return float(tt[0]), float(tt[1]), int(tt[2]), clocks

    for i in range(max(args.warmup, 3)):
        step(i)
    drain()
    barrier()
    assert results[: max(args.warmup, 3)].eq(ss.DEVICE_NONE).all().item(), "needle must be absent"

    K = args.steps
    ms_total, kern_avg_ms, launches, clocks = timed_run(K, True)
    value = total * K / (ms_total * 1e-3) / 1e9
    peak, peak_src = peak_hbm()
    kern_bytes = owned + k - 1 if owned else 0  # bytes one launch must examine (not found => all of them)
    achieved = kern_bytes / (kern_avg_ms * 1e-3) / 1e9
    traffic = committed_traffic()

    # ---- sustained: a longer run of the same step, so the power-capped rate is on the record ----
    sustained = None
    if args.sustained_steps > 0:
        ms_s, _, _, clk_s = timed_run(args.sustained_steps, False)
        sustained = {"steps": args.sustained_steps, "ms_per_step": round(ms_s / args.sustained_steps, 5),
                     "value": round(total * args.sustained_steps / (ms_s * 1e-3) / 1e9, 3), "unit": UNIT,
                     "frac_of_peak_per_gpu": round(total * args.sustained_steps / (ms_s * 1e-3) / 1e9 / world / peak, 4),
                     "clocks": clk_s}

    # ---- unpipelined latency of one complete search (scan + exchange + result on the host) ----
    latency = {}
    for ex in exchanges:
        for _ in range(3):
            search_once(ex)
        barrier()
        ts = []
        for _ in range(15):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            search_once(ex)
            ts.append(time.perf_counter() - t0)
        tl = torch.tensor([statistics.median(ts) * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        latency[ex] = round(float(tl[0]), 4)

    # ---- found needles: how long until the answer, planted at 25 % / 50 % / the last k bytes ----
    found = {}
    for label, pos in (("at_25pct", total // 4 + 7), ("at_50pct", total // 2 + 7), ("last_k_bytes", total - k)):
        tok = plant(pos)
        torch.cuda.synchronize()
        row = {"offset": pos}
        for ex in exchanges:
            assert search_once(ex) == pos
            barrier()
            ts = []
            for _ in range(5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                search_once(ex)
                ts.append(time.perf_counter() - t0)
            tl = torch.tensor([min(ts) * 1e3], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tl, op=dist.ReduceOp.MAX)
            row[f"ms_{ex}"] = round(float(tl[0]), 4)
        found[label] = row
        unplant(tok)
        torch.cuda.synchronize()
    found["note"] = ("host wall clock of one synchronous search, max over ranks; the GPUs left of the match must "
                     "finish their shards (nothing earlier exists only once they have looked), the ones to its right "
                     "are stopped through the peer stop words (exchange 'peer') or run to the end (exchange 'nccl')")

    # ---- e2e: host buffer -> C ABI -> result, copies inside the timed region ----------------
    def measure_e2e():
        eg = args.e2e_gib or args.gib
        n_host = min(int(eg * (1 << 30)), S)
        line = None
        if rank == 0:
            total_host = n_host * world
            t0 = time.perf_counter()
            try:
                host = torch.empty(total_host, dtype=torch.uint8, pin_memory=True)
            except Exception:
                n_host = min(n_host, 1 << 30)
                total_host = n_host * world
                host = torch.empty(total_host, dtype=torch.uint8, pin_memory=True)
            pin_s = time.perf_counter() - t0
            # the same tiling as the device-resident workload, generated on the GPU piece by piece
            tmp = torch.empty(n_host, dtype=torch.uint8, device="cuda")
            for r in range(world):
                ss.fill_tiled(tmp, r * n_host, src)
                host[r * n_host:(r + 1) * n_host].copy_(tmp)
            del tmp
            torch.cuda.synchronize()
            if world > 1:
                ctx = ss.Context(devices=list(range(world)))
                call = lambda: ctx.find_in_host(searcher, host)  # noqa: E731
                call_name = ("ss_b200_find_in_host_multi: ONE pinned host slice striped over all GPUs by the library "
                             "(chunk i -> GPU i mod N, a copy/scan ring per GPU), called by rank 0")
                h2d_peak = concurrent_h2d_gbs(torch, list(range(world)))
            else:
                ctx = None
                call = lambda: searcher.find_in(host)  # noqa: E731
                call_name = "ss_b200_find_in_host (pinned host haystack, ring of 3 device buffers)"
                h2d_peak = [round(ss.measure_h2d(1 << 30, 3), 2)]
            # positive parity first: the needle planted in the host slice (restored afterwards)
            hv = host.numpy()
            e2e_parity = {}
            nd_np = np.frombuffer(needle, np.uint8)
            for label, pos in (("last_k_bytes", total_host - k), ("middle", (total_host // 2) - k // 2), ("early", 4242)):
                saved = hv[pos:pos + k].copy()
                hv[pos:pos + k] = nd_np
                got = call()
                hv[pos:pos + k] = saved
                e2e_parity[label] = {"expected": pos, "got": got}
                assert got == pos, f"e2e parity: planted at {pos}, got {got}"
            for _ in range(2):
                assert call() is None
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                r_ = call()
            dt = time.perf_counter() - t0
            assert r_ is None
            if ctx is not None:
                st = ctx.last_host_stats()
                h2d, chunks = st["h2d_bytes"], st["chunks"]
            else:
                chunk = 64 << 20
                chunks = max(1, -(-(n_host - k + 1) // chunk))
                h2d = n_host + (chunks - 1) * (k - 1)
            val = total_host * args.e2e_steps / dt / 1e9
            line = {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(8 * chunks), "host_bytes": total_host, "host_bytes_per_gpu": n_host,
                    "steps": args.e2e_steps, "pin_seconds": round(pin_s, 2),
                    "timer": "host wall clock around the synchronous C-ABI call on rank 0",
                    "call": call_name, "parity": e2e_parity,
                    "roofline": {"bound": "pcie", "achieved": round(val, 2), "peak": round(sum(h2d_peak), 2),
                                 "unit": UNIT, "frac": round(val / sum(h2d_peak), 4),
                                 "peak_source": "pinned cudaMemcpyAsync host->device, 1 GiB per device, all devices "
                                                "copying at once, measured in this run",
                                 "per_device_peak": h2d_peak}}
            if world == 1:
                # the other data paths of the same call, for the record: pinned input read in place over PCIe,
                # and an ordinary (pageable) buffer staged through the pinned ring by the copy pool
                modes = {}
                for mode, nm in ((2, "in_place_ldg"), (3, "in_place_tma")):
                    ss.set_host_path(mode, 0, -1)
                    sub = host[:min(n_host, 2 << 30)]
                    assert searcher.find_in(sub) is None
                    t0 = time.perf_counter()
                    for _ in range(2):
                        searcher.find_in(sub)
                    modes[nm] = round(sub.numel() * 2 / (time.perf_counter() - t0) / 1e9, 3)
                ss.set_host_path(0, 0, -1)
                sizes = {}
                for mib in (1, 16, 256):
                    sub = host[:mib << 20]
                    row = {}
                    for mode, nm in ((1, "dma_ring"), (2, "in_place_ldg")):
                        ss.set_host_path(mode, 0, -1)
                        assert searcher.find_in(sub) is None
                        reps = 30 if mib <= 16 else 5
                        t0 = time.perf_counter()
                        for _ in range(reps):
                            searcher.find_in(sub)
                        row[nm] = round(sub.numel() * reps / (time.perf_counter() - t0) / 1e9, 3)
                    sizes[f"{mib}MiB"] = row
                ss.set_host_path(0, 0, -1)
                line["other_data_paths_gbs"] = modes
                line["by_slice_size_gbs"] = sizes
                n_pg = min(n_host, 2 << 30)
                pageable = np.empty(n_pg, np.uint8)
                pageable[:] = hv[:n_pg]
                assert searcher.find_in(pageable) is None
                t0 = time.perf_counter()
                for _ in range(3):
                    searcher.find_in(pageable)
                line["pageable_host_buffer_gbs"] = round(n_pg * 3 / (time.perf_counter() - t0) / 1e9, 3)
                del pageable
            if ctx is not None:
                ctx.close()
            del host
        host_barrier()
        if world > 1:
            # for comparison: one process per GPU, each searching its own pinned slice at the same time
            # (round 1's e2e); what a host gets WITHOUT the context call
            n_pp = min(n_host, 2 << 30)
            hp = torch.empty(n_pp, dtype=torch.uint8, pin_memory=True)
            hp.copy_(shard[:n_pp])
            torch.cuda.synchronize()
            assert searcher.find_in(hp) is None
            host_barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                searcher.find_in(hp)
            dt = time.perf_counter() - t0
            td = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
            if rank == 0 and line is not None:
                line["one_process_per_gpu_gbs"] = round(n_pp * world * 3 / float(td[0]) / 1e9, 3)
            del hp
        return line

    e2e = None
    if not args.no_e2e:
        try:
            e2e = measure_e2e()
        except Exception as e:  # noqa: BLE001
            e2e = {"error": f"{type(e).__name__}: {e}"}
            host_barrier()

    # context blocks: a failure here must not cost the headline line its numbers
    extras = {}
    if not args.no_extras:
        if rank == 0 and world == 1:
            try:
                extras = extras_single_gpu(ss, torch, i386, shard, args)
            except Exception as e:  # noqa: BLE001
                extras = {"error": f"{type(e).__name__}: {e}"}
        del shard
        torch.cuda.empty_cache()
        # the north-star's batched many-haystack mode at this N (its own full-size parity check inside)
        try:
            mm = many_mode_run(args, ss, torch, dist, world, rank, local, min(args.steps, 40), with_clocks=False)
            if rank == 0:
                extras["many_haystack_mode"] = {kk: mm[kk] for kk in ("value", "unit", "ms_per_step", "steps", "config",
                                                                      "parity", "present_needle_gbs_per_gpu",
                                                                      "gpu_launches")}
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                extras["many_haystack_mode"] = {"error": f"{type(e).__name__}: {e}"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        torch.cuda.empty_cache()
        try:
            cpu = cpu_baseline(i386, needle, args.cpu_sample_gib)
        except Exception as e:  # noqa: BLE001
            cpu = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        cfg = workload_config(args, world)
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_total / K, 5), "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": cfg,
            "implementation": {
                "sharding": ("single GPU" if world == 1 else
                             "contiguous start-position ranges + k-1 byte right halo; " +
                             ("first offsets MIN-reduced through peer mailboxes: 8-byte NVLink stores fused into the "
                              "scan epilogue + a one-warp min kernel (no collective call)" if use_peer else
                              "NCCL all_reduce(MIN) of the 8-byte first offset per step, overlapped with the next "
                              "step's scan (async_op, all waited for inside the timed region)")),
                "kernel_variant": {0: "auto", 1: "ldg", 2: "tma"}[args.variant],
            },
            "gpu_launches": launches,
            "clocks": clocks,
            "e2e": e2e,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": UNIT,
                         "frac": round(achieved / peak, 4), "peak_source": peak_src,
                         "frac_of_nominal_8000": round(achieved / 8000.0, 4),
                         "read_only_stream_gbs": committed_read_peak(),
                         "frac_of_read_only_stream": (round(achieved / committed_read_peak(), 4)
                                                      if committed_read_peak() else None),
                         "kernel_ms": round(kern_avg_ms, 5), "algorithmic_bytes_per_launch": kern_bytes,
                         # the committed ncu capture is of the default 8 GiB launch: it says nothing about
                         # launches of another size
                         "traffic": ((traffic or {}).get("dram_bytes_per_launch")
                                     if abs(kern_bytes - (8 << 30)) < (1 << 20) else None),
                         "traffic_source": ((traffic or {}).get("source")
                                            if abs(kern_bytes - (8 << 30)) < (1 << 20) else None)},
            "cpu_baseline": cpu,
            "parity": parity,
            "sustained": sustained,
            "single_search_latency_ms": latency,
            "found_needle": found,
        }
        if extras:
            line["extras"] = extras
        emit(line)
    if peer is not None:
        peer.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
