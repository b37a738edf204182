#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers the roofline argument needs.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--json profiles/traffic.json] > profiles/<name>.txt
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max",
]


def main():
    rep = sys.argv[1]
    out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of {rep}")
    traffic = None
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"\nkernel: {name}")
        vals = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals[k] = (r[i], units[i])
                print(f"  {k:86s} {r[i]:>18s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.1:
                    stalls.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        print("  warp stall reasons (warps stalled per issue-active cycle): " +
              ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)))
        try:
            def to_bytes(key):
                v, u = vals[key]
                m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
                return float(v) * m
            rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
            t, tu = vals["gpu__time_duration.sum"]
            sec = float(t) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[tu]
            print(f"  => DRAM traffic per launch {rd + wr:.0f} B (read {rd:.0f} + write {wr:.0f}); "
                  f"{(rd + wr) / sec / 1e9:.1f} GB/s under the profiler (serialised, cold cache)")
            traffic = {"dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
                       "kernel": name, "source": rep.replace("gpurun_out/", "profiles/<summary of> ")}
        except Exception as e:  # noqa: BLE001
            print(f"  (traffic not derived: {e})")
    if out_json and traffic:
        with open(out_json, "w") as f:
            json.dump(traffic, f, indent=1)


if __name__ == "__main__":
    main()
