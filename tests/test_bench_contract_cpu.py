"""CPU checks of bench.py: the reference arm runs here (it is the CPU path) and prints the contract's
JSON line; our arm must refuse to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    try:
        import torch

        return not torch.cuda.is_available()
    except Exception:
        return True


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "2", "--warmup", "1", "--gib", "0.0625"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads([x for x in r.stdout.splitlines() if x.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("haystack GB/s scanned")
    assert line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["vs_baseline"] is None and line["dtype"] == "u8" and "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    if not _no_gpu():
        import pytest

        pytest.skip("only meaningful on a machine without a GPU")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_full_size_expectations_match_brute_force():
    """bench.py checks its full-size count-mode and many-haystack results against closed-form CPU
    expectations derived from one period of the i386 tiling; pin those formulas on small cases."""
    import numpy as np

    sys.path.insert(0, ROOT)
    import bench
    import oracle

    with open(os.path.join(ROOT, "data", "i386.txt"), "rb") as f:
        i386 = f.read()
    m = len(i386)
    rng = np.random.default_rng(3)
    for nd in (b"segment", b"the", b"ipsum", b"descriptor table", b"e\n"):
        P, mm = bench.periodic_matches(i386, nd)
        assert mm == m
        for gs in (0, 54321, 5 * m - 3):
            lens = rng.integers(0, 16384, 120)
            off = np.zeros(121, np.int64)
            np.cumsum(lens, out=off[1:])
            tiled = bench.tiled_host(i386, int(off[-1]), gs).tobytes()
            brute = np.array([tiled[off[h]:off[h + 1]].find(nd) >= 0 for h in range(120)], np.uint8)
            assert np.array_equal(bench.expected_set_flags(i386, nd, gs, off), brute), (nd, gs)
        n = 2 * m + 4321
        expect = int(sum((n - len(nd) - int(p)) // m + 1 for p in P if p <= n - len(nd)))
        assert expect == oracle.count(bench.tiled_host(i386, n), nd), nd
    assert oracle.count(b"aaaaa", b"aa") == 4  # overlapping occurrences count
