"""Builds and runs tests/cpp/test_host_mirror.cpp: the reference's own test shapes driven through the
C++ host mirror (include/sliceslice_b200.hpp) of DynamicAvx2Searcher / Avx2Searcher over the C ABI.
The KAT tables are generated from tests/golden/kats.json (transcribed from src/lib.rs:299-331, :422-544)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")


def _c(s: str) -> str:
    return '"' + "".join(f"\\x{b:02x}" for b in s.encode()) + '"'


@pytest.fixture(scope="module")
def binary(kats):
    from sliceslice_rs_b200 import build

    lib = build.build()
    os.makedirs(BUILD, exist_ok=True)
    with open(os.path.join(BUILD, "kats_table.inc"), "w") as f:
        for k in kats["kats"]:
            off = -1 if k["offset"] is None else k["offset"]
            f.write(f'{{{_c(k["group"])}, {_c(k["haystack"])}, {_c(k["needle"])}, {str(k["found"]).lower()}, {off}}},\n')
    with open(os.path.join(BUILD, "memchr_table.inc"), "w") as f:
        for k in kats["memchr"]:
            f.write(f'{{{_c(k["haystack"])}, {_c(k["needle"])}, {str(k["found"]).lower()}}},\n')
    exe = os.path.join(BUILD, "test_host_mirror")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", BUILD,
           os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"), "-o", exe, lib,
           f"-Wl,-rpath,{os.path.dirname(lib)}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


@pytest.fixture(scope="module")
def grep_binary():
    from sliceslice_rs_b200 import build

    lib = build.build()
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "grep")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "grep.cpp"), "-o", exe, lib, f"-Wl,-rpath,{os.path.dirname(lib)}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_grep_example_builds_and_rejects_bad_arguments(grep_binary):
    r = subprocess.run([grep_binary], capture_output=True, text=True)
    assert r.returncode == 101 and "<backend> <needle> <file>" in r.stderr
    r = subprocess.run([grep_binary, "avx512", "x", os.path.join(ROOT, "data", "needle")], capture_output=True, text=True)
    assert r.returncode == 101 and "Invalid backend" in r.stderr
    # the empty needle panics for the strict searcher before any device work (src/x86.rs:285,300)
    r = subprocess.run([grep_binary, "b200", "", os.path.join(ROOT, "data", "needle")], capture_output=True, text=True)
    assert r.returncode == 101 and "panicked" in r.stderr


@pytest.mark.gpu
def test_grep_example_on_i386(grep_binary):
    # examples/grep.rs:42-57 output format
    f = os.path.join(ROOT, "data", "i386.txt")
    for backend in ("b200", "DynamicB200"):
        r = subprocess.run([grep_binary, backend, "segmentation", f], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.strip() == f'Searching for segmentation in "{f}": true', r.stdout + r.stderr
        r = subprocess.run([grep_binary, backend, "ipsum", f], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.strip().endswith(": false")


def test_cpp_mirror_constructor_contract(binary):
    r = subprocess.run([binary, "ctor"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok:" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_reference_test_shapes(binary):
    r = subprocess.run([binary, "all", os.path.join(ROOT, "data", "i386.txt"), os.path.join(ROOT, "data", "words.txt")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok:" in r.stdout


@pytest.fixture(scope="module")
def ctx_binary():
    from sliceslice_rs_b200 import build

    lib = build.build()
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "test_ctx")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
           os.path.join(ROOT, "tests", "cpp", "test_ctx.cpp"), "-o", exe, lib, "-L/usr/local/cuda/lib64", "-lcudart",
           f"-Wl,-rpath,{os.path.dirname(lib)}", "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_cpp_ctx_test_builds(ctx_binary):
    assert os.path.exists(ctx_binary)


@pytest.mark.gpu
def test_cpp_multi_gpu_context(ctx_binary):
    """tests/cpp/test_ctx.cpp: the multi-GPU context from a compiled host, on every GPU of the box --
    sharded haystack through the three exchanges, one host slice striped over all devices, many-haystack
    mode; expectations from a naive search."""
    r = subprocess.run([ctx_binary, os.path.join(ROOT, "data", "i386.txt")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok:" in r.stdout


def test_swar_identities_exhaustive():
    """All 2^32 words: the any-zero-byte test never misses or invents a candidate, and the exact mask
    marks exactly the zero bytes.  The formulas are checked to be the ones the kernels compile."""
    src = open(os.path.join(ROOT, "sliceslice_rs_b200", "csrc", "ss_filter.cuh")).read()
    assert "return (x - 0x01010101u) & ~x;" in src
    assert "return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);" in src
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "test_swar")
    subprocess.run(["gcc", "-O3", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_swar.c")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout


def test_kernel_arithmetic_emulated_on_cpu():
    """tests/cpp/test_filter_host.cu: the filter / refinement code the kernels compile (ss_filter.cuh,
    __host__ __device__) driven chunk by chunk on the CPU against a naive search -- no false negatives with
    any extra-anchor kind, verified positions == naive positions."""
    from sliceslice_rs_b200 import build

    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "test_filter_host")
    cmd = [build.nvcc(), "-std=c++17", "-O2", "--extended-lambda", "--expt-relaxed-constexpr", "-o", exe,
           os.path.join(ROOT, "tests", "cpp", "test_filter_host.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe, "300000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout
