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
