"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the header declares,
enforces the reference's constructor contract without touching a device, and fails loudly
(no CPU fallback) when asked to search without a GPU."""
import ctypes as C
import os
import re

import pytest

import sliceslice_rs_b200 as ss

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    try:
        import torch

        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.fixture(scope="module", autouse=True)
def built():
    from sliceslice_rs_b200 import build

    build.build()  # nvcc cross-compiles sm_100a without a GPU


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "sliceslice_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ss_b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 28
    raw = C.CDLL(ss.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in include/sliceslice_b200.h but not exported"
    assert ss.lib().ss_b200_abi_version() == 1


def test_no_torch_or_python_symbols_in_the_abi():
    import subprocess

    out = subprocess.run(["nm", "-D", "--undefined-only", ss.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out.lower() and "Py_" not in out and "c10" not in out


def test_constructor_contract_matches_reference(kats):
    # src/x86.rs:533-565 (+ :470-475): same outcomes as the oracle's reading, through the C ABI
    for c in kats["ctor"]:
        cls = ss.DynamicB200Searcher if c["searcher"] == "dynamic" else ss.B200Searcher
        n = c["needle"].encode()

        def make():
            return cls.new(n) if c["position"] is None else cls.with_position(n, c["position"])

        if c["outcome"] == "panic":
            with pytest.raises(ss.SearcherPanic):
                make()
        else:
            s = make()
            assert s.needle == n
            s.close()


def test_default_position_is_last_byte():
    s = ss.DynamicB200Searcher.new(b"ipsum")
    assert s.position == 4  # len.wrapping_sub(1), src/x86.rs:457
    s.close()
    s = ss.DynamicB200Searcher.with_position(b"ipsum", 2)
    assert s.position == 2
    s.close()


def test_empty_needle_is_found_without_a_device():
    # DynamicAvx2Searcher::N0 => true, even for an empty haystack (src/x86.rs:470,500): decided on the host
    s = ss.DynamicB200Searcher.new(b"")
    assert s.search_in(b"") is True
    assert s.find_in(b"abc") == 0
    # haystack shorter than the needle never reaches the device either (src/x86.rs:357-359)
    t = ss.DynamicB200Searcher.new(b"abcd")
    assert t.search_in(b"abc") is False
    assert ss.DynamicB200Searcher.new(b"a").search_in(b"") is False  # src/lib.rs:131-133


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a machine without a GPU")
def test_search_fails_loudly_without_gpu():
    s = ss.DynamicB200Searcher.new(b"ipsum")
    with pytest.raises(ss.B200Error):
        s.search_in(b"Lorem ipsum dolor sit amet")
    with pytest.raises(ss.B200Error):
        ss.DeviceHaystack.upload(b"Lorem ipsum dolor sit amet")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sliceslice_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
