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
    assert ss.lib().ss_b200_abi_version() == 2


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


# ---- second-anchor selection: a host-side rule, testable without a device ----------

def _rarest_model(needle: bytes, hist):
    """The documented rule of ss_b200_rarest_position, restated: cost = count * (16 if p < 16 else 17),
    minimum over p in [1, min(len - 1, 2032)], ties to the larger p."""
    if len(needle) < 2:
        return 0
    best, best_cost = 1, None
    for p in range(1, min(len(needle) - 1, 2032) + 1):
        cost = int(hist[needle[p]]) * (16 if p < 16 else 17)
        if best_cost is None or cost <= best_cost:
            best, best_cost = p, cost
    return best


def test_rarest_position_follows_the_histogram():
    import numpy as np

    rng = np.random.default_rng(7)
    assert ss.rarest_position(b"") == 0
    assert ss.rarest_position(b"x") == 0
    flat = np.ones(256, np.uint64)
    # equally frequent bytes: the reference's default (last byte, src/x86.rs:457) while it is below 16 ...
    assert ss.rarest_position(b"ipsum", flat) == 4
    assert ss.rarest_position(b"0123456789abcdef", flat) == 15
    # ... and the last index of the first 16 beyond that (positions >= 16 are charged 17/16)
    assert ss.rarest_position(bytes(range(40)), flat) == 15
    # a clearly rarer byte further out still wins
    h = np.full(256, 1000, np.uint64)
    h[ord("Z")] = 3
    assert ss.rarest_position(b"a" * 30 + b"Z" + b"a" * 10, h) == 30
    # second anchors beyond the staged halo are never picked
    far = b"a" * 3000 + b"Z"
    assert ss.rarest_position(far, h) <= 2032
    for _ in range(300):
        k = int(rng.integers(2, 80))
        needle = bytes(rng.integers(0, 256, k, dtype=np.uint8))
        hist = rng.integers(0, 1 << 40, 256).astype(np.uint64)
        assert ss.rarest_position(needle, hist) == _rarest_model(needle, hist)
    # counts near 2^64 do not wrap the cost
    big = np.full(256, (1 << 64) - 1, np.uint64)
    big[7] = 5
    assert ss.rarest_position(bytes([1, 2, 7, 3]), big) == 2


def test_rarest_position_default_table_prefers_unusual_bytes():
    # built-in background table: rare letters / punctuation beat common letters and the space
    assert ss.rarest_position(b"the quiz") == 7  # 'z'
    assert ss.rarest_position(b"e e e e#e e") == 7
    s = ss.DynamicB200Searcher.with_rarest_position(b"consecteturadipi")
    assert s.position == ss.rarest_position(b"consecteturadipi") and 1 <= s.position < 16
    s.close()
    with pytest.raises(ss.SearcherPanic):
        ss.B200Searcher.with_rarest_position(b"")  # Avx2Searcher::new(empty) panics, src/x86.rs:285,300
    assert ss.DynamicB200Searcher.with_rarest_position(b"").search_in(b"") is True


# ---- round 2's boundary additions: everything that can be pinned without a device -----------------

def test_setters_validate_their_arguments():
    L = ss.lib()
    assert L.ss_b200_set_host_path(0, 0, -1) == ss.OK
    for bad in ((4, 0, -1), (-1, 0, -1), (0, -1, -1), (0, 5000, -1), (0, 0, -2), (0, 0, 1000)):
        assert L.ss_b200_set_host_path(*bad) == ss.E_ARG, bad
    assert L.ss_b200_set_launch_pdl(2) == ss.E_ARG and L.ss_b200_set_launch_pdl(1) == ss.OK
    assert L.ss_b200_set_sync_service(2, 0) == ss.E_ARG and L.ss_b200_set_sync_service(1, -5) == ss.E_ARG
    assert L.ss_b200_set_sync_service(1, 0) == ss.OK
    assert L.ss_b200_set_scan_variant(3) == ss.E_ARG and L.ss_b200_set_scan_variant(0) == ss.OK
    assert L.ss_b200_strerror(ss.E_NCCL).decode().lower().startswith("nccl")


def test_no_environment_knobs_in_the_library():
    # the reference reads no runtime configuration; every knob here is a ss_b200_set_* call
    csrc = os.path.join(ROOT, "sliceslice_rs_b200", "csrc")
    for f in os.listdir(csrc):
        assert "getenv" not in open(os.path.join(csrc, f)).read(), f


def test_nccl_is_loaded_on_demand_not_linked():
    import subprocess

    needed = subprocess.run(["readelf", "-d", ss.LIB_PATH], capture_output=True, text=True).stdout
    assert "nccl" not in needed.lower()  # dlopen at ss_b200_ctx_set_exchange(NCCL), SS_B200_E_NCCL if absent
    v = ss.nccl_version()  # the image has libnccl.so.2: loadable without a device
    assert v >= 20000


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a machine without a GPU")
def test_context_fails_loudly_without_gpu():
    with pytest.raises(ss.B200Error):
        ss.Context()
    with pytest.raises(ss.B200Error):
        ss.measure_h2d(1 << 20, 1)
    # thread bookkeeping works without a device: nothing held, nothing to release
    assert ss.thread_footprint() == (0, 0)
    ss.thread_release()


def test_thread_local_lanes_are_released_without_a_device():
    import threading

    out = []

    def work():
        out.append(ss.thread_footprint())
        ss.thread_release()

    t = threading.Thread(target=work)
    t.start()
    t.join()
    assert out == [(0, 0)]


def test_rust_shim_declares_every_header_entry():
    """rust/sliceslice-b200 cannot be compiled here (no rustc); what CAN be pinned is that its `sys` module
    declares every entry point of the header, with the same number of parameters."""
    hdr = open(os.path.join(ROOT, "include", "sliceslice_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {name: ([] if args.strip() == "void" else args.split(","))
              for name, args in re.findall(r"\b(ss_b200_\w+)\s*\(([^;{]*?)\)\s*;", hdr)}
    assert len(protos) >= 70
    rs = open(os.path.join(ROOT, "rust", "sliceslice-b200", "src", "lib.rs")).read()
    sys_block = rs[rs.index("pub mod sys {"):]
    sys_block = sys_block[:sys_block.index("\n}\n") + 3]
    decls = {name: ([] if not args.strip() else args.split(","))
             for name, args in re.findall(r"pub fn (ss_b200_\w+)\(([^)]*)\)", sys_block)}
    assert set(decls) == set(protos), sorted(set(protos) ^ set(decls))
    for name, params in protos.items():
        assert len(decls[name]) == len(params), name
