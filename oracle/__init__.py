"""ctypes front for the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Loads ``oracle/liboracle.so`` (C restatement of ``DynamicAvx2Searcher``, see
``sliceslice_oracle.c``) and, when present, ``oracle/_ref/libsse4strstr_ref.so``
(the reference's vendored ``avx2_strstr_v2`` compiled from /root/reference by
``oracle/Makefile``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module.
Nothing under ``sliceslice_rs_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libsse4strstr_ref.so")

NPOS = (1 << 64) - 1
OK, E_POSITION, E_EMPTY_NEEDLE = 0, 1, 2

_u8p = C.POINTER(C.c_uint8)
_lib = None
_ref = None


def build(force: bool = False) -> None:
    """Compile the C restatement (and oracle/_ref when /root/reference exists)."""
    src = os.path.join(HERE, "sliceslice_oracle.c")
    stale = (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)
    need_ref = os.path.isdir("/root/reference") and not os.path.exists(REF_PATH)
    if force or stale or need_ref:
        subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        sz = C.c_size_t
        L.ss_oracle_naive_find.restype = sz
        L.ss_oracle_naive_find.argtypes = [C.c_void_p, sz, C.c_void_p, sz]
        L.ss_oracle_check_ctor.restype = C.c_int
        L.ss_oracle_check_ctor.argtypes = [sz, sz, C.c_int]
        for name in ("ss_oracle_dynamic_avx2_find", "ss_oracle_avx2_find"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, sz, C.c_void_p, sz, sz, C.POINTER(sz)]
        L.ss_oracle_dynamic_avx2_find_mt.restype = C.c_int
        L.ss_oracle_dynamic_avx2_find_mt.argtypes = [C.c_void_p, sz, C.c_void_p, sz, sz, C.c_int, C.POINTER(sz)]
        L.ss_oracle_count_candidates.restype = sz
        L.ss_oracle_count_candidates.argtypes = [C.c_void_p, sz, C.c_void_p, sz, sz]
        L.ss_oracle_long_sweep.restype = None
        L.ss_oracle_long_sweep.argtypes = [C.c_void_p, C.c_void_p, sz, C.c_void_p, sz, C.c_int, C.c_int, C.c_void_p]
        L.ss_oracle_short_sweep.restype = C.c_uint64
        L.ss_oracle_short_sweep.argtypes = [C.c_void_p, C.c_void_p, sz, C.c_int, C.c_void_p]
        L.ss_oracle_pairs.restype = None
        L.ss_oracle_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, sz,
                                      C.c_int, C.c_void_p]
        L.ss_oracle_fill_random.restype = None
        L.ss_oracle_fill_random.argtypes = [C.c_void_p, C.c_uint64, sz, C.c_uint64]
        L.ss_oracle_fill_tiled.restype = None
        L.ss_oracle_fill_tiled.argtypes = [C.c_void_p, C.c_uint64, sz, C.c_void_p, sz]
        _lib = L
    return _lib


def ref_lib():
    """The reference's vendored avx2_strstr_v2 (oracle/_ref) or None."""
    global _ref
    if _ref is None and os.path.exists(REF_PATH):
        R = C.CDLL(REF_PATH)
        R.avx2_strstr_v2.restype = C.c_size_t
        R.avx2_strstr_v2.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        _ref = R
    return _ref


def _buf(b):
    """bytes / bytearray / numpy uint8 -> (address, length, keepalive)."""
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b, dtype=np.uint8)
        return a.ctypes.data, a.size, a
    a = np.frombuffer(bytes(b), dtype=np.uint8) if len(b) else np.zeros(1, np.uint8)
    return a.ctypes.data, len(b), a


class OracleError(AssertionError):
    """The reference would have panicked at construction."""

    def __init__(self, code):
        super().__init__({E_POSITION: "position out of range", E_EMPTY_NEEDLE: "empty needle"}.get(code, str(code)))
        self.code = code


def naive_find(hay, needle):
    hp, n, _h = _buf(hay)
    np_, k, _n = _buf(needle)
    r = lib().ss_oracle_naive_find(hp, n, np_, k)
    return None if r == NPOS else r


def find(hay, needle, position=None, dynamic=True, threads=1):
    """DynamicAvx2Searcher::with_position(needle, position).search_in(hay) -> first offset | None."""
    hp, n, _h = _buf(hay)
    np_, k, _n = _buf(needle)
    if position is None:
        position = (k - 1) & ((1 << 64) - 1)
    out = C.c_size_t(0)
    if threads > 1:
        rc = lib().ss_oracle_dynamic_avx2_find_mt(hp, n, np_, k, position, threads, C.byref(out))
    else:
        fn = lib().ss_oracle_dynamic_avx2_find if dynamic else lib().ss_oracle_avx2_find
        rc = fn(hp, n, np_, k, position, C.byref(out))
    if rc != OK:
        raise OracleError(rc)
    return None if out.value == NPOS else out.value


def search_in(hay, needle, position=None, dynamic=True):
    return find(hay, needle, position, dynamic) is not None


def count(hay, needle):
    """Number of occurrences (overlapping ones included): the reference's search restarted one byte
    after every match -- what `search_in` would report if its loop did not return at the first
    verified candidate (src/lib.rs:242-244).  Checker for the count mode."""
    import numpy as np

    a = np.frombuffer(bytes(hay), np.uint8) if not isinstance(hay, np.ndarray) else hay
    if len(needle) == 0:
        raise ValueError("the empty needle has no occurrence count")
    total, start = 0, 0
    while True:
        r = find(a[start:], needle)
        if r is None:
            return total
        total += 1
        start += r + 1


def count_candidates(hay, needle, position=None):
    hp, n, _h = _buf(hay)
    np_, k, _n = _buf(needle)
    return lib().ss_oracle_count_candidates(hp, n, np_, k, k - 1 if position is None else position)


def ref_find(hay, needle):
    """avx2_strstr_v2 from the reference tree (k >= 2 only; pads for its over-read)."""
    R = ref_lib()
    if R is None:
        return NotImplemented
    k = len(needle)
    assert k >= 2, "vendored k==1 arm uses strchr on NUL-terminated input"
    padded = np.zeros(len(hay) + 64 + k, np.uint8)
    padded[: len(hay)] = np.frombuffer(bytes(hay), np.uint8) if not isinstance(hay, np.ndarray) else hay
    np_, _, _n = _buf(needle)
    r = R.avx2_strstr_v2(padded.ctypes.data, len(hay), np_, k)
    return None if r == NPOS else r


def csr(items):
    """list[bytes] -> (blob uint8[], offsets uint64[n+1])."""
    off = np.zeros(len(items) + 1, np.uint64)
    off[1:] = np.cumsum([len(x) for x in items], dtype=np.uint64)
    blob = np.frombuffer(b"".join(items), np.uint8) if int(off[-1]) else np.zeros(1, np.uint8)
    return np.ascontiguousarray(blob), off


def long_sweep(needles, hay, naive=False, threads=1):
    blob, off = csr(needles)
    hp, n, _h = _buf(hay)
    out = np.empty(len(needles), np.uint64)
    lib().ss_oracle_long_sweep(blob.ctypes.data, off.ctypes.data, len(needles), hp, n, int(naive), threads,
                               out.ctypes.data)
    return out


def short_sweep(words, naive=False, want_bitmap=True):
    blob, off = csr(words)
    w = len(words)
    npairs = w * (w + 1) // 2
    bm = np.zeros((npairs + 31) // 32, np.uint32) if want_bitmap else None
    m = lib().ss_oracle_short_sweep(blob.ctypes.data, off.ctypes.data, w, int(naive),
                                    bm.ctypes.data if want_bitmap else None)
    return int(m), bm


def pairs(needles, hays, pair_needle, pair_hay, naive=False):
    nb, no = csr(needles)
    hb, ho = csr(hays)
    pn = np.ascontiguousarray(pair_needle, np.uint32)
    ph = np.ascontiguousarray(pair_hay, np.uint32)
    out = np.empty(pn.size, np.uint64)
    lib().ss_oracle_pairs(nb.ctypes.data, no.ctypes.data, hb.ctypes.data, ho.ctypes.data, pn.ctypes.data,
                          ph.ctypes.data, pn.size, int(naive), out.ctypes.data)
    return out


def fill_random(global_start, length, seed):
    out = np.empty(length, np.uint8)
    lib().ss_oracle_fill_random(out.ctypes.data, global_start, length, seed)
    return out


def fill_tiled(global_start, length, src):
    s = np.ascontiguousarray(np.frombuffer(bytes(src), np.uint8) if not isinstance(src, np.ndarray) else src)
    out = np.empty(length, np.uint8)
    lib().ss_oracle_fill_tiled(out.ctypes.data, global_start, length, s.ctypes.data, s.size)
    return out
