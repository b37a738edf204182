"""sliceslice_rs_b200 -- B200-native single-pattern substring search behind the
``DynamicAvx2Searcher`` surface of cloudflare/sliceslice-rs.

This package is the Python host side ABOVE the C ABI (``include/sliceslice_b200.h``,
implemented by ``libsliceslice_b200.so``: hand-written sm_100a CUDA, no torch types).
It mirrors the reference's searcher interface for the one accelerated path::

    reference (src/x86.rs)                              here
    DynamicAvx2Searcher::new(needle)            ->  DynamicB200Searcher.new(needle)
    DynamicAvx2Searcher::with_position(n, p)    ->  DynamicB200Searcher.with_position(n, p)
    searcher.search_in(haystack) -> bool        ->  searcher.search_in(haystack) -> bool
    Avx2Searcher::{new, with_position}          ->  B200Searcher.{new, with_position}
    (panic at construction)                     ->  raises SearcherPanic

plus ``find_in`` (the index at which the reference's scan returns true = leftmost
occurrence).  PyTorch is used only as plumbing: device memory, streams, and
``torch.distributed`` for the multi-GPU min-reduction (``sharded.py``).

There is NO CPU fallback: if the shared library is missing or no CUDA device is usable,
searches raise; nothing in this package imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

__all__ = ["DynamicB200Searcher", "B200Searcher", "DeviceHaystack", "SearcherPanic", "B200Error", "lib",
           "NPOS", "DEVICE_NONE", "fill_random", "fill_tiled", "set_scan_variant", "set_scan_tuning",
           "launch_count", "Batch", "set_extra_anchors", "HaystackSet", "rarest_position", "Context",
           "ShardedHaystack", "ContextHaystackSet", "set_host_path", "set_launch_pdl", "measure_h2d",
           "thread_release", "thread_footprint", "set_sync_service", "E_NCCL", "EXCHANGE_HOST", "EXCHANGE_PEER", "EXCHANGE_NCCL"]

PKG = os.path.dirname(os.path.abspath(__file__))
# SS_B200_LIB: load another build of the same ABI instead (A/B measurements of kernel changes)
LIB_PATH = os.environ.get("SS_B200_LIB") or os.path.join(PKG, "libsliceslice_b200.so")
NPOS = (1 << 64) - 1
DEVICE_NONE = 0x7FFFFFFFFFFFFFFF
OK, E_POSITION, E_EMPTY_NEEDLE, E_ARG, E_CUDA, E_NOMEM, E_NCCL = range(7)
EXCHANGE_HOST, EXCHANGE_PEER, EXCHANGE_NCCL = range(3)

_lib = None


class SearcherPanic(AssertionError):
    """The reference panics here (src/x86.rs:300 ``assert!(position < needle.size())``,
    :473 ``assert_eq!(position, 0)``, :285 empty needle for ``Avx2Searcher``)."""


class B200Error(RuntimeError):
    """CUDA / argument failure reported by the C ABI (never raised by the reference)."""


def lib() -> C.CDLL:
    """Load libsliceslice_b200.so (built in-tree by ``python -m sliceslice_rs_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(f"{LIB_PATH} is missing: build it with `python -m sliceslice_rs_b200.build` "
                        "(the CUDA path is the only path; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    sz, u64, vp, i32 = C.c_size_t, C.c_uint64, C.c_void_p, C.c_int
    pp = C.POINTER(vp)
    sig = {
        "ss_b200_strerror": (C.c_char_p, [i32]),
        "ss_b200_last_error": (C.c_char_p, []),
        "ss_b200_abi_version": (i32, []),
        "ss_b200_searcher_new": (i32, [vp, sz, pp]),
        "ss_b200_searcher_with_position": (i32, [vp, sz, sz, pp]),
        "ss_b200_searcher_new_strict": (i32, [vp, sz, pp]),
        "ss_b200_searcher_with_position_strict": (i32, [vp, sz, sz, pp]),
        "ss_b200_rarest_position": (i32, [vp, sz, vp, C.POINTER(sz)]),
        "ss_b200_searcher_new_rarest": (i32, [vp, sz, vp, pp]),
        "ss_b200_byte_histogram_device_async": (i32, [vp, sz, sz, vp, vp]),
        "ss_b200_haystack_byte_histogram": (i32, [vp, sz, vp]),
        "ss_b200_searcher_free": (None, [vp]),
        "ss_b200_searcher_needle_len": (sz, [vp]),
        "ss_b200_searcher_position": (sz, [vp]),
        "ss_b200_haystack_upload": (i32, [vp, sz, pp]),
        "ss_b200_haystack_from_device": (i32, [vp, sz, pp]),
        "ss_b200_haystack_free": (None, [vp]),
        "ss_b200_haystack_len": (sz, [vp]),
        "ss_b200_haystack_device_ptr": (vp, [vp]),
        "ss_b200_search_in": (i32, [vp, vp, C.POINTER(C.c_uint8)]),
        "ss_b200_find_in": (i32, [vp, vp, C.POINTER(sz)]),
        "ss_b200_search_in_host": (i32, [vp, vp, sz, C.POINTER(C.c_uint8)]),
        "ss_b200_find_in_host": (i32, [vp, vp, sz, C.POINTER(sz)]),
        "ss_b200_find_in_device_async": (i32, [vp, vp, sz, u64, sz, vp, vp, vp]),
        "ss_b200_count_in_device_async": (i32, [vp, vp, sz, sz, vp, vp, vp]),
        "ss_b200_mailbox_create": (i32, [i32, pp]),
        "ss_b200_mailbox_free": (i32, [vp]),
        "ss_b200_ipc_export": (i32, [vp, vp]),
        "ss_b200_ipc_open": (i32, [vp, pp]),
        "ss_b200_ipc_close": (i32, [vp]),
        "ss_b200_find_in_device_exchange_async": (i32, [vp, vp, sz, u64, sz, vp, vp, i32, i32, u64, vp, vp]),
        "ss_b200_search_many_async": (i32, [vp, vp, vp, sz, sz, vp, vp, vp]),
        "ss_b200_hayset_create": (i32, [vp, vp, sz, sz, vp, pp]),
        "ss_b200_hayset_free": (None, [vp]),
        "ss_b200_hayset_len": (sz, [vp]),
        "ss_b200_hayset_search_async": (i32, [vp, vp, vp, vp, vp]),
        "ss_b200_batch_create": (i32, [vp, vp, sz, vp, vp, sz, pp]),
        "ss_b200_batch_free": (None, [vp]),
        "ss_b200_batch_search_pairs": (i32, [vp, vp, vp, sz, vp, vp]),
        "ss_b200_batch_search_triangular": (i32, [vp, vp, C.POINTER(u64)]),
        "ss_b200_batch_find_all_in": (i32, [vp, vp, vp]),
        "ss_b200_fill_random": (i32, [vp, sz, u64, u64, vp]),
        "ss_b200_fill_tiled": (i32, [vp, sz, u64, vp, sz, vp]),
        "ss_b200_set_scan_variant": (i32, [i32]),
        "ss_b200_set_scan_tuning": (i32, [i32, i32, i32, i32]),
        "ss_b200_set_extra_anchors": (i32, [i32]),
        "ss_b200_launch_count": (u64, []),
        "ss_b200_set_launch_pdl": (i32, [i32]),
        "ss_b200_set_sync_service": (i32, [i32, i32]),
        "ss_b200_set_host_path": (i32, [i32, i32, i32]),
        "ss_b200_measure_h2d": (i32, [sz, i32, C.POINTER(C.c_double)]),
        "ss_b200_thread_release": (i32, []),
        "ss_b200_thread_footprint": (i32, [C.POINTER(sz), C.POINTER(sz)]),
        "ss_b200_batch_search_pairs_async": (i32, [vp, vp, vp, sz, vp, vp, vp]),
        "ss_b200_batch_search_triangular_async": (i32, [vp, vp, vp, vp]),
        "ss_b200_batch_find_all_in_device_async": (i32, [vp, vp, sz, vp, vp]),
        "ss_b200_pack_flags_async": (i32, [vp, sz, sz, vp, sz, vp]),
        "ss_b200_ctx_create": (i32, [i32, vp, pp]),
        "ss_b200_ctx_free": (None, [vp]),
        "ss_b200_ctx_device_count": (i32, [vp]),
        "ss_b200_ctx_device": (i32, [vp, i32]),
        "ss_b200_ctx_set_exchange": (i32, [vp, i32]),
        "ss_b200_ctx_nccl_version": (i32, [C.POINTER(i32)]),
        "ss_b200_sharded_upload": (i32, [vp, vp, sz, sz, pp]),
        "ss_b200_sharded_from_device": (i32, [vp, vp, vp, vp, pp]),
        "ss_b200_sharded_free": (None, [vp]),
        "ss_b200_sharded_len": (sz, [vp]),
        "ss_b200_sharded_shard": (i32, [vp, i32, pp, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]),
        "ss_b200_search_sharded": (i32, [vp, vp, vp, C.POINTER(C.c_uint8), C.POINTER(sz)]),
        "ss_b200_find_sharded": (i32, [vp, vp, vp, C.POINTER(sz)]),
        "ss_b200_find_in_host_multi": (i32, [vp, vp, vp, sz, C.POINTER(sz)]),
        "ss_b200_search_in_host_multi": (i32, [vp, vp, vp, sz, C.POINTER(C.c_uint8)]),
        "ss_b200_ctx_last_host_stats": (i32, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(i32)]),
        "ss_b200_ctx_hayset_upload": (i32, [vp, vp, vp, sz, pp]),
        "ss_b200_ctx_hayset_free": (None, [vp]),
        "ss_b200_ctx_hayset_len": (sz, [vp]),
        "ss_b200_ctx_hayset_part": (i32, [vp, i32, C.POINTER(sz), C.POINTER(sz)]),
        "ss_b200_ctx_hayset_search": (i32, [vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)  # AttributeError here == header/library mismatch: fail loudly
        f.restype = res
        f.argtypes = args
    _lib = L
    return L



def _check(rc: int) -> None:
    if rc == OK:
        return
    if rc in (E_POSITION, E_EMPTY_NEEDLE):
        raise SearcherPanic(lib().ss_b200_strerror(rc).decode())
    detail = lib().ss_b200_last_error().decode() if rc in (E_CUDA, E_NOMEM, E_NCCL, E_ARG) else ""
    raise B200Error(f"{lib().ss_b200_strerror(rc).decode()} {detail}".strip())


def _host_view(b):
    """bytes-like / numpy -> (address, length, keepalive) without copying when possible."""
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b).view(np.uint8).reshape(-1)
        return a.ctypes.data, a.size, a
    mv = memoryview(b)
    if mv.nbytes == 0:
        z = np.zeros(1, np.uint8)
        return z.ctypes.data, 0, z
    a = np.frombuffer(mv, dtype=np.uint8)
    return a.ctypes.data, a.size, (a, b)


def _is_torch_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


class DeviceHaystack:
    """A haystack resident in HBM (``ss_b200_haystack``): uploaded (owned) or borrowed from a
    CUDA uint8 tensor / raw device pointer."""

    def __init__(self, handle, keepalive=None):
        self._h = handle
        self._keep = keepalive

    @classmethod
    def upload(cls, data) -> "DeviceHaystack":
        addr, n, keep = _host_view(data)
        h = C.c_void_p()
        _check(lib().ss_b200_haystack_upload(addr, n, C.byref(h)))
        return cls(h)

    @classmethod
    def from_tensor(cls, t) -> "DeviceHaystack":
        if not t.is_cuda or t.dtype.itemsize != 1 or not t.is_contiguous():
            raise B200Error("from_tensor needs a contiguous 1-byte CUDA tensor")
        import torch

        torch.cuda.current_stream(t.device).synchronize()  # searches run on the library's own stream
        h = C.c_void_p()
        _check(lib().ss_b200_haystack_from_device(t.data_ptr(), t.numel(), C.byref(h)))
        return cls(h, keepalive=t)

    @classmethod
    def from_pointer(cls, dptr: int, length: int, keepalive=None) -> "DeviceHaystack":
        h = C.c_void_p()
        _check(lib().ss_b200_haystack_from_device(dptr, length, C.byref(h)))
        return cls(h, keepalive)

    def __len__(self) -> int:
        return lib().ss_b200_haystack_len(self._h)

    def byte_histogram(self, sample_bytes: int = 0) -> np.ndarray:
        """256 byte counts of the haystack (``sample_bytes`` = 0: every byte; else an evenly spaced
        sample of about that many bytes) -- the input of ``with_rarest_position``."""
        hist = np.zeros(256, np.uint64)
        _check(lib().ss_b200_haystack_byte_histogram(self._h, sample_bytes, hist.ctypes.data))
        return hist

    @property
    def device_ptr(self) -> int:
        return lib().ss_b200_haystack_device_ptr(self._h) or 0

    def close(self) -> None:
        if self._h:
            lib().ss_b200_haystack_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _SearcherBase:
    _STRICT = False

    def __init__(self, handle, needle: bytes):
        self._s = handle
        self._needle = needle

    # -- construction: same names, argument meaning and failure behaviour as the reference ----
    @classmethod
    def new(cls, needle):
        """``::new(needle)``: second anchor = last byte (src/x86.rs:454-459 / :282-287)."""
        nb = bytes(needle)
        addr, n, keep = _host_view(nb)
        h = C.c_void_p()
        fn = lib().ss_b200_searcher_new_strict if cls._STRICT else lib().ss_b200_searcher_new
        _check(fn(addr, n, C.byref(h)))
        return cls(h, nb)

    @classmethod
    def with_position(cls, needle, position: int):
        """``::with_position(needle, position)`` (src/x86.rs:468-493 / :297-316)."""
        nb = bytes(needle)
        if position < 0:
            raise SearcherPanic("position must be an unsigned index")
        addr, n, keep = _host_view(nb)
        h = C.c_void_p()
        fn = lib().ss_b200_searcher_with_position_strict if cls._STRICT else lib().ss_b200_searcher_with_position
        _check(fn(addr, n, position, C.byref(h)))
        return cls(h, nb)

    @classmethod
    def with_rarest_position(cls, needle, hist=None):
        """``with_position(needle, p)`` with p chosen by ``rarest_position``: the index
        whose byte is rarest under ``hist`` (256 counts, e.g. ``DeviceHaystack.byte_histogram()``;
        None = built-in background table).  Same results as any other position (src/lib.rs:375-378)."""
        return cls.with_position(needle, rarest_position(needle, hist))

    # -- accessors (the reference's private Searcher trait, src/lib.rs:289-293) ---------------
    @property
    def needle(self) -> bytes:
        return self._needle

    @property
    def position(self) -> int:
        return lib().ss_b200_searcher_position(self._s)

    # -- the hot call ------------------------------------------------------------------------
    def find_in(self, haystack) -> Optional[int]:
        """Index at which the reference's scan returns true (leftmost occurrence) or None."""
        out = C.c_size_t(0)
        if isinstance(haystack, DeviceHaystack):
            _check(lib().ss_b200_find_in(self._s, haystack._h, C.byref(out)))
        elif _is_torch_tensor(haystack) and haystack.is_cuda:
            # the synchronous C entry scans on the library's own stream, so from_tensor first waits for
            # whatever is still writing the tensor on torch's current stream (find_in_async stays
            # stream-ordered instead)
            hs = DeviceHaystack.from_tensor(haystack)  # waits for torch's current stream
            _check(lib().ss_b200_find_in(self._s, hs._h, C.byref(out)))
            hs.close()
        else:
            if _is_torch_tensor(haystack):
                if not haystack.is_contiguous() or haystack.dtype.itemsize != 1:
                    raise B200Error("host tensor must be contiguous and 1 byte per element")
                addr, n, keep = haystack.data_ptr(), haystack.numel(), haystack
            else:
                addr, n, keep = _host_view(haystack)
            _check(lib().ss_b200_find_in_host(self._s, addr, n, C.byref(out)))
        return None if out.value == NPOS else out.value

    def search_in(self, haystack) -> bool:
        """``search_in(&self, haystack: &[u8]) -> bool`` (src/x86.rs:523-525)."""
        return self.find_in(haystack) is not None

    inlined_search_in = search_in  # src/x86.rs:496-519; inlining is a Rust codegen concern only

    def find_in_async(self, hay, result, workspace, base_offset: int = 0, start_limit: Optional[int] = None,
                      stream=None) -> None:
        """Stream-ordered scan of a CUDA uint8 tensor (``ss_b200_find_in_device_async``).

        ``result``: 1-element int64/uint64 CUDA tensor receiving base_offset + first offset or
        DEVICE_NONE; ``workspace``: >= 16 zero bytes of CUDA memory (kept zero by the kernel)."""
        import torch

        if stream is None:
            stream = torch.cuda.current_stream(hay.device)
        sp = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
        lim = NPOS if start_limit is None else int(start_limit)
        _check(lib().ss_b200_find_in_device_async(self._s, hay.data_ptr(), hay.numel(), int(base_offset), lim,
                                                  workspace.data_ptr(), result.data_ptr(), sp))

    def count_in_async(self, hay, count, workspace, start_limit: Optional[int] = None, stream=None) -> None:
        """Stream-ordered count of all occurrences (overlapping included) in a CUDA uint8 tensor;
        ``count``: 1-element int64 CUDA tensor, ``workspace``: >= 32 zero bytes of CUDA memory."""
        import torch

        if stream is None:
            stream = torch.cuda.current_stream(hay.device)
        sp = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
        lim = NPOS if start_limit is None else int(start_limit)
        _check(lib().ss_b200_count_in_device_async(self._s, hay.data_ptr(), hay.numel(), lim, workspace.data_ptr(),
                                                   count.data_ptr(), sp))

    def search_many_async(self, hayset: "HaystackSet", flags=None, stream=None):
        """One pass over a device-resident set of haystacks (``ss_b200_search_many_async``):
        ``flags[h] = search_in(haystack h)`` as uint8.  Returns the flags tensor (stream-ordered)."""
        import torch

        if flags is None:
            flags = torch.empty(len(hayset), dtype=torch.uint8, device=hayset.blob.device)
        if stream is None:
            stream = torch.cuda.current_stream(hayset.blob.device)
        sp = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
        ws = hayset.workspace_for(sp)
        if hayset.prepared:
            _check(lib().ss_b200_hayset_search_async(self._s, hayset.handle(stream), flags.data_ptr(),
                                                     ws.data_ptr(), sp))
        else:
            _check(lib().ss_b200_search_many_async(self._s, hayset.blob.data_ptr(), hayset.offsets.data_ptr(),
                                                   len(hayset), hayset.blob_len, flags.data_ptr(),
                                                   ws.data_ptr(), sp))
        return flags

    def close(self) -> None:
        if self._s:
            lib().ss_b200_searcher_free(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DynamicB200Searcher(_SearcherBase):
    """Drop-in for ``sliceslice::x86::DynamicAvx2Searcher`` (src/x86.rs:405-526)."""
    _STRICT = False


class B200Searcher(_SearcherBase):
    """Drop-in for ``sliceslice::x86::Avx2Searcher`` (src/x86.rs:266-383): empty needle panics."""
    _STRICT = True


class HaystackSet:
    """A set of haystacks resident in HBM as one blob + uint64 offsets (many-haystack mode).

    ``prepared`` (default): searches go through ``ss_b200_hayset`` -- lookup hints built once on the
    first search, so a match finds its haystack in one or two probes; ``prepared=False`` uses the
    hint-free ``ss_b200_search_many_async``.  Same flags either way."""

    def __init__(self, haystacks, device="cuda", prepared: bool = True):
        import torch

        blob, off = _csr(haystacks)
        self.n = len(haystacks)
        self.blob_len = int(off[-1])
        self.blob = torch.from_numpy(blob.copy()).to(device)
        self.offsets = torch.from_numpy(off.astype(np.int64)).to(device)
        self._ws_by_stream = {}
        self.prepared = prepared
        self._hs = None

    def workspace_for(self, stream_ptr: int):
        """The self-resetting scan workspace is per stream: searches of one set issued on different
        streams must not share its key / ticket words."""
        ws = self._ws_by_stream.get(stream_ptr)
        if ws is None:
            import torch

            with torch.cuda.device(self.blob.device):
                ws = torch.zeros(32, dtype=torch.uint8, device=self.blob.device)
                torch.cuda.current_stream().synchronize()  # zeroed before any other stream uses it
            self._ws_by_stream[stream_ptr] = ws
        return ws

    def handle(self, stream=None):
        """The ``ss_b200_hayset`` of this set, created (hints built in stream order) on first use."""
        if self._hs is None:
            import torch

            if stream is None:
                stream = torch.cuda.current_stream(self.blob.device)
            sp = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
            h = C.c_void_p()
            _check(lib().ss_b200_hayset_create(self.blob.data_ptr(), self.offsets.data_ptr(), self.n, self.blob_len,
                                               sp, C.byref(h)))
            # later searches may come in on other streams: make the hints visible to all of them
            if hasattr(stream, "synchronize"):
                stream.synchronize()
            else:
                torch.cuda.synchronize(self.blob.device)
            self._hs = h
        return self._hs

    def close(self) -> None:
        if self._hs:
            lib().ss_b200_hayset_free(self._hs)
            self._hs = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_device(cls, blob, offsets, prepared: bool = True) -> "HaystackSet":
        """Wrap a set that already lives in HBM: ``blob`` = concatenated haystack bytes (CUDA uint8
        tensor), ``offsets`` = n+1 ascending int64 byte offsets on the same device, offsets[0] == 0
        and offsets[n] == len(blob)."""
        import torch

        if not (blob.is_cuda and offsets.is_cuda and blob.dtype == torch.uint8 and offsets.dtype == torch.int64
                and blob.is_contiguous() and offsets.is_contiguous() and offsets.numel() >= 1):
            raise B200Error("from_device needs a contiguous CUDA uint8 blob and CUDA int64 offsets")
        self = cls.__new__(cls)
        self.n = offsets.numel() - 1
        self.blob_len = blob.numel()
        self.blob = blob
        self.offsets = offsets
        self._ws_by_stream = {}
        self.prepared = prepared
        self._hs = None
        return self

    def __len__(self) -> int:
        return self.n


class Batch:
    """Device-resident needle and haystack sets for the batched modes (``ss_b200_batch``)."""

    def __init__(self, needles, haystacks):
        self.n_needles, self.n_haystacks = len(needles), len(haystacks)
        nb, no = _csr(needles)
        hb, ho = _csr(haystacks)
        h = C.c_void_p()
        _check(lib().ss_b200_batch_create(nb.ctypes.data, no.ctypes.data, len(needles), hb.ctypes.data,
                                          ho.ctypes.data, len(haystacks), C.byref(h)))
        self._b = h

    def search_pairs(self, pair_needle, pair_hay, want_offsets: bool = True):
        pn = np.ascontiguousarray(pair_needle, np.uint32)
        ph = np.ascontiguousarray(pair_hay, np.uint32)
        bm = np.zeros((pn.size + 31) // 32, np.uint32)
        off = np.empty(pn.size, np.uint64) if want_offsets else None
        _check(lib().ss_b200_batch_search_pairs(self._b, pn.ctypes.data, ph.ctypes.data, pn.size, bm.ctypes.data,
                                                off.ctypes.data if want_offsets else None))
        return bm, off

    def search_triangular(self):
        w = self.n_needles
        npairs = w * (w + 1) // 2
        bm = np.zeros((npairs + 31) // 32, np.uint32)
        m = C.c_uint64(0)
        _check(lib().ss_b200_batch_search_triangular(self._b, bm.ctypes.data, C.byref(m)))
        return bm, m.value

    def find_all_in(self, haystack: DeviceHaystack):
        out = np.empty(self.n_needles, np.uint64)
        _check(lib().ss_b200_batch_find_all_in(self._b, haystack._h, out.ctypes.data))
        return out

    def close(self):
        if self._b:
            lib().ss_b200_batch_free(self._b)
            self._b = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rarest_position(needle, hist=None) -> int:
    """Second-anchor index ``ss_b200_rarest_position`` picks for ``needle`` (0 for len < 2)."""
    nb = bytes(needle)
    addr, n, keep = _host_view(nb)
    hp = None
    if hist is not None:
        hist = np.ascontiguousarray(hist, np.uint64)
        if hist.size != 256:
            raise B200Error("hist must hold 256 counts")
        hp = hist.ctypes.data
    out = C.c_size_t(0)
    _check(lib().ss_b200_rarest_position(addr, n, hp, C.byref(out)))
    return out.value


def _csr(items):
    off = np.zeros(len(items) + 1, np.uint64)
    if len(items):
        off[1:] = np.cumsum([len(x) for x in items], dtype=np.uint64)
    blob = np.frombuffer(b"".join(bytes(x) for x in items), np.uint8) if int(off[-1]) else np.zeros(1, np.uint8)
    return np.ascontiguousarray(blob), off


def fill_random(t, global_start: int, seed: int, stream=None) -> None:
    """Fill a CUDA uint8 tensor with the splitmix64 byte stream of BASELINE configs 4/5."""
    import torch

    sp = (stream or torch.cuda.current_stream(t.device)).cuda_stream
    _check(lib().ss_b200_fill_random(t.data_ptr(), t.numel(), global_start, seed, sp))


def fill_tiled(t, global_start: int, src, stream=None) -> None:
    """t[i] = src[(global_start + i) % len(src)] for CUDA uint8 tensors (config 2')."""
    import torch

    sp = (stream or torch.cuda.current_stream(t.device)).cuda_stream
    _check(lib().ss_b200_fill_tiled(t.data_ptr(), t.numel(), global_start, src.data_ptr(), src.numel(), sp))


def set_scan_variant(variant: int) -> None:
    _check(lib().ss_b200_set_scan_variant(variant))


def set_scan_tuning(ctas_per_sm: int = 0, unroll: int = 0, tile_kib: int = 0, stages: int = 0) -> None:
    _check(lib().ss_b200_set_scan_tuning(ctas_per_sm, unroll, tile_kib, stages))


def set_extra_anchors(n: int = -1) -> None:
    _check(lib().ss_b200_set_extra_anchors(n))


def launch_count() -> int:
    return lib().ss_b200_launch_count()


def set_launch_pdl(on: bool = True) -> None:
    _check(lib().ss_b200_set_launch_pdl(1 if on else 0))


def set_sync_service(on: bool = True, idle_us: int = 0) -> None:
    """Resident kernel for synchronous searches over short device-resident haystacks
    (``ss_b200_set_sync_service``); ``idle_us`` = 0 keeps the current idle time."""
    _check(lib().ss_b200_set_sync_service(1 if on else 0, idle_us))


def set_host_path(mode: int = 0, chunk_mib: int = 0, copy_threads: int = -1) -> None:
    """Host-slice path knobs (``ss_b200_set_host_path``): mode 0 auto / 1 DMA ring / 2 in place (direct
    loads) / 3 in place (TMA); chunk size in MiB (0 = sized from the slice); memcpy workers for pageable
    input (-1 auto, 0 = driver staging)."""
    _check(lib().ss_b200_set_host_path(mode, chunk_mib, copy_threads))


def measure_h2d(nbytes: int = 1 << 30, reps: int = 3) -> float:
    """Pinned host->device cudaMemcpyAsync bandwidth of the current device in GB/s."""
    out = C.c_double(0.0)
    _check(lib().ss_b200_measure_h2d(nbytes, reps, C.byref(out)))
    return out.value


def thread_release() -> None:
    """Free the calling thread's streams, result slot and staging ring (``ss_b200_thread_release``)."""
    _check(lib().ss_b200_thread_release())


def thread_footprint():
    d, p = C.c_size_t(0), C.c_size_t(0)
    _check(lib().ss_b200_thread_footprint(C.byref(d), C.byref(p)))
    return d.value, p.value


def _host_addr(haystack):
    if _is_torch_tensor(haystack):
        if haystack.is_cuda or not haystack.is_contiguous() or haystack.dtype.itemsize != 1:
            raise B200Error("host tensor must be a contiguous CPU tensor of 1-byte elements")
        return haystack.data_ptr(), haystack.numel(), haystack
    return _host_view(haystack)


class ShardedHaystack:
    """One haystack as contiguous shards of start positions, one per device of a :class:`Context`
    (``ss_b200_sharded``)."""

    def __init__(self, handle, keepalive=None):
        self._h = handle
        self._keep = keepalive

    def __len__(self) -> int:
        return lib().ss_b200_sharded_len(self._h)

    def shard(self, i: int):
        """(device pointer, start, owned, span) of shard i."""
        p, a, b, c = C.c_void_p(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        _check(lib().ss_b200_sharded_shard(self._h, i, C.byref(p), C.byref(a), C.byref(b), C.byref(c)))
        return p.value or 0, a.value, b.value, c.value

    def close(self) -> None:
        if self._h:
            lib().ss_b200_sharded_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ContextHaystackSet:
    """A set of haystacks partitioned over the devices of a :class:`Context` (``ss_b200_ctx_hayset``)."""

    def __init__(self, handle):
        self._h = handle

    def __len__(self) -> int:
        return lib().ss_b200_ctx_hayset_len(self._h)

    def part(self, i: int):
        lo, hi = C.c_size_t(), C.c_size_t()
        _check(lib().ss_b200_ctx_hayset_part(self._h, i, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def close(self) -> None:
        if self._h:
            lib().ss_b200_ctx_hayset_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """Every GPU of the box from one process (``ss_b200_ctx``): sharded device-resident haystacks,
    one host slice striped over all PCIe links, and the many-haystack mode partitioned over the devices."""

    def __init__(self, ndev: int = 0, devices=None, exchange: int = EXCHANGE_HOST):
        h = C.c_void_p()
        arr = None
        if devices is not None:
            ndev = len(devices)
            arr = (C.c_int * ndev)(*devices)
        _check(lib().ss_b200_ctx_create(ndev, arr, C.byref(h)))
        self._c = h
        if exchange != EXCHANGE_HOST:
            self.set_exchange(exchange)

    @property
    def device_count(self) -> int:
        return lib().ss_b200_ctx_device_count(self._c)

    @property
    def devices(self):
        return [lib().ss_b200_ctx_device(self._c, i) for i in range(self.device_count)]

    def set_exchange(self, kind: int) -> None:
        _check(lib().ss_b200_ctx_set_exchange(self._c, kind))

    def upload_sharded(self, data, halo: int = 4096) -> ShardedHaystack:
        addr, n, keep = _host_addr(data)
        h = C.c_void_p()
        _check(lib().ss_b200_sharded_upload(self._c, addr, n, halo, C.byref(h)))
        return ShardedHaystack(h)

    def sharded_from_tensors(self, tensors, owned) -> ShardedHaystack:
        """Borrow one CUDA uint8 tensor per device of the context (tensor d on device d holds
        ``owned[d]`` start positions followed by its right halo)."""
        n = self.device_count
        if len(tensors) != n or len(owned) != n:
            raise B200Error("one tensor and one owned count per device of the context")
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
        ow = (C.c_size_t * n)(*[int(x) for x in owned])
        sp = (C.c_size_t * n)(*[t.numel() for t in tensors])
        h = C.c_void_p()
        _check(lib().ss_b200_sharded_from_device(self._c, ptrs, ow, sp, C.byref(h)))
        return ShardedHaystack(h, keepalive=list(tensors))

    def find_sharded(self, searcher, sharded: ShardedHaystack) -> Optional[int]:
        out = C.c_size_t(0)
        _check(lib().ss_b200_find_sharded(self._c, searcher._s, sharded._h, C.byref(out)))
        return None if out.value == NPOS else out.value

    def search_sharded(self, searcher, sharded: ShardedHaystack) -> bool:
        found, off = C.c_uint8(0), C.c_size_t(0)
        _check(lib().ss_b200_search_sharded(self._c, searcher._s, sharded._h, C.byref(found), C.byref(off)))
        return bool(found.value)

    def find_in_host(self, searcher, haystack) -> Optional[int]:
        """``search_in(&[u8])`` with one host slice striped over every device of the context."""
        addr, n, keep = _host_addr(haystack)
        out = C.c_size_t(0)
        _check(lib().ss_b200_find_in_host_multi(self._c, searcher._s, addr, n, C.byref(out)))
        return None if out.value == NPOS else out.value

    def last_host_stats(self) -> dict:
        a, b, c, m = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        _check(lib().ss_b200_ctx_last_host_stats(self._c, C.byref(a), C.byref(b), C.byref(c), C.byref(m)))
        return {"h2d_bytes": a.value, "chunks": b.value, "chunk_bytes": c.value, "mode": m.value}

    def upload_haystack_set(self, haystacks) -> ContextHaystackSet:
        blob, off = _csr(haystacks)
        h = C.c_void_p()
        _check(lib().ss_b200_ctx_hayset_upload(self._c, blob.ctypes.data, off.ctypes.data, len(haystacks), C.byref(h)))
        return ContextHaystackSet(h)

    def search_haystack_set(self, searcher, hset: ContextHaystackSet) -> np.ndarray:
        flags = np.zeros(len(hset), np.uint8)
        _check(lib().ss_b200_ctx_hayset_search(self._c, searcher._s, hset._h, flags.ctypes.data))
        return flags

    def close(self) -> None:
        if self._c:
            lib().ss_b200_ctx_free(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_version() -> int:
    v = C.c_int(0)
    _check(lib().ss_b200_ctx_nccl_version(C.byref(v)))
    return v.value
