"""Multi-GPU modes: one process per GPU, ``torch.distributed`` as plumbing.

The path shards embarrassingly (start positions are independent, reference
src/lib.rs:263-274 carries no state between chunks except the early return):

* **one huge haystack** -- rank r owns the start positions of the contiguous byte range
  ``[r*S, (r+1)*S)`` and holds ``S + k - 1`` bytes (right halo only).  Each rank runs the K1
  scan with ``base_offset = r*S`` and ``start_limit = S``; the only exchange is an 8-byte
  ``all_reduce(MIN)`` over the global first offsets (``DEVICE_NONE`` = INT64_MAX = not found),
  which yields both the OR of the per-shard match flags and the exact leftmost offset.
  NCCL has no bitwise OR (nccl.h: sum, prod, max, min, avg), hence MIN.
* **many haystacks** -- haystack index ranges per rank (length-balanced), needles replicated.  Every
  haystack lives on one rank, so the ranks' flags are disjoint: each rank bit-packs its slice into its
  own bit range of one global bitmap (``ss_b200_pack_flags_async``) and the bitmaps are combined with
  ``all_reduce(SUM)`` on int32 words -- a sum of disjoint bits is the bitwise OR NCCL lacks -- at 1 bit
  per haystack (the first version MAX-reduced 1 byte per haystack x world).

The partition arithmetic and the reductions are backend-agnostic (tested with gloo,
world_size 2, on CPU); the scans themselves only run on CUDA.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

from . import DEVICE_NONE


def shard_bounds(total_len: int, k: int, world: int, rank: int, align: int = 16) -> Tuple[int, int, int]:
    """Contiguous shard of a haystack of ``total_len`` bytes for a needle of length ``k``.

    Returns ``(start, owned, span)``: the shard owns start positions ``[start, start+owned)``
    and must hold bytes ``[start, start+span)`` (owned + right halo of k-1, clipped).
    ``owned`` is a multiple of ``align`` except for the last non-empty shard."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = -(-total_len // world)
    per = -(-per // align) * align
    start = min(rank * per, total_len)
    owned = min(per, total_len - start)
    span = min(owned + max(k, 1) - 1, total_len - start)
    return start, owned, span


def partition_by_length(lengths: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Split haystack indices into ``world`` contiguous ranges with balanced total bytes."""
    n = len(lengths)
    total = sum(lengths)
    out, i, acc = [], 0, 0
    for r in range(world):
        target = total * (r + 1) / world
        j = i
        while j < n and (acc + lengths[j] <= target or r == world - 1):
            acc += lengths[j]
            j += 1
        out.append((i, j))
        i = j
    return out


def reduce_first_offset(local, group=None) -> Optional[int]:
    """all_reduce(MIN) over per-rank global first offsets (1-element int64 tensor holding the
    offset or DEVICE_NONE).  Returns the global leftmost offset or None; `local` is updated."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local, op=dist.ReduceOp.MIN, group=group)
    v = int(local.item())
    return None if v == DEVICE_NONE else v


def reduce_flags(flags, group=None):
    """all_reduce(MAX) over per-haystack uint8 match flags (each haystack lives on one rank).
    Kept for byte-flag consumers and CPU backends; the CUDA path uses :func:`reduce_packed_flags`."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
    return flags


def pack_flags_reference(local_flags, first_bit: int, total_bits: int):
    """numpy statement of ``ss_b200_pack_flags_async``: int32 words of a ``total_bits`` bitmap with bit
    ``first_bit + i`` set iff ``local_flags[i] != 0`` (all other bits zero)."""
    import numpy as np

    bits = np.zeros(((total_bits + 31) // 32) * 32, np.uint8)
    lf = np.asarray(local_flags)
    bits[first_bit:first_bit + lf.size] = lf != 0
    return np.packbits(bits, bitorder="little").view(np.int32)


def unpack_flags(words, total_bits: int):
    """Inverse of the packing: int32 bitmap words (numpy or CPU tensor) -> ``total_bits`` uint8 flags."""
    import numpy as np

    w = words.numpy() if hasattr(words, "numpy") else np.asarray(words)
    return np.unpackbits(np.ascontiguousarray(w).view(np.uint8), bitorder="little")[:total_bits]


def reduce_packed_flags(words, group=None, async_op: bool = False):
    """all_reduce(SUM) over int32 bitmap words whose set bits are disjoint between ranks == bitwise OR."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return dist.all_reduce(words, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return None


def pack_flags_async(local_flags, first_bit: int, words, total_bits: int, stream=None) -> None:
    """``ss_b200_pack_flags_async`` on CUDA tensors: zero ``words`` (int32, ceil(total_bits/32)) and set
    this rank's bit range from its uint8 flags."""
    import torch

    from . import _check, lib

    if stream is None:
        stream = torch.cuda.current_stream(words.device)
    _check(lib().ss_b200_pack_flags_async(local_flags.data_ptr() if local_flags.numel() else None,
                                          local_flags.numel(), int(first_bit), words.data_ptr(), int(total_bits),
                                          stream.cuda_stream))


class ShardedSearch:
    """Rank-local state for repeated searches over one sharded haystack.

    ``shard``: CUDA uint8 tensor holding this rank's bytes (owned + halo);
    ``start``/``owned``: from :func:`shard_bounds`."""

    def __init__(self, shard, start: int, owned: int, group=None, exchange: str = "nccl"):
        """``exchange``: "nccl" = 8-byte all_reduce(MIN) per search (works with any backend);
        "peer" = fused into the scan epilogue through :class:`PeerExchange` (CUDA only)."""
        import torch

        self.shard, self.start, self.owned, self.group = shard, start, owned, group
        self.workspace = torch.zeros(32, dtype=torch.uint8, device=shard.device)
        self.result = torch.full((1,), DEVICE_NONE, dtype=torch.int64, device=shard.device)
        self.peer = PeerExchange(group) if exchange == "peer" else None

    def find_async(self, searcher, stream=None):
        """Enqueue scan + exchange on ``stream`` (default: the current stream); returns the device
        result tensor.  The collective is issued under the same stream as the scan, so the MIN never
        reads the result word before the scan has written it."""
        import torch
        import torch.distributed as dist

        if stream is None:
            stream = torch.cuda.current_stream(self.shard.device)
        if self.peer is not None:
            return self.peer.find_async(searcher, self.shard, self.start, self.owned, self.workspace, self.result,
                                        stream=stream)
        searcher.find_in_async(self.shard, self.result, self.workspace, base_offset=self.start,
                               start_limit=self.owned, stream=stream)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(self.result, op=dist.ReduceOp.MIN, group=self.group)
        return self.result

    def find(self, searcher) -> Optional[int]:
        v = int(self.find_async(searcher).item())
        return None if v == DEVICE_NONE else v

    def find_many(self, searchers) -> List[Optional[int]]:
        """Pipelined batch of independent searches over the same sharded haystack: the 8-byte
        MIN-allreduce of search i runs on NCCL's stream while search i+1 already scans, so the
        exchange never stalls the scan stream.  Returns the global leftmost offsets."""
        import torch
        import torch.distributed as dist

        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        results = torch.full((len(searchers),), DEVICE_NONE, dtype=torch.int64, device=self.shard.device)
        works = []
        for i, s in enumerate(searchers):
            s.find_in_async(self.shard, results[i:i + 1], self.workspace, base_offset=self.start,
                            start_limit=self.owned)
            if multi:
                works.append(dist.all_reduce(results[i:i + 1], op=dist.ReduceOp.MIN, group=self.group, async_op=True))
        for w in works:
            w.wait()
        return [None if v == DEVICE_NONE else v for v in results.tolist()]

    def search(self, searcher) -> bool:
        return self.find(searcher) is not None


class PeerExchange:
    """Mailboxes for the fused exchange (``ss_b200_find_in_device_exchange_async``): every rank owns
    a small cudaMalloc buffer, exports it through CUDA IPC, and maps everybody else's.  The handles
    travel once through ``torch.distributed`` (any backend); afterwards a sharded search needs no
    collective call at all -- the scan kernel's epilogue writes the result into peer HBM."""

    def __init__(self, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _check, lib

        self.group = group
        multi = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if multi else 1
        self.rank = dist.get_rank(group) if multi else 0
        self.seq = 0
        own = C.c_void_p()
        _check(lib().ss_b200_mailbox_create(self.world, C.byref(own)))
        self._own = own
        handle = (C.c_uint8 * 64)()
        _check(lib().ss_b200_ipc_export(own, handle))
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, bytes(handle), group=group)
        self._opened = []
        ptrs = (C.c_void_p * self.world)()
        for r in range(self.world):
            if r == self.rank:
                ptrs[r] = own.value
            else:
                p = C.c_void_p()
                buf = (C.c_uint8 * 64).from_buffer_copy(handles[r])
                _check(lib().ss_b200_ipc_open(buf, C.byref(p)))
                self._opened.append(p)
                ptrs[r] = p.value
        self.ptrs = ptrs
        if self.world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=group)  # nobody posts before every mailbox is mapped and emptied

    def find_async(self, searcher, shard, start: int, owned: int, workspace, result, stream=None):
        import torch

        from . import NPOS, _check, lib

        if stream is None:
            stream = torch.cuda.current_stream(shard.device)
        self.seq += 1
        _check(lib().ss_b200_find_in_device_exchange_async(
            searcher._s, shard.data_ptr(), shard.numel(), int(start), NPOS if owned is None else int(owned),
            workspace.data_ptr(), self.ptrs, self.world, self.rank, self.seq, result.data_ptr(), stream.cuda_stream))
        return result

    def close(self):
        from . import lib

        for p in self._opened:
            lib().ss_b200_ipc_close(p)
        self._opened = []
        if self._own:
            lib().ss_b200_mailbox_free(self._own)
            self._own = None


class ShardedHaystackSet:
    """Many-haystack mode over several GPUs: haystack index ranges per rank (length-balanced,
    :func:`partition_by_length`), needles replicated.  Each rank scans its own haystacks, bit-packs its
    flags into its bit range of the global bitmap, and the bitmaps are OR-ed with ``all_reduce(SUM)``
    (disjoint bits; 1 bit per haystack on the wire)."""

    def __init__(self, haystacks, rank: int = 0, world: int = 1, group=None, device="cuda"):
        import torch

        from . import HaystackSet

        self.total = len(haystacks)
        self.group = group
        self.lo, self.hi = partition_by_length([len(h) for h in haystacks], world)[rank]
        self.local = HaystackSet(haystacks[self.lo:self.hi], device=device)
        self.local_flags = torch.zeros(max(self.hi - self.lo, 1), dtype=torch.uint8, device=device)
        self.words = torch.zeros((self.total + 31) // 32, dtype=torch.int32, device=device)

    def search_async(self, searcher, stream=None):
        """Bitmap words (int32 CUDA tensor): bit h = searcher.search_in(haystack h) for every haystack of
        the global set.  Everything is enqueued on ``stream`` (default: the current stream)."""
        import torch

        if stream is None:
            stream = torch.cuda.current_stream(self.words.device)
        n_local = self.hi - self.lo
        if n_local > 0:
            searcher.search_many_async(self.local, self.local_flags[:n_local], stream=stream)
        pack_flags_async(self.local_flags[:n_local], self.lo, self.words, self.total, stream=stream)
        with torch.cuda.stream(stream):
            reduce_packed_flags(self.words, self.group)
        return self.words

    def search(self, searcher):
        return unpack_flags(self.search_async(searcher).cpu(), self.total).astype(bool)
