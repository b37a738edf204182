"""CPU tests of the multi-GPU host logic: shard arithmetic and the MIN / MAX reductions, on the
gloo backend with world_size 2.  The per-rank scan is replaced by the oracle (the checker) so the
test pins exactly what sharded.py adds: partition + halo + reduction."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from sliceslice_rs_b200 import DEVICE_NONE
from sliceslice_rs_b200.sharded import (pack_flags_reference, partition_by_length, reduce_first_offset, reduce_flags,
                                        reduce_packed_flags, shard_bounds, unpack_flags)


def test_shard_bounds_cover_every_start_position_once():
    rng = random.Random(3)
    for _ in range(300):
        total = rng.randrange(0, 5000)
        k = rng.randrange(0, 40)
        world = rng.choice([1, 2, 3, 4, 8])
        covered = 0
        prev_end = 0
        for r in range(world):
            start, owned, span = shard_bounds(total, k, world, r)
            assert start == prev_end or owned == 0
            assert start + span <= total
            assert span >= min(owned, total - start)
            if owned and start + owned < total:
                assert owned % 16 == 0
                assert span == min(owned + max(k, 1) - 1, total - start)
            prev_end = start + owned
            covered += owned
        assert covered == total


def test_sharded_find_equals_global_find_single_process():
    rng = random.Random(11)
    for _ in range(400):
        n = rng.randrange(1, 400)
        k = rng.randrange(1, 9)
        h = bytes(rng.randrange(3) + 97 for _ in range(n))
        nd = bytes(rng.randrange(3) + 97 for _ in range(k))
        world = rng.choice([2, 4, 8])
        best = None
        for r in range(world):
            start, owned, span = shard_bounds(n, k, world, r)
            local = oracle.find(h[start:start + span], nd) if span >= k else None
            if local is not None and local < owned:
                best = start + local if best is None else min(best, start + local)
        exp = h.find(nd)
        assert best == (None if exp < 0 else exp)


def test_partition_by_length_balances_and_covers():
    lens = [5, 1, 1, 1, 10, 2, 2, 7, 3]
    parts = partition_by_length(lens, 3)
    assert parts[0][0] == 0 and parts[-1][1] == len(lens)
    assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, hay, needle, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k = len(needle)
        start, owned, span = shard_bounds(len(hay), k, world, rank)
        local = oracle.find(hay[start:start + span], needle) if span >= k else None
        v = DEVICE_NONE if (local is None or local >= owned) else start + local
        t = torch.tensor([v], dtype=torch.int64)
        got = reduce_first_offset(t)
        # batch mode: haystack set partitioned by index range, flags MAX-reduced
        hays = [hay[i:i + 50] for i in range(0, min(len(hay), 1000), 50)]
        lo, hi = partition_by_length([len(x) for x in hays], world)[rank]
        flags = torch.zeros(len(hays), dtype=torch.uint8)
        for i in range(lo, hi):
            flags[i] = 1 if oracle.search_in(hays[i], needle) else 0
        # the cheaper form of the same OR: each rank packs its own slice into its bit range of one
        # bitmap (all other bits zero) and the bitmaps are SUM-reduced -- disjoint bits, so SUM == OR
        words = torch.from_numpy(pack_flags_reference(flags[lo:hi].numpy(), lo, len(hays)).copy())
        reduce_packed_flags(words)
        reduce_flags(flags)
        if rank == 0:
            ret["offset"] = got
            ret["flags"] = flags.tolist()
            ret["packed_flags"] = unpack_flags(words, len(hays)).tolist()
            ret["exp_flags"] = [1 if needle in x else 0 for x in hays]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["absent", "rank0", "rank1", "straddle", "both"])
def test_gloo_world2_min_reduce(case):
    rng = np.random.default_rng(42)
    n = 4096
    hay = bytearray(rng.integers(97, 100, n, dtype=np.uint8).tobytes())
    needle = b"XYZW"
    if case in ("rank0", "both"):
        hay[100:104] = needle
    if case in ("rank1", "both"):
        hay[3000:3004] = needle
    if case == "straddle":
        hay[2046:2050] = needle  # shard boundary at 2048: found only thanks to the halo
    hay = bytes(hay)
    exp = hay.find(needle)
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(2, _free_port(), hay, needle, ret), nprocs=2, join=True)
        assert ret["offset"] == (None if exp < 0 else exp)
        assert ret["flags"] == ret["exp_flags"]
        assert ret["packed_flags"] == ret["exp_flags"]


def test_packed_flags_are_disjoint_and_round_trip():
    rng = np.random.default_rng(7)
    for total in (1, 31, 32, 33, 1000, 4097):
        flags = (rng.integers(0, 3, total) == 0).astype(np.uint8) * rng.integers(1, 255, total).astype(np.uint8)
        for world in (1, 2, 3, 8):
            cuts = sorted(rng.integers(0, total + 1, world - 1).tolist())
            bounds = [0] + cuts + [total]
            acc = np.zeros((total + 31) // 32, np.int64)
            for r in range(world):
                lo, hi = bounds[r], bounds[r + 1]
                w = pack_flags_reference(flags[lo:hi], lo, total)
                assert not (acc.astype(np.uint32) & w.view(np.uint32)).any()  # disjoint: SUM == OR, no carries
                acc += w.view(np.uint32)
            assert np.array_equal(unpack_flags(acc.astype(np.uint32), total), (flags != 0).astype(np.uint8))
