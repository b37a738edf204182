"""GPU parity tests: the sm_100a path, called through the C ABI (ctypes -> libsliceslice_b200.so),
against the CPU oracle (C restatement of DynamicAvx2Searcher), the reference's golden vectors
(tests/golden, transcribed from src/lib.rs:299-331, :422-544, src/x86.rs:6-14, :533-565) and the
corpus numbers of tests/i386.rs.  Bit-exact: same found flag and same first-match offset.

Every test runs the product path only; the oracle is the checker."""
import hashlib
import os
import random
import threading

import numpy as np
import pytest

import oracle
import sliceslice_rs_b200 as ss

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

VARIANTS = [1, 2]  # 1 = direct LDG, 2 = TMA-staged ring
SEED_HAY = 0x5EEDB20000000001


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: the CUDA path is the only path (no CPU fallback)")
    assert os.path.exists(ss.LIB_PATH), "libsliceslice_b200.so must be built in-tree"
    ss.lib()
    yield
    ss.set_scan_variant(0)
    ss.set_scan_tuning(0, 0, 0, 0)
    ss.set_host_path(0, 0, -1)


@pytest.fixture(params=VARIANTS, ids=["ldg", "tma"])
def variant(request):
    ss.set_scan_variant(request.param)
    yield request.param
    ss.set_scan_variant(0)


def _dev(b) -> "torch.Tensor":
    a = np.frombuffer(bytes(b), np.uint8) if not isinstance(b, np.ndarray) else b
    return torch.from_numpy(a.copy()).cuda()


def _expect(h: bytes, nd: bytes):
    e = h.find(nd)
    return None if e < 0 else e


# ------------------------------------------------------------------------------------------
# golden vectors of the reference's own tests


def test_kats_every_position(kats, variant):
    # src/lib.rs:365-381: every pair, every position in 0..needle.len(), literal expected bools
    for c in kats["kats"]:
        h, nd = c["haystack"].encode(), c["needle"].encode()
        hs = ss.DeviceHaystack.upload(h)
        for pos in range(len(nd)):
            for cls in (ss.DynamicB200Searcher, ss.B200Searcher):
                s = cls.with_position(nd, pos)
                assert s.search_in(hs) is c["found"], (h, nd, pos)
                assert s.find_in(hs) == c["offset"], (h, nd, pos)
                assert s.find_in(hs) == oracle.find(h, nd, pos, dynamic=cls is ss.DynamicB200Searcher)
                s.close()
        s = ss.DynamicB200Searcher.new(nd)
        assert s.search_in(hs) is c["found"]
        assert s.search_in(h) is c["found"]  # host-slice entry (ss_b200_search_in_host)
        s.close()
        hs.close()


def test_memchr_kats_and_doctests(kats, variant):
    for c in kats["memchr"] + kats["doctest"]:
        h, nd = c["haystack"].encode(), c["needle"].encode()
        s = ss.DynamicB200Searcher.new(nd)
        hs = ss.DeviceHaystack.upload(h)
        assert s.search_in(hs) is c["found"]
        assert s.find_in(hs) == c["offset"]
        assert s.find_in(h) == c["offset"]


def test_ctor_contract_on_gpu_box(kats):
    for c in kats["ctor"]:
        cls = ss.DynamicB200Searcher if c["searcher"] == "dynamic" else ss.B200Searcher
        n = c["needle"].encode()
        make = (lambda: cls.new(n)) if c["position"] is None else (lambda: cls.with_position(n, c["position"]))
        if c["outcome"] == "panic":
            with pytest.raises(ss.SearcherPanic):
                make()
        else:
            make().close()


def test_edge_semantics(variant):
    dyn = ss.DynamicB200Searcher
    e = ss.DeviceHaystack.upload(b"")
    assert dyn.new(b"").search_in(e) is True  # N0, src/x86.rs:470,500
    assert dyn.new(b"").find_in(ss.DeviceHaystack.upload(b"abc")) == 0
    assert dyn.new(b"a").search_in(e) is False  # src/lib.rs:131-133
    assert dyn.new(b"abcd").search_in(ss.DeviceHaystack.upload(b"abc")) is False  # src/x86.rs:357-359
    assert dyn.new(b"abc").find_in(ss.DeviceHaystack.upload(b"abc")) == 0  # n == k => haystack == needle
    assert dyn.new(b"abd").find_in(ss.DeviceHaystack.upload(b"abc")) is None
    # strict one-byte needle takes the two-anchor path with position 0 (Avx2Searcher<[u8;1]>)
    assert ss.B200Searcher.new(b"c").find_in(ss.DeviceHaystack.upload(b"abcabc")) == 2
    # NUL bytes and bytes >= 0x80 are ordinary bytes
    h = bytes([0, 0xFF, 0, 0x80, 0x7F, 0, 0xFF, 0xFF])
    assert dyn.new(bytes([0xFF, 0xFF])).find_in(ss.DeviceHaystack.upload(h)) == 6
    assert dyn.new(bytes([0])).find_in(ss.DeviceHaystack.upload(h)) == 0
    assert dyn.new(bytes([0x80, 0x7F, 0])).find_in(ss.DeviceHaystack.upload(h)) == 3


# ------------------------------------------------------------------------------------------
# corpus sweeps (tests/i386.rs:46-70; workload bench/benches/i386.rs:246-257, :118-131)


def test_ipsum_absent_from_i386(i386, variant):
    hs = ss.DeviceHaystack.upload(i386)
    for pos in range(5):
        assert ss.DynamicB200Searcher.with_position(b"ipsum", pos).search_in(hs) is False


def test_long_sweep_api_faithful(corpus, i386, words, variant):
    # one search_in per needle, exactly as bench/benches/i386.rs:252-256
    hs = ss.DeviceHaystack.upload(i386)
    searchers = [ss.DynamicB200Searcher.new(w) for w in words]
    offs = [s.find_in(hs) for s in searchers]
    assert all(o is not None for o in offs)
    assert offs == corpus["long"]["first_offsets"]
    assert sum(offs) == corpus["long"]["sum_first_offsets"] == 809985317
    for s in searchers:
        s.close()


def test_long_sweep_other_positions(corpus, i386, words):
    hs = ss.DeviceHaystack.upload(i386)
    rng = random.Random(7)
    for idx in rng.sample(range(len(words)), 400):
        w = words[idx]
        pos = rng.randrange(len(w))
        assert ss.DynamicB200Searcher.with_position(w, pos).find_in(hs) == corpus["long"]["first_offsets"][idx]


def test_long_sweep_batched_single_launch(corpus, i386, words):
    hs = ss.DeviceHaystack.upload(i386)
    b = ss.Batch(words, [])
    before = ss.launch_count()
    offs = b.find_all_in(hs)
    assert ss.launch_count() - before == 1
    assert offs.tolist() == corpus["long"]["first_offsets"]


def test_long_sweep_lossy_utf8_haystack(corpus, i386, words):
    # tests/i386.rs:63 searches String::from_utf8_lossy(I386) (958 733 bytes)
    lossy = i386.decode("utf-8", errors="replace").encode("utf-8")
    assert len(lossy) == corpus["long_lossy_utf8"]["len"]
    hs = ss.DeviceHaystack.upload(lossy)
    b = ss.Batch(words, [])
    offs = b.find_all_in(hs)
    exp = oracle.long_sweep(words, lossy, threads=8)
    assert np.array_equal(offs, exp)
    assert int((offs != ss.NPOS).sum()) == corpus["long_lossy_utf8"]["found"]


def test_short_sweep_triangular(corpus, sorted_words):
    b = ss.Batch(sorted_words, sorted_words)
    bm, matches = b.search_triangular()
    assert matches == corpus["short"]["matches"] == 39105
    assert int(np.unpackbits(bm.view(np.uint8)).sum()) == 39105
    assert hashlib.sha256(bm.tobytes()).hexdigest() == corpus["short"]["bitmap_sha256"]


def test_pairs_mode_vs_oracle(sorted_words):
    rng = np.random.default_rng(3)
    extra = [b"", b"a", b"the", b"x" * 40, bytes(range(256))]
    needles = sorted_words[::7] + extra
    hays = sorted_words[::5] + extra + [b"lorem ipsum dolor sit amet " * 40]
    pn = rng.integers(0, len(needles), 200_000, dtype=np.uint32)
    ph = rng.integers(0, len(hays), 200_000, dtype=np.uint32)
    b = ss.Batch(needles, hays)
    bm, off = b.search_pairs(pn, ph)
    exp = oracle.pairs(needles, hays, pn, ph)
    assert np.array_equal(off, exp)
    bits = np.unpackbits(bm.view(np.uint8), bitorder="little")[: pn.size]
    assert np.array_equal(bits.astype(bool), exp != oracle.NPOS)


def test_random_bench_size_grid(variant):
    # bench/benches/random.rs:12-99: prefixes of data/needle in prefixes of data/haystack for the
    # sizes {1,5,10,20,50,100,1000}, every needle size against every not-smaller haystack size (28 pairs)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    haystack = open(os.path.join(root, "data", "haystack"), "rb").read()
    needle = open(os.path.join(root, "data", "needle"), "rb").read()
    sizes = [1, 5, 10, 20, 50, 100, 1000]
    found = []
    for i, ns in enumerate(sizes):
        nd = needle[:ns]
        s = ss.DynamicB200Searcher.new(nd)
        for hsz in sizes[i:]:
            h = haystack[:hsz]
            got = s.find_in(ss.DeviceHaystack.upload(h))
            assert got == oracle.find(h, nd) == _expect(h, nd), (ns, hsz)
            assert s.search_in(h) is (got is not None)  # host-slice entry
            if got is not None:
                found.append((ns, hsz))
    assert found == [(1, 100), (1, 1000)]  # the only true pairs


# ------------------------------------------------------------------------------------------
# offset-pinning and randomized differential tests


def test_mula_dollar_hash_offsets(variant):
    # vendored unittests.cpp:13-51: "$x..x#" between paddings; offset must equal the left padding
    for size in list(range(1, 40, 3)) + [63, 200]:
        needle = b"$" + b"x" * size + b"#"
        for pre in (0, 1, 15, 16, 17, 31, 32, 33, 47, 255, 256, 4095):
            for post in (0, 1, 31, 47):
                h = b"_" * pre + needle + b"_" * post
                hs = ss.DeviceHaystack.upload(h)
                for position in (0, len(needle) // 2, len(needle) - 1):
                    assert ss.DynamicB200Searcher.with_position(needle, position).find_in(hs) == pre
                hs.close()


@pytest.mark.parametrize("alphabet", [2, 3, 26])
def test_randomized_differential(alphabet, variant):
    # small alphabets make filter candidates, failed verifies and overlapping matches frequent;
    # device sub-slices at byte offsets 0..31 exercise every head alignment
    rng = random.Random(99 + alphabet)
    pool = torch.empty(4096, dtype=torch.uint8, device="cuda")
    for it in range(2500):
        n = rng.randrange(0, 400) if it % 5 else rng.randrange(0, 40)
        k = rng.randrange(0, 40) if it % 3 else rng.randrange(0, 6)
        h = bytes(rng.randrange(alphabet) + 97 for _ in range(n))
        if k and n >= k and rng.random() < 0.4:
            st = rng.randrange(0, n - k + 1)
            nd = h[st:st + k]
        else:
            nd = bytes(rng.randrange(alphabet) + 97 for _ in range(k))
        pos = 0 if k <= 1 else rng.randrange(k)
        a = rng.randrange(32)
        pool[a:a + n] = torch.frombuffer(bytearray(h), dtype=torch.uint8).cuda() if n else pool[a:a]
        # poison the bytes around the slice so an out-of-range read would produce false matches
        if k:
            fill = nd[0]
            pool[max(a - 16, 0):a] = fill
            pool[a + n:a + n + 48] = fill
        s = ss.DynamicB200Searcher.with_position(nd, pos)
        got = s.find_in(pool[a:a + n])
        exp = oracle.find(h, nd, pos)
        assert exp == _expect(h, nd)
        assert got == exp, (h, nd, pos, a, got, exp)
        s.close()


@pytest.mark.parametrize("ne", [0, 1])
def test_extra_anchor_settings_do_not_change_results(ne, variant):
    # extra anchors (adaptive, see AdaptiveFilter in ss_device.cuh) only thin out the candidates
    ss.set_extra_anchors(ne)
    try:
        rng = random.Random(500 + ne)
        pool = torch.empty(70000, dtype=torch.uint8, device="cuda")
        for it in range(600):
            alphabet = rng.choice([2, 3, 4])
            n = rng.randrange(0, 600) if it % 4 else rng.randrange(30000, 66000)
            k = rng.randrange(1, 48)
            h = bytes(rng.randrange(alphabet) + 97 for _ in range(n))
            if n >= k and rng.random() < 0.5:
                st = rng.randrange(0, n - k + 1)
                nd = h[st:st + k]
            else:
                nd = bytes(rng.randrange(alphabet) + 97 for _ in range(k))
            pos = 0 if k <= 1 else rng.randrange(k)
            a = rng.randrange(32)
            if n:
                pool[a:a + n] = torch.frombuffer(bytearray(h), dtype=torch.uint8).cuda()
            pool[a + n:a + n + 64] = nd[0]
            s = ss.DynamicB200Searcher.with_position(nd, pos)
            assert s.find_in(pool[a:a + n]) == oracle.find(h, nd, pos) == _expect(h, nd), (h[:80], nd, pos, a)
            s.close()
    finally:
        ss.set_extra_anchors(-1)


def test_adversarial_all_candidates(variant):
    # every position passes the filter (src/x86.rs:252-255 motivates `position`)
    n = 1 << 20
    h = b"a" * n
    hs = ss.DeviceHaystack.upload(h)
    assert ss.DynamicB200Searcher.new(b"a" * 31 + b"b" + b"a").find_in(hs) is None
    assert ss.DynamicB200Searcher.with_position(b"a" * 31 + b"b" + b"a", 31).find_in(hs) is None
    h2 = bytearray(h)
    h2[n - 100] = ord("b")
    hs2 = ss.DeviceHaystack.upload(bytes(h2))
    assert ss.DynamicB200Searcher.new(b"aaab" + b"a" * 20).find_in(hs2) == n - 103
    assert ss.DynamicB200Searcher.new(b"a" * 8).find_in(hs2) == 0


@pytest.mark.parametrize("k", [1, 2, 5, 16, 17, 33, 64, 65, 100, 300, 3000])
def test_tile_boundary_straddles(k, variant):
    # 24 MiB of generator bytes (alphabet 0..254); needle contains 0xFF so it is absent until planted.
    # Plant it so that it straddles every kind of internal boundary of both kernels.
    n = 24 << 20
    t = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    ss.fill_random(t[:n], 0, SEED_HAY)
    host = oracle.fill_random(0, n, SEED_HAY)
    assert np.array_equal(t[:n].cpu().numpy(), host), "device generator != oracle generator"
    rng = random.Random(k)
    nd = bytearray(rng.randrange(255) for _ in range(k))
    nd[min(1, k - 1)] = 0xFF
    nd = bytes(nd)
    positions = sorted({0, k // 2, k - 1, min(k - 1, 16), min(k - 1, 15), min(k - 1, 17)})
    searchers = [ss.DynamicB200Searcher.with_position(nd, p) for p in positions]
    hay = t[:n]
    for s in searchers:
        assert s.find_in(hay) is None
    edges = [0, 1, 16 - k // 2, 4096, 16384, 32768, 65536, 131072, 1 << 20, (8 << 20), n - k]
    spots = sorted({max(0, min(n - k, e - d)) for e in edges for d in (0, 1, k // 2, k - 1, k)}, reverse=True)
    ndt = torch.frombuffer(bytearray(nd), dtype=torch.uint8).cuda()
    for spot in spots:  # descending: each new plant becomes the leftmost occurrence
        hay[spot:spot + k] = ndt
        host[spot:spot + k] = np.frombuffer(nd, np.uint8)
        for s in searchers:
            assert s.find_in(hay) == spot, (k, s.position, spot)
    assert oracle.find(host, nd) == spots[-1]


def test_unaligned_device_slices_long(variant):
    n = (9 << 20) + 123
    t = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    ss.fill_random(t, 0, SEED_HAY)
    nd = bytes([7, 0xFF, 9, 11, 13])
    s = ss.DynamicB200Searcher.new(nd)
    ndt = torch.frombuffer(bytearray(nd), dtype=torch.uint8).cuda()
    for a in (1, 3, 8, 15, 17, 31):
        sl = t[a:a + n - 40]
        assert s.find_in(sl) is None
        t[a + n - 45:a + n - 40] = ndt  # match ends exactly at the slice end
        assert s.find_in(sl) == n - 45
        assert s.find_in(t[a:a + n - 41]) is None  # one byte shorter: last byte missing
        ss.fill_random(t, 0, SEED_HAY)


def test_long_needles_fall_back_cleanly(variant):
    # second anchor further away than a TMA stage can carry (SS_TMA_HALO_MAX) -> LDG path
    n = 12 << 20
    t = torch.empty(n, dtype=torch.uint8, device="cuda")
    ss.fill_random(t, 0, SEED_HAY)
    host = oracle.fill_random(0, n, SEED_HAY)
    for k in (2049, 5000, 70000):
        spot = (10 << 20) + 5
        nd = bytes(host[spot:spot + k])
        for pos in (0, 63, 2047, 2048, k - 1):
            assert ss.DynamicB200Searcher.with_position(nd, pos).find_in(t) == oracle.find(host, nd, pos) == spot


def test_early_exit_returns_leftmost_of_many(variant):
    # thousands of matches: the atomic/early-exit logic must still return the leftmost one
    n = 64 << 20
    t = torch.empty(n, dtype=torch.uint8, device="cuda")
    src = _dev(b"abcdefghijklmnopqrstuvwxyz0123456789-")
    ss.fill_tiled(t, 0, src)
    s = ss.DynamicB200Searcher.new(b"xyz0123")
    assert s.find_in(t) == 23
    t[: 40 << 20] = 0x2E
    first = (40 << 20)
    exp = first + (23 - first % 37) % 37
    assert s.find_in(t) == exp
    host = t.cpu().numpy()
    assert oracle.find(host, b"xyz0123") == exp


# ------------------------------------------------------------------------------------------
# host-slice entry, async entry, shards, threads


def test_host_path_chunked(variant):
    ss.set_host_path(0, 1, -1)  # 1 MiB chunks: several chunks, ring wrap-around, chunk-boundary straddles
    n = (5 << 20) + 77
    host = oracle.fill_random(0, n, SEED_HAY)
    nd = bytes([1, 0xFF, 3, 4, 5, 6, 7])
    s = ss.DynamicB200Searcher.new(nd)
    assert s.find_in(host) is None
    for spot in (n - 7, (3 << 20) - 3, (1 << 20) - 6, (1 << 20) - 7, 5):
        host[spot:spot + 7] = np.frombuffer(nd, np.uint8)
        assert s.find_in(host) == spot
        assert s.search_in(host) is True
    assert oracle.find(host, nd) == 5
    ss.set_host_path(0, 0, -1)


def test_host_path_short_slices_in_place():
    # host slices up to 32 KiB are searched in place from a mapped pinned copy (no DMA); lengths on both
    # sides of that limit and of the 16-byte chunking, unaligned host pointers, needle at the very end
    rng = np.random.default_rng(11)
    base = rng.integers(97, 101, 70000, dtype=np.uint8)
    for n in (1, 2, 15, 16, 17, 31, 32, 33, 4095, 4096, 32767, 32768, 32769, 40000):
        for shift in (0, 1, 5):
            h = base[shift:shift + n].copy()
            hb = h.tobytes()
            for k in (1, 2, 5, 17, 40):
                if k > n:
                    continue
                nd_absent = bytes([0xEE]) * k
                s = ss.DynamicB200Searcher.new(nd_absent)
                assert s.find_in(h) is None and s.search_in(hb) is False
                h2 = h.copy()
                h2[n - k:] = 0xEE  # planted as the last k bytes
                assert s.find_in(h2) == n - k == oracle.find(h2, nd_absent), (n, shift, k)
                s.close()
            nd = hb[n // 2:n // 2 + 6]
            s = ss.DynamicB200Searcher.new(nd)
            assert s.find_in(base[shift:shift + n]) == oracle.find(hb, nd) == _expect(hb, nd), (n, shift)
            s.close()
    # back-to-back calls reuse the same buffer: a shorter slice after a longer one must not see stale bytes
    s = ss.DynamicB200Searcher.new(b"zzzz")
    assert s.find_in(b"a" * 1000 + b"zzzz") == 1000
    assert s.find_in(b"a" * 998 + b"zz") is None
    assert s.find_in(b"a" * 20) is None
    s.close()


def test_host_path_pageable_staging_pool():
    # large pageable host slices go through the pinned staging ring filled by the copy pool
    n = (150 << 20) + 1234  # several 32 MiB chunks
    host = oracle.fill_random(0, n, SEED_HAY)  # plain numpy memory: pageable
    nd = bytes([9, 0xFF, 8, 7, 6, 5, 4, 3, 2])
    s = ss.DynamicB200Searcher.new(nd)
    assert s.find_in(host) is None
    for spot in (n - 9, (96 << 20) - 4, (64 << 20) - 9, (32 << 20) - 8, (32 << 20) - 9, 77):
        host[spot:spot + 9] = np.frombuffer(nd, np.uint8)
        assert s.find_in(host) == spot
    pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[:] = host
    assert s.find_in(pinned) == 77  # pinned memory takes the direct path


def test_async_entry_base_offset_and_start_limit(variant):
    n = 3 << 20
    t = torch.empty(n, dtype=torch.uint8, device="cuda")
    ss.fill_random(t, 0, SEED_HAY)
    nd = bytes([5, 0xFF, 6])
    t[1000:1003] = torch.tensor(list(nd), dtype=torch.uint8, device="cuda")
    s = ss.DynamicB200Searcher.new(nd)
    ws = torch.zeros(16, dtype=torch.uint8, device="cuda")
    res = torch.zeros(1, dtype=torch.int64, device="cuda")
    s.find_in_async(t, res, ws, base_offset=1 << 40)
    assert int(res.item()) == (1 << 40) + 1000
    s.find_in_async(t, res, ws, base_offset=0, start_limit=1000)  # position 1000 is not owned
    assert int(res.item()) == ss.DEVICE_NONE
    s.find_in_async(t, res, ws, base_offset=0, start_limit=1001)
    assert int(res.item()) == 1000
    assert int(ws.view(torch.int64).abs().sum().item()) == 0  # workspace restored to zero
    # back-to-back launches on one stream share the workspace without a memset
    outs = torch.zeros(8, dtype=torch.int64, device="cuda")
    for i in range(8):
        s.find_in_async(t[i * 100:], outs[i:i + 1], ws, base_offset=i * 100)
    assert outs.tolist() == [1000] * 8


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_sharded_logic_on_one_gpu(world, variant):
    # shard arithmetic is rank-count agnostic: emulate `world` ranks on one device, MIN over results
    from sliceslice_rs_b200.sharded import shard_bounds

    n = (16 << 20) + 5
    t = torch.empty(n, dtype=torch.uint8, device="cuda")
    ss.fill_random(t, 0, SEED_HAY)
    host = oracle.fill_random(0, n, SEED_HAY)
    ws = torch.zeros(16, dtype=torch.uint8, device="cuda")
    for k in (1, 4, 16, 64):
        nd = bytearray(random.Random(k).randrange(255) for _ in range(k))
        nd[min(1, k - 1)] = 0xFF
        nd = bytes(nd)
        s = ss.DynamicB200Searcher.new(nd)
        per = shard_bounds(n, k, world, 0)[1]
        plants = [None, n - k, per - k // 2 if world > 1 else 77, 12345]
        for plant in plants:
            if plant is not None:
                plant = max(0, min(n - k, plant))
                t[plant:plant + k] = torch.frombuffer(bytearray(nd), dtype=torch.uint8).cuda()
                host[plant:plant + k] = np.frombuffer(nd, np.uint8)
            res = torch.full((world,), -1, dtype=torch.int64, device="cuda")
            for r in range(world):
                start, owned, span = shard_bounds(n, k, world, r)
                s.find_in_async(t[start:start + span], res[r:r + 1], ws, base_offset=start, start_limit=owned)
            got = int(res.min().item())
            exp = oracle.find(host, nd)
            assert (None if got == ss.DEVICE_NONE else got) == exp
        ss.fill_random(t, 0, SEED_HAY)
        host = oracle.fill_random(0, n, SEED_HAY)


def _count_overlapping(h: bytes, nd: bytes) -> int:
    c, i = 0, h.find(nd)
    while i >= 0:
        c += 1
        i = h.find(nd, i + 1)
    return c


def test_count_mode(i386, variant):
    # the same scan without the early return: every occurrence, overlapping ones included
    ws = torch.zeros(32, dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    big = i386 * 12  # 10 MiB: long enough for the TMA ring, matches straddle the copies' seams
    t = _dev(big)
    for nd in (b"e", b"th", b"the", b"ing ", b"segment", b"ipsum", b"descriptor table", b"  ", b"\n\n"):
        for pos in sorted({0, len(nd) - 1}):
            s = ss.DynamicB200Searcher.with_position(nd, pos)
            s.count_in_async(t, cnt, ws)
            assert int(cnt.item()) == _count_overlapping(big, nd), nd
            s.count_in_async(t, cnt, ws, start_limit=len(i386))  # only start positions of the first copy
            assert int(cnt.item()) == _count_overlapping(big[: len(i386) + len(nd) - 1], nd), nd
    a = _dev(b"a" * 100000)
    s = ss.DynamicB200Searcher.new(b"aaaa")
    s.count_in_async(a, cnt, ws)
    assert int(cnt.item()) == 100000 - 3  # overlapping occurrences
    s.count_in_async(a[:3], cnt, ws)
    assert int(cnt.item()) == 0
    with pytest.raises(ss.B200Error):
        ss.DynamicB200Searcher.new(b"").count_in_async(a, cnt, ws)


def _sampled_hist(data: bytes, sample_bytes: int) -> np.ndarray:
    """CPU restatement of the documented sampling rule of ss_b200_byte_histogram_device_async."""
    a = np.frombuffer(data, np.uint8)
    total = (len(a) + 4095) // 4096
    stride = 1
    if sample_bytes and sample_bytes < len(a):
        stride = total // ((sample_bytes + 4095) // 4096)
    idx = np.arange(0, total, stride)
    parts = [a[g * 4096:(g + 1) * 4096] for g in idx]
    return np.bincount(np.concatenate(parts) if parts else a[:0], minlength=256).astype(np.uint64)


def test_byte_histogram_and_rarest_position(i386, corpus):
    # histogram of the device-resident haystack -> rarest needle byte as second anchor
    big = i386 * 5 + i386[:12345]
    for data in (i386, big, i386[:4095], i386[:4097], b"a", b""):
        for shift in (0, 1, 7):
            t = _dev(b"\0" * shift + data)[shift:]  # unaligned device pointer
            hs = ss.DeviceHaystack.from_tensor(t)
            got = hs.byte_histogram()
            assert np.array_equal(got, np.bincount(np.frombuffer(data, np.uint8), minlength=256).astype(np.uint64))
            for sample in (4096, 100_000, 1 << 20):
                assert np.array_equal(hs.byte_histogram(sample), _sampled_hist(data, sample)), (len(data), sample)
            hs.close()
    hs = ss.DeviceHaystack.upload(i386)
    hist = hs.byte_histogram()
    checked = 0
    for nd in (b"consecteturadipi", b"segmentation", b"the", b" of the ", b"ipsum", b"descriptor table", b"80386",
               b"x", b"zq", i386[70000:70300], i386[500:540]):
        s = ss.DynamicB200Searcher.with_rarest_position(nd, hist)
        p = s.position
        if len(nd) >= 2:
            # the chosen byte is the rarest one the rule may pick (positions >= 16 cost 17/16)
            cost = lambda q: int(hist[nd[q]]) * (16 if q < 16 else 17)
            assert 1 <= p < len(nd) and cost(p) == min(cost(q) for q in range(1, len(nd)))
        assert s.find_in(hs) == oracle.find(i386, nd), nd  # results never depend on the position
        s.close()
        checked += 1
    assert checked == 11
    # stream-ordered form: 256 uint64 in device memory
    d_hist = torch.zeros(256, dtype=torch.int64, device="cuda")
    t = _dev(i386)
    rc = ss.lib().ss_b200_byte_histogram_device_async(t.data_ptr(), t.numel(), 0, d_hist.data_ptr(),
                                                      torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    assert np.array_equal(d_hist.cpu().numpy().astype(np.uint64), hist)
    hs.close()
    # the default (sample_bytes == 0) is a 16 MiB sample: exact up to that size, the documented granule rule
    # beyond it; sample_bytes >= len still counts every byte
    long = (i386 * 25)[: (20 << 20) + 999]
    hs = ss.DeviceHaystack.upload(long)
    assert np.array_equal(hs.byte_histogram(), _sampled_hist(long, 16 << 20))
    assert np.array_equal(hs.byte_histogram(len(long)),
                          np.bincount(np.frombuffer(long, np.uint8), minlength=256).astype(np.uint64))
    hs.close()


def test_many_haystack_mode_vs_oracle(sorted_words, i386, variant):
    # one needle against a device-resident SET of haystacks in one pass; per haystack == search_in()
    rng = random.Random(21)
    hays = list(sorted_words) + [b"", b"a", b"", i386[:70000], b"xyz" * 5000, b"", i386[300000:340000], b"q"]
    rng.shuffle(hays)
    hs = ss.HaystackSet(hays)                        # prepared: lookup hints (ss_b200_hayset)
    hs_plain = ss.HaystackSet(hays, prepared=False)  # hint-free ss_b200_search_many_async
    needles = [b"", b"a", b"e", b"th", b"the", b"tion", b"segment", b"ipsum", b"zq", b"xyzx", b"interrupt",
               b"descriptor table", i386[1000:1040], i386[69990:70010], i386[339980:340000], b"q"]
    for nd in needles:
        for pos in sorted({0, len(nd) // 2, max(len(nd) - 1, 0)}):
            s = ss.DynamicB200Searcher.with_position(nd, pos) if nd else ss.DynamicB200Searcher.new(nd)
            got = s.search_many_async(hs).cpu().numpy().astype(bool)
            exp = oracle.pairs([nd], hays, np.zeros(len(hays), np.uint32), np.arange(len(hays), dtype=np.uint32))
            assert np.array_equal(got, exp != oracle.NPOS), nd
            assert got.tolist() == [h.find(nd) >= 0 for h in hays]
            assert np.array_equal(s.search_many_async(hs_plain).cpu().numpy().astype(bool), got), nd
            s.close()


def test_prepared_set_hints_with_awkward_boundaries():
    # haystack boundaries on, just before and just after the 4 KiB hint granules; runs of empty
    # haystacks (equal offsets) at granule boundaries; one haystack spanning many granules
    rng = random.Random(5)
    lens = [4096, 0, 0, 4095, 1, 0, 4097, 8191, 1, 0, 0, 0, 40000, 3, 4093, 0, 12288, 5, 0]
    lens += [rng.choice([0, 1, 2, 7, 100, 4096, 5000]) for _ in range(400)]
    alphabet = b"ab"
    hays = [bytes(rng.choice(alphabet) for _ in range(n)) for n in lens]
    hs = ss.HaystackSet(hays)
    hs_plain = ss.HaystackSet(hays, prepared=False)
    for nd in (b"a", b"ab", b"abba", b"aaaaaaaa", b"abababababab", b"bbbbbbbbbbbbbbbbbbbb", b"b" * 40):
        s = ss.DynamicB200Searcher.new(nd)
        got = s.search_many_async(hs).cpu().numpy().astype(bool)
        assert got.tolist() == [h.find(nd) >= 0 for h in hays], nd
        assert np.array_equal(s.search_many_async(hs_plain).cpu().numpy().astype(bool), got), nd
        s.close()
    # an empty set and a set of empty haystacks
    assert ss.DynamicB200Searcher.new(b"a").search_many_async(ss.HaystackSet([b"", b"", b""])).cpu().tolist() == [0, 0, 0]
    assert ss.DynamicB200Searcher.new(b"").search_many_async(ss.HaystackSet([b"", b"x"])).cpu().tolist() == [1, 1]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_sharded_haystack_set_on_one_gpu(world, sorted_words):
    # emulate `world` ranks on one device: OR of the per-rank flag arrays == the global answer
    from sliceslice_rs_b200.sharded import ShardedHaystackSet

    hays = sorted_words[::3] + [b"", b"the quick brown fox jumps over the lazy dog" * 300]
    for nd in (b"the", b"ing", b"ipsum", b"o"):
        s = ss.DynamicB200Searcher.new(nd)
        acc = np.zeros(len(hays), bool)
        covered = 0
        for r in range(world):
            sh = ShardedHaystackSet(hays, rank=r, world=world)
            covered += sh.hi - sh.lo
            got = sh.search(s)
            assert not got[: sh.lo].any() and not got[sh.hi:].any()  # a rank only sets its own slice
            acc |= got
        assert covered == len(hays)
        assert acc.tolist() == [h.find(nd) >= 0 for h in hays]


def test_find_in_waits_for_pending_tensor_writes(variant):
    # find_in(tensor) scans on the library's stream: it must wait for writes still queued on torch's
    # current stream (found by tools/fuzz_gpu.py: a large copy followed at once by a search)
    n = 48 << 20
    rng = np.random.default_rng(11)
    pool = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    s = ss.DynamicB200Searcher.with_position(b"bbab", 0)
    for rep in range(6):
        h = rng.integers(0, 2, size=n, dtype=np.uint8) + 97
        src = torch.from_numpy(h).cuda()
        pool[8:8 + n] = src  # asynchronous device-to-device copy on torch's stream
        assert s.find_in(pool[8:8 + n]) == h.tobytes().find(b"bbab")


def test_const_handles_from_many_threads(i386, words):
    # reference searchers are Send + Sync (src/x86.rs:266-271)
    hs = ss.DeviceHaystack.upload(i386)
    sample = words[::40]
    searchers = [ss.DynamicB200Searcher.new(w) for w in sample]
    exp = [i386.find(w) for w in sample]
    errs = []

    def work():
        try:
            torch.cuda.set_device(0)
            for _ in range(3):
                got = [s.find_in(hs) for s in searchers]
                assert got == exp
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work) for _ in range(6)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


# ------------------------------------------------------------------------------------------
# BASELINE full sizes: size-independent properties


def test_full_size_8gib_random_properties():
    # config 4: 8 GiB generator haystack, needle lengths {1,4,16,64}; absent -> None over the whole
    # scan; planted at the very end -> offset n-k (> 2^32); generator checked against the oracle on
    # slices (head, a 2^32 straddle, tail)
    n = 8 << 30
    t = torch.empty(n, dtype=torch.uint8, device="cuda")
    ss.fill_random(t, 0, SEED_HAY)
    for a, ln in ((0, 1 << 16), ((1 << 32) - 4097, 8200), (n - 70001, 70001)):
        assert np.array_equal(t[a:a + ln].cpu().numpy(), oracle.fill_random(a, ln, SEED_HAY))
    for variant in VARIANTS:
        ss.set_scan_variant(variant)
        for k in (1, 4, 16, 64):
            nd = bytearray(random.Random(1000 + k).randrange(255) for _ in range(k))
            nd[min(1, k - 1)] = 0xFF
            nd = bytes(nd)
            s = ss.DynamicB200Searcher.new(nd)
            assert s.find_in(t) is None
            tail = t[n - k:].clone()
            t[n - k:] = torch.frombuffer(bytearray(nd), dtype=torch.uint8).cuda()
            assert s.find_in(t) == n - k
            mid = (1 << 32) - k // 2 - 1  # straddles the 32-bit offset boundary
            keep = t[mid:mid + k].clone()
            t[mid:mid + k] = torch.frombuffer(bytearray(nd), dtype=torch.uint8).cuda()
            assert s.find_in(t) == mid
            t[mid:mid + k] = keep
            t[n - k:] = tail
        # k == 1 present byte: tiny offset, equals the oracle on the head slice
        head = oracle.fill_random(0, 1 << 16, SEED_HAY)
        assert ss.DynamicB200Searcher.new(b"A").find_in(t) == oracle.find(head, b"A")
    ss.set_scan_variant(0)
    del t
    torch.cuda.empty_cache()


def test_full_size_i386_tiled_absent_needles(corpus, i386):
    # config 2': i386.txt tiled to 8 GiB; the four needles are absent from the text and the seam
    n = 8 << 30
    src = _dev(i386)
    t = torch.empty(n, dtype=torch.uint8, device="cuda")
    ss.fill_tiled(t, 0, src)
    m = len(i386)
    for a, ln in ((0, 4096), (m - 100, 300), (n - 5000, 5000), ((1 << 32) - 50, 100)):
        assert np.array_equal(t[a:a + ln].cpu().numpy(), oracle.fill_tiled(a, ln, i386))
    for nd in corpus["absent_needles"]:
        two = i386 + i386
        assert oracle.find(two, nd.encode()) is None
        for variant in VARIANTS:
            ss.set_scan_variant(variant)
            assert ss.DynamicB200Searcher.new(nd.encode()).search_in(t) is False
    # a present word is found in the first copy at the golden offset, even in an 8 GiB haystack
    ss.set_scan_variant(0)
    assert ss.DynamicB200Searcher.new(b"segmentation").find_in(t) == i386.find(b"segmentation")
    del t
    torch.cuda.empty_cache()
