"""GPU tests of round 2's boundary additions, all through the C ABI: the multi-GPU context (sharded
device-resident haystack with the three exchanges, one host slice striped over every device, the
many-haystack mode partitioned over the devices), the host-slice engine's data paths (DMA ring, in
place, pageable staging; early stop; ring sized from the slice), the peer-mailbox exchange at world 1,
the bit-packed flags, the stream-ordered batch entries, and the multi-rank torchrun check.

Every test runs with however many GPUs the box has (1 on the single-GPU tier, up to 8); the oracle is
the checker."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
import sliceslice_rs_b200 as ss
from sliceslice_rs_b200 import sharded

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED_HAY = 0x5EEDB20000000001


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: the CUDA path is the only path (no CPU fallback)")
    ss.lib()
    yield
    ss.set_scan_variant(0)
    ss.set_host_path(0, 0, -1)


def _ndev():
    return min(torch.cuda.device_count(), 8)


def _exchanges(ctx):
    out = [ss.EXCHANGE_HOST]
    if ctx.device_count > 1:
        out += [ss.EXCHANGE_PEER, ss.EXCHANGE_NCCL]
    return out


def _plants(n, k, per, ndev):
    """SURVEY 8d C5 plants, descending so that each new plant is the global leftmost: the last k bytes;
    straddling every shard boundary (found only thanks to the halo); early in the last shard; inside
    shard 0 (the MIN must pick it over later shards)."""
    spots = {n - k, min(n - k, 4242), min(n - k, (ndev - 1) * per + 99)}
    for d in range(1, ndev):
        spots.add(max(0, min(n - k, d * per - k // 2 - 1)))
        spots.add(max(0, min(n - k, d * per - 1)))
        spots.add(max(0, min(n - k, d * per)))
    return sorted(spots, reverse=True)


@pytest.mark.parametrize("ndev", sorted({1, _ndev()}))
def test_ctx_sharded_upload_and_search(ndev):
    ctx = ss.Context(ndev)
    assert ctx.device_count == ndev and ctx.devices == list(range(ndev))
    n = (24 << 20) + 12345
    host = oracle.fill_random(0, n, SEED_HAY)
    for k in (1, 4, 16, 64, 300):
        rng = np.random.default_rng(1000 + k)
        nd = bytearray(rng.integers(0, 255, k, dtype=np.uint8).tobytes())
        nd[min(1, k - 1)] = 0xFF  # absent from the generator's alphabet
        nd = bytes(nd)
        s = ss.DynamicB200Searcher.new(nd)
        h = host.copy()
        sh = ctx.upload_sharded(h, halo=512)
        assert len(sh) == n
        per = sh.shard(0)[2]
        for ex in _exchanges(ctx):
            ctx.set_exchange(ex)
            assert ctx.find_sharded(s, sh) is None and ctx.search_sharded(s, sh) is False
        sh.close()
        for spot in _plants(n, k, per, ndev):
            h[spot:spot + k] = np.frombuffer(nd, np.uint8)
            sh = ctx.upload_sharded(h, halo=512)
            for ex in _exchanges(ctx):
                ctx.set_exchange(ex)
                assert ctx.find_sharded(s, sh) == spot == oracle.find(h, nd), (k, spot, ex)
            sh.close()
        s.close()
    # trivial outcomes take the same route: empty needle => found at 0; needle longer than the haystack
    sh = ctx.upload_sharded(host[:1000], halo=64)
    assert ctx.find_sharded(ss.DynamicB200Searcher.new(b""), sh) == 0
    assert ctx.find_sharded(ss.DynamicB200Searcher.new(b"x" * 40), ctx.upload_sharded(b"x" * 39, halo=64)) is None
    if ndev > 1:
        with pytest.raises(ss.B200Error):  # the halo bounds the needle length
            ctx.find_sharded(ss.DynamicB200Searcher.new(b"y" * 200), sh)
    ctx.close()


def test_ctx_sharded_from_device_tensors():
    ndev = _ndev()
    ctx = ss.Context(ndev)
    S, k = (4 << 20), 9
    nd = bytes([7, 0xFF, 1, 2, 3, 4, 5, 6, 8])
    total = S * ndev
    tensors, owned = [], []
    for d in range(ndev):
        span = min(S + k - 1, total - d * S)
        with torch.cuda.device(d):
            t = torch.empty(span, dtype=torch.uint8, device=f"cuda:{d}")
            ss.fill_random(t, d * S, SEED_HAY)
            torch.cuda.synchronize()
        tensors.append(t)
        owned.append(S if d < ndev - 1 else span - k + 1)
    sh = ctx.sharded_from_tensors(tensors, owned)
    assert len(sh) == total
    s = ss.DynamicB200Searcher.new(nd)
    ndt = np.frombuffer(nd, np.uint8)
    for ex in _exchanges(ctx):
        ctx.set_exchange(ex)
        assert ctx.find_sharded(s, sh) is None
    for spot in _plants(total, k, S, ndev):
        for d in range(ndev):  # write the plant into every shard that holds some of its bytes (halo!)
            lo, hi = max(spot, d * S), min(spot + k, d * S + tensors[d].numel())
            if lo < hi:
                tensors[d][lo - d * S:hi - d * S] = torch.from_numpy(ndt[lo - spot:hi - spot].copy()).to(f"cuda:{d}")
        torch.cuda.synchronize()
        for ex in _exchanges(ctx):
            ctx.set_exchange(ex)
            assert ctx.find_sharded(s, sh) == spot, (spot, ex)
    ctx.close()


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_find_in_host_multi_pinned(mode):
    ctx = ss.Context(_ndev())
    ndev = ctx.device_count
    n = (40 << 20) + 333
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    ref = oracle.fill_random(0, n, SEED_HAY)
    host.numpy()[:] = ref
    nd = bytes([9, 0xFF, 8, 7, 6, 5, 4, 3, 2, 1, 0])
    k = len(nd)
    s = ss.DynamicB200Searcher.new(nd)
    for chunk_mib in (0, 1):
        ss.set_host_path(mode, chunk_mib, -1)
        assert ctx.find_in_host(s, host) is None
        st = ctx.last_host_stats()
        assert st["mode"] == (1 if mode in (0, 1) else 2), st
        if chunk_mib:
            assert st["chunks"] == -(-(n - k + 1) // (1 << 20))
    ss.set_host_path(mode, 1, -1)
    chunk = 1 << 20
    spots = sorted({n - k, 5 * chunk - 4, 3 * chunk - k, 3 * chunk - k + 1, ndev * chunk + 17, 77}, reverse=True)
    for spot in spots:
        host.numpy()[spot:spot + k] = np.frombuffer(nd, np.uint8)
        ref[spot:spot + k] = np.frombuffer(nd, np.uint8)
        got = ctx.find_in_host(s, host)
        assert got == spot == oracle.find(ref, nd), (mode, spot, got)
        # early stop: a match in an early chunk ends the feeding long before the end of the slice
        if spot < 6 * chunk:
            assert ctx.last_host_stats()["chunks"] <= spot // chunk + 1 + 3 * ndev
    # the single-device entry goes through the same engine
    assert s.find_in(host) == 77
    ss.set_host_path(0, 0, -1)
    ctx.close()


def test_find_in_host_multi_pageable_and_short():
    ctx = ss.Context(_ndev())
    n = (70 << 20) + 99
    host = oracle.fill_random(0, n, SEED_HAY)  # numpy memory: pageable => staged through the pinned ring
    nd = bytes([3, 0xFF, 3, 3, 3])
    s = ss.DynamicB200Searcher.new(nd)
    assert ctx.find_in_host(s, host) is None
    assert ctx.last_host_stats()["mode"] == 11
    for spot in (n - 5, (33 << 20) - 2, 12):
        host[spot:spot + 5] = np.frombuffer(nd, np.uint8)
        assert ctx.find_in_host(s, host) == spot
    # short slices, the trivial outcomes and every length around the 32 KiB in-place limit
    for m in (0, 1, 4, 5, 6, 4096, 32768, 32769, 100000):
        h = bytes(host[100:100 + m])
        e = h.find(nd)
        assert ctx.find_in_host(s, h) == (None if e < 0 else e)
        h2 = h[:max(0, m - 5)] + nd if m >= 5 else h
        e = h2.find(nd)
        assert ctx.find_in_host(s, h2) == (None if e < 0 else e), m
    assert ctx.find_in_host(ss.DynamicB200Searcher.new(b""), b"") == 0
    ctx.close()


def test_ring_is_sized_from_the_slice_and_released():
    ss.thread_release()
    assert ss.thread_footprint() == (0, 0)
    s = ss.DynamicB200Searcher.new(b"\xff\xfe")
    small = np.zeros(6 << 20, np.uint8)
    assert s.find_in(small) is None
    dev1, pin1 = ss.thread_footprint()
    assert 0 < dev1 <= 3 * ((6 << 20) + 4096), dev1  # one chunk covers the slice: no 3 x 64 MiB ring
    big = torch.zeros(600 << 20, dtype=torch.uint8, pin_memory=True)
    assert s.find_in(big) is None
    dev2, _ = ss.thread_footprint()
    assert dev1 < dev2 <= 3 * ((64 << 20) + 4096)
    ss.thread_release()
    assert ss.thread_footprint() == (0, 0)
    assert s.find_in(small) is None  # the next call rebuilds what it needs


def test_lane_released_at_thread_exit():
    import threading

    free0 = torch.cuda.mem_get_info()[0]
    out = {}

    def work():
        s = ss.DynamicB200Searcher.new(b"\xff\xfd")
        out["r"] = s.find_in(np.zeros(48 << 20, np.uint8))
        out["fp"] = ss.thread_footprint()

    for _ in range(3):
        t = threading.Thread(target=work)
        t.start()
        t.join()
        assert out["r"] is None and out["fp"][0] > 0
    torch.cuda.synchronize()
    # three dead threads must not hold three rings (3 x 3 x 8 MiB) any more
    assert free0 - torch.cuda.mem_get_info()[0] < (40 << 20)


def test_ctx_hayset_vs_oracle(sorted_words, i386):
    ctx = ss.Context(_ndev())
    rng = np.random.default_rng(5)
    hays = [b""] + [i386[a:a + int(l)] for a, l in zip(rng.integers(0, 800000, 3000), rng.integers(0, 6000, 3000))]
    hs = ctx.upload_haystack_set(hays)
    assert len(hs) == len(hays)
    lo = 0
    for d in range(ctx.device_count):
        a, b = hs.part(d)
        assert a == lo and b >= a
        lo = b
    assert lo == len(hays)
    for nd in (b"segment", b"the", b"ipsum", b"x", b"", b"descriptor table", i386[5000:5100]):
        s = ss.DynamicB200Searcher.new(nd)
        got = ctx.search_haystack_set(s, hs)
        exp = np.array([oracle.search_in(h, nd) for h in hays], np.uint8)
        assert np.array_equal(got, exp), nd
    ctx.close()


@pytest.mark.parametrize("variant", [1, 2])
def test_many_mode_boundary_rows_and_count_from_filter_words(i386, variant):
    """The staged variant's many-haystack step places matches with a shared-memory row of haystack
    boundaries (fallback to the plain lookup beyond 32 boundaries), and its count mode counts needles of up
    to three bytes straight from the filter words while they are frequent: unaligned blobs, runs of tiny
    and empty haystacks, haystacks longer than a tile, needles longer than the register window."""
    import random

    ss.set_scan_variant(variant)
    rng = random.Random(77)
    lens = []
    while sum(lens) < (9 << 20):
        r = rng.random()
        if r < 0.15:
            lens += [rng.choice([0, 0, 1, 2, 3, 5, 9]) for _ in range(rng.randrange(1, 120))]  # clusters of tiny ones
        elif r < 0.9:
            lens.append(rng.randrange(0, 16384))
        else:
            lens.append(rng.randrange(30000, 90000))
    text = i386 * (sum(lens) // len(i386) + 2)
    for shift in (0, 5):
        off = np.zeros(len(lens) + 1, np.int64)
        np.cumsum(lens, out=off[1:])
        blob = torch.frombuffer(bytearray(b"\0" * shift + text[:int(off[-1])]), dtype=torch.uint8).cuda()[shift:]
        hays = [text[int(a):int(b)] for a, b in zip(off[:-1], off[1:])]
        hs = ss.HaystackSet.from_device(blob, torch.from_numpy(off).cuda())
        plain = ss.HaystackSet.from_device(blob, hs.offsets, prepared=False)
        for nd in (b"the", b"e", b"segment", b"ipsum", b"descriptor table", b"the 80386 provides a", b"\n\n"):
            s = ss.DynamicB200Searcher.new(nd)
            got = s.search_many_async(hs).cpu().numpy().astype(bool)
            assert got.tolist() == [nd in h for h in hays], (nd, shift)
            assert np.array_equal(s.search_many_async(plain).cpu().numpy().astype(bool), got)
        # count mode on the same (unaligned) bytes, with start limits that end inside a tile
        ws = torch.zeros(32, dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        n = blob.numel()
        for nd in (b"e", b" ", b"th", b"he", b"the", b"ing", b"zqj", b"\xff"):
            s = ss.DynamicB200Searcher.new(nd)
            for lim in (None, n // 2 + 12345, 70001):
                s.count_in_async(blob, cnt, ws, start_limit=lim)
                m = n if lim is None else lim + len(nd) - 1
                assert int(cnt.item()) == oracle.count(text[:min(m, n)], nd), (nd, lim, shift)
    ss.set_scan_variant(0)


def test_resident_service_kernel_for_synchronous_calls(i386, words):
    """Synchronous find_in over a short device-resident haystack goes through the resident kernel (no
    launch per call): same answers as the launched kernels, across idle retirements, haystack rewrites
    between calls, alignments, needle lengths up to 64 and beyond (which fall back to a launch), more
    threads than service slots, and thread_release with a grid resident."""
    import threading
    import time

    hs = ss.DeviceHaystack.upload(i386)
    sample = words[::23] + [b"ipsum", b"x", b"\n\n", i386[1000:1064], i386[5000:5065], i386[857000:857300]]
    ss.set_sync_service(False)
    launched = [ss.DynamicB200Searcher.new(w).find_in(hs) for w in sample]
    assert launched == [(lambda e: None if e < 0 else e)(i386.find(w)) for w in sample]
    ss.set_sync_service(True, 100)
    l0 = ss.launch_count()
    searchers = [ss.DynamicB200Searcher.new(w) for w in sample]
    t_loop = time.perf_counter()
    for rep in range(3):
        assert [s.find_in(hs) for s in searchers] == launched
        for pos_rule in (0, 1):  # other second anchors
            got = [ss.DynamicB200Searcher.with_position(w, min(pos_rule, len(w) - 1)).find_in(hs) for w in sample[:40]]
            assert got == launched[:40]
        time.sleep(0.01)  # longer than the idle time: the grid retires and the next call starts a new one
    served = 3 * (len(sample) + 80)
    if (time.perf_counter() - t_loop - 0.03) / served < 50e-6:  # (under compute-sanitizer a call outlasts the idle time)
        assert ss.launch_count() - l0 < served // 4, "most synchronous calls must not launch anything"
    # the haystack changes between calls: the resident grid must see the new bytes (no stale L1 lines)
    t = torch.frombuffer(bytearray(i386[:300000]), dtype=torch.uint8).cuda()
    s = ss.DynamicB200Searcher.new(b"\x01\x02needle\x03")
    assert s.find_in(t) is None
    for spot in (299990, 150001, 77, 0):
        t[spot:spot + 9] = torch.tensor(list(b"\x01\x02needle\x03"), dtype=torch.uint8, device="cuda")
        assert s.find_in(t) == spot
    # unaligned views, tiny haystacks, needle == haystack
    for shift in (0, 1, 7, 15):
        for n in (1, 2, 9, 16, 17, 31, 33, 100, 4097):
            v = t[shift:shift + n]
            hb = bytes(v.cpu().numpy())
            for nd in (hb[:1], hb[-2:], hb[n // 2:n // 2 + 5], hb, b"\xfe\xfd"):
                if not nd:
                    continue
                e = hb.find(nd)
                assert ss.DynamicB200Searcher.new(nd).find_in(v) == (None if e < 0 else e), (shift, n, nd)
    # more threads than service slots per device: the extra ones launch kernels, answers stay the same
    errs = []

    def work(idx):
        try:
            for rep in range(20):
                for w, e in list(zip(sample, launched))[idx::8]:
                    assert ss.DynamicB200Searcher.new(w).find_in(hs) == e
            ss.thread_release()  # with a grid resident
        except Exception as ex:  # noqa: BLE001
            errs.append(repr(ex))

    th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    # implicit synchronisations (cudaFree inside close()) only wait for the idle time
    t0 = time.perf_counter()
    assert searchers[0].find_in(hs) == launched[0]
    hs2 = ss.DeviceHaystack.upload(i386[:4096])
    hs2.close()
    assert time.perf_counter() - t0 < 1.0
    ss.thread_release()
    assert searchers[1].find_in(hs) == launched[1]  # rebuilt on demand


def test_peer_exchange_world1_mailbox_kernels():
    """ss_b200_find_in_device_exchange_async with world == 1: the scan's epilogue posts into the rank's
    own mailbox and mailbox_min_kernel collects it -- the whole fused-exchange code path on one GPU."""
    px = sharded.PeerExchange()
    assert px.world == 1 and px.rank == 0
    n = 3 << 20
    t = torch.empty(n, dtype=torch.uint8, device="cuda")
    ss.fill_random(t, 0, SEED_HAY)
    ws = torch.zeros(32, dtype=torch.uint8, device="cuda")
    res = torch.zeros(1, dtype=torch.int64, device="cuda")
    nd = bytes([5, 0xFF, 6, 7])
    s = ss.DynamicB200Searcher.new(nd)
    for variant in (1, 2):
        ss.set_scan_variant(variant)
        for i in range(9):  # more searches than mailbox rows: the rows are emptied and reused
            px.find_async(s, t, 1000, n - 3, ws, res)
            assert int(res.item()) == ss.DEVICE_NONE
    t[n - 4:] = torch.tensor(list(nd), dtype=torch.uint8, device="cuda")
    px.find_async(s, t, 1000, n - 3, ws, res)
    assert int(res.item()) == 1000 + n - 4
    px.find_async(s, t, 1000, n - 4, ws, res)  # the match starts beyond start_limit: not ours
    assert int(res.item()) == ss.DEVICE_NONE
    px.find_async(ss.DynamicB200Searcher.new(b""), t, 1000, n, ws, res)  # trivial outcome posted too
    assert int(res.item()) == 1000
    ss.set_scan_variant(0)
    px.close()


def test_pack_flags_matches_the_numpy_statement():
    rng = np.random.default_rng(3)
    for total, lo, n in ((1, 0, 1), (70, 3, 40), (5000, 0, 5000), (5000, 17, 4000), (100000, 33333, 33334),
                         (100000, 64, 6400), (64, 64, 0)):
        flags = ((rng.integers(0, 3, n) == 0) * rng.integers(1, 256, n)).astype(np.uint8)
        words = torch.full(((total + 31) // 32,), -1, dtype=torch.int32, device="cuda")
        sharded.pack_flags_async(torch.from_numpy(flags).cuda(), lo, words, total)
        exp = sharded.pack_flags_reference(flags, lo, total)
        assert np.array_equal(words.cpu().numpy(), exp), (total, lo, n)
        assert np.array_equal(sharded.unpack_flags(words.cpu(), total)[lo:lo + n], (flags != 0).astype(np.uint8))


def test_sharded_haystack_set_packed_flags_one_gpu(sorted_words):
    hays = sorted_words[::3]
    for world in (1, 2, 3, 8):
        acc = np.zeros((len(hays) + 31) // 32, np.int64)
        for rank in range(world):
            hs = sharded.ShardedHaystackSet(hays, rank=rank, world=world)
            s = ss.DynamicB200Searcher.new(b"ing")
            acc += hs.search_async(s).cpu().numpy().view(np.uint32)
        got = sharded.unpack_flags(acc.astype(np.uint32), len(hays))
        assert got.tolist() == [1 if b"ing" in h else 0 for h in hays]


def test_batch_stream_ordered_entries(corpus, i386, words, sorted_words):

    L = ss.lib()
    st = torch.cuda.Stream()
    sp = st.cuda_stream
    hay = torch.frombuffer(bytearray(i386), dtype=torch.uint8).cuda()
    b = ss.Batch(words + [b""], [])
    out = torch.zeros(len(words) + 1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ss._check(L.ss_b200_batch_find_all_in_device_async(b._b, hay.data_ptr(), hay.numel(), out.data_ptr(), sp))
    st.synchronize()
    o = out.cpu().numpy()
    assert int(o[:-1].sum()) == corpus["long"]["sum_first_offsets"] and o[-1] == 0
    tri = ss.Batch(sorted_words, sorted_words)
    w = len(sorted_words)
    npairs = w * (w + 1) // 2
    bm = torch.zeros((npairs + 31) // 32, dtype=torch.int32, device="cuda")
    m = torch.zeros(1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ss._check(L.ss_b200_batch_search_triangular_async(tri._b, bm.data_ptr(), m.data_ptr(), sp))
    st.synchronize()
    assert int(m.item()) == corpus["short"]["matches"]
    ref_bm, ref_m = tri.search_triangular()
    assert np.array_equal(bm.cpu().numpy().view(np.uint32), ref_bm) and ref_m == int(m.item())
    pn = torch.arange(0, 500, dtype=torch.int32, device="cuda")
    ph = torch.arange(500, 1000, dtype=torch.int32, device="cuda")
    pbm = torch.zeros(16, dtype=torch.int32, device="cuda")
    poff = torch.zeros(500, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ss._check(L.ss_b200_batch_search_pairs_async(tri._b, pn.data_ptr(), ph.data_ptr(), 500, pbm.data_ptr(),
                                                 poff.data_ptr(), sp))
    st.synchronize()
    exp = [sorted_words[500 + i].find(sorted_words[i]) for i in range(500)]
    assert [(-1 if v < 0 else int(v)) for v in poff.cpu().numpy()] == exp


def test_upload_then_search_small_haystack_race_free():
    # a small pageable upload followed at once by a search on the library's own stream
    rng = np.random.default_rng(9)
    for i in range(300):
        n = int(rng.integers(100, 60000))
        h = rng.integers(97, 100, n, dtype=np.uint8)
        nd = bytes([0xEE]) * 70  # > 64 bytes: the needle takes the device-copy path too
        h[n - 70:] = 0xEE
        hs = ss.DeviceHaystack.upload(h)
        s = ss.DynamicB200Searcher.new(nd)
        assert s.find_in(hs) == n - 70, i
        hs.close()
        s.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2+ GPUs (runs on the multi-GPU tier)")
def test_multi_rank_sharded_and_many_modes_under_torchrun():
    """One rank per GPU over NCCL: sharded single haystack through both exchanges (NCCL all_reduce(MIN)
    and peer mailboxes) with the SURVEY 8d C5 plants, pipelined find_many, the many-haystack mode with
    bit-packed flags -- tools/check_sharded_multi_gpu.py, compared with bytes.find on every rank."""
    n = _ndev()
    import socket

    sk = socket.socket()
    sk.bind(("127.0.0.1", 0))
    port = sk.getsockname()[1]
    sk.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", "check_sharded_multi_gpu.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and f"multi-gpu check: ok (world {n})" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
