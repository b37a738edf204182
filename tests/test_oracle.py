"""CPU tests: pin the oracle (C restatement of DynamicAvx2Searcher) against the reference's
own golden vectors (tests/golden/kats.json, corpus.json) and against independent voices."""
import hashlib
import os
import random

import numpy as np
import pytest

import oracle


def test_kats_every_position(kats):
    # src/lib.rs:365-381: for every position in 0..needle.len() the result equals the naive one,
    # and the literal expectation asserted by the reference holds.
    for kat in kats["kats"] + kats["doctest"] + kats["memchr"]:
        h, n = kat["haystack"].encode(), kat["needle"].encode()
        assert oracle.naive_find(h, n) == kat["offset"]
        for position in range(len(n)):
            if len(n) == 1 and position != 0:
                continue
            got = oracle.find(h, n, position)
            assert got == kat["offset"], (h, n, position)
            assert (got is not None) == kat["found"]
            if len(n) >= 1:
                assert oracle.find(h, n, position, dynamic=False) == kat["offset"]


def test_ctor_contract(kats):
    # src/x86.rs:533-565 (+ :470-475)
    for c in kats["ctor"]:
        n = c["needle"].encode()
        dyn = c["searcher"] == "dynamic"
        if c["outcome"] == "panic":
            with pytest.raises(oracle.OracleError):
                oracle.find(b"foobar", n, c["position"], dynamic=dyn)
        else:
            oracle.find(b"foobar", n, c["position"], dynamic=dyn)


def test_empty_needle_and_memchr_edges():
    assert oracle.find(b"", b"") == 0  # N0 => true even on an empty haystack (src/x86.rs:500)
    assert oracle.find(b"abc", b"") == 0
    assert oracle.find(b"", b"a") is None  # src/lib.rs:131-133
    assert oracle.find(b"", b"ab") is None
    assert oracle.find(b"a", b"ab") is None


def test_long_sweep_matches_golden(corpus, i386, words):
    assert hashlib.sha256(i386).hexdigest() == corpus["i386_sha256"]
    got = oracle.long_sweep(words, i386, naive=False, threads=4)
    exp = np.array([o if o >= 0 else oracle.NPOS for o in corpus["long"]["first_offsets"]], dtype=np.uint64)
    assert np.array_equal(got, exp)
    assert int(got.sum()) == corpus["long"]["sum_first_offsets"] == 809985317
    assert corpus["long"]["found"] == 4585
    # examined bytes = sum(min(off + k, n))
    ex = sum(min(int(o) + len(w), len(i386)) for o, w in zip(got, words))
    assert ex == corpus["long"]["examined_bytes"] == 810016020


def test_ipsum_absent_and_candidates(corpus, i386):
    assert oracle.find(i386, b"ipsum") is None
    assert oracle.count_candidates(i386, b"ipsum") == corpus["ipsum"]["filter_candidates"] == 242
    for nd in corpus["absent_needles"]:
        assert oracle.find(i386, nd.encode()) is None
        assert oracle.find(i386 + i386, nd.encode()) is None  # absent across the tiling seam too


def test_short_sweep_matches_golden(corpus, sorted_words):
    m, bm = oracle.short_sweep(sorted_words, naive=False)
    assert m == corpus["short"]["matches"] == 39105
    assert hashlib.sha256(bm.tobytes()).hexdigest() == corpus["short"]["bitmap_sha256"]
    assert int(sum(bin(int(x)).count("1") for x in bm[bm != 0])) == m


def test_vendored_reference_voice(i386, words):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    rng = random.Random(7)
    for w in rng.sample(words, 300):
        if len(w) >= 2:
            assert oracle.ref_find(i386, w) == oracle.find(i386, w)


def test_mula_dollar_hash_offsets():
    # offset-pinning pattern of the vendored unittests.cpp:13-51: needle "$x..x#" planted
    # between paddings; the found offset must equal the left padding length.
    for size in list(range(1, 40)) + [63]:
        needle = b"$" + b"x" * size + b"#"
        for pre in (0, 1, 2, 15, 16, 17, 31, 32, 33, 47):
            for post in (0, 1, 31, 32, 47):
                h = b"_" * pre + needle + b"_" * post
                for position in (0, len(needle) // 2, len(needle) - 1):
                    assert oracle.find(h, needle, position) == pre


@pytest.mark.parametrize("alphabet", [2, 3, 26])
def test_randomized_differential(alphabet):
    # restatement == naive leftmost for every needle length / position / haystack length;
    # small alphabets make candidates and overlaps frequent
    rng = random.Random(1234 + alphabet)
    for _ in range(6000):
        n = rng.randrange(0, 120)
        k = rng.randrange(0, 22)
        h = bytes(rng.randrange(alphabet) + 97 for _ in range(n))
        if k and n >= k and rng.random() < 0.4:
            s = rng.randrange(0, n - k + 1)
            nd = h[s:s + k]
        else:
            nd = bytes(rng.randrange(alphabet) + 97 for _ in range(k))
        pos = 0 if k <= 1 else rng.randrange(k)
        exp = h.find(nd)
        assert oracle.find(h, nd, pos) == (None if exp < 0 else exp), (h, nd, pos)


def test_multithreaded_driver_equals_single():
    rng = np.random.default_rng(5)
    h = rng.integers(97, 100, size=300_000, dtype=np.uint8)
    for k in (2, 5, 17):
        nd = bytes(h[250_000:250_000 + k]) if k < 17 else b"z" * k
        assert oracle.find(h, nd, threads=8) == oracle.find(h, nd)


def test_generators_are_deterministic():
    a = oracle.fill_random(0, 4096, 0x5EEDB20000000001)
    b = oracle.fill_random(1000, 100, 0x5EEDB20000000001)
    assert np.array_equal(a[1000:1100], b)
    assert 0xFF not in a
    t = oracle.fill_tiled(5, 20, b"abcdefg")
    assert bytes(t) == (b"abcdefg" * 5)[5:25]


def test_random_bench_size_grid():
    # bench/benches/random.rs:12-99 workload: only needle[..1] in haystack[..100] / [..1000] match
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    haystack = open(os.path.join(root, "data", "haystack"), "rb").read()
    needle = open(os.path.join(root, "data", "needle"), "rb").read()
    sizes = [1, 5, 10, 20, 50, 100, 1000]
    found = []
    for i, ns in enumerate(sizes):
        for hsz in sizes[i:]:
            r = oracle.find(haystack[:hsz], needle[:ns])
            assert r == (None if haystack[:hsz].find(needle[:ns]) < 0 else haystack[:hsz].find(needle[:ns]))
            if r is not None:
                found.append((ns, hsz))
    assert found == [(1, 100), (1, 1000)]


def test_count_restarts_the_reference_search_after_every_match(i386):
    # checker of the count mode: overlapping occurrences included; compared with a naive count
    import random

    def naive(h, nd):
        c, i = 0, h.find(nd)
        while i >= 0:
            c += 1
            i = h.find(nd, i + 1)
        return c

    assert oracle.count(b"aaaaa", b"aa") == 4
    assert oracle.count(b"abcabc", b"abcd") == 0 and oracle.count(b"", b"a") == 0
    for nd in (b"e", b"th", b"segment", b"ipsum", b"descriptor table", b"  "):
        assert oracle.count(i386[:200000], nd) == naive(i386[:200000], nd), nd
    rng = random.Random(9)
    for _ in range(200):
        h = bytes(rng.choice(b"ab") for _ in range(rng.randrange(0, 300)))
        nd = bytes(rng.choice(b"ab") for _ in range(rng.randrange(1, 6)))
        assert oracle.count(h, nd) == naive(h, nd)
