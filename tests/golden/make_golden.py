#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

Writes
  kats.json    the reference's own known-answer vectors for the path, transcribed
               from its test sources (inputs + the literal expected bool the
               reference asserts), plus the leftmost offset from two independent
               voices: Python's bytes.find and -- where it applies (k >= 2) -- the
               reference's vendored avx2_strstr_v2 compiled from /root/reference
               (oracle/_ref).
  corpus.json  golden numbers for the two corpus sweeps of tests/i386.rs and
               bench/benches/i386.rs over data/words.txt and data/i386.txt:
               per-needle first offsets, their sum, the short-sweep match count and
               a sha256 of its result bitmap.

The CUDA path and the C oracle are NOT used to produce expected values here; the
voices are CPython's bytes.find and the reference's own C++ code.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

LOREM55 = b"Lorem ipsum dolor sit amet, consectetur adipiscing elit"
LOREM187 = (b"Lorem ipsum dolor sit amet, consectetur adipiscing elit. Maecenas commodo posuere orci a consectetur. "
            b"Ut mattis turpis ut auctor consequat. Aliquam iaculis fringilla mi, nec aliquet purus")
FOO74 = b"foo bar baz qux quux quuz corge grault garply waldo fred plugh xyzzy thud"

# (group, haystack, needle, expected bool asserted by the reference) -- src/lib.rs:422-544
KATS = [
    # search_same  src/lib.rs:422-438
    ("same", b"x", b"x", True),
    ("same", b"xy", b"xy", True),
    ("same", b"foo", b"foo", True),
    ("same", LOREM55, LOREM55, True),
    ("same", LOREM187, LOREM187, True),
    # search_different  src/lib.rs:440-461
    ("different", b"x", b"y", False),
    ("different", b"xy", b"xz", False),
    ("different", b"bar", b"foo", False),
    ("different", LOREM55, b"foo", False),
    ("different", LOREM187, b"foo", False),
    ("different", LOREM187, FOO74, False),
    # search_prefix  src/lib.rs:463-482
    ("prefix", b"xy", b"x", True),
    ("prefix", b"foobar", b"foo", True),
    ("prefix", LOREM55, b"Lorem", True),
    ("prefix", LOREM187, b"Lorem", True),
    ("prefix", LOREM187, LOREM55, True),
    # search_suffix  src/lib.rs:484-503
    ("suffix", b"xy", b"y", True),
    ("suffix", b"foobar", b"bar", True),
    ("suffix", LOREM55, b"elit", True),
    ("suffix", LOREM187, b"purus", True),
    ("suffix", LOREM187, b"Aliquam iaculis fringilla mi, nec aliquet purus", True),
    # search_multiple  src/lib.rs:505-521
    ("multiple", b"xx", b"x", True),
    ("multiple", b"xyxy", b"xy", True),
    ("multiple", b"foobarfoo", b"foo", True),
    ("multiple", LOREM55, b"it", True),
    ("multiple", LOREM187, b"conse", True),
    # search_middle  src/lib.rs:523-544
    ("middle", b"xyz", b"y", True),
    ("middle", b"wxyz", b"xy", True),
    ("middle", b"foobarfoo", b"bar", True),
    ("middle", LOREM55, b"consectetur", True),
    ("middle", LOREM187, b"orci", True),
    ("middle", LOREM187, b"Maecenas commodo posuere orci a consectetur", True),
]

# MemchrSearcher KATs  src/lib.rs:303-331  (haystack, needle byte, expected)
MEMCHR_KATS = [
    (b"f", b"f", True),
    (b"foo", b"b", False),
    (b"foobar", b"f", True),
    (b"foobar", b"r", True),
    (b"foobarfoo", b"o", True),
    (b"foobarfoo", b"b", True),
]

# doctest  src/x86.rs:6-14 and README.md:16-25
DOCTEST = [
    (LOREM55, b"ipsum", True),
    (b"foo bar baz qux quux quuz corge grault garply waldo fred", b"ipsum", False),
]

# constructor contract  src/x86.rs:533-565 (+ :470-475 for the dynamic N0/N1 arms)
CTOR = [
    # (searcher, needle, position or null for new(), "ok" | "panic")
    ("avx2", "foo", 3, "panic"),
    ("dynamic", "foo", 3, "panic"),
    ("avx2", "", None, "panic"),
    ("dynamic", "", None, "ok"),
    ("dynamic", "", 7, "ok"),
    ("dynamic", "f", 0, "ok"),
    ("dynamic", "f", 1, "panic"),
    ("dynamic", "foo", 2, "ok"),
    ("avx2", "foo", 0, "ok"),
]


def py_find(h, n):
    r = h.find(n)
    return None if r < 0 else r


def main():
    import oracle

    oracle.build()
    have_ref = oracle.ref_lib() is not None
    print("oracle/_ref available:", have_ref)

    def voices(h, n):
        off = py_find(h, n)
        if have_ref and len(n) >= 2 and len(h) >= len(n):
            r = oracle.ref_find(h, n)
            assert r == off, (h, n, r, off)
        return off

    kats = []
    for group, h, n, exp in KATS:
        off = voices(h, n)
        assert (off is not None) == exp
        kats.append({"group": group, "haystack": h.decode(), "needle": n.decode(), "found": exp, "offset": off})
    mem = []
    for h, n, exp in MEMCHR_KATS:
        off = py_find(h, n)
        assert (off is not None) == exp
        mem.append({"haystack": h.decode(), "needle": n.decode(), "found": exp, "offset": off})
    doc = []
    for h, n, exp in DOCTEST:
        off = voices(h, n)
        assert (off is not None) == exp
        doc.append({"haystack": h.decode(), "needle": n.decode(), "found": exp, "offset": off})
    ctor = [{"searcher": s, "needle": n, "position": p, "outcome": o} for s, n, p, o in CTOR]
    with open(os.path.join(HERE, "kats.json"), "w") as f:
        json.dump({"source": "cloudflare/sliceslice-rs src/lib.rs:303-331,422-544; src/x86.rs:6-14,533-565",
                   "ref_voice_checked": have_ref, "kats": kats, "memchr": mem, "doctest": doc, "ctor": ctor},
                  f, indent=1)

    # ---- corpus sweeps ------------------------------------------------------------
    i386 = open(os.path.join(ROOT, "data", "i386.txt"), "rb").read()
    words = [w for w in open(os.path.join(ROOT, "data", "words.txt"), "rb").read().split(b"\n") if w]
    offs = []
    for w in words:
        o = py_find(i386, w)
        if have_ref and len(w) >= 2:
            assert oracle.ref_find(i386, w) == o
        offs.append(-1 if o is None else o)
    n = len(i386)
    found = sum(o >= 0 for o in offs)
    sum_off = sum(o for o in offs if o >= 0)
    examined = sum(min(o + len(w), n) if o >= 0 else n for o, w in zip(offs, words))

    # short sweep: stable sort by (len, file order); pairs (i, j >= i)
    order = sorted(range(len(words)), key=lambda i: (len(words[i]), i))
    sw = [words[i] for i in order]
    W = len(sw)
    npairs = W * (W + 1) // 2
    bm = np.zeros((npairs + 31) // 32, np.uint32)
    p = 0
    matches = 0
    hay_bytes = 0
    for i in range(W):
        nd = sw[i]
        for j in range(i, W):
            hay_bytes += len(sw[j])
            if nd in sw[j]:
                matches += 1
                bm[p >> 5] |= np.uint32(1 << (p & 31))
            p += 1
    lossy = i386.decode("utf-8", errors="replace").encode("utf-8")  # String::from_utf8_lossy, tests/i386.rs:63
    lossy_found = sum(1 for w in words if w in lossy)
    ipsum_cands = sum(1 for i in range(n - 4) if i386[i] == ord("i") and i386[i + 4] == ord("m"))
    corpus = {
        "i386_sha256": hashlib.sha256(i386).hexdigest(), "i386_len": n,
        "words_sha256": hashlib.sha256(open(os.path.join(ROOT, "data", "words.txt"), "rb").read()).hexdigest(),
        "n_words": len(words),
        "long": {"found": found, "sum_first_offsets": sum_off, "max_offset": max(offs),
                 "examined_bytes": examined, "nominal_bytes": len(words) * n, "first_offsets": offs},
        "long_lossy_utf8": {"len": len(lossy), "found": lossy_found},
        "ipsum": {"found": py_find(i386, b"ipsum") is not None, "filter_candidates": ipsum_cands},
        "absent_needles": {nd.decode(): (py_find(i386, nd) is None and py_find(i386 + i386, nd) is None)
                           for nd in (b"ipsum", b"zq", b"ipsumdol", b"consecteturadipi")},
        "short": {"pairs": npairs, "haystack_bytes": hay_bytes, "matches": matches,
                  "bitmap_sha256": hashlib.sha256(bm.tobytes()).hexdigest()},
        "ref_voice_checked": have_ref,
    }
    with open(os.path.join(HERE, "corpus.json"), "w") as f:
        json.dump(corpus, f)
    print({k: v for k, v in corpus["long"].items() if k != "first_offsets"})
    print(corpus["short"], corpus["ipsum"], corpus["absent_needles"], corpus["long_lossy_utf8"])


if __name__ == "__main__":
    main()
