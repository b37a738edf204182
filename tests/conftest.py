"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(ROOT, "tests", "golden", "kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def corpus():
    with open(os.path.join(ROOT, "tests", "golden", "corpus.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def i386():
    with open(os.path.join(ROOT, "data", "i386.txt"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def words():
    with open(os.path.join(ROOT, "data", "words.txt"), "rb") as f:
        return [w for w in f.read().split(b"\n") if w]


@pytest.fixture(scope="session")
def sorted_words(words):
    # bench/benches/i386.rs:21 sorts by length (unstable); we fix the order as (len, file order)
    return [words[i] for i in sorted(range(len(words)), key=lambda i: (len(words[i]), i))]
