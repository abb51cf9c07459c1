"""CPU checks of bench.py's host-side helpers (no GPU, no oracle)."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_split_by_weight(bench):
    # the 8-GPU boxes of this pool: four links at ~half the bandwidth of the other four
    assert bench.split_by_weight(64, [23.3] * 4 + [35.4] * 4) == [6, 6, 6, 6, 10, 10, 10, 10]
    assert bench.split_by_weight(64, [25.3, 26.1, 25.4, 25.5, 51.3, 50.2, 50.5, 49.5]) == [5, 6, 5, 5, 11, 11, 11, 10]
    assert bench.split_by_weight(64, [55.4, 55.6]) == [32, 32]
    g = np.random.default_rng(0)
    for _ in range(200):
        n, total = int(g.integers(1, 9)), int(g.integers(0, 200))
        w = g.uniform(0.0, 60.0, n)
        s = bench.split_by_weight(total, w)
        assert sum(s) == total and min(s) >= 0 and len(s) == n
        ideal = total * np.maximum(w, 1e-9) / np.maximum(w, 1e-9).sum()
        assert np.all(np.abs(np.asarray(s) - ideal) < 1.0 + 1e-9)          # largest remainders: never off by a whole unit


def test_digest_is_order_and_content_sensitive(bench):
    a, b = np.arange(10, dtype=np.int32), np.arange(10, dtype=np.int32)[::-1].copy()
    assert bench.sha16(a) == bench.sha16(a.copy()) and len(bench.sha16(a)) == 16
    assert bench.sha16(a) != bench.sha16(b) and bench.sha16(a, b) != bench.sha16(b, a)


def test_all_pairs_order(bench):
    p = bench.all_pairs(5)
    assert [tuple(x) for x in p] == [(i, j) for i in range(5) for j in range(i + 1, 5)]      # diasss2.cpp:88-97
