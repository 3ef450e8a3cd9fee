"""Pins the oracle's random primitives: Philox4x32-10 against the Random123 known-answer vectors and the
fixed-point exponential draw against libm."""
import math
import random

import numpy as np

from oracle import philox as px


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds (counter words, key words -> output words)
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = px.philox4x32_10(*ctr, *key)  # Python-integer path (single counter)
        assert tuple(int(v) for v in got) == want
        got = px.philox4x32_10(*(np.array([c, c], dtype=np.uint64) for c in ctr), *key)  # array path
        assert tuple(int(v[1]) for v in got) == want


def test_philox_is_vectorised_consistently():
    c0 = np.arange(7, dtype=np.uint64)
    a = px.philox4x32_10(c0, 5, 9, 11, 123, 456)
    for i in range(7):
        b = px.philox4x32_10(i, 5, 9, 11, 123, 456)
        assert [int(x[i]) for x in a] == [int(x) for x in b]


def test_exp_draw_fx_matches_log():
    rng = random.Random(5)
    worst = 0.0
    for r in [0, 1, 2, 3, 255, 256, 1 << 24, 1 << 31, (1 << 32) - 1] + [rng.getrandbits(32) for _ in range(20000)]:
        got = px.exp_draw_fx(r) / 2.0**56
        want = -math.log((r + 0.5) / 2.0**32)
        worst = max(worst, abs(got - want))
    assert worst < 3e-6  # table interpolation error bound


def test_exp_draw_fx_is_exponential():
    rng = np.random.default_rng(1)
    r = rng.integers(0, 1 << 32, size=200000, dtype=np.uint64)
    e = np.array([px.exp_draw_fx(int(v)) for v in r[:50000]], dtype=np.float64) / 2.0**56
    assert abs(e.mean() - 1.0) < 0.02
    assert abs((e > 1.0).mean() - math.exp(-1)) < 0.01


def test_lam_fx():
    assert px.lam_fx(0.0) == 0
    assert px.lam_fx(1.0) == px.LAM_MAX
    v = px.lam_fx(1e-3)
    assert abs(v / 2.0**56 - (-math.log1p(-float(np.float32(1e-3))))) < 1e-15


def test_exp_draw_q26_matches_log():
    rng = random.Random(5)
    worst = 0.0
    for r in [0, 1, 2, 3, 255, 256, 1 << 24, 1 << 31, (1 << 32) - 1] + [rng.getrandbits(32) for _ in range(20000)]:
        got = px.exp_draw_q26(r) / 2.0**26
        want = -math.log((r | 1) / 2.0**32)
        worst = max(worst, abs(got - want))
        assert 0 <= px.exp_draw_q26(r) < 1 << 31
    assert worst < 3e-6  # table interpolation error bound


def test_rate_and_gap_arithmetic_of_the_dem_sampler():
    assert px.rate_of(0.0) is None and px.rate_of(-1.0) is None and px.rate_of(1e-30) is None
    assert px.rate_of(1.0) == (0, 0) and px.gap_of(12345, (0, 0)) == 0  # p = 1: an event at every shot
    rng = random.Random(7)
    for p in (1e-9, 1e-6, 1e-3, 0.02, 0.3, 0.5, 0.999, float(np.float32(1) - np.float32(2.0**-24))):
        inv, sh = px.rate_of(p)
        assert (1 << 31) <= inv < (1 << 32) and 0 <= sh <= 62
        lam = -math.log1p(-float(np.float32(p)))
        assert abs(inv / 2.0 ** (sh - 26) * lam - 1.0) < 1e-9  # INV * 2^(26 - SH) = 1 / lambda
        for _ in range(2000):
            w = rng.getrandbits(32)
            exact = px.exp_draw_q26(w) / 2.0**26 / lam
            assert abs(px.gap_of(w, (inv, sh)) - exact) <= 1.0 + 1e-6 * exact


def test_q26_gaps_are_geometric():
    """floor(Exp(1) / lambda) with lambda = -log1p(-p) is Geometric(p): P(gap = k) = p (1 - p)^k."""
    p = 0.1
    rate = px.rate_of(p)
    rng = np.random.default_rng(3)
    n = 200000
    gaps = np.array([px.gap_of(int(w), rate) for w in rng.integers(0, 1 << 32, size=n, dtype=np.uint64)])
    pf = float(np.float32(p))
    for k in (0, 1, 2, 5, 10, 20):
        want = pf * (1 - pf) ** k
        got = float((gaps == k).mean())
        assert abs(got - want) < 5 * math.sqrt(want * (1 - want) / n), (k, got, want)
    assert abs(gaps.mean() - (1 - pf) / pf) < 5 * math.sqrt((1 - pf) / pf**2 / n)
