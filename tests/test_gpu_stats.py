"""Statistical parity for 0 < p < 1 (north_star): per-detector / per-observable flip rates and adjacent-pair
correlations of the CUDA sampler must match the reference within 5 sigma (two-sample binomial, the method of
/root/reference/src/stim/cmd/command_sample.test.cc:39-71) over >= 10^7 shots.

Reference counts: tests/golden/stats_ref.json (2^24 shots of the unmodified reference, tools/gen_stats_ref.py)."""
import json
import os

import numpy as np
import pytest

import stim_b200
from conftest import ROOT

pytestmark = pytest.mark.gpu

with open(os.path.join(ROOT, "tests", "golden", "stats_ref.json")) as f:
    REF = json.load(f)

N_GPU = 1 << 24
CHUNK = 1 << 21


def _text(name):
    if REF[name]["circuit"] is not None:
        return REF[name]["circuit"]
    with open(os.path.join(ROOT, "tests", "golden", "circuits", name + ".stim")) as f:
        return f.read()


def _check(k_gpu, n_gpu, k_ref, n_ref, what):
    k_gpu, k_ref = np.asarray(k_gpu, dtype=np.float64), np.asarray(k_ref, dtype=np.float64)
    p = (k_gpu + k_ref) / (n_gpu + n_ref)
    sigma = np.sqrt(np.maximum(p * (1 - p), 1e-12) * (1.0 / n_gpu + 1.0 / n_ref))
    z = (k_gpu / n_gpu - k_ref / n_ref) / sigma
    worst = int(np.argmax(np.abs(z)))
    assert np.all(np.abs(z) <= 5.0), f"{what}: column {worst} deviates {z[worst]:.2f} sigma ({k_gpu[worst]} vs {k_ref[worst]})"
    return z


@pytest.mark.parametrize("name", sorted(REF))
def test_rates_and_pair_correlations_match_reference(name):
    ref = REF[name]
    text = _text(name)
    circ = stim_b200.Circuit(text)
    n_bits = ref["n_bits"]
    if ref["mode"] == "detect":
        sampler = circ.compile_detector_sampler(seed=20251017)
        draw = lambda n: sampler.sample(n, bit_packed=True, append_observables=True)  # noqa: E731
    else:
        sampler = circ.compile_sampler(seed=20251017)  # default reference sample (e.g. MPAD 1 records a constant 1)
        draw = lambda n: sampler.sample(n, bit_packed=True)  # noqa: E731
    single = np.zeros(n_bits, dtype=np.int64)
    pair = np.zeros(n_bits - 1, dtype=np.int64)
    for _ in range(N_GPU // CHUNK):
        packed = draw(CHUNK)
        bits = np.unpackbits(packed, axis=1, bitorder="little")[:, :n_bits]
        single += bits.sum(axis=0, dtype=np.int64)
        pair += (bits[:, :-1] & bits[:, 1:]).sum(axis=0, dtype=np.int64)
    z1 = _check(single, N_GPU, ref["single"], ref["n_ref"], name + " rates")
    z2 = _check(pair, N_GPU, ref["pair"], ref["n_ref"], name + " pair correlations")
    # the z-scores themselves should look standard normal, not merely bounded
    assert np.sqrt(np.mean(z1**2)) < 1.6 and np.sqrt(np.mean(z2**2)) < 1.6


def test_flip_count_api_matches_sampled_bits():
    """gstim_detector_flip_counts (the quantity an NCCL allreduce would sum) == popcounts of the sampled bits."""
    text = _text("c2_surface_x_d5_r5")
    circ = stim_b200.Circuit(text)
    a = circ.compile_detector_sampler(seed=77)
    b = circ.compile_detector_sampler(seed=77)
    shots = 100_000
    counts = a.flip_counts(shots)
    bits = np.unpackbits(b.sample(shots, bit_packed=True, append_observables=True), axis=1, bitorder="little")
    np.testing.assert_array_equal(counts.astype(np.int64), bits[:, : counts.size].sum(axis=0, dtype=np.int64))
