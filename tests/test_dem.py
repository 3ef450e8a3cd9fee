"""Detector-error-model sampler (reference: src/stim/simulators/dem_sampler.inl:52-130,
src/stim/cmd/command_sample_dem.cc).

CPU: the oracle (oracle/dem_oracle.py) reproduces the reference's `stim sample_dem` outputs on deterministic models
(tests/golden/dem_cases.json, tools/gen_dem_golden.py); the library's parser agrees with it and raises ValueError on bad
models. GPU: the CUDA sampler equals the oracle bit for bit on noisy models, reproduces the reference's bytes in every
result format on the deterministic ones, and matches the reference's per-detector / pair / observable flip statistics of
the c2 and c3 (d=25, 356 321 mechanisms) models within 5 sigma over 2^24 shots."""
import base64
import gzip
import json
import os

import numpy as np
import pytest

import stim_b200
from conftest import ROOT
from oracle import dem_oracle as do

with open(os.path.join(ROOT, "tests", "golden", "dem_cases.json")) as f:
    CASES = json.load(f)
IDS = [c["name"] for c in CASES]

NOISY = """
error(0.125) D0 D1
error(0.03) D1 D2 ^ D3 L0
repeat 4 {
    error(0.25) D0 L1
    error(0.5) D1 D2
    error(0.002) D0 D3
    shift_detectors 2
}
error(1) D0
error(0.9) D1 L0 L2
detector D5
"""


def _bits_01(raw, n_bits):
    lines = raw.decode().split("\n")[:-1]
    return np.array([[int(ch) for ch in ln] for ln in lines], dtype=np.uint8).reshape(len(lines), n_bits)


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_reproduces_reference_sample_dem(case):
    out = case["outputs"]["01"]
    D, L, errors = do.parse_dem(case["dem"])
    m = stim_b200.DetectorErrorModel(case["dem"])
    assert (m.num_detectors, m.num_observables, m.num_errors) == (D, L, len(errors))
    dets, obs, errs = do.sample(case["dem"], out["shots"], seed=3, K=32)
    np.testing.assert_array_equal(dets, _bits_01(base64.b64decode(out["det"]), D))
    np.testing.assert_array_equal(obs, _bits_01(base64.b64decode(out["obs"]), L))
    np.testing.assert_array_equal(errs, _bits_01(base64.b64decode(out["err"]), len(errors)))


@pytest.mark.parametrize("text", ["error D0", "error(1.5) D0", "error(0.1) Q3", "repeat 2 {\nerror(0.1) D0\n", "}", "bogus(1) D0",
                                  "shift_detectors", "repeat 0 {\n}"])
def test_bad_models_raise_value_error(text):
    with pytest.raises(ValueError):
        stim_b200.DetectorErrorModel(text)


def test_fixture_models_have_the_reference_sizes():
    with open(os.path.join(ROOT, "tests", "golden", "dem", "c2_surface_x_d5_r5.dem")) as f:
        c2 = stim_b200.DetectorErrorModel(f.read())
    assert (c2.num_detectors, c2.num_observables) == (120, 1)
    c3 = stim_b200.DetectorErrorModel(gzip.open(os.path.join(ROOT, "tests", "golden", "dem", "c3_surface_z_d25_r25.dem.gz"), "rt").read())
    assert (c3.num_detectors, c3.num_observables, c3.num_errors) == (15600, 1, 356321)


# ---------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("shots", [1, 300, 4096 + 77])
def test_cuda_dem_sampler_equals_oracle(shots):
    s = stim_b200.DetectorErrorModel(NOISY).compile_sampler(seed=99)
    dets, obs, errs = s.sample(shots, return_errors=True)
    wd, wo, we = do.sample(NOISY, shots, 99, 32)
    np.testing.assert_array_equal(dets.astype(np.uint8), wd)
    np.testing.assert_array_equal(obs.astype(np.uint8), wo)
    np.testing.assert_array_equal(errs.astype(np.uint8), we)
    # the stream continues across calls, bit-packed output describes the same shots
    if shots % 128 == 0 or shots > 4096:
        return
    s2 = stim_b200.DetectorErrorModel(NOISY).compile_sampler(seed=99)
    pd, po, pe = s2.sample(shots, bit_packed=True, return_errors=True)
    np.testing.assert_array_equal(np.unpackbits(pd, axis=1, bitorder="little")[:, : wd.shape[1]], wd)
    np.testing.assert_array_equal(np.unpackbits(po, axis=1, bitorder="little")[:, : wo.shape[1]], wo)
    np.testing.assert_array_equal(np.unpackbits(pe, axis=1, bitorder="little")[:, : we.shape[1]], we)
    assert s2.sample(5)[2] is None


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_cuda_dem_sampler_reproduces_reference_bytes(case, tmp_path):
    m = stim_b200.DetectorErrorModel(case["dem"])
    for fmt, out in case["outputs"].items():
        paths = {k: tmp_path / f"{k}.{fmt}" for k in ("det", "obs", "err")}
        m.compile_sampler(seed=4).sample_write(out["shots"], det_out_file=paths["det"], det_out_format=fmt, obs_out_file=paths["obs"],
                                              obs_out_format=fmt, err_out_file=paths["err"], err_out_format=fmt)
        for k in ("det", "obs", "err"):
            assert paths[k].read_bytes() == base64.b64decode(out[k]), (case["name"], fmt, k)
    with pytest.raises(ValueError):
        m.compile_sampler(seed=4).sample_write(5, det_out_file=tmp_path / "x", det_out_format="ptb64")


@pytest.mark.gpu
def test_command_line_mirror_of_sample_dem(tmp_path):
    import subprocess
    import sys

    case = CASES[2]
    src = tmp_path / "m.dem"
    src.write_text(case["dem"])
    out = case["outputs"]["dets"]
    r = subprocess.run([sys.executable, "-m", "stim_b200", "sample_dem", "--shots", str(out["shots"]), "--in", str(src), "--out",
                        str(tmp_path / "d"), "--out_format", "dets", "--obs_out", str(tmp_path / "o"), "--obs_out_format", "dets"],
                       env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert (tmp_path / "d").read_bytes() == base64.b64decode(out["det"])
    assert (tmp_path / "o").read_bytes() == base64.b64decode(out["obs"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,fixture", [("dem_c2_surface_x_d5_r5", "c2_surface_x_d5_r5.dem"),
                                          ("dem_c3_surface_z_d25_r25", "c3_surface_z_d25_r25.dem.gz")])
def test_dem_statistics_match_reference(name, fixture):
    from test_gpu_stats_big import check

    ref = np.load(os.path.join(ROOT, "tests", "golden", "stats_big", name + ".npz"))
    src = os.path.join(ROOT, "tests", "golden", "dem", fixture)
    text = gzip.open(src, "rt").read() if src.endswith(".gz") else open(src).read()
    s = stim_b200.DetectorErrorModel(text).compile_sampler(seed=20261018)
    n = 1 << 24
    single, pair = s.bit_counts(n)
    D, L, n_ref = int(ref["D"]), int(ref["L"]), int(ref["n_ref"])
    assert single.size == D + L
    check(single[:D], n, ref["single"], n_ref, name + " detector rates")
    check(pair[: D - 1], n, ref["pair"], n_ref, name + " adjacent-pair correlations")
    check(single[D:], n, ref["obs"], n_ref, name + " observable rates")


@pytest.mark.gpu
@pytest.mark.parametrize("shots", [1, 300, 4096 + 77])
def test_cuda_dem_sampler_on_the_event_engine_equals_oracle(shots):
    """Without return_errors the model is sampled by the event engine: its table must be the model itself (one site per
    mechanism, response = targets) and the shots must equal oracle/sparse_oracle.sample on that table."""
    from oracle import sparse_oracle as so

    s = stim_b200.DetectorErrorModel(NOISY).compile_sampler(seed=7)
    D, L, errors = do.parse_dem(NOISY)
    t = s.response_table()
    site = 0
    for c in t["classes"]:
        for k in range(int(c[21])):
            e = int(t["site_group"][site])
            p, tg = errors[e]
            want = sorted(x if not isinstance(x, tuple) else D + x[1] for x in tg)
            want = [v for v in set(want) if want.count(v) % 2]
            assert so.entry_ids(t, int(c[22]) + k) == sorted(want), e
            site += 1
    assert site == sum(1 for p, _ in errors if p > 0)
    dets, obs, errs = s.sample(shots)
    assert errs is None
    rows = so.sample(t, t["slices"], t["tile_shots"], 7, 0, shots, D + L)
    np.testing.assert_array_equal(dets.astype(np.uint8), rows[:, :D])
    np.testing.assert_array_equal(obs.astype(np.uint8), rows[:, D:])
    # packed output and the stream continuing across calls
    s2 = stim_b200.DetectorErrorModel(NOISY).compile_sampler(seed=7)
    pd, po, _ = s2.sample(shots, bit_packed=True)
    np.testing.assert_array_equal(np.unpackbits(pd, axis=1, bitorder="little")[:, :D], rows[:, :D])
    np.testing.assert_array_equal(np.unpackbits(po, axis=1, bitorder="little")[:, :L], rows[:, D:])
    more = s2.sample(200)[0]
    first = (shots + 127) // 128 * 128
    np.testing.assert_array_equal(more.astype(np.uint8), so.sample(t, t["slices"], t["tile_shots"], 7, first, 200, D + L)[:, :D])


REPLAY = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "dem_replay_cases.json")))


@pytest.mark.gpu
@pytest.mark.parametrize("case", REPLAY, ids=[c["name"] for c in REPLAY])
def test_replaying_recorded_errors_reproduces_the_reference(case, tmp_path):
    """sample(recorded_errors_to_replay=...) / --replay_err_in (dem_sampler.inl:52-130): the errors the reference sampled from a
    noisy model give the detectors and observables the reference derived from them."""
    dem = stim_b200.DetectorErrorModel(case["dem"])
    D, L, E = dem.num_detectors, dem.num_observables, dem.num_errors
    shots = case["shots"]
    raw = {k: np.frombuffer(base64.b64decode(case[k]), dtype=np.uint8) for k in ("det", "obs", "err")}
    errs = raw["err"].reshape(shots, (E + 7) // 8)
    want_d = raw["det"].reshape(shots, (D + 7) // 8)
    want_o = raw["obs"].reshape(shots, (L + 7) // 8)
    sampler = dem.compile_sampler(seed=3)
    d, o, e = sampler.sample(shots, bit_packed=True, return_errors=True, recorded_errors_to_replay=errs)
    np.testing.assert_array_equal(d, want_d)
    np.testing.assert_array_equal(o, want_o)
    np.testing.assert_array_equal(e, errs)
    d2, o2, e2 = sampler.sample(shots, recorded_errors_to_replay=np.unpackbits(errs, axis=1, bitorder="little", count=E).astype(np.bool_))
    np.testing.assert_array_equal(np.packbits(d2, axis=1, bitorder="little") if D else d2.astype(np.uint8), want_d)
    np.testing.assert_array_equal(np.packbits(o2, axis=1, bitorder="little") if L else o2.astype(np.uint8), want_o)
    assert e2 is None
    # files, through the command line mirror
    import subprocess
    import sys

    (tmp_path / "m.dem").write_text(case["dem"])
    (tmp_path / "e.b8").write_bytes(raw["err"].tobytes())
    r = subprocess.run([sys.executable, "-m", "stim_b200", "sample_dem", "--shots", str(shots), "--in", str(tmp_path / "m.dem"),
                        "--replay_err_in", str(tmp_path / "e.b8"), "--replay_err_in_format", "b8", "--out", str(tmp_path / "d.b8"),
                        "--out_format", "b8", "--obs_out", str(tmp_path / "o.b8"), "--obs_out_format", "b8"],
                       capture_output=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stderr.decode()
    assert (tmp_path / "d.b8").read_bytes() == raw["det"].tobytes()
    assert (tmp_path / "o.b8").read_bytes() == raw["obs"].tobytes()
    with pytest.raises(ValueError):
        sampler.sample(shots + 1, recorded_errors_to_replay=errs)
