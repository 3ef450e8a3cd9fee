"""GPU tests against the reference's own outputs (tests/golden/reference_outputs.json): deterministic circuits must
be reproduced byte for byte in every result format, through the same API surface the reference offers
(sample / sample_write of the compiled samplers)."""
import os

import numpy as np
import pytest

import stim_b200
from conftest import ROOT
from golden_util import arrange, case_ids, expected_bits, load_cases, output_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["interp", "events"])
def engine(request, monkeypatch):
    """Every golden case runs under both engines: the reference's deterministic outputs do not depend on how the shots
    are sampled ("events" = the event engine wherever the circuit is eligible for it)."""
    monkeypatch.setenv("GSTIM_ENGINE", request.param)
    return request.param

CASES = load_cases()
DETECT = [c for c in CASES if c["mode"] == "detect"]
SAMPLE = [c for c in CASES if c["mode"] == "sample"]


def _ref_array(case):
    return np.array([int(ch) for ch in case["reference_sample"]], dtype=np.bool_)


@pytest.mark.parametrize("case", DETECT, ids=case_ids(DETECT))
def test_detector_sampler_reproduces_reference_bytes(case, tmp_path):
    circ = stim_b200.Circuit(case["circuit"])
    flags = case["flags"]
    kw = dict(append_observables="--append_observables" in flags, prepend_observables="--prepend_observables" in flags)
    for fmt, out in case["outputs"].items():
        path = tmp_path / f"out.{fmt}"
        circ.compile_detector_sampler(seed=5).sample_write(out["shots"], filepath=str(path), format=fmt, **kw)
        assert path.read_bytes() == output_bytes(case, fmt), (case["name"], fmt)
    # the in-memory API describes the same shots
    shots = next(iter(case["outputs"].values()))["shots"]
    dets, obs = circ.compile_detector_sampler(seed=6).sample(shots, separate_observables=True)
    got = arrange(dets.astype(np.uint8), obs.astype(np.uint8), flags)
    np.testing.assert_array_equal(got, expected_bits(case, got.shape[1])[:shots])


@pytest.mark.parametrize("case", SAMPLE, ids=case_ids(SAMPLE))
def test_measurement_sampler_reproduces_reference_bytes(case, tmp_path):
    circ = stim_b200.Circuit(case["circuit"])
    ref = _ref_array(case)
    for fmt, out in case["outputs"].items():
        path = tmp_path / f"out.{fmt}"
        circ.compile_sampler(seed=5, reference_sample=ref).sample_write(out["shots"], filepath=str(path), format=fmt)
        assert path.read_bytes() == output_bytes(case, fmt), (case["name"], fmt)
    shots = next(iter(case["outputs"].values()))["shots"]
    got = circ.compile_sampler(seed=6, reference_sample=np.packbits(ref, bitorder="little")).sample(shots)
    np.testing.assert_array_equal(got.astype(np.uint8), expected_bits(case, got.shape[1])[:shots])
    # default reference sample (host stabilizer simulation) must agree with the reference's TableauSimulator
    got2 = circ.compile_sampler(seed=7).sample(shots, bit_packed=True)
    np.testing.assert_array_equal(np.unpackbits(got2, axis=1, bitorder="little")[:, : ref.size], expected_bits(case, ref.size)[:shots])


def test_obs_out_file_and_flag_errors(tmp_path):
    case = next(c for c in DETECT if c["name"] == "detector_sampler_no_obs")
    circ = stim_b200.Circuit(case["circuit"])
    s = circ.compile_detector_sampler(seed=1)
    p1, p2 = tmp_path / "d.01", tmp_path / "o.01"
    s.sample_write(5, filepath=str(p1), format="01", obs_out_filepath=str(p2), obs_out_format="01")
    assert p1.read_bytes() == output_bytes(case, "01")
    assert p2.read_bytes() == b"0001\n" * 5  # observable 3 <- rec[-2] = X_ERROR(1) flipped qubit 0
    with pytest.raises(IndexError):  # frame_simulator_util.inl:127-129 throws std::out_of_range
        s.sample_write(5, filepath=str(p1), format="01", obs_out_filepath=str(p2), append_observables=True)
    with pytest.raises(ValueError):  # measure_record_writer.h:123-125
        s.sample_write(5, filepath=str(p1), format="ptb64")
    with pytest.raises(ValueError):
        s.sample_write(5, filepath=str(p1), format="bogus")


def _c3_variant(knob):
    """The full-size benchmark circuit made deterministic: every DEPOLARIZE off, one family of flips at p = 1."""
    import re

    with open(os.path.join(ROOT, "tests", "golden", "circuits", "c3_surface_z_d25_r25.stim")) as f:
        text = f.read()
    text = re.sub(r"DEPOLARIZE([12])\(0\.001\)", r"DEPOLARIZE\1(0)", text)
    lines = text.split("\n")
    out = []
    for i, ln in enumerate(lines):
        if ln.strip().startswith("X_ERROR(0.001)"):
            nxt = next((l.strip() for l in lines[i + 1:] if l.strip()), "")
            before_measure = nxt.startswith("M")
            on = (knob == "measure") == before_measure
            ln = ln.replace("X_ERROR(0.001)", "X_ERROR(1)" if on else "X_ERROR(0)")
        out.append(ln)
    return "\n".join(out)


def _c5_variant(knob):
    """BASELINE.json configs[4] (d=51 r=51, REPEAT kept) made deterministic: DEPOLARIZE off, one family of flips at p = 1:
    "measure" = the X_ERROR / Z_ERROR directly in front of a measurement, "reset" = the ones behind a reset."""
    import re

    with open(os.path.join(ROOT, "tests", "golden", "circuits", "c5_surface_x_d51_r51.stim")) as f:
        text = f.read()
    text = re.sub(r"DEPOLARIZE([12])\(0\.001\)", r"DEPOLARIZE\1(0)", text)
    lines = text.split("\n")
    out = []
    for i, ln in enumerate(lines):
        m = re.match(r"\s*([XZ])_ERROR\(0\.001\)", ln)
        if m:
            nxt = next((l.strip() for l in lines[i + 1:] if l.strip()), "")
            before_measure = nxt.startswith("M")
            on = (knob == "measure") == before_measure
            ln = ln.replace("_ERROR(0.001)", "_ERROR(1)" if on else "_ERROR(0)")
        out.append(ln)
    return "\n".join(out)


def test_d51_deterministic_noise_rows_equal_reference():
    """c5 (5201 qubits, 132 600 detectors, lookback 5201, REPEAT 50): every shot must equal the row the reference
    produces (tests/golden/c5_det_rows.json, tools/gen_c3_rows.py); a noiseless run gives all-zero detectors."""
    import json
    import re

    with open(os.path.join(ROOT, "tests", "golden", "c5_det_rows.json")) as f:
        rows = json.load(f)
    for knob in ("measure", "reset"):
        want = np.frombuffer(bytes.fromhex(rows[knob]), dtype=np.uint8)
        got = stim_b200.Circuit(_c5_variant(knob)).compile_detector_sampler(seed=2).sample(
            1024 + 77, bit_packed=True, append_observables=True)
        assert got.shape[1] == want.size == 16576
        assert (got == want[None, :]).all(), knob
    with open(os.path.join(ROOT, "tests", "golden", "circuits", "c5_surface_x_d51_r51.stim")) as f:
        noiseless = re.sub(r"\((0\.001)\)", "(0)", f.read())
    a = stim_b200.Circuit(noiseless).compile_detector_sampler(seed=1).sample(2048, bit_packed=True, append_observables=True)
    assert a.shape == (2048, 16576) and not a.any()


def test_full_size_noiseless_and_deterministic_noise():
    """BASELINE.json configs[2] size: a noiseless run gives all-zero detectors; with probability-1 flips every shot
    must equal the row the reference produces (fixture c3_det_rows.json)."""
    import json
    import re

    with open(os.path.join(ROOT, "tests", "golden", "circuits", "c3_surface_z_d25_r25.stim")) as f:
        noisy = f.read()
    noiseless = re.sub(r"\((0\.001)\)", "(0)", noisy)
    shots = 1 << 16
    a = stim_b200.Circuit(noiseless).compile_detector_sampler(seed=1).sample(shots, bit_packed=True, append_observables=True)
    assert a.shape == (shots, 1951) and not a.any()
    with open(os.path.join(ROOT, "tests", "golden", "c3_det_rows.json")) as f:
        rows = json.load(f)
    for knob in ("measure", "reset"):
        want = np.frombuffer(bytes.fromhex(rows[knob]), dtype=np.uint8)
        got = stim_b200.Circuit(_c3_variant(knob)).compile_detector_sampler(seed=2).sample(
            4096 + 77, bit_packed=True, append_observables=True)
        assert got.shape[1] == want.size
        assert (got == want[None, :]).all(), knob


def test_command_line_mirror_reproduces_reference_bytes(tmp_path):
    """`python -m stim_b200 detect|sample` = `stim detect|sample` (command_detect.cc:23-79, command_sample.cc:25-71):
    same flags, same bytes, errors on stderr with exit status 1."""
    import subprocess
    import sys

    env = dict(os.environ, PYTHONPATH=ROOT)
    n_run = 0
    for case in DETECT[:6] + SAMPLE[:4]:
        src = tmp_path / "c.stim"
        src.write_text(case["circuit"])
        for fmt, out in list(case["outputs"].items())[:2]:
            dst = tmp_path / f"cli.{fmt}"
            cmd = [sys.executable, "-m", "stim_b200", case["mode"], "--shots", str(out["shots"]), "--in", str(src),
                   "--out", str(dst), "--out_format", fmt, "--seed", "3"] + [f for f in case["flags"] if f.startswith("--")]
            r = subprocess.run(cmd, env=env, capture_output=True, cwd=str(tmp_path))
            assert r.returncode == 0, r.stderr.decode()
            assert dst.read_bytes() == output_bytes(case, fmt), (case["name"], fmt)
            n_run += 1
    assert n_run >= 10
    bad = tmp_path / "bad.stim"
    bad.write_text("M 0\nDETECTOR rec[-2]\n")
    r = subprocess.run([sys.executable, "-m", "stim_b200", "detect", "--shots", "1", "--in", str(bad)], env=env, capture_output=True)
    assert r.returncode == 1 and b"rec[-2]" in r.stderr.replace(b"\x1b[31m", b"") or r.returncode == 1
