"""Measurements -> detection events. CPU: the oracle restatement reproduces the reference CLI's `stim m2d`
outputs (tests/golden/m2d_cases.json, tools/gen_m2d_golden.py: seeded random measurement + sweep data, with and without
the reference sample, with and without sweep bits). GPU: the CUDA converter (gstim_m2d_convert through the Python mirror of
stim.CompiledMeasurementsToDetectionEventsConverter) reproduces the same bytes, and equals the oracle on the headline
circuit's 16 225 measurements."""
import json
import os

import numpy as np
import pytest

import stim_b200
from conftest import ROOT
from oracle import m2d_oracle

with open(os.path.join(ROOT, "tests", "golden", "m2d_cases.json")) as f:
    CASES = json.load(f)


def bits01(text, width=None):
    rows = text.split("\n")[:-1] if text else []
    a = np.array([[int(ch) for ch in r] for r in rows], dtype=np.uint8)
    if a.ndim == 1:
        a = a.reshape(len(rows), 0)
    return a


def variants(case):
    for key, want in case["outputs"].items():
        opts = dict(kv.split("=") for kv in key.split(","))
        yield key, opts["skip"] == "1", opts["sweep"] == "1", want


def reference_sample_bits(text, M):
    from stim_b200._reference_sample import reference_sample_bits as packed

    return np.unpackbits(packed(text, M), bitorder="little")[:M]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_reference_m2d(case):
    meas, sweep = bits01(case["measurements"]), bits01(case["sweep"])
    shots = case["shots"]
    meas = meas.reshape(shots, -1)
    sweep = sweep.reshape(shots, -1)
    for key, skip, use_sweep, want in variants(case):
        ref = None if skip else reference_sample_bits(case["circuit"], meas.shape[1])
        got = m2d_oracle.convert(case["circuit"], meas, sweep if use_sweep else None, ref, append_observables=True)
        exp = bits01(want).reshape(shots, -1)
        np.testing.assert_array_equal(got, exp, err_msg=f"{case['name']} {key}")


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_converter_reproduces_reference_m2d(case):
    shots = case["shots"]
    meas = bits01(case["measurements"]).reshape(shots, -1).astype(np.bool_)
    sweep = bits01(case["sweep"]).reshape(shots, -1).astype(np.bool_)
    circ = stim_b200.Circuit(case["circuit"])
    for key, skip, use_sweep, want in variants(case):
        conv = circ.compile_m2d_converter(skip_reference_sample=skip)
        assert conv.num_measurements == meas.shape[1] and conv.num_sweep_bits == sweep.shape[1]
        exp = bits01(want).reshape(shots, -1).astype(np.bool_)
        got = conv.convert(measurements=meas, sweep_bits=sweep if use_sweep else None, append_observables=True)
        np.testing.assert_array_equal(got, exp, err_msg=f"{case['name']} {key}")
        # packed input, separate observables, packed output
        D, L = conv.num_detectors, conv.num_observables
        d2, o2 = conv.convert(measurements=np.packbits(meas, axis=1, bitorder="little"),
                              sweep_bits=np.packbits(sweep, axis=1, bitorder="little") if use_sweep else None,
                              separate_observables=True, bit_packed=True)
        assert d2.dtype == np.uint8 and d2.shape == (shots, (D + 7) // 8) and o2.shape == (shots, (L + 7) // 8)
        np.testing.assert_array_equal(np.unpackbits(d2, axis=1, bitorder="little", count=D), exp[:, :D].astype(np.uint8))
        np.testing.assert_array_equal(np.unpackbits(o2, axis=1, bitorder="little", count=L), exp[:, D:].astype(np.uint8))


@pytest.mark.gpu
def test_converter_argument_errors():
    conv = stim_b200.Circuit("M 0 1\nDETECTOR rec[-1]\n").compile_m2d_converter()
    m = np.zeros((4, 2), dtype=np.bool_)
    with pytest.raises(ValueError):  # pybind.cc:86-90
        conv.convert(measurements=m)
    with pytest.raises(ValueError):
        conv.convert(measurements=np.zeros((4, 3), dtype=np.bool_), append_observables=False)
    with pytest.raises(ValueError):
        conv.convert(measurements=m, sweep_bits=np.zeros((3, 0), dtype=np.bool_), append_observables=False)
    assert conv.convert(measurements=m, append_observables=False).shape == (4, 1)


@pytest.mark.gpu
def test_headline_circuit_samples_convert_back_to_their_detection_events():
    """Round trip at full width (c3: 16 225 measurements -> 15 600 detectors + 1 observable): measurement samples of the
    noisy circuit, converted, equal the oracle's conversion; and the converted detection fraction matches detect's."""
    with open(os.path.join(ROOT, "tests", "golden", "circuits", "c3_surface_z_d25_r25.stim")) as f:
        text = f.read()
    circ = stim_b200.Circuit(text)
    shots = 300
    meas = circ.compile_sampler(seed=11).sample(shots)
    conv = circ.compile_m2d_converter()
    dets, obs = conv.convert(measurements=meas, separate_observables=True)
    ref = reference_sample_bits(text, meas.shape[1])
    od, oo = m2d_oracle.convert(text, meas[:64], None, ref)
    np.testing.assert_array_equal(dets[:64].astype(np.uint8), od)
    np.testing.assert_array_equal(obs[:64].astype(np.uint8), oo)
    direct = circ.compile_detector_sampler(seed=12).sample(4096)
    assert abs(dets.mean() - direct.mean()) < 0.003


@pytest.mark.gpu
def test_command_line_mirror_of_m2d(tmp_path):
    import subprocess
    import sys

    case = next(c for c in CASES if c["name"] == "sweep_feedback_repeat")
    (tmp_path / "c.stim").write_text(case["circuit"])
    (tmp_path / "m.01").write_text(case["measurements"])
    (tmp_path / "s.01").write_text(case["sweep"])
    r = subprocess.run([sys.executable, "-m", "stim_b200", "m2d", "--circuit", str(tmp_path / "c.stim"), "--in", str(tmp_path / "m.01"),
                        "--sweep", str(tmp_path / "s.01"), "--append_observables"], capture_output=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout.decode() == case["outputs"]["skip=0,sweep=1"]
    # b8 input, dets output with prefixes
    meas = np.packbits(bits01(case["measurements"]).reshape(case["shots"], -1), axis=1, bitorder="little")
    (tmp_path / "m.b8").write_bytes(meas.tobytes())
    r = subprocess.run([sys.executable, "-m", "stim_b200", "m2d", "--circuit", str(tmp_path / "c.stim"), "--in", str(tmp_path / "m.b8"),
                        "--in_format", "b8", "--out_format", "dets", "--append_observables", "--skip_reference_sample"],
                       capture_output=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr.decode()
    exp = bits01(case["outputs"]["skip=1,sweep=0"]).reshape(case["shots"], -1)
    D = stim_b200.Circuit(case["circuit"]).num_detectors
    want = "".join("shot" + "".join(f" {'D' if j < D else 'L'}{j if j < D else j - D}" for j in np.flatnonzero(row)) + "\n" for row in exp)
    assert r.stdout.decode() == want


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["r8", "hits", "dets", "b8"])
def test_convert_file_reads_and_writes_every_format(tmp_path, fmt):
    """convert_file (measurements_to_detection_events.pybind.cc, command_m2d.cc): measurement and sweep data in `fmt`, the
    detection events written in `fmt` and the observables separately — decoded again they equal the reference's 01 output."""
    case = next(c for c in CASES if c["name"] == "sweep_feedback_repeat")
    circ = stim_b200.Circuit(case["circuit"])
    conv = circ.compile_m2d_converter()
    shots = case["shots"]
    meas = bits01(case["measurements"]).reshape(shots, -1).astype(np.bool_)
    sweep = bits01(case["sweep"]).reshape(shots, -1).astype(np.bool_)
    stim_b200.write_shot_data_file(data=meas, path=str(tmp_path / "m"), format=fmt, num_measurements=meas.shape[1])
    stim_b200.write_shot_data_file(data=sweep, path=str(tmp_path / "s"), format=fmt, num_measurements=sweep.shape[1])
    conv.convert_file(measurements_filepath=str(tmp_path / "m"), measurements_format=fmt, sweep_bits_filepath=str(tmp_path / "s"),
                      sweep_bits_format=fmt, detection_events_filepath=str(tmp_path / "d"), detection_events_format=fmt,
                      obs_out_filepath=str(tmp_path / "o"), obs_out_format=fmt)
    D, L = circ.num_detectors, circ.num_observables
    dets = stim_b200.read_shot_data_file(path=str(tmp_path / "d"), format=fmt, num_detectors=D)
    obs = stim_b200.read_shot_data_file(path=str(tmp_path / "o"), format=fmt, num_observables=L)
    want = bits01(case["outputs"]["skip=0,sweep=1"]).reshape(shots, -1).astype(np.bool_)  # (detectors + appended observables)
    np.testing.assert_array_equal(dets, want[:, :D])
    np.testing.assert_array_equal(obs, want[:, D:])
