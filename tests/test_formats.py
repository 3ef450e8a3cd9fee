"""The six shot-data formats on the input side (stim_b200/_formats.py) and the output side (writers.cc) against files written by
the unmodified reference CLI (`stim convert`, tools/gen_formats_golden.py -> tests/golden/formats_cases.json), plus the
read_shot_data_file / write_shot_data_file mirrors of /root/reference/src/stim/io/read_write.pybind.cc. Host code only."""
import base64
import json
import os

import numpy as np
import pytest

import stim_b200
from stim_b200 import _formats

CASES = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "formats_cases.json")))


def _bits(case):
    n = case["num_measurements"] + case["num_detectors"] + case["num_observables"]
    rows = np.frombuffer(base64.b64decode(case["bits"]), dtype=np.uint8).reshape(case["shots"], (n + 7) // 8)
    return n, rows


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"m{c['num_measurements']}d{c['num_detectors']}l{c['num_observables']}x{c['shots']}")
def test_readers_decode_reference_files_and_writers_reproduce_them(case, tmp_path):
    n, rows = _bits(case)
    nm, nd, no = case["num_measurements"], case["num_detectors"], case["num_observables"]
    for fmt, b64 in case["files"].items():
        data = base64.b64decode(b64)
        got = _formats.read_shots(data, fmt, n, num_measurements=nm, num_detectors=nd, num_observables=no)
        np.testing.assert_array_equal(got, rows, err_msg=fmt)
        path = tmp_path / f"x.{fmt}"
        stim_b200.write_shot_data_file(data=rows, path=str(path), format=fmt, num_measurements=nm or None,
                                       num_detectors=nd or None, num_observables=no or None)
        assert path.read_bytes() == data, fmt
        back = stim_b200.read_shot_data_file(path=str(path), format=fmt, bit_packed=True, num_measurements=nm, num_detectors=nd,
                                             num_observables=no)
        np.testing.assert_array_equal(back, rows, err_msg=fmt)
        if no:
            dets, obs = stim_b200.read_shot_data_file(path=str(path), format=fmt, num_measurements=nm, num_detectors=nd,
                                                      num_observables=no, separate_observables=True)
            full = np.unpackbits(rows, axis=1, bitorder="little", count=n).astype(np.bool_)
            np.testing.assert_array_equal(dets, full[:, : nm + nd])
            np.testing.assert_array_equal(obs, full[:, nm + nd:])


def test_reader_errors_match_the_reference_messages():
    with pytest.raises(ValueError, match="b8 data ended in middle of record"):
        _formats.read_shots(b"\x00\x00\x00", "b8", 16)
    with pytest.raises(ValueError, match="hit index is too large"):
        _formats.read_shots(b"0,9\n", "hits", 9)
    with pytest.raises(ValueError, match="comma-separated integers"):
        _formats.read_shots(b"0;1\n", "hits", 9)
    with pytest.raises(ValueError, match="End of file before end of r8 data"):
        _formats.read_shots(bytes([3]), "r8", 9)
    with pytest.raises(ValueError, match="jumped past expected end"):
        _formats.read_shots(bytes([3, 7]), "r8", 9)
    with pytest.raises(ValueError, match="didn't start with 'shot'"):
        _formats.read_shots(b"M0\n", "dets", 9)
    with pytest.raises(ValueError, match="larger than expected"):
        _formats.read_shots(b"shot M9\n", "dets", 9)
    with pytest.raises(ValueError, match="larger than expected"):
        _formats.read_shots(b"shot D0\n", "dets", 9)
    with pytest.raises(ValueError, match="middle of a ptb64 record"):
        _formats.read_shots(b"\x00" * 12, "ptb64", 2)
    with pytest.raises(ValueError, match="Must specify"):
        stim_b200.read_shot_data_file(path="/dev/null", format="01")
    with pytest.raises(ValueError, match="num_measurements and"):
        stim_b200.write_shot_data_file(data=np.zeros((1, 3), np.bool_), path="/dev/null", format="01", num_measurements=1, num_detectors=2)
    # a hit listed twice toggles back; blank hits lines are empty shots; r8 255-runs
    np.testing.assert_array_equal(_formats.read_shots(b"1,1,2\n\n", "hits", 8), np.array([[4], [0]], np.uint8))
    long = bytes([255, 45, 0]) + bytes([255, 46])  # one 1 at position 300 (+ terminator), then an empty 301-bit record
    got = _formats.read_shots(long + bytes([255, 46]), "r8", 301)
    assert got.shape == (3, 38) and got[0, 37] == 1 << 4 and got.sum() == 1 << 4
