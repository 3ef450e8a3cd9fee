"""`python -m stim_b200 convert` against the unmodified reference CLI (`stim convert`,
/root/reference/src/stim/cmd/command_convert.cc): same bytes on --out and --obs_out and the same exit status for every way of
describing the record layout. Goldens: tests/golden/convert_cases.json (tools/gen_convert_golden.py). Host-only."""
import base64
import json
import os
import sys

import pytest

import stim_b200.__main__ as cli

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "convert_cases.json")))


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_convert_reproduces_the_reference_cli(case, tmp_path, capsys):
    (tmp_path / "c.stim").write_text(GOLD["circuit"])
    (tmp_path / "m.dem").write_text(GOLD["dem"])
    (tmp_path / "in.dat").write_bytes(base64.b64decode(case["input"]))
    flags = [{"@CIRCUIT": str(tmp_path / "c.stim"), "@DEM": str(tmp_path / "m.dem")}.get(f, f) for f in case["flags"]]
    flags += ["--in", str(tmp_path / "in.dat"), "--out", str(tmp_path / "out.dat")]
    if case["obs_out"] is not None:
        flags += ["--obs_out", str(tmp_path / "obs.dat")]
    rc = cli.main(["convert"] + flags)
    assert rc == case["rc"]
    if rc != 0:
        assert capsys.readouterr().err.startswith("\033[31m")
        return
    assert (tmp_path / "out.dat").read_bytes() == base64.b64decode(case["stdout"])
    if case["obs_out"] is not None:
        assert (tmp_path / "obs.dat").read_bytes() == base64.b64decode(case["obs_out"])


def test_convert_reads_stdin_and_writes_stdout(tmp_path):
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "stim_b200", "convert", "--in_format", "01", "--out_format", "dets", "--num_measurements",
                        "1", "--num_detectors", "2", "--num_observables", "1"], input=b"0101\n1100\n", capture_output=True, cwd=root)
    assert r.returncode == 0 and r.stdout == b"shot D0 L0\nshot M0 D0\n"


# Known answers restated from the reference's own tests (/root/reference/src/stim/cmd/command_convert.test.cc:22-51, 53-100,
# 308-313, 334-344): the same records in four input formats, the `--flag=value` spelling included.
MEAS_CIRCUIT = "X 0\nM 0 1\nDETECTOR rec[-2]\nDETECTOR rec[-1]\nOBSERVABLE_INCLUDE(2) rec[-1]\n"
DET_CIRCUIT = ("CX 0 2 1 2\nM 2\nCX rec[-1] 2\nDETECTOR rec[-1]\nTICK\n" + "CX 0 2 1 2\nM 2\nCX rec[-1] 2\nDETECTOR rec[-1] rec[-2]\nTICK\n" * 2
               + "M 0 1\nDETECTOR rec[-1] rec[-2] rec[-3]\nOBSERVABLE_INCLUDE(0) rec[-1]\n")


def _run(tmp_path, flags, data, circuit=None):
    (tmp_path / "in.dat").write_bytes(data)
    if circuit is not None:
        (tmp_path / "c.stim").write_text(circuit)
        flags = flags + ["--circuit", str(tmp_path / "c.stim")]
    rc = cli.main(["convert"] + flags + ["--in", str(tmp_path / "in.dat"), "--out", str(tmp_path / "out.dat")])
    return rc, (tmp_path / "out.dat").read_bytes() if rc == 0 else None


@pytest.mark.parametrize("fmt,data", [("01", b"00\n01\n10\n11\n"), ("b8", bytes([0, 2, 1, 3])), ("hits", b"\n1\n0\n0,1\n"),
                                      ("r8", bytes([2, 1, 0, 0, 1, 0, 0, 0]))])
def test_reference_known_answers_measurements_to_dets(tmp_path, fmt, data):
    assert _run(tmp_path, ["--in_format", fmt, "--out_format", "dets", "--types=M"], data, MEAS_CIRCUIT) == (
        0, b"shot\nshot M1\nshot M0\nshot M0 M1\n")


@pytest.mark.parametrize("fmt,data", [("01", b"00000\n11000\n01100\n00110\n00010\n00011\n"), ("b8", bytes([0, 3, 6, 12, 8, 24])),
                                      ("hits", b"\n0,1\n1,2\n2,3\n3\n3,4\n"),
                                      ("r8", bytes([5, 0, 0, 3, 1, 0, 2, 2, 0, 1, 3, 1, 3, 0, 0]))])
def test_reference_known_answers_detections_and_observables_to_dets(tmp_path, fmt, data):
    assert _run(tmp_path, ["--in_format", fmt, "--out_format", "dets", "--types=DL"], data, DET_CIRCUIT) == (
        0, b"shot\nshot D0 D1\nshot D1 D2\nshot D2 D3\nshot D3\nshot D3 L0\n")


def test_reference_known_answers_wide_records_and_refusals(tmp_path):
    assert _run(tmp_path, ["--in_format=b8", "--out_format=b8", "--bits_per_shot=2048"], b"\x6b" * 256) == (0, b"\x6b" * 256)
    assert _run(tmp_path, ["--in_format=r8", "--out_format=b8"], b"")[0] == 1
    assert _run(tmp_path, ["--in_format=01", "--out_format", "dets", "--bits_per_shot=2"], b"")[0] == 1


def test_writing_to_stdout_appends_to_a_redirected_file(tmp_path):
    """The reference writes to its `stdout` stream; opening "/dev/stdout" again would truncate a file that stdout was
    appended to (`>> log`) or that several commands share."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    log = tmp_path / "log.txt"
    log.write_bytes(b"first line\n")
    for data in (b"0,2\n", b"1\n"):
        with open(log, "ab") as f:
            r = subprocess.run([sys.executable, "-m", "stim_b200", "convert", "--in_format", "hits", "--out_format", "01",
                                "--num_measurements", "4"], input=data, stdout=f, cwd=root)
        assert r.returncode == 0
    assert log.read_bytes() == b"first line\n1010\n0100\n"


def test_convert_also_writes_ptb64(tmp_path):
    """Beyond the reference (its per-record converter refuses to WRITE ptb64): the bytes equal the ptb64 files of
    tests/golden/formats_cases.json, which the reference CLI reads back into the original records (tools/gen_formats_golden.py)."""
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "formats_cases.json")))
    n = 0
    for c in cases:
        if "ptb64" not in c["files"]:
            continue
        counts = [f"--num_measurements={c['num_measurements']}", f"--num_detectors={c['num_detectors']}",
                  f"--num_observables={c['num_observables']}"]
        rc, out = _run(tmp_path, ["--in_format=01", "--out_format=ptb64"] + counts, base64.b64decode(c["files"]["01"]))
        assert rc == 0 and out == base64.b64decode(c["files"]["ptb64"])
        n += 1
    assert n >= 5
    assert _run(tmp_path, ["--in_format=01", "--out_format=ptb64", "--num_measurements=3"], b"010\n")[0] == 1  # shots % 64 != 0
