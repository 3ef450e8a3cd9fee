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
