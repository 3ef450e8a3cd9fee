"""The boundary against its real caller: oracle/_ref/stim_gstim is `stim detect` / `stim sample` built from the REFERENCE's own
host code (argument parser, Circuit::from_file, Circuit::str(), TableauSimulator reference sample — objects compiled from
/root/reference by oracle/Makefile) with only the two batch drivers replaced by the shim of INTEGRATION.md §1
(tests/integration/frame_simulator_gstim.h), which calls libgstim.so through the C ABI.

Replayed here: every deterministic case of tests/golden/reference_outputs.json — they include the cases of
/root/reference/src/stim/cmd/command_detect.test.cc:21-183 and command_sample.test.cc — in every result format; the bytes
must equal what the unmodified reference CLI printed. Error behaviour: exit status 1 and the reference's message."""
import os
import subprocess

import pytest

from conftest import ROOT
from golden_util import case_ids, load_cases, output_bytes

pytestmark = pytest.mark.gpu

SHIM = os.path.join(ROOT, "oracle", "_ref", "stim_gstim")
CASES = load_cases()


def run_shim(*args, stdin=b""):
    return subprocess.run([SHIM, *args], input=stdin, capture_output=True, timeout=300)


@pytest.fixture(autouse=True)
def _need_shim():
    if not os.path.exists(SHIM):
        pytest.skip("oracle/_ref/stim_gstim is not built (make -C oracle shim)")


@pytest.mark.parametrize("index,case", list(enumerate(CASES)), ids=case_ids(CASES))
def test_reference_host_code_over_the_c_abi_reproduces_reference_bytes(index, case):
    # (every process start initialises CUDA: b8 always, plus one of the other formats in rotation)
    fmts = sorted(case["outputs"])
    pick = {f for f in fmts if f == "b8"} | {fmts[index % len(fmts)]}
    for fmt in sorted(pick):
        out = case["outputs"][fmt]
        r = run_shim("detect" if case["mode"] == "detect" else "sample", "--shots", str(out["shots"]), "--out_format", fmt,
                     *case["flags"], "--seed", "7", stdin=case["circuit"].encode())
        assert r.returncode == 0, r.stderr.decode()
        assert r.stdout == output_bytes(case, fmt), (case["name"], fmt)


def test_obs_out_and_error_behaviour(tmp_path):
    case = next(c for c in CASES if c["name"] == "detector_sampler_no_obs")
    obs = tmp_path / "obs.01"
    r = run_shim("detect", "--shots", "5", "--obs_out", str(obs), "--seed", "1", stdin=case["circuit"].encode())
    assert r.returncode == 0 and r.stdout == output_bytes(case, "01") and obs.read_bytes() == b"0001\n" * 5
    # ptb64 needs a multiple of 64 shots (measure_record_writer.h:123-125): invalid_argument -> status 1, red message
    r = run_shim("detect", "--shots", "5", "--out_format", "ptb64", stdin=case["circuit"].encode())
    assert r.returncode == 1 and b"multiple of 64" in r.stderr
    # a lookback before the beginning of time (measure_record_batch.inl:83-94): out_of_range
    r = run_shim("detect", "--shots", "5", stdin=b"M 0\nDETECTOR rec[-2]\n")
    assert r.returncode == 1 and b"before the beginning of time" in r.stderr
    # the reference's own parser rejects a bad circuit before the library sees it
    r = run_shim("sample", "--shots", "5", stdin=b"NOT_A_GATE 0\n")
    assert r.returncode == 1 and b"NOT_A_GATE" in r.stderr
