import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_STIM = os.path.join(ROOT, "oracle", "_ref", "stim")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The shared library is a build product; make sure it exists before anything imports stim_b200.
    from stim_b200 import build as _build

    _build.build()


def have_ref() -> bool:
    return os.path.exists(REF_STIM)


def ref_stim(*args, stdin: bytes = b"") -> bytes:
    """Runs the reference CLI built by oracle/Makefile (test infrastructure only)."""
    r = subprocess.run([REF_STIM, *args], input=stdin, capture_output=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr.decode())
    return r.stdout


def gen_circuit(code, task, distance, rounds, p=0.0, **knobs) -> str:
    """Generated benchmark circuits come from committed fixtures when present, else from oracle/_ref."""
    name = f"{code}_{task}_d{distance}_r{rounds}_p{p}" + "".join(f"_{k}{v}" for k, v in sorted(knobs.items())) + ".stim"
    path = os.path.join(ROOT, "tests", "golden", "circuits", name)
    if os.path.exists(path):
        with open(path) as f:
            return f.read()
    if not have_ref():
        pytest.skip("fixture circuit missing and oracle/_ref/stim not built")
    args = ["gen", "--code", code, "--task", task, "--distance", str(distance), "--rounds", str(rounds)]
    allk = dict(after_clifford_depolarization=p, before_round_data_depolarization=p, before_measure_flip_probability=p,
                after_reset_flip_probability=p)
    allk.update(knobs)
    for k, v in allk.items():
        args += [f"--{k}", str(v)]
    return ref_stim(*args).decode()


@pytest.fixture(scope="session")
def gpu_available():
    from stim_b200 import _native

    return _native.lib().gstim_device_count() > 0
