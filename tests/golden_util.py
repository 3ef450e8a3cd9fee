"""Helpers shared by the golden-vector tests: load tests/golden/reference_outputs.json (outputs of the unmodified
reference CLI on deterministic circuits, see tools/gen_golden.py) and decode the packed formats."""
import base64
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_cases():
    with open(os.path.join(ROOT, "tests", "golden", "reference_outputs.json")) as f:
        return json.load(f)


def case_ids(cases):
    return [c["name"] for c in cases]


def output_bytes(case, fmt) -> bytes:
    return base64.b64decode(case["outputs"][fmt]["data"])


def expected_bits(case, n_bits):
    """uint8 [shots, n_bits] decoded from the case's b8 (or 01) reference output."""
    if "b8" in case["outputs"]:
        shots = case["outputs"]["b8"]["shots"]
        raw = np.frombuffer(output_bytes(case, "b8"), dtype=np.uint8)
        nb = (n_bits + 7) // 8
        assert raw.size == shots * nb, (raw.size, shots, nb)
        if n_bits == 0:
            return np.zeros((shots, 0), dtype=np.uint8)
        return np.unpackbits(raw.reshape(shots, nb), axis=1, bitorder="little")[:, :n_bits]
    text = output_bytes(case, "01").decode().split("\n")[:-1]
    return np.array([[int(ch) for ch in line] for line in text], dtype=np.uint8).reshape(len(text), n_bits)


def arrange(dets, obs, flags):
    """dets/obs uint8 arrays -> the column order the CLI flags ask for."""
    if "--append_observables" in flags:
        return np.concatenate([dets, obs], axis=1)
    if "--prepend_observables" in flags:
        return np.concatenate([obs, dets], axis=1)
    return dets


def dets_prefixes(case, n_det, n_obs):
    """(prefix1, prefix2, transition) the reference uses for the "dets" format (frame_simulator_util.inl:149-187)."""
    if case["mode"] == "sample":
        return b"M", b"M", 0
    if "--append_observables" in case["flags"]:
        return b"D", b"L", n_det
    if "--prepend_observables" in case["flags"]:
        return b"L", b"D", n_obs
    return b"D", b"L", n_det
