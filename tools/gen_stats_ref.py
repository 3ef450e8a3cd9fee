"""Generates tests/golden/stats_ref.json: per-bit flip counts and adjacent-pair (AND) counts of the unmodified reference
(oracle/_ref/stim detect / sample) over N_REF shots, for the statistical-parity tests (0 < p < 1 circuits have no golden
bytes: the reference's seeded stream is not part of its API contract). Needs oracle/_ref/stim."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
STIM = os.path.join(ROOT, "oracle", "_ref", "stim")
N_REF = 1 << 24
CHUNK = 1 << 20

DENSE = """
R 0 1 2 3 4 5
RX 6 7
PAULI_CHANNEL_1(0.05, 0.1, 0.15) 0 1 6
PAULI_CHANNEL_2(0.01, 0.02, 0.03, 0.01, 0.02, 0.03, 0.01, 0.02, 0.03, 0.01, 0.02, 0.03, 0.01, 0.02, 0.03) 2 3 4 5
DEPOLARIZE1(0.3) 0 7
DEPOLARIZE2(0.4) 1 2
X_ERROR(0.6) 3
Y_ERROR(0.02) 4 4
E(0.2) X0 Y1 Z6
ELSE_CORRELATED_ERROR(0.5) X2 Z7
ELSE_CORRELATED_ERROR(0.25) Y5
HERALDED_ERASE(0.1) 0 1
HERALDED_PAULI_CHANNEL_1(0.05, 0.1, 0.15, 0.2) 2 6
CX 0 1 2 3
CZ 4 5
MPP(0.05) Z0*Z1 X6*X7
MXX(0.1) 6 7
MR(0.2) 2
M(0.03) 0 1 2 3 4 5
MX 6 7
MPAD(0.35) 0 1
DETECTOR rec[-1]
DETECTOR rec[-2]
DETECTOR rec[-3]
DETECTOR rec[-4]
DETECTOR rec[-5] rec[-6]
DETECTOR rec[-7] rec[-8] rec[-9]
DETECTOR rec[-10]
DETECTOR rec[-11]
DETECTOR rec[-12]
DETECTOR rec[-13]
DETECTOR rec[-14]
DETECTOR rec[-15]
DETECTOR rec[-16]
DETECTOR rec[-17]
DETECTOR rec[-18]
OBSERVABLE_INCLUDE(0) rec[-1] rec[-3] rec[-10]
OBSERVABLE_INCLUDE(1) rec[-5]
"""


def circuits():
    with open(os.path.join(ROOT, "tests", "golden", "circuits", "c2_surface_x_d5_r5.stim")) as f:
        c2 = f.read()
    with open(os.path.join(ROOT, "tests", "golden", "circuits", "c1_rep_d3_r10.stim")) as f:
        c1 = f.read()
    return {"c1_rep_d3_r10": ("detect", c1), "c2_surface_x_d5_r5": ("detect", c2), "dense_noise_detect": ("detect", DENSE),
            "dense_noise_sample": ("sample", DENSE)}


def counts_of(mode, text, n_bits):
    nb = (n_bits + 7) // 8
    single = np.zeros(n_bits, dtype=np.int64)
    pair = np.zeros(max(n_bits - 1, 0), dtype=np.int64)
    for i in range(N_REF // CHUNK):
        args = [STIM, mode, "--shots", str(CHUNK), "--out_format", "b8", "--seed", str(1000 + i)]
        if mode == "detect":
            args.append("--append_observables")
        raw = subprocess.run(args, input=text.encode(), capture_output=True, check=True).stdout
        bits = np.unpackbits(np.frombuffer(raw, dtype=np.uint8).reshape(CHUNK, nb), axis=1, bitorder="little")[:, :n_bits]
        single += bits.sum(axis=0, dtype=np.int64)
        pair += (bits[:, :-1] & bits[:, 1:]).sum(axis=0, dtype=np.int64)
    return single, pair


def main():
    import stim_b200

    out = {}
    for name, (mode, text) in circuits().items():
        c = stim_b200.Circuit(text)
        n_bits = c.num_detectors + c.num_observables if mode == "detect" else c.num_measurements
        s, p = counts_of(mode, text, n_bits)
        out[name] = dict(mode=mode, n_ref=N_REF, n_bits=n_bits, single=s.tolist(), pair=p.tolist(),
                         circuit=None if name.startswith("c") else text)
        print(name, n_bits, "mean rate", float(s.mean()) / N_REF)
    with open(os.path.join(ROOT, "tests", "golden", "stats_ref.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
