"""Golden vectors for the measurement -> detection-event converter: outputs of the unmodified reference CLI `stim m2d`
(oracle/_ref/stim) on seeded random measurement / sweep data. Writes tests/golden/m2d_cases.json.

    python tools/gen_m2d_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STIM = os.path.join(ROOT, "oracle", "_ref", "stim")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def stim(*args, stdin=b""):
    r = subprocess.run([STIM, *args], input=stdin, capture_output=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr.decode())
    return r.stdout


SWEEP_CIRCUIT = """
R 0 1 2 3
CX sweep[0] 0
CZ sweep[1] 1
H 1 3
CY sweep[2] 2
XCZ 3 sweep[3]
CX 0 1 2 3
M 0 1
CX rec[-1] 2
CZ sweep[5] 3 2 rec[-2]
MR 2 3
DETECTOR rec[-1] rec[-3]
DETECTOR rec[-2]
X_ERROR(0.25) 0
REPEAT 3 {
    H 0
    CX sweep[4] 0 0 1
    SPP X0*Z1 Z1
    MX 0
    M !1
    DETECTOR rec[-1] rec[-2]
    OBSERVABLE_INCLUDE(1) rec[-2]
}
MPP X0*Z1 !Y2
OBSERVABLE_INCLUDE(0) rec[-1] rec[-2]
OBSERVABLE_INCLUDE(2) X0 rec[-1]
DETECTOR
"""


def cases():
    from test_gpu_parity import ALL_OPS

    d3 = stim("gen", "--code", "surface_code", "--task", "rotated_memory_x", "--distance", "3", "--rounds", "3",
              "--after_clifford_depolarization", "0.01").decode()
    rep = stim("gen", "--code", "repetition_code", "--task", "memory", "--distance", "5", "--rounds", "4",
               "--before_measure_flip_probability", "0.1").decode()
    yield "sweep_feedback_repeat", SWEEP_CIRCUIT, 40
    yield "all_ops", ALL_OPS, 24
    yield "surface_x_d3_r3", d3, 32
    yield "repetition_d5_r4", rep, 17
    yield "no_detectors", "M 0 1\nOBSERVABLE_INCLUDE(0) rec[-1]\n", 5
    yield "empty", "", 3


def count(text, what):
    from oracle import frame_oracle as fo

    o = fo.FrameOracle(text, 0, 1, 1).run()
    sweep = 0
    for name, args, targets in fo.flatten(o.ops):
        for t in targets:
            if t != fo.T_COMB and (t & fo.T_SWEEP):
                sweep = max(sweep, (t & fo.T_VAL) + 1)
    return {"M": len(o.rec), "D": len(o.dets), "L": (max(o.obs) + 1) if o.obs else 0, "S": sweep}[what]


def to01(a):
    return ("".join("".join(str(int(v)) for v in row) + "\n" for row in a)).encode()


def main():
    rng = np.random.default_rng(20261018)
    out = []
    for name, text, shots in cases():
        M, S = count(text, "M"), count(text, "S")
        meas = rng.integers(0, 2, size=(shots, M), dtype=np.uint8)
        sweep = rng.integers(0, 2, size=(shots, S), dtype=np.uint8)
        entry = dict(name=name, circuit=text, shots=shots, measurements=to01(meas).decode(), sweep=to01(sweep).decode(), outputs={})
        with tempfile.TemporaryDirectory() as tmp:
            cpath, spath = os.path.join(tmp, "c.stim"), os.path.join(tmp, "s.01")
            open(cpath, "w").write(text)
            open(spath, "wb").write(to01(sweep))
            for skip in (False, True):
                for use_sweep in ((False, True) if S else (False,)):
                    args = ["m2d", "--in_format", "01", "--out_format", "01", "--circuit", cpath, "--append_observables"]
                    if skip:
                        args.append("--skip_reference_sample")
                    if use_sweep:
                        args += ["--sweep", spath, "--sweep_format", "01"]
                    key = f"skip={int(skip)},sweep={int(use_sweep)}"
                    entry["outputs"][key] = stim(*args, stdin=to01(meas)).decode()
        out.append(entry)
    path = os.path.join(ROOT, "tests", "golden", "m2d_cases.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print(f"wrote {len(out)} cases to {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    if not os.path.exists(STIM):
        sys.exit("oracle/_ref/stim is missing: run `make -C oracle ref` first")
    main()
