"""Generates tests/golden/reference_outputs.json by running the UNMODIFIED reference CLI (oracle/_ref/stim, built by
oracle/Makefile from /root/reference) on deterministic circuits. Every expected output in the fixture is therefore
the reference's own answer; determinism (noise probabilities in {0, 1}, deterministic measurements) makes it
seed-independent, which the script double-checks by sampling with two seeds.

    python tools/gen_golden.py          (needs oracle/_ref/stim; the fixture itself is committed)

Circuits marked "src" are transcribed from the reference's own tests (file:line given).
"""
import base64
import json
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STIM = os.path.join(ROOT, "oracle", "_ref", "stim")

NOISE_PREFIXES = ("X_ERROR", "Y_ERROR", "Z_ERROR", "DEPOLARIZE", "PAULI_CHANNEL", "E(", "ELSE_CORRELATED_ERROR",
                  "CORRELATED_ERROR", "HERALDED", "I_ERROR", "II_ERROR")


def run(*args, stdin=""):
    r = subprocess.run([STIM, *args], input=stdin.encode(), capture_output=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr.decode())
    return r.stdout


def gen(code, task, d, r, **knobs):
    args = ["gen", "--code", code, "--task", task, "--distance", str(d), "--rounds", str(r)]
    for k, v in knobs.items():
        args += [f"--{k}", str(v)]
    return run(*args).decode()


def classical_circuit(rng, n, depth, basis):
    """Random reversible-classical circuit with probability-1 errors: every measurement is deterministic."""
    reset, meas, err = {"Z": ("R", "M", "X_ERROR"), "X": ("RX", "MX", "Z_ERROR"), "Y": ("RY", "MY", "X_ERROR")}[basis]
    lines = [f"{reset} " + " ".join(map(str, range(n)))]
    n_meas = 0
    for _ in range(depth):
        k = rng.randrange(6)
        a, b = rng.sample(range(n), 2)
        if k == 0:
            lines.append(f"{err}(1) {a}")
        elif k == 1 and basis == "Z":
            lines.append(f"CX {a} {b}")
        elif k == 1 and basis == "X":
            lines.append(f"XCZ {a} {b}" if rng.random() < 0.5 else f"CX {b} {a}")
        elif k == 2:
            lines.append(f"SWAP {a} {b}")
        elif k == 3:
            lines.append(f"{meas} {a}")
            n_meas += 1
            if n_meas >= 2 and rng.random() < 0.7:
                lines.append(f"DETECTOR rec[-1] rec[-{rng.randrange(1, n_meas + 1)}]")
        elif k == 4:
            lines.append({"Z": "MR", "X": "MRX", "Y": "MRY"}[basis] + f" {a}")
            n_meas += 1
            lines.append("DETECTOR rec[-1]")
        elif k == 5 and n_meas and basis == "Z":
            lines.append(f"CX rec[-{rng.randrange(1, n_meas + 1)}] {a}")
    lines.append(f"{meas} " + " ".join(map(str, range(n))))
    n_meas += n
    for i in range(n):
        lines.append(f"DETECTOR rec[-{i + 1}]")
    lines.append(f"OBSERVABLE_INCLUDE(0) rec[-1] rec[-{n}]")
    lines.append(f"OBSERVABLE_INCLUDE(2) rec[-2]")
    return "\n".join(lines) + "\n"


CASES = []


def detect_case(name, circuit, src=None, formats=("b8",), shots=70, flags=("--append_observables",)):
    CASES.append(dict(name=name, circuit=circuit, src=src, mode="detect", formats=list(formats), shots=shots, flags=list(flags)))


def sample_case(name, circuit, src=None, formats=("b8",), shots=70):
    CASES.append(dict(name=name, circuit=circuit, src=src, mode="sample", formats=list(formats), shots=shots, flags=[]))


ALL_FORMATS = ("01", "b8", "r8", "hits", "dets")

# --- transcribed from the reference's tests -------------------------------------------------------------------
sample_case("x_then_measure", "X 0\nM 1\nM 0\nM 2\nM 3\n", src="src/stim/simulators/frame_simulator.test.cc:289-330",
            formats=ALL_FORMATS + ("ptb64",), shots=64)
sample_case("big_circuit_measurements", "".join(f"X {k}\n" for k in range(0, 1250, 3)) + "".join(f"M {k}\n" for k in range(1250)),
            src="src/stim/simulators/frame_simulator.test.cc:332-371", formats=("01", "b8"), shots=96)
sample_case("run_length_formats", "X 100 500 501 551 1200\n" + "".join(f"M {k}\n" for k in range(1250)),
            src="src/stim/simulators/frame_simulator.test.cc:373-407", formats=("b8", "hits", "dets", "r8"), shots=3)
detect_case("correlated_error_chain", """
E(1) X0
ELSE_CORRELATED_ERROR(1) X1
ELSE_CORRELATED_ERROR(1) X2
E(0) X3
ELSE_CORRELATED_ERROR(1) X4
ELSE_CORRELATED_ERROR(1) X5
M 0 1 2 3 4 5
DETECTOR rec[-6]
DETECTOR rec[-5]
DETECTOR rec[-4]
DETECTOR rec[-3]
DETECTOR rec[-2]
DETECTOR rec[-1]
""", src="src/stim/simulators/frame_simulator.test.cc:425-549 (E / ELSE_CORRELATED_ERROR chains)", formats=ALL_FORMATS)
detect_case("classical_control", """
X_ERROR(1) 0
M !0
CX rec[-1] 1
CY rec[-1] 2
CZ rec[-1] 3
M 1 2 3
RX 4
CZ 4 rec[-4]
MX 4
DETECTOR rec[-1]
DETECTOR rec[-2]
DETECTOR rec[-3]
DETECTOR rec[-4]
DETECTOR rec[-5]
OBSERVABLE_INCLUDE(1) rec[-4] rec[-1]
""", src="src/stim/simulators/frame_simulator.test.cc:645-825 (classical control)", formats=ALL_FORMATS)
detect_case("repeat_block_10000", """
X_ERROR(1) 0
REPEAT 10000 {
    M 0
    DETECTOR rec[-1]
    X_ERROR(1) 0
}
M 0
OBSERVABLE_INCLUDE(0) rec[-1]
""", src="src/stim/simulators/frame_simulator.test.cc:861-976 (REPEAT 10000)", formats=("b8", "r8"), shots=40)
detect_case("detector_sampler_pybind", """
X_ERROR(1) 0
M 0 1
DETECTOR rec[-1]
DETECTOR rec[-2]
OBSERVABLE_INCLUDE(3) rec[-2]
""", src="src/stim/py/compiled_detector_sampler_pybind_test.py:23-80", formats=ALL_FORMATS + ("ptb64",), shots=64,
            flags=("--append_observables",))
detect_case("detector_sampler_prepend", """
X_ERROR(1) 0
M 0 1
DETECTOR rec[-1]
DETECTOR rec[-2]
OBSERVABLE_INCLUDE(3) rec[-2]
""", src="src/stim/cmd/command_detect.test.cc:21-183 (--prepend_observables)", formats=("01", "b8", "dets"), shots=5,
            flags=("--prepend_observables",))
detect_case("detector_sampler_no_obs", """
X_ERROR(1) 0
M 0 1
DETECTOR rec[-1]
DETECTOR rec[-2]
OBSERVABLE_INCLUDE(3) rec[-2]
""", src="src/stim/cmd/command_detect.test.cc:21-183", formats=("01", "b8", "hits", "r8"), shots=5, flags=())
detect_case("measure_reset_bases", """
RX 0
RY 1
R 2
Z_ERROR(1) 0
X_ERROR(1) 1 2
MX 0
MY 1
M 2
DETECTOR rec[-3]
DETECTOR rec[-2]
DETECTOR rec[-1]
MRX 0
MRY 1
MR 2
DETECTOR rec[-3]
DETECTOR rec[-2]
DETECTOR rec[-1]
Y_ERROR(1) 0 1 2
MX 0
MY 1
M 2
DETECTOR rec[-3]
DETECTOR rec[-2]
DETECTOR rec[-1]
MR 2 2
DETECTOR rec[-2]
DETECTOR rec[-1]
M(1) 2
MR(1) 2
MX(0) 0
DETECTOR rec[-3]
DETECTOR rec[-2]
DETECTOR rec[-1]
""", src="src/stim/simulators/frame_simulator.test.cc:978-1175 (M/MR/R in X/Y/Z bases, repeated targets)", formats=ALL_FORMATS)
detect_case("mpad_mxx_mpp", """
R 0 1 2 3
X_ERROR(1) 0
MPAD 0 1 1
MZZ 0 1 2 3
MPP Z0*Z1 Z2*Z3 Z0
RX 4 5
Z_ERROR(1) 4
MXX 4 5
MPP X4*X5
RY 6 7
X_ERROR(1) 6
MYY 6 7
MPP Y6*Y7 !Y6*Y7
DETECTOR rec[-1]
DETECTOR rec[-2]
DETECTOR rec[-3]
DETECTOR rec[-4]
DETECTOR rec[-5]
DETECTOR rec[-6]
DETECTOR rec[-7]
DETECTOR rec[-8]
DETECTOR rec[-9]
DETECTOR rec[-10]
DETECTOR rec[-11] rec[-12] rec[-13]
""", src="src/stim/simulators/frame_simulator.test.cc:1489-1542 (MPAD, MXX/MYY/MZZ)", formats=ALL_FORMATS)
detect_case("observable_pauli_targets", """
R 0 1
RX 2
X_ERROR(1) 0
Z_ERROR(1) 2
OBSERVABLE_INCLUDE(0) Z0
OBSERVABLE_INCLUDE(1) Z1
OBSERVABLE_INCLUDE(2) X2
OBSERVABLE_INCLUDE(3) Z0 Z1
M 0
DETECTOR rec[-1]
""", src="src/stim/simulators/frame_simulator.test.cc:1685-1743 (OBSERVABLE_INCLUDE Pauli targets)", formats=("01", "b8"))
detect_case("pauli_channels_deterministic", """
R 0 1 2 3 4 5
PAULI_CHANNEL_1(1, 0, 0) 0
PAULI_CHANNEL_1(0, 1, 0) 1
PAULI_CHANNEL_1(0, 0, 1) 2
PAULI_CHANNEL_2(0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0) 3 4
X_ERROR(0) 5
Y_ERROR(1) 5
M 0 1 2 3 4 5
DETECTOR rec[-6]
DETECTOR rec[-5]
DETECTOR rec[-4]
DETECTOR rec[-3]
DETECTOR rec[-2]
DETECTOR rec[-1]
""", src="single-term PAULI_CHANNEL is deterministic (SURVEY 8c)", formats=("01", "b8"))

# --- deterministic families generated by the reference's own generators (flip knobs = 1) ----------------------------
for code, task, d, r in [("repetition_code", "memory", 3, 4), ("repetition_code", "memory", 9, 30),
                         ("surface_code", "rotated_memory_x", 3, 3), ("surface_code", "rotated_memory_z", 5, 6),
                         ("surface_code", "unrotated_memory_x", 3, 2), ("surface_code", "unrotated_memory_z", 5, 3),
                         ("color_code", "memory_xyz", 3, 3), ("color_code", "memory_xyz", 7, 5),
                         ("surface_code", "rotated_memory_z", 11, 11)]:
    for knob in ("before_measure_flip_probability", "after_reset_flip_probability"):
        detect_case(f"{code}_{task}_d{d}_r{r}_{knob}", gen(code, task, d, r, **{knob: 1}),
                    src="generated: stim gen (knob=1 emits X_ERROR(1)/Z_ERROR(1), circuit_gen_params.cc:7-15,59-80)",
                    formats=("b8",), shots=130)
detect_case("surface_d5_all_formats", gen("surface_code", "rotated_memory_x", 5, 4, before_measure_flip_probability=1),
            formats=ALL_FORMATS + ("ptb64",), shots=128)

# --- BASELINE.json config 4, MPP variant with deterministic noise (tools/gen_variants.py): b8 and ptb64 bytes -------------
with open(os.path.join(ROOT, "tests", "golden", "circuits", "c4v_color_d15_r15_mpp_det.stim")) as _f:
    detect_case("c4v_color_d15_r15_mpp_deterministic", _f.read(),
                src="generated color_code:memory_xyz d=15 r=15 with the ancilla cycle replaced by MPP (tools/gen_variants.py)",
                formats=("b8", "ptb64"), shots=128)

# --- random reversible-classical circuits -------------------------------------------------------------------------
_rng = random.Random(20240601)
for basis in "ZXY":
    for i in range(4):
        detect_case(f"classical_{basis}_{i}", classical_circuit(_rng, 6 + 3 * i, 60 + 40 * i, basis), formats=("b8",), shots=33)
for i in range(3):
    c = classical_circuit(_rng, 8, 80, "Z")
    sample_case(f"classical_sample_{i}", c, formats=("b8", "01"), shots=65)


# --- every Clifford gate against the reference (restates frame_simulator.test.cc:58-104 through the CLI) -----------
# G^60 is the identity for every named Clifford, so "prepare in basis B, apply G 59 times, flip with a probability-1
# Pauli P, apply G once more, measure in basis B" has deterministic detectors that spell out whether G P G^-1 anticommutes
# with B on each target: over B in {X, Y, Z} and P in {X, Y, Z} on each target this pins the whole frame action of G.
CLIFFORD_1Q = ["I", "X", "Y", "Z", "H", "H_XY", "H_YZ", "H_NXY", "H_NXZ", "H_NYZ", "S", "S_DAG", "SQRT_X", "SQRT_X_DAG", "SQRT_Y",
               "SQRT_Y_DAG", "C_XYZ", "C_ZYX", "C_NXYZ", "C_XNYZ", "C_XYNZ", "C_NZYX", "C_ZNYX", "C_ZYNX"]
CLIFFORD_2Q = ["CX", "CY", "CZ", "XCX", "XCY", "XCZ", "YCX", "YCY", "YCZ", "SWAP", "ISWAP", "ISWAP_DAG", "CXSWAP", "SWAPCX", "CZSWAP",
               "SQRT_XX", "SQRT_XX_DAG", "SQRT_YY", "SQRT_YY_DAG", "SQRT_ZZ", "SQRT_ZZ_DAG", "II"]
SPP_PRODUCTS = ["X0", "Y0", "Z0", "X0*X1", "Z0*Z1", "Y0*Y1", "X0*Z1", "X0*Y1*Z2", "Z0*Y1*X2"]


def clifford_probe(apply_line, n_targets):
    """apply_line(q0) -> instruction text acting on qubits q0 .. q0 + n_targets - 1."""
    prep = {"Z": ("R", "M"), "X": ("RX", "MX"), "Y": ("RY", "MY")}
    lines, q = [], 0
    for basis in "ZXY":
        for victim in range(n_targets):
            for pauli in "XYZ":
                qs = " ".join(str(q + t) for t in range(n_targets))
                lines.append(f"{prep[basis][0]} {qs}")
                lines.append("REPEAT 59 {\n    " + apply_line(q) + "\n}")
                lines.append(f"{pauli}_ERROR(1) {q + victim}")
                lines.append(apply_line(q))
                lines.append(f"{prep[basis][1]} {qs}")
                for t in range(n_targets):
                    lines.append(f"DETECTOR rec[-{n_targets - t}]")
                q += n_targets
    return "\n".join(lines) + "\n"


def _reference_knows(line):
    try:
        run("detect", "--shots", "1", stdin=line + "\n")
        return True
    except RuntimeError:
        return False


if os.path.exists(STIM):
    for g in CLIFFORD_1Q:
        if _reference_knows(f"{g} 0"):
            detect_case(f"clifford_probe_{g}", clifford_probe(lambda q, g=g: f"{g} {q}", 1),
                        src="src/stim/simulators/frame_simulator.test.cc:58-104 (gate action vs tableau), as CLI probes", shots=5)
    for g in CLIFFORD_2Q:
        if _reference_knows(f"{g} 0 1"):
            detect_case(f"clifford_probe_{g}", clifford_probe(lambda q, g=g: f"{g} {q} {q + 1}", 2),
                        src="src/stim/simulators/frame_simulator.test.cc:58-104 (gate action vs tableau), as CLI probes", shots=5)
    for name in ("SPP", "SPP_DAG"):
        for prod in SPP_PRODUCTS:
            n = prod.count("*") + 1

            def line(q, prod=prod, name=name):
                import re as _re
                return name + " " + _re.sub(r"([XYZ])(\d)", lambda m: m.group(1) + str(q + int(m.group(2))), prod)

            detect_case(f"clifford_probe_{name}_{prod.replace('*', '')}", clifford_probe(line, n),
                        src="src/stim/simulators/frame_simulator.inl:705-716 + gate_decomposition.cc:163-243, as CLI probes", shots=5)


def noiseless(circuit):
    return "\n".join(ln for ln in circuit.split("\n") if not ln.strip().upper().startswith(NOISE_PREFIXES)) + "\n"


def main():
    out = []
    for c in CASES:
        entry = dict(name=c["name"], src=c["src"], circuit=c["circuit"], mode=c["mode"], shots=c["shots"], flags=c["flags"],
                     outputs={})
        for fmt in c["formats"]:
            shots = c["shots"]
            if fmt == "ptb64":
                shots = (shots + 63) // 64 * 64
            args = [c["mode"], "--shots", str(shots), "--out_format", fmt, *c["flags"]]
            a = run(*args, "--seed", "1", stdin=c["circuit"])
            b = run(*args, "--seed", "2", stdin=c["circuit"])
            if a != b:
                raise RuntimeError(f"case {c['name']} is not deterministic in format {fmt}")
            entry["outputs"][fmt] = dict(shots=shots, data=base64.b64encode(a).decode())
        if c["mode"] == "sample":
            ref = run("sample", "--shots", "1", "--out_format", "01", "--seed", "3", stdin=noiseless(c["circuit"])).decode().strip()
            ref2 = run("sample", "--shots", "1", "--out_format", "01", "--seed", "4", stdin=noiseless(c["circuit"])).decode().strip()
            assert ref == ref2
            entry["reference_sample"] = ref
        out.append(entry)
    path = os.path.join(ROOT, "tests", "golden", "reference_outputs.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print(f"wrote {len(out)} cases to {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    if not os.path.exists(STIM):
        sys.exit("oracle/_ref/stim is missing: run `make -C oracle ref` first")
    main()
