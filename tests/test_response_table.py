"""CPU tests of the event engine's response table (stim_b200/csrc/response.cc, host only): every (site, outcome) entry
the library derives by BACKWARD sensitivity propagation must equal the set of output bits that flip when that single
event is injected into the FORWARD frame oracle (oracle/frame_oracle.py, pinned against the reference CLI's goldens)."""
import os

import numpy as np
import pytest

import stim_b200
from conftest import ROOT, gen_circuit
from oracle import sparse_oracle as so
from test_gpu_parity import ALL_OPS

LONG_CHAIN = "E(0.01) X0\n" + "".join(f"ELSE_CORRELATED_ERROR(0.01) X{k}\n" for k in range(1, 17)) + "M 0\nDETECTOR rec[-1]\n"


def check_table(text, mode="detectors", entries=None):
    t = stim_b200.response_table(text, mode)
    assert t["info"]["eligible"] == 1, t["info"]["why_not"]
    want = so.responses_by_injection(text, t, mode, entries)
    assert len(want) > 0
    bad = [(e, so.entry_ids(t, e), ids) for e, ids in want.items() if so.entry_ids(t, e) != ids]
    assert not bad, bad[:5]
    return t


@pytest.mark.parametrize("mode", ["detectors", "measurements"])
def test_every_instruction_matches_forward_injection(mode):
    t = check_table(ALL_OPS, mode)
    assert (t["classes"][:, 4] == 3).any()  # the E / ELSE_CORRELATED_ERROR chain is one categorical site
    assert t["info"]["overflow_words"] > 0  # responses longer than four bits go through the overflow list
    if mode == "measurements":
        assert (t["site_group"] & 0x80000000).any()  # collapse randomisation that reaches an output is a p = 1/2 site


GENERATED = [
    ("repetition_code", "memory", 3, 10, 0.02),
    ("surface_code", "rotated_memory_z", 3, 3, 0.02),
    ("surface_code", "rotated_memory_x", 5, 5, 0.01),
    ("surface_code", "unrotated_memory_z", 3, 2, 0.05),
    ("color_code", "memory_xyz", 3, 3, 0.02),
    ("color_code", "memory_xyz", 5, 2, 0.01),
]


@pytest.mark.parametrize("code,task,d,r,p", GENERATED)
def test_generated_circuits_match_forward_injection(code, task, d, r, p):
    t = check_table(gen_circuit(code, task, d, r, p))
    # QEC memory circuits have deterministic detectors: no collapse bit reaches an output
    assert not (t["site_group"] & 0x80000000).any()


def test_measurement_mode_of_a_memory_circuit():
    check_table(gen_circuit("surface_code", "rotated_memory_x", 3, 3, 0.02), "measurements")


def test_feedback_and_repeat_blocks():
    text = """
    R 0 1 2
    X_ERROR(0.1) 0 1
    REPEAT 4 {
        CX 0 2 1 2
        DEPOLARIZE2(0.05) 0 2
        MR 2
        CX rec[-1] 0
        CZ rec[-1] 1
        DETECTOR rec[-1]
        HERALDED_ERASE(0.1) 1
        DETECTOR rec[-1] rec[-2]
    }
    M 0 1
    DETECTOR rec[-1] rec[-2] rec[-4]
    OBSERVABLE_INCLUDE(0) rec[-1]
    OBSERVABLE_INCLUDE(1) X0 Z1
    """
    check_table(text)
    check_table(text, "measurements")


def test_headline_circuit_table():
    """c3 (d = 25): table statistics, and a random sample of entries against forward injection."""
    with open(os.path.join(ROOT, "tests", "golden", "circuits", "c3_surface_z_d25_r25.stim")) as f:
        text = f.read()
    t = stim_b200.response_table(text)
    info = t["info"]
    assert info["eligible"] == 1 and info["num_sites"] == 124299 and info["num_classes"] == 3
    assert abs(info["events_per_shot"] - 124.299) < 0.01 and info["max_response"] == 4 and info["overflow_words"] == 0
    rng = np.random.default_rng(5)
    pick = set(int(v) for v in rng.choice(int(info["num_entries"]), size=384, replace=False))
    want = so.responses_by_injection(text, t, "detectors", pick)
    assert len(want) == 384
    for e, ids in want.items():
        assert so.entry_ids(t, e) == ids, e


def test_else_chains():
    """A chain is one site with outcome probabilities p_i prod_{j<i} (1 - p_j); more than 16 elements are not eligible."""
    text = "E(0.25) X0\nELSE_CORRELATED_ERROR(0.5) X1\nELSE_CORRELATED_ERROR(1) X2\nM 0 1 2\nDETECTOR rec[-3]\nDETECTOR rec[-2]\nDETECTOR rec[-1]\n"
    t = check_table(text)
    c = t["classes"][0]
    assert len(t["classes"]) == 1 and c[4] == 3 and c[5] == 3 and c[2] == 0  # always fires (the last element has p = 1)
    assert abs(int(c[6]) / 2**32 - 0.25) < 1e-6 and abs(int(c[7]) / 2**32 - 0.625) < 1e-6
    slices = np.array([[0, 128, int(c[22]), 0]], dtype=np.uint32)
    out = so.sample(t, slices, 128, seed=3, first_shot=0, n_shots=1 << 13, n_outputs=3)
    assert (out.sum(axis=1) == 1).all()  # exactly one element applies in every shot
    assert np.abs(out.mean(axis=0) - [0.25, 0.375, 0.375]).max() < 0.02
    t = stim_b200.response_table(LONG_CHAIN)
    assert t["info"]["eligible"] == 0 and "ELSE_CORRELATED_ERROR" in t["info"]["why_not"]


def test_oracle_sampler_statistics_on_a_tiny_circuit():
    """The sampler restatement itself: flip rates of a two-site toy table match the probabilities."""
    text = "X_ERROR(0.25) 0\nM 0\nDETECTOR rec[-1]\nX_ERROR(0.5) 1\nM 1\nDETECTOR rec[-1]\n"
    t = stim_b200.response_table(text)
    # slices as the engine builds them for a 128-shot tile: one slice per class covering all its sites
    slices = np.array([[ci, int(c[21]) * 128, int(c[22]), 0] for ci, c in enumerate(t["classes"])], dtype=np.uint32)
    out = so.sample(t, slices, 128, seed=7, first_shot=0, n_shots=1 << 14, n_outputs=2)
    rates = out.mean(axis=0)
    assert abs(rates[0] - 0.25) < 0.02 and abs(rates[1] - 0.5) < 0.02


def test_dense_classes_use_packed_bernoulli_words():
    """p = 1/2 and 1/4 need one and two random words per 32 trials, float32(0.3) 24 of them (still cheaper than 0.3 draws per
    trial); p = 0.1 stays with the geometric walk; collapse bits that reach an output are fair coins. The oracle's word
    method has the right rates."""
    text = "X_ERROR(0.5) 0\nX_ERROR(0.25) 1\nX_ERROR(0.3) 2\nDEPOLARIZE1(0.75) 3\nX_ERROR(0.1) 4\nM 0 1 2 3 4\n" + "".join(
        f"DETECTOR rec[-{k}]\n" for k in (5, 4, 3, 2, 1))
    t = check_table(text)
    by_p = {round(-np.expm1(-float(int(c[0]) | (int(c[1]) << 32)) * 2.0**-56), 3): int(c[23]) for c in t["classes"]}
    assert by_p[0.5] == 1 << 31 and by_p[0.25] == 1 << 30 and by_p[0.3] == 0x4CCCCD00 and by_p[0.75] == 3 << 30 and by_p[0.1] == 0
    slices = np.array([[ci, int(c[21]) * 128, int(c[22]), 0] for ci, c in enumerate(t["classes"])], dtype=np.uint32)
    out = so.sample(t, slices, 128, seed=11, first_shot=0, n_shots=1 << 13, n_outputs=5)
    np.testing.assert_allclose(out.mean(axis=0), [0.5, 0.25, 0.3, 0.5, 0.1], atol=0.025)  # DEPOLARIZE1(0.75): X or Y flips the result
    m = stim_b200.response_table("H 0\nM 0\nDETECTOR rec[-1]\n")
    assert len(m["classes"]) == 1 and int(m["classes"][0][23]) == 1 << 31 and (m["site_group"] & 0x80000000).all()


@pytest.mark.parametrize("name", ["c2_surface_x_d5_r5", "c3_surface_z_d25_r25"])
def test_round_structure_folds_the_table_exactly(name):
    """The periods the device table is folded by (response.h ResponsePeriod): every site of a folded range equals its image
    in the first period with the detector ids shifted; c3 folds 8x (one of 25 rounds is stored)."""
    with open(os.path.join(ROOT, "tests", "golden", "circuits", name + ".stim")) as f:
        text = f.read()
    t = stim_b200.response_table(text)
    D = stim_b200.Circuit(text).num_detectors
    rng = np.random.default_rng(1)
    kept = 0
    for c, (a, p, n, delta) in zip(t["classes"], t["periods"].astype(np.int64)):
        n_out, n_sites, entry0 = int(c[5]), int(c[21]), int(c[22])
        kept += n_out * (n_sites - (n - p if p else 0))
        if p == 0:
            continue
        assert n >= 3 * p and a + n <= n_sites and delta > 0
        sites = np.arange(a + p, a + n)
        if sites.size > 3000:
            sites = np.concatenate([sites[:50], sites[-50:], rng.choice(sites, 2900, replace=False)])
        for s in sites:
            k, img = divmod(int(s) - a, p)
            for o in range(n_out):
                want = [v + k * delta if v < D else v for v in so.entry_ids(t, entry0 + (a + img) * n_out + o)]
                assert so.entry_ids(t, entry0 + int(s) * n_out + o) == want, (int(c[4]), int(s), o)
    if name.startswith("c3"):
        assert kept * 8 < len(t["entries"])


def _random_circuit(rng, n_qubits=6, n_ops=60):
    """Random circuits over every kind of instruction the lowering knows (gates in random order, resets and measurements in
    all bases with repeated targets, feedback, MPP / SPP / pair measurements, every noise channel, heralded errors, E / ELSE
    chains, REPEAT blocks, detectors and observables with record and Pauli targets)."""
    one = ["H", "S", "S_DAG", "SQRT_X", "SQRT_X_DAG", "SQRT_Y", "SQRT_Y_DAG", "H_XY", "H_YZ", "C_XYZ", "C_ZYX", "X", "Y", "Z", "I"]
    two = ["CX", "CY", "CZ", "XCX", "XCY", "XCZ", "YCX", "YCY", "YCZ", "SWAP", "ISWAP", "ISWAP_DAG", "CXSWAP", "SWAPCX", "CZSWAP",
           "SQRT_XX", "SQRT_XX_DAG", "SQRT_YY", "SQRT_YY_DAG", "SQRT_ZZ", "SQRT_ZZ_DAG"]
    meas = ["M", "MX", "MY", "MR", "MRX", "MRY"]
    lines, n_meas = [], 0
    q = lambda k=1: [int(v) for v in rng.choice(n_qubits, size=k, replace=False)]

    def emit(depth):
        nonlocal n_meas
        kind = rng.integers(0, 14)
        if kind == 0:
            lines.append(f"{rng.choice(one)} " + " ".join(map(str, q(rng.integers(1, 4)))))
        elif kind == 1:
            a = q(2)
            lines.append(f"{rng.choice(two)} {a[0]} {a[1]}")
        elif kind == 2:
            ts = [int(v) for v in rng.integers(0, n_qubits, size=rng.integers(1, 4))]  # (targets may repeat)
            name = rng.choice(meas)
            arg = f"({rng.choice([0.1, 0.25])})" if rng.random() < 0.3 else ""
            lines.append(f"{name}{arg} " + " ".join(("!" if rng.random() < 0.2 else "") + str(t) for t in ts))
            n_meas += len(ts)
        elif kind == 3:
            lines.append(f"{rng.choice(['R', 'RX', 'RY'])} " + " ".join(map(str, q(rng.integers(1, 3)))))
        elif kind == 4 and n_meas:
            k = int(rng.integers(1, min(n_meas, 5) + 1))
            g = rng.choice(["CX", "CY", "CZ"])
            lines.append(f"{g} rec[-{k}] {q()[0]}" if g != "CZ" or rng.random() < 0.5 else f"CZ {q()[0]} rec[-{k}]")
        elif kind == 5:
            a = q(3)
            ps = rng.choice(list("XYZ"), size=3)
            lines.append(f"MPP {ps[0]}{a[0]}*{ps[1]}{a[1]} {'!' if rng.random() < 0.3 else ''}{ps[2]}{a[2]}")
            n_meas += 2
        elif kind == 6:
            a = q(2)
            ps = rng.choice(list("XYZ"), size=2)
            lines.append(f"{rng.choice(['SPP', 'SPP_DAG'])} {ps[0]}{a[0]}*{ps[1]}{a[1]}")
        elif kind == 7:
            a = q(2)
            lines.append(f"{rng.choice(['MXX', 'MYY', 'MZZ'])} {a[0]} {a[1]}")
            n_meas += 1
        elif kind == 8:
            name = rng.choice(["X_ERROR", "Y_ERROR", "Z_ERROR", "DEPOLARIZE1"])
            lines.append(f"{name}({rng.choice([0.01, 0.3, 0.5])}) " + " ".join(map(str, q(rng.integers(1, 4)))))
        elif kind == 9:
            a = q(2)
            if rng.random() < 0.5:
                lines.append(f"DEPOLARIZE2({rng.choice([0.02, 0.4])}) {a[0]} {a[1]}")
            else:
                ps = rng.dirichlet(np.ones(15)) * 0.3
                lines.append("PAULI_CHANNEL_2(" + ", ".join(f"{v:.4f}" for v in ps) + f") {a[0]} {a[1]}")
        elif kind == 10:
            r = rng.random()
            if r < 0.4:
                lines.append(f"PAULI_CHANNEL_1(0.05, 0.1, 0.02) {q()[0]}")
            elif r < 0.7:
                lines.append(f"HERALDED_ERASE(0.2) {q()[0]}")
                n_meas += 1
            else:
                lines.append(f"HERALDED_PAULI_CHANNEL_1(0.02, 0.05, 0.1, 0.03) {q()[0]}")
                n_meas += 1
        elif kind == 11:
            a = q(2)
            lines.append(f"E(0.1) X{a[0]} Z{a[1]}")
            for _ in range(rng.integers(0, 3)):
                lines.append(f"ELSE_CORRELATED_ERROR(0.2) {rng.choice(list('XYZ'))}{q()[0]}")
        elif kind == 12 and n_meas:
            ks = sorted({int(v) for v in rng.integers(1, min(n_meas, 6) + 1, size=rng.integers(1, 4))})
            if rng.random() < 0.7:
                lines.append("DETECTOR " + " ".join(f"rec[-{k}]" for k in ks))
            else:
                extra = f" {rng.choice(list('XYZ'))}{q()[0]}" if rng.random() < 0.5 else ""
                lines.append(f"OBSERVABLE_INCLUDE({rng.integers(0, 3)}) " + " ".join(f"rec[-{k}]" for k in ks) + extra)
        elif kind == 13 and depth == 0 and n_meas >= 2:
            reps = int(rng.integers(2, 4))
            lines.append(f"REPEAT {reps} {{")
            before, m0 = len(lines), n_meas
            for _ in range(rng.integers(2, 6)):
                emit(1)
            lines.append("}")
            n_meas = m0 + (n_meas - m0) * reps
            if before == len(lines) - 1:
                lines.insert(before, f"H {q()[0]}")

    for _ in range(n_ops):
        emit(0)
    lines.append("M " + " ".join(map(str, range(n_qubits))))
    lines.append("DETECTOR rec[-1] rec[-2]")
    lines.append("OBSERVABLE_INCLUDE(0) rec[-3]")
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("seed", range(24))
def test_random_circuits_match_forward_injection(seed):
    """Differential test of the backward pass against the forward frame oracle on random instruction sequences."""
    rng = np.random.default_rng(1000 + seed)
    text = _random_circuit(rng)
    for mode in ("detectors", "measurements"):
        t = stim_b200.response_table(text, mode)
        if not t["info"]["eligible"]:
            assert "ELSE_CORRELATED_ERROR" in t["info"]["why_not"] or "large" in t["info"]["why_not"]
            continue
        want = so.responses_by_injection(text, t, mode)
        bad = [(e, so.entry_ids(t, e), ids) for e, ids in want.items() if so.entry_ids(t, e) != ids]
        assert not bad, (mode, bad[:3], text)
