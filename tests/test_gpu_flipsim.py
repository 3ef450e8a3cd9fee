"""GPU tests of the interactive flip simulator (stim_b200.FlipSimulator over stim_b200/csrc/flipsim.cu).

* every deterministic golden case of the reference CLI (tests/golden/reference_outputs.json): one `do` of the whole
  circuit reproduces the reference's detection events / measurement flips, with and without stabilizer randomization,
  and so does feeding the circuit one instruction at a time (the interactive use the API exists for);
* the frame oracle on a circuit with every instruction, randomization off, noise at p in {0, 1};
* the scenarios of the reference's own tests (/root/reference/src/stim/simulators/frame_simulator_pybind_test.py:8-640)
  restated against this mirror: sizes, indexing and errors, Pauli frames, broadcast_pauli_errors, Bernoulli samples;
* statistics of the noisy instructions against the bulk sampler."""
import numpy as np
import pytest

import stim_b200
from conftest import gen_circuit
from golden_util import case_ids, expected_bits, load_cases
from oracle import frame_oracle as fo

pytestmark = pytest.mark.gpu

CASES = load_cases()
DETECT = [c for c in CASES if c["mode"] == "detect"]
SAMPLE = [c for c in CASES if c["mode"] == "sample"]


def top_level_pieces(text):
    """The circuit cut into top-level instructions / REPEAT blocks."""
    pieces, depth, cur = [], 0, []
    for ln in text.split("\n"):
        s = ln.split("#", 1)[0].strip()
        if not s:
            continue
        cur.append(ln)
        depth += s.count("{") - s.count("}")
        if depth == 0:
            pieces.append("\n".join(cur))
            cur = []
    return pieces


@pytest.mark.parametrize("case", DETECT, ids=case_ids(DETECT))
def test_detector_flips_reproduce_reference(case):
    shots = next(iter(case["outputs"].values()))["shots"]
    circ = stim_b200.Circuit(case["circuit"])
    D, L = circ.num_detectors, circ.num_observables
    flags = case["flags"]
    n_cols = D + (L if ("--append_observables" in flags or "--prepend_observables" in flags) else 0)
    want = expected_bits(case, n_cols)[:shots]
    for randomize, piecewise in ((False, False), (True, False), (True, True)):
        s = stim_b200.FlipSimulator(batch_size=shots, disable_stabilizer_randomization=not randomize, seed=5)
        if piecewise:
            for piece in top_level_pieces(case["circuit"]):
                s.do(piece)
        else:
            s.do(case["circuit"])
        assert s.num_detectors == D and s.num_measurements == circ.num_measurements
        dets = s.get_detector_flips().T.astype(np.uint8).reshape(shots, D)
        obs = np.zeros((shots, L), dtype=np.uint8)
        obs[:, :s.num_observables] = s.get_observable_flips().T
        got = np.concatenate([dets, obs], axis=1) if "--append_observables" in flags else np.concatenate(
            [obs, dets], axis=1) if "--prepend_observables" in flags else dets
        np.testing.assert_array_equal(got, want, err_msg=f"{case['name']} {randomize} {piecewise}")


@pytest.mark.parametrize("case", SAMPLE, ids=case_ids(SAMPLE))
def test_measurement_flips_reproduce_reference(case):
    shots = next(iter(case["outputs"].values()))["shots"]
    ref = np.array([int(ch) for ch in case["reference_sample"]], dtype=np.uint8)
    want = expected_bits(case, ref.size)[:shots] ^ ref[None, :]
    for randomize in (False, True):
        s = stim_b200.FlipSimulator(batch_size=shots, disable_stabilizer_randomization=not randomize, seed=9)
        s.do(case["circuit"])
        np.testing.assert_array_equal(s.get_measurement_flips().T.astype(np.uint8), want, err_msg=case["name"])


def test_every_instruction_matches_the_frame_oracle_without_randomness():
    from test_gpu_parity import ALL_OPS
    import re

    # noise made deterministic: every probability argument 1 or 0 alternately (single-argument channels only; the
    # multi-argument channels are switched off)
    k = [0]

    def repl(m):
        k[0] += 1
        return f"{m.group(1)}({k[0] % 2})"

    text = re.sub(r"\b(X_ERROR|Y_ERROR|Z_ERROR|DEPOLARIZE1|DEPOLARIZE2|E|ELSE_CORRELATED_ERROR|I_ERROR|II_ERROR)\(([0-9.]+)\)", repl, ALL_OPS)
    text = re.sub(r"HERALDED_ERASE\([^)]*\)", "HERALDED_ERASE(0)", text)  # (an erasure applies a RANDOM Pauli even at p = 1)
    text = re.sub(r"\b(MR|MRY|M|MPP|MYY|MPAD)\((0\.[0-9]+)\)", lambda m: f"{m.group(1)}(1)", text)
    text = re.sub(r"PAULI_CHANNEL_1\([^)]*\)", "PAULI_CHANNEL_1(0, 1, 0)", text)
    text = re.sub(r"PAULI_CHANNEL_2\([^)]*\)", "PAULI_CHANNEL_2(0,0,0,0,0,0,1,0,0,0,0,0,0,0,0)", text)
    text = re.sub(r"HERALDED_PAULI_CHANNEL_1\([^)]*\)", "HERALDED_PAULI_CHANNEL_1(0, 0, 0, 1)", text)
    text = text.replace("DEPOLARIZE1(1)", "DEPOLARIZE1(0)").replace("DEPOLARIZE2(1)", "DEPOLARIZE2(0)")
    shots = 70

    class Quiet(fo.FrameOracle):
        randomize = False

    o = Quiet(text, 0, 1, 1).run()
    s = stim_b200.FlipSimulator(batch_size=shots, disable_stabilizer_randomization=True, seed=1)
    for piece in top_level_pieces(text):
        s.do(piece)
    np.testing.assert_array_equal(s.get_measurement_flips().T.astype(np.uint8), o.measurement_flips()[:shots])
    np.testing.assert_array_equal(s.get_detector_flips().T.astype(np.uint8), o.detectors()[:shots])
    np.testing.assert_array_equal(s.get_observable_flips().T.astype(np.uint8), o.observables()[:shots])


def test_sizes_indexing_and_errors():
    s = stim_b200.FlipSimulator(batch_size=11)
    assert (s.num_measurements, s.num_qubits, s.batch_size) == (0, 0, 11)
    s.do("X_ERROR(1) 100\nM 100")
    assert (s.num_measurements, s.num_qubits, s.batch_size) == (1, 101, 11)
    np.testing.assert_array_equal(s.get_measurement_flips(record_index=0), [True] * 11)
    s.do("X_ERROR(1) 25\nM 24 25\nDETECTOR rec[-1]\nDETECTOR rec[-2]\nDETECTOR rec[-1] rec[-3]\nOBSERVABLE_INCLUDE(2) rec[-1]")
    assert (s.num_measurements, s.num_detectors, s.num_observables) == (3, 3, 3)
    np.testing.assert_array_equal(s.get_detector_flips(), [[True] * 11, [False] * 11, [False] * 11])
    np.testing.assert_array_equal(s.get_detector_flips(detector_index=-3), [True] * 11)
    assert s.get_detector_flips(detector_index=1, instance_index=-1) is False
    np.testing.assert_array_equal(s.get_detector_flips(instance_index=4), [True, False, False])
    np.testing.assert_array_equal(s.get_observable_flips(), [[False] * 11, [False] * 11, [True] * 11])
    np.testing.assert_array_equal(s.get_measurement_flips(bit_packed=True), [[0xFF, 0x07], [0, 0], [0xFF, 0x07]])
    np.testing.assert_array_equal(s.get_measurement_flips(record_index=-1, bit_packed=True), [0xFF, 0x07])
    for bad in (dict(record_index=3), dict(record_index=-4), dict(instance_index=11), dict(instance_index=-12)):
        with pytest.raises(IndexError):
            s.get_measurement_flips(**bad)
    with pytest.raises(IndexError):  # a lookback before the first measurement
        stim_b200.FlipSimulator(batch_size=4).do("M 0\nDETECTOR rec[-2]")
    with pytest.raises(ValueError):
        s.do("NOT_A_GATE 0")
    s.clear()
    assert (s.num_measurements, s.num_detectors, s.num_observables, s.num_qubits) == (0, 0, 0, 101)


def test_stabilizer_randomization_and_pauli_frames():
    s = stim_b200.FlipSimulator(batch_size=256, num_qubits=10, disable_stabilizer_randomization=True)
    assert s.peek_pauli_flips() == ["+" + "_" * 10] * 256
    s.do("R 19")
    assert s.peek_pauli_flips() == ["+" + "_" * 20] * 256
    s = stim_b200.FlipSimulator(batch_size=256, num_qubits=10, seed=3)
    v = np.array([list(p[1:]) for p in s.peek_pauli_flips()])
    assert v.shape == (256, 10) and np.all((v == "_") | (v == "Z")) and 0.2 < np.mean(v == "Z") < 0.8
    s.do("R 19")
    v = np.array([list(p[1:]) for p in s.peek_pauli_flips()])
    assert v.shape == (256, 20) and np.all((v == "_") | (v == "Z")) and 0.2 < np.mean(v == "Z") < 0.8
    # set_pauli_flip / peek (frame_simulator_pybind_test.py:153-217)
    s = stim_b200.FlipSimulator(batch_size=2, disable_stabilizer_randomization=True, num_qubits=3)
    assert s.peek_pauli_flips() == ["+___", "+___"]
    s.set_pauli_flip("X", qubit_index=2, instance_index=1)
    assert s.peek_pauli_flips() == ["+___", "+__X"]
    s.set_pauli_flip(3, qubit_index=1, instance_index=0)
    s.set_pauli_flip("Y", qubit_index=0, instance_index=-1)
    assert s.peek_pauli_flips() == ["+_Z_", "+Y_X"] and s.peek_pauli_flips(instance_index=1) == "+Y_X"
    s.set_pauli_flip("I", qubit_index=2, instance_index=1)
    assert s.peek_pauli_flips(instance_index=-1) == "+Y__"
    s.set_pauli_flip("X", qubit_index=5, instance_index=0)
    assert s.num_qubits == 6 and s.peek_pauli_flips(instance_index=0) == "+_Z___X"
    with pytest.raises(ValueError):
        s.set_pauli_flip("Q", qubit_index=0, instance_index=0)
    with pytest.raises(IndexError):
        s.set_pauli_flip("X", qubit_index=0, instance_index=2)
    # frames propagate through gates
    s = stim_b200.FlipSimulator(batch_size=3, disable_stabilizer_randomization=True, num_qubits=2)
    s.set_pauli_flip("X", qubit_index=0, instance_index=1)
    s.do("CX 0 1\nH 0")
    assert s.peek_pauli_flips() == ["+__", "+ZX", "+__"]


def test_broadcast_pauli_errors_bernoulli_and_append():
    s = stim_b200.FlipSimulator(batch_size=2, num_qubits=3, disable_stabilizer_randomization=True)
    s.broadcast_pauli_errors(pauli="X", mask=np.asarray([[True, False], [False, False], [True, True]]))
    assert s.peek_pauli_flips() == ["+X_X", "+__X"]
    s.broadcast_pauli_errors(pauli="Z", mask=np.asarray([[False, True], [False, False], [True, True]]))
    assert s.peek_pauli_flips() == ["+X_Y", "+Z_Y"]
    s.broadcast_pauli_errors(pauli="Y", mask=np.asarray([[True, False], [False, True], [False, True]]))
    assert s.peek_pauli_flips() == ["+Z_Y", "+ZY_"]
    s.broadcast_pauli_errors(pauli="I", mask=np.asarray([[True, True], [False, True], [True, True]]))
    assert s.peek_pauli_flips() == ["+Z_Y", "+ZY_"]
    with pytest.raises(ValueError):
        s.broadcast_pauli_errors(pauli="X", mask=np.asarray([[True, False, False]]))
    with pytest.raises(ValueError):
        s.broadcast_pauli_errors(pauli="X", mask=np.asarray([[True, False]]), p=1.5)
    # p < 1: the right rate, and Y flips x and z with the same coin
    s = stim_b200.FlipSimulator(batch_size=4096, num_qubits=4, disable_stabilizer_randomization=True, seed=11)
    mask = np.zeros((4, 4096), dtype=np.bool_)
    mask[1] = True
    mask[3, ::2] = True
    s.broadcast_pauli_errors(pauli="Y", mask=mask, p=0.25)
    xs, zs, *_ = s.to_numpy(output_xs=True, output_zs=True)
    np.testing.assert_array_equal(xs, zs)
    assert not xs[0].any() and not xs[2].any() and not xs[3, 1::2].any()
    assert abs(xs[1].mean() - 0.25) < 0.03 and abs(xs[3, ::2].mean() - 0.25) < 0.05
    # generate_bernoulli_samples
    v = s.generate_bernoulli_samples(20000, p=0.3)
    assert v.dtype == np.bool_ and v.shape == (20000,) and abs(v.mean() - 0.3) < 0.02
    v = s.generate_bernoulli_samples(1001, p=1, bit_packed=True)
    assert v.dtype == np.uint8 and v.shape == (126,) and np.unpackbits(v, bitorder="little")[:1001].all() and v[-1] == 1
    assert not s.generate_bernoulli_samples(64, p=0).any()
    # append_measurement_flips + detectors over them
    s = stim_b200.FlipSimulator(batch_size=5, disable_stabilizer_randomization=True)
    s.append_measurement_flips(np.array([[0, 1, 0, 0, 1], [0, 0, 1, 0, 1]], dtype=np.bool_))
    s.do("DETECTOR rec[-1] rec[-2]")
    assert s.num_measurements == 2
    np.testing.assert_array_equal(s.get_detector_flips(detector_index=0), [False, True, True, False, False])
    # to_numpy shapes / transpose / packing
    xs, zs, ms, ds, os_ = s.to_numpy(transpose=True, bit_packed=True, output_measure_flips=True, output_detector_flips=True)
    assert xs is None and zs is None and os_ is None and ms.shape == (5, 1) and ds.shape == (5, 1)
    with pytest.raises(ValueError):
        s.to_numpy()


def test_noisy_instructions_match_the_bulk_sampler_statistically():
    text = gen_circuit("surface_code", "rotated_memory_x", 5, 5, 0.01)
    shots = 1 << 17
    s = stim_b200.FlipSimulator(batch_size=shots, seed=21)
    for piece in top_level_pieces(text):
        s.do(piece)
    dets = s.get_detector_flips()
    obs = s.get_observable_flips()
    bulk_d, bulk_o = stim_b200.Circuit(text).compile_detector_sampler(seed=22).sample(shots, separate_observables=True)
    k1, k2 = dets.sum(axis=1).astype(np.float64), bulk_d.sum(axis=0).astype(np.float64)
    p = (k1 + k2) / (2 * shots)
    z = (k1 - k2) / shots / np.sqrt(np.maximum(p * (1 - p), 1e-12) * 2 / shots)
    assert np.abs(z).max() < 5.0 and abs(np.sqrt(np.mean(z**2)) - 1) < 0.25
    assert abs(obs.mean() - bulk_o.mean()) < 0.004
    # adjacent-pair correlations of the flips
    c1 = (dets[:-1] & dets[1:]).sum(axis=1).astype(np.float64)
    c2 = (bulk_d[:, :-1] & bulk_d[:, 1:]).sum(axis=0).astype(np.float64)
    p = (c1 + c2) / (2 * shots)
    z = (c1 - c2) / shots / np.sqrt(np.maximum(p * (1 - p), 1e-12) * 2 / shots)
    assert np.abs(z).max() < 5.0


def test_copy_is_independent_and_copy_rng_continues_the_same_stream():
    """FlipSimulator.copy (frame_simulator.pybind.cc:1476-1488)."""
    s1 = stim_b200.FlipSimulator(batch_size=256, num_qubits=3, seed=11, disable_stabilizer_randomization=True)
    s1.do("X_ERROR(0.4) 0 1 2\nM 0 1\nDETECTOR rec[-1]\nOBSERVABLE_INCLUDE(0) rec[-2]")
    s2 = s1.copy(copy_rng=True)
    s3 = s1.copy(seed=5)
    for s in (s2, s3):
        assert (s.batch_size, s.num_qubits, s.num_measurements, s.num_detectors, s.num_observables) == (256, 3, 2, 1, 1)
        np.testing.assert_array_equal(s.get_measurement_flips(), s1.get_measurement_flips())
        np.testing.assert_array_equal(s.get_detector_flips(), s1.get_detector_flips())
        np.testing.assert_array_equal(s.get_observable_flips(), s1.get_observable_flips())
        assert [str(p) for p in s.peek_pauli_flips()] == [str(p) for p in s1.peek_pauli_flips()]
    more = "DEPOLARIZE1(0.3) 0 1 2\nM 0 1 2"
    s1.do(more)
    s2.do(more)
    s3.do(more)
    np.testing.assert_array_equal(s2.get_measurement_flips(), s1.get_measurement_flips())      # same stream
    assert not np.array_equal(s3.get_measurement_flips(), s1.get_measurement_flips())          # its own stream
    assert s3.num_measurements == 5
    s2.clear()
    assert s1.num_measurements == 5 and s2.num_measurements == 0                                # independent storage
    with pytest.raises(ValueError, match="incompatible"):
        s1.copy(copy_rng=True, seed=1)
