"""GPU parity: the CUDA path (through the C ABI) must equal the oracle bit for bit — noisy circuits
included — for the same (seed, shot offset, columns-per-block)."""
import numpy as np
import pytest

import stim_b200
from conftest import gen_circuit
from oracle import frame_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _interpreter_engine(monkeypatch):
    """These tests compare with the frame oracle's random stream, which is the interpreter's (the event engine draws a
    different stream and has its own bit-exact tests in test_gpu_events.py)."""
    monkeypatch.setenv("GSTIM_ENGINE", "interp")


def oracle_for(sampler, text, shots, seed, mode, offset_before):
    K = sampler.last_block_columns()
    assert K >= 1
    return fo.sample(text, shots, seed, K, mode, col0=offset_before // 128)


def check_detectors(text, shots, seed):
    s = stim_b200.Circuit(text).compile_detector_sampler(seed=seed)
    off = s.shot_offset
    dets, obs = s.sample(shots, separate_observables=True)
    assert dets.dtype == np.bool_ and obs.dtype == np.bool_
    od, oo = oracle_for(s, text, shots, seed, "detectors", off)
    assert dets.shape == od.shape and obs.shape == oo.shape
    np.testing.assert_array_equal(dets.astype(np.uint8), od)
    np.testing.assert_array_equal(obs.astype(np.uint8), oo)
    return s


def check_measurements(text, shots, seed):
    s = stim_b200.Circuit(text).compile_sampler(seed=seed, skip_reference_sample=True)
    off = s.shot_offset
    m = s.sample(shots)
    om = oracle_for(s, text, shots, seed, "measurements", off)
    assert m.shape == om.shape
    np.testing.assert_array_equal(m.astype(np.uint8), om)
    return s


GENERATED = [
    ("repetition_code", "memory", 3, 10, 0.02),
    ("repetition_code", "memory", 5, 4, 0.3),
    ("surface_code", "rotated_memory_z", 3, 3, 0.02),
    ("surface_code", "rotated_memory_x", 5, 5, 0.01),
    ("surface_code", "unrotated_memory_z", 3, 2, 0.05),
    ("color_code", "memory_xyz", 3, 3, 0.02),
    ("color_code", "memory_xyz", 5, 2, 0.01),
]


@pytest.mark.parametrize("code,task,d,r,p", GENERATED)
@pytest.mark.parametrize("shots", [1, 200, 3000])
def test_generated_detectors_match_oracle(code, task, d, r, p, shots):
    check_detectors(gen_circuit(code, task, d, r, p), shots, seed=1234 + shots)


def test_many_noise_batches_use_global_event_counters():
    """> 2048 noise batches: the event counters / segment offsets no longer fit shared memory (GSTIM_EV_SMEM_MAX)."""
    text = gen_circuit("repetition_code", "memory", 3, 700, 0.01)
    s = check_detectors(text, 300, seed=4321)
    assert s.stats.num_batches > 2048


def test_producers_run_ahead_over_many_shot_blocks(monkeypatch):
    """One column per block -> every CTA executes several shot blocks, so the producer warps reuse both event
    buffers (FREE / FULL barrier hand-over) while the interpreter is still busy with the previous block."""
    monkeypatch.setenv("GSTIM_KMAX", "1")
    monkeypatch.setenv("GSTIM_G_LOG2", "0")  # one lane per item, otherwise K is at least the lanes per item
    text = gen_circuit("surface_code", "rotated_memory_z", 3, 2, 0.03)
    s = check_detectors(text, 128 * 148 * 3 + 77, seed=77)
    assert s.last_block_columns() == 1


@pytest.mark.parametrize("code,task,d,r,p", GENERATED[:5])
def test_generated_measurements_match_oracle(code, task, d, r, p):
    check_measurements(gen_circuit(code, task, d, r, p), 1500, seed=99)


def test_successive_calls_continue_the_stream():
    text = gen_circuit("surface_code", "rotated_memory_z", 3, 3, 0.02)
    s = stim_b200.Circuit(text).compile_detector_sampler(seed=5)
    a = s.sample(700, append_observables=True)
    off1 = s.shot_offset
    assert off1 >= 700 and off1 % 128 == 0
    K1 = s.last_block_columns()
    b = s.sample(700, append_observables=True)
    K2 = s.last_block_columns()
    oa = np.concatenate(fo.sample(text, 700, 5, K1, "detectors", col0=0), axis=1)
    ob = np.concatenate(fo.sample(text, 700, 5, K2, "detectors", col0=off1 // 128), axis=1)
    np.testing.assert_array_equal(a.astype(np.uint8), oa)
    np.testing.assert_array_equal(b.astype(np.uint8), ob)
    assert not np.array_equal(a, b)


def test_same_seed_same_results_different_seed_differs():
    text = gen_circuit("surface_code", "rotated_memory_x", 5, 5, 0.01)
    a = stim_b200.Circuit(text).compile_detector_sampler(seed=11).sample(2000, bit_packed=True)
    b = stim_b200.Circuit(text).compile_detector_sampler(seed=11).sample(2000, bit_packed=True)
    c = stim_b200.Circuit(text).compile_detector_sampler(seed=12).sample(2000, bit_packed=True)
    np.testing.assert_array_equal(a, b)
    assert not np.array_equal(a, c)


ALL_OPS = """
QUBIT_COORDS(0, 1) 0
R 0 1 2 3 4 5 6 7
RX 8 9
RY 10 11
X_ERROR(0.2) 0 1 2 3
Z_ERROR(0.2) 4 5 8 9
Y_ERROR(0.1) 10 11 0
H 0 1 8
S 2 3
S_DAG 4
SQRT_X 5 6
SQRT_X_DAG 7
SQRT_Y 8
SQRT_Y_DAG 9
H_XY 10
H_YZ 11
H_NXY 0
H_NXZ 1
H_NYZ 2
C_XYZ 3 4
C_ZYX 5
C_NXYZ 6
C_XNYZ 7
C_XYNZ 8
C_NZYX 9
C_ZNYX 10
C_ZYNX 11
I 0
X 1
Y 2
Z 3
TICK
DEPOLARIZE1(0.1) 0 1 2 3 4 5 6 7 8 9 10 11
CX 0 1 2 3
CY 4 5 6 7
CZ 8 9 10 11
XCX 0 2 1 3
XCY 4 6
XCZ 5 7
YCX 8 10
YCY 9 11
YCZ 0 11
SWAP 1 10
ISWAP 2 9
ISWAP_DAG 3 8
CXSWAP 4 7
SWAPCX 5 6
CZSWAP 0 6
SQRT_XX 1 7
SQRT_XX_DAG 2 8
SQRT_YY 3 9
SQRT_YY_DAG 4 10
SQRT_ZZ 5 11
SQRT_ZZ_DAG 0 1
II 2 3
DEPOLARIZE2(0.1) 0 1 2 3 4 5 6 7 8 9 10 11
PAULI_CHANNEL_1(0.05, 0.1, 0.15) 0 1 2 3
PAULI_CHANNEL_2(0.01, 0.02, 0.03, 0.01, 0.02, 0.03, 0.01, 0.02, 0.03, 0.01, 0.02, 0.03, 0.01, 0.02, 0.03) 4 5 6 7
I_ERROR(0.1) 0
II_ERROR(0.1) 0 1
E(0.2) X0 Y1 Z2
ELSE_CORRELATED_ERROR(0.3) Z3 X4
ELSE_CORRELATED_ERROR(0.4) Y5
HERALDED_ERASE(0.2) 6 7
HERALDED_PAULI_CHANNEL_1(0.05, 0.1, 0.15, 0.2) 8 9
M 0 !1
MX 2
MY 3
MR(0.1) 4 5
MRX 6
MRY(0.05) 7
M(0.25) 8 8
MR 9 9
MPP X0*Y1*Z2 Z3*Z4
MPP(0.1) !X5*X6 Y7
MPP X0*X0
MXX 0 1 2 3
MYY(0.05) 4 5
MZZ 6 7 7 8
MPAD 0 1
MPAD(0.3) 1
SPP X0*Y1 Z2
SPP_DAG Z3*Z4
CX rec[-1] 0 rec[-3] 1
CY rec[-2] 2
CZ rec[-4] 3 4 rec[-5]
XCZ 5 rec[-6]
YCZ 6 rec[-7]
CX sweep[0] 1
DETECTOR rec[-1] rec[-2]
DETECTOR(1, 2, 3) rec[-3]
DETECTOR
OBSERVABLE_INCLUDE(0) rec[-1] rec[-5]
OBSERVABLE_INCLUDE(2) X0 Y1 Z2 rec[-2]
REPEAT 3 {
    H 0 1 2
    CX 0 1 1 2
    DEPOLARIZE2(0.05) 0 1
    MR 1 2
    DETECTOR rec[-1] rec[-3]
    SHIFT_COORDS(0, 1)
    OBSERVABLE_INCLUDE(1) rec[-2]
}
M 0 1 2 3 4 5 6 7 8 9 10 11
DETECTOR rec[-1] rec[-12] rec[-20]
OBSERVABLE_INCLUDE(0) rec[-3]
"""


@pytest.mark.parametrize("shots", [300, 5000])
def test_every_instruction_detectors(shots):
    check_detectors(ALL_OPS, shots, seed=2024)


def test_every_instruction_measurements():
    check_measurements(ALL_OPS, 2500, seed=77)


def test_output_layouts_are_consistent():
    """bool vs bit_packed vs prepend/append/separate vs caller buffers describe the same shots."""
    text = gen_circuit("surface_code", "rotated_memory_z", 3, 3, 0.05)
    c = stim_b200.Circuit(text)
    D, L = c.num_detectors, c.num_observables
    shots = 777

    def fresh():
        return c.compile_detector_sampler(seed=3)

    base_d, base_o = fresh().sample(shots, separate_observables=True)
    np.testing.assert_array_equal(fresh().sample(shots), base_d)
    np.testing.assert_array_equal(fresh().sample(shots, append_observables=True), np.concatenate([base_d, base_o], axis=1))
    np.testing.assert_array_equal(fresh().sample(shots, prepend_observables=True), np.concatenate([base_o, base_d], axis=1))
    pk = fresh().sample(shots, bit_packed=True, append_observables=True)
    assert pk.dtype == np.uint8 and pk.shape == (shots, (D + L + 7) // 8)
    np.testing.assert_array_equal(np.unpackbits(pk, axis=1, bitorder="little")[:, : D + L],
                                  np.concatenate([base_d, base_o], axis=1).astype(np.uint8))
    pd, po = fresh().sample(shots, bit_packed=True, separate_observables=True)
    np.testing.assert_array_equal(np.unpackbits(pd, axis=1, bitorder="little")[:, :D], base_d.astype(np.uint8))
    np.testing.assert_array_equal(np.unpackbits(po, axis=1, bitorder="little")[:, :L], base_o.astype(np.uint8))
    # caller supplied buffers, including a strided view
    big = np.zeros((shots, D + 5), dtype=np.bool_)
    view = big[:, 2: 2 + D]
    ob = np.zeros((shots, L), dtype=np.bool_)
    r = fresh().sample(shots, dets_out=view, obs_out=ob)
    assert r is view
    np.testing.assert_array_equal(view, base_d)
    np.testing.assert_array_equal(ob, base_o)
    assert not big[:, :2].any() and not big[:, 2 + D:].any()
    with pytest.raises(ValueError):
        fresh().sample(shots, dets_out=np.zeros((shots, D + 1), dtype=np.bool_))
    with pytest.raises(ValueError):
        fresh().sample(shots, dets_out=np.zeros((shots, D), dtype=np.uint8))
    with pytest.raises(ValueError):
        fresh().sample(shots, separate_observables=True, append_observables=True)


def test_device_resident_consumer_hook_matches_oracle():
    """the sinter-shaped consumer (glue/sample/src/sinter/_decoding/_stim_then_decode_sampler.py:162-185:
    sample bit-packed dets + separate observables, decode, count shots whose prediction differs) with the shots staying on
    the GPU (sample_torch), and with page-locked host arrays (sample_pinned). Bits are checked against the ORACLE."""
    import torch

    text = gen_circuit("surface_code", "rotated_memory_z", 5, 5, 0.01)
    seed, shots = 11, 5000
    a = stim_b200.Circuit(text).compile_detector_sampler(seed=seed)
    dets_t, obs_t = a.sample_torch(shots)
    assert dets_t.is_cuda and dets_t.dtype == torch.uint8 and obs_t.shape == (shots, 1)
    want_d, want_o = fo.sample(text, shots, seed, a.last_block_columns(), "detectors")
    D = want_d.shape[1]
    np.testing.assert_array_equal(np.unpackbits(dets_t.cpu().numpy(), axis=1, bitorder="little")[:, :D], want_d)
    np.testing.assert_array_equal(obs_t.cpu().numpy() & 1, want_o)
    # "decoder" that always predicts no logical flip, evaluated next to the data: errors = shots with a flipped observable
    predictions = torch.zeros_like(obs_t)
    num_errors = int((predictions != obs_t).any(dim=1).sum().item())
    assert num_errors == int(want_o.any(axis=1).sum()) and 0 < num_errors < shots
    # detection-event weight histogram on the device (what a GPU decoder's front end consumes)
    lut = torch.tensor([bin(v).count("1") for v in range(256)], dtype=torch.int32, device=dets_t.device)
    weights = lut[dets_t.long()].sum(dim=1)
    np.testing.assert_array_equal(weights.cpu().numpy(), want_d.sum(axis=1))
    # appended form continues the stream: compare with the oracle at the advanced offset
    off = a.shot_offset
    both = a.sample_torch(shots, separate_observables=False)
    d2, o2 = fo.sample(text, shots, seed, a.last_block_columns(), "detectors", col0=off // 128)
    np.testing.assert_array_equal(np.unpackbits(both.cpu().numpy(), axis=1, bitorder="little")[:, : D + 1], np.concatenate([d2, o2], axis=1))
    # page-locked host arrays (DMA target of the direct copy path)
    b = stim_b200.Circuit(text).compile_detector_sampler(seed=seed)
    dp, op = b.sample_pinned(shots)
    assert dp.dtype == np.uint8 and dp.shape == (shots, (D + 7) // 8)
    np.testing.assert_array_equal(np.unpackbits(dp, axis=1, bitorder="little")[:, :D], want_d)
    np.testing.assert_array_equal(op & 1, want_o)


def test_samplers_of_different_sizes_interleave():
    """The dynamic shared-memory attribute is per kernel and device, not per sampler (round-1 advisor finding): a small
    sampler created after a large one must not break the large one's next launch."""
    big_text = gen_circuit("surface_code", "rotated_memory_z", 5, 5, 0.01)
    small_text = gen_circuit("repetition_code", "memory", 3, 10, 0.02)
    big = stim_b200.Circuit(big_text).compile_detector_sampler(seed=5)
    n_big = 128 * 148 * int(big.stats.max_columns)
    first = big.sample(n_big, bit_packed=True)
    small = stim_b200.Circuit(small_text).compile_detector_sampler(seed=6)
    small.sample(100)
    big.shot_offset = 0
    again = big.sample(n_big, bit_packed=True)
    np.testing.assert_array_equal(first, again)
