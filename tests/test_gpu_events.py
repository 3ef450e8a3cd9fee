"""GPU parity of the event-driven engine (stim_b200/csrc/sparse.cu): bit for bit against oracle/sparse_oracle.sample
(the restatement of its sampling on the exported response table) for the same (seed, shot offset); the table itself is
checked against forward injection on the frame oracle in tests/test_response_table.py (CPU), and the engine as a whole
against the reference CLI's deterministic outputs (test_gpu_golden.py runs under both engines) and 2^24-shot reference
statistics (test_gpu_stats_big.py)."""
import numpy as np
import pytest

import stim_b200
from conftest import gen_circuit
from oracle import sparse_oracle as so
from test_gpu_parity import ALL_OPS

pytestmark = pytest.mark.gpu

LONG_CHAIN = "E(0.01) X0\n" + "".join(f"ELSE_CORRELATED_ERROR(0.01) X{k}\n" for k in range(1, 17)) + "M 0\nDETECTOR rec[-1]\n"


def oracle_rows(s, seed, first_shot, shots, n_outputs):
    t = s.response_table()
    return so.sample(t, t["slices"], s.engine_info()["tile_shots"], seed, first_shot, shots, n_outputs)


def check_detectors(text, shots, seed, **kw):
    s = stim_b200.Circuit(text).compile_detector_sampler(seed=seed, engine="events")
    D, L = int(s.stats.num_detectors), int(s.stats.num_observables)
    off = s.shot_offset
    dets, obs = s.sample(shots, separate_observables=True, **kw)
    assert s.engine_info()["last_engine"] == "events"
    want = oracle_rows(s, seed, off, shots, D + L)
    np.testing.assert_array_equal(dets.astype(np.uint8), want[:, :D])
    np.testing.assert_array_equal(obs.astype(np.uint8), want[:, D:])
    return s, want


GENERATED = [
    ("repetition_code", "memory", 3, 10, 0.02),
    ("surface_code", "rotated_memory_z", 3, 3, 0.02),
    ("surface_code", "rotated_memory_x", 5, 5, 0.01),
    ("color_code", "memory_xyz", 3, 3, 0.02),
    ("color_code", "memory_xyz", 5, 2, 0.01),
]


@pytest.mark.parametrize("code,task,d,r,p", GENERATED)
@pytest.mark.parametrize("shots", [1, 200, 1500])
def test_generated_detectors_match_oracle(code, task, d, r, p, shots):
    check_detectors(gen_circuit(code, task, d, r, p), shots, seed=4321 + shots)


def test_every_instruction_detectors_match_oracle():
    s, _ = check_detectors(ALL_OPS, 700, seed=2025)
    info = s.engine_info()
    assert info["eligible"] == 1 and info["max_response"] >= 5 and info["overflow_words"] > 0


def test_every_instruction_measurements_match_oracle():
    seed = 31
    s = stim_b200.Circuit(ALL_OPS).compile_sampler(seed=seed, skip_reference_sample=True, engine="events")
    M = int(s.stats.num_measurements)
    m = s.sample(600)
    assert s.engine_info()["last_engine"] == "events"
    np.testing.assert_array_equal(m.astype(np.uint8), oracle_rows(s, seed, 0, 600, M))
    # the reference sample is XORed in (rows start from it)
    ref = np.zeros(M, dtype=np.bool_)
    ref[::3] = True
    s2 = stim_b200.Circuit(ALL_OPS).compile_sampler(seed=seed, reference_sample=ref, engine="events")
    np.testing.assert_array_equal(s2.sample(600), m ^ ref[None, :])


def test_long_else_chain_is_not_eligible():
    s = stim_b200.Circuit(LONG_CHAIN).compile_detector_sampler(seed=1)
    info = s.engine_info()
    assert info["eligible"] == 0 and "ELSE" in info["why_not"]
    with pytest.raises(ValueError):
        s.set_engine("events")
    s.sample(10)
    assert s.engine_info()["last_engine"] == "interp"


def test_layouts_strides_and_stream_continuation():
    text = gen_circuit("surface_code", "rotated_memory_x", 5, 5, 0.01)
    seed = 77
    s = stim_b200.Circuit(text).compile_detector_sampler(seed=seed, engine="events")
    D, L = int(s.stats.num_detectors), int(s.stats.num_observables)
    S = s.engine_info()["tile_shots"]
    a = s.sample(333, append_observables=True)
    off1 = s.shot_offset
    assert off1 % 128 == 0 and off1 >= 333
    b = s.sample(130, prepend_observables=True, bit_packed=True)
    off2 = s.shot_offset
    buf = np.zeros((257, 40), dtype=np.uint8)  # strided rows
    c = s.sample(257, bit_packed=True, dets_out=buf[:, 3:3 + (D + 7) // 8])
    wa = oracle_rows(s, seed, 0, 333, D + L)
    wb = oracle_rows(s, seed, off1, 130, D + L)
    wc = oracle_rows(s, seed, off2, 257, D + L)
    np.testing.assert_array_equal(a.astype(np.uint8), wa)
    np.testing.assert_array_equal(np.unpackbits(b, axis=1, bitorder="little")[:, :D + L], np.concatenate([wb[:, D:], wb[:, :D]], axis=1))
    np.testing.assert_array_equal(np.unpackbits(c, axis=1, bitorder="little")[:, :D], wc[:, :D])
    assert not buf[:, :3].any() and not buf[:, 3 + (D + 7) // 8:].any()
    assert S & (S - 1) == 0


def test_device_results_counts_and_files_agree(tmp_path):
    import torch

    text = gen_circuit("color_code", "memory_xyz", 5, 2, 0.01)
    seed = 9
    mk = lambda: stim_b200.Circuit(text).compile_detector_sampler(seed=seed, engine="events")
    shots = 1000
    s = mk()
    D, L = int(s.stats.num_detectors), int(s.stats.num_observables)
    host = s.sample(shots, append_observables=True, bit_packed=True)
    dets, obs = mk().sample_torch(shots)
    both = mk().sample_torch(shots, separate_observables=False)
    np.testing.assert_array_equal(both.cpu().numpy(), host)
    bits = np.unpackbits(host, axis=1, bitorder="little")[:, :D + L]
    np.testing.assert_array_equal(np.unpackbits(dets.cpu().numpy(), axis=1, bitorder="little")[:, :D], bits[:, :D])
    np.testing.assert_array_equal(np.unpackbits(obs.cpu().numpy(), axis=1, bitorder="little")[:, :L], bits[:, D:])
    single, pair = mk().bit_counts(shots)
    np.testing.assert_array_equal(single, bits.sum(axis=0).astype(np.uint64))
    np.testing.assert_array_equal(pair, (bits[:, :-1] & bits[:, 1:]).sum(axis=0).astype(np.uint64))
    # files: every format describes the same shots (1024 shots: ptb64 needs a multiple of 64)
    host2 = mk().sample(1024, append_observables=True, bit_packed=True)
    for fmt in ["b8", "01", "ptb64"]:
        p = tmp_path / f"x.{fmt}"
        mk().sample_write(1024, filepath=str(p), format=fmt, append_observables=True)
        raw = p.read_bytes()
        if fmt == "b8":
            assert raw == host2.tobytes()
        elif fmt == "01":
            want = "".join("".join(map(str, r)) + "\n" for r in np.unpackbits(host2, axis=1, bitorder="little")[:, :D + L])
            assert raw.decode() == want
        else:
            b2 = np.unpackbits(host2, axis=1, bitorder="little")[:, :D + L]
            got = np.frombuffer(raw, dtype=np.uint8).reshape(1024 // 64, D + L, 8)
            gb = np.unpackbits(got, axis=2, bitorder="little")  # [group, bit, shot in group]
            np.testing.assert_array_equal(gb.transpose(0, 2, 1).reshape(1024, D + L), b2)
    del torch


def test_large_rows_use_small_tiles():
    """d = 51 (16.6 KB per shot): four shots per tile; deterministic variant checked against the all-zero expectation
    plus one injected p = 1 flip."""
    import os
    import re

    from conftest import ROOT

    with open(os.path.join(ROOT, "tests", "golden", "circuits", "c5_surface_x_d51_r51.stim")) as f:
        text = f.read()
    text = re.sub(r"\(0\.001\)", "(0)", text)
    s = stim_b200.Circuit(text).compile_detector_sampler(seed=3, engine="events")
    assert s.engine_info()["eligible"] == 1
    out = s.sample(300, bit_packed=True, append_observables=True)
    assert not out.any()


def test_d71_surface_code_beyond_the_interpreters_frame_limit():
    """10 081 qubits: the frame of one 128-shot column (323 KB) no longer fits in shared memory, so the interpreter
    refuses the circuit; the event engine needs no frame at all (9 rounds here: 5.7 KB of detection events per shot).
    Checked against 2^13 shots of the reference CLI (rates of the busiest detectors and the mean detection fraction)."""
    import os
    import subprocess

    from conftest import REF_STIM, have_ref

    if not have_ref():
        pytest.skip("oracle/_ref/stim not built")
    text = gen_circuit("surface_code", "rotated_memory_z", 71, 9, 0.001)
    circ = stim_b200.Circuit(text)
    s = circ.compile_detector_sampler(seed=71)
    info = s.engine_info()
    assert info["eligible"] == 1 and info["tile_shots"] <= 16 and int(s.stats.active_qubits) == 10081
    with pytest.raises(ValueError):
        circ.compile_detector_sampler(seed=71, engine="interp").sample(1)
    shots = 1 << 15
    D = int(s.stats.num_detectors)
    got = s.sample(shots, bit_packed=True)
    assert s.engine_info()["last_engine"] == "events"
    n_ref = 1 << 13
    r = subprocess.run([REF_STIM, "detect", "--shots", str(n_ref), "--out_format", "b8", "--seed", "5"], input=text.encode(),
                       capture_output=True, check=True)
    ref = np.frombuffer(r.stdout, dtype=np.uint8).reshape(n_ref, (D + 7) // 8)
    k_gpu = np.unpackbits(got, axis=1, bitorder="little")[:, :D].sum(axis=0, dtype=np.int64)
    k_ref = np.unpackbits(ref, axis=1, bitorder="little")[:, :D].sum(axis=0, dtype=np.int64)
    p = (k_gpu + k_ref) / (shots + n_ref)
    z = (k_gpu / shots - k_ref / n_ref) / np.sqrt(np.maximum(p * (1 - p), 1e-12) * (1 / shots + 1 / n_ref))
    busy = (k_gpu + k_ref) > 200
    assert busy.sum() > 1000 and np.abs(z[busy]).max() < 6.0
    assert abs(np.sqrt(np.mean(z[busy] ** 2)) - 1.0) < 0.1
    assert abs(k_gpu.sum() / shots - k_ref.sum() / n_ref) < 0.01 * k_ref.sum() / n_ref


def test_multi_device_sampler_equals_single_device():
    """gstim_create_from_text_multi: one sample() call sharded over two GPUs of this process returns exactly the shots
    a single-device sampler returns (the event engine's stream depends on the seed and the global shot index only)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    text = gen_circuit("surface_code", "rotated_memory_x", 5, 5, 0.01)
    shots = 100_000 + 77
    one = stim_b200.Circuit(text).compile_detector_sampler(seed=3, engine="events")
    two = stim_b200.Circuit(text).compile_detector_sampler(seed=3, engine="events", device=[0, 1])
    a = one.sample(shots, bit_packed=True, append_observables=True)
    b = two.sample(shots, bit_packed=True, append_observables=True)
    np.testing.assert_array_equal(a, b)
    assert two.shot_offset == one.shot_offset
    a2, ao = one.sample(1000, separate_observables=True)
    b2, bo = two.sample(1000, separate_observables=True)
    np.testing.assert_array_equal(a2, b2)
    np.testing.assert_array_equal(ao, bo)
    m1 = stim_b200.Circuit(text).compile_sampler(seed=4, engine="events").sample(5000, bit_packed=True)
    m2 = stim_b200.Circuit(text).compile_sampler(seed=4, engine="events", device=[0, 1]).sample(5000, bit_packed=True)
    np.testing.assert_array_equal(m1, m2)


def test_dense_noise_takes_the_packed_word_path_and_matches_oracle_and_interpreter():
    """Packed Bernoulli words for dense sites (north_star; probability_util.cc:74-132): bit for bit against the oracle, and
    statistically against the interpreter's geometric walk on the same circuit."""
    text = """
    R 0 1 2 3 4 5
    X_ERROR(0.5) 0 1
    DEPOLARIZE1(0.75) 2
    DEPOLARIZE2(0.5) 3 4
    Z_ERROR(0.25) 5
    H 5
    CX 0 1 2 3
    M 0 1 2 3 4 5
    DETECTOR rec[-6]
    DETECTOR rec[-5]
    DETECTOR rec[-4] rec[-3]
    DETECTOR rec[-3]
    DETECTOR rec[-2]
    DETECTOR rec[-1]
    OBSERVABLE_INCLUDE(0) rec[-6] rec[-1]
    """
    s, _ = check_detectors(text, 3000, seed=8)
    cl = s.response_table()["classes"]
    assert (cl[:, 23] != 0).sum() >= 4
    shots = 1 << 20
    ev, _ = stim_b200.Circuit(text).compile_detector_sampler(seed=1, engine="events").bit_counts(shots)
    it, _ = stim_b200.Circuit(text).compile_detector_sampler(seed=2, engine="interp").bit_counts(shots)
    p = (ev + it) / (2.0 * shots)
    z = (ev.astype(np.float64) - it.astype(np.float64)) / shots / np.sqrt(np.maximum(p * (1 - p), 1e-12) * 2 / shots)
    assert np.abs(z).max() < 5.0, z
    np.testing.assert_allclose(ev[:2] / shots, [0.5, 0.5], atol=0.003)


def test_sinter_shaped_consumer_counts_on_the_device():
    """stim_b200.sinter_hook.StimThenDecodeSampler = sinter's sample -> decode -> count loop with the shots left on the GPU.
    Its counts must equal the reference loop restated in numpy (_stim_then_decode_sampler.py:162-223) over the SAME shots
    (same seed -> same stream) pulled to the host through sample()."""
    import torch

    from stim_b200.sinter_hook import StimThenDecodeSampler

    text = gen_circuit("surface_code", "rotated_memory_x", 5, 5, 0.01)
    circ = stim_b200.Circuit(text)
    D, L = circ.num_detectors, circ.num_observables
    shots = 50_000
    rng = np.random.default_rng(3)
    post = np.zeros((D + 7) // 8, dtype=np.uint8)
    post[:2] = 0x0F  # post-select on the first few detectors
    weights = rng.integers(0, 256, size=(D + 7) // 8, dtype=np.uint8)

    def decoder(dets):
        # a stand-in decoder that runs on the device: predicts the observable as a fixed parity of the detection events
        assert isinstance(dets, torch.Tensor) and dets.is_cuda
        w = torch.as_tensor(weights, device=dets.device)
        v = dets & w
        par = torch.zeros(dets.shape[0], dtype=torch.uint8, device=dets.device)
        for b in range(8):
            par ^= ((v >> b) & 1).sum(dim=1).to(torch.uint8) & 1
        return par.reshape(-1, 1)

    stats = StimThenDecodeSampler(circ, decoder, count_detection_events=True, postselection_mask=post, seed=44).sample(shots)
    # the reference loop on the host
    dets, obs = circ.compile_detector_sampler(seed=44).sample(shots, bit_packed=True, separate_observables=True)
    n_events = sum(int(np.count_nonzero(dets & (1 << b))) for b in range(8))
    discarded = np.any(dets & post, axis=1)
    kept_d, kept_o = dets[~discarded], obs[~discarded]
    pred = np.zeros(kept_d.shape[0], dtype=np.uint8)
    v = kept_d & weights
    for b in range(8):
        pred ^= (((v >> b) & 1).sum(axis=1) & 1).astype(np.uint8)
    errors = int(np.count_nonzero(np.any((pred.reshape(-1, 1) ^ kept_o) != 0, axis=1)))
    assert stats.shots == shots and stats.discards == int(discarded.sum()) and stats.errors == errors
    assert stats.custom_counts["detection_events"] == n_events and stats.custom_counts["detectors_checked"] == D * shots
    assert 0 < stats.discards < shots and 0 < stats.errors < shots


def test_sparse_host_delivery_equals_dense_delivery(monkeypatch):
    """GSTIM_D2H=sparse: rows cross PCIe as records of their non-zero bytes and are rebuilt by host threads; the caller sees
    the same arrays as with the dense copy (packed, unpacked, separate observables, strided rows, several chunks)."""
    text = gen_circuit("surface_code", "rotated_memory_x", 5, 5, 0.01)
    circ = stim_b200.Circuit(text)
    D = circ.num_detectors
    shots = 70_001
    monkeypatch.setenv("GSTIM_D2H", "dense")
    want_p = circ.compile_detector_sampler(seed=5, engine="events").sample(shots, bit_packed=True, append_observables=True)
    want_d, want_o = circ.compile_detector_sampler(seed=5, engine="events").sample(shots, separate_observables=True)
    monkeypatch.setenv("GSTIM_D2H", "sparse")
    monkeypatch.setenv("GSTIM_STAGE_MB", "1")  # many chunks: the three-stage pipeline wraps around
    got_p = circ.compile_detector_sampler(seed=5, engine="events").sample(shots, bit_packed=True, append_observables=True)
    np.testing.assert_array_equal(got_p, want_p)
    got_d, got_o = circ.compile_detector_sampler(seed=5, engine="events").sample(shots, separate_observables=True)
    np.testing.assert_array_equal(got_d, want_d)
    np.testing.assert_array_equal(got_o, want_o)
    buf = np.full((shots, 40), 0xAA, dtype=np.uint8)
    circ.compile_detector_sampler(seed=5, engine="events").sample(shots, bit_packed=True, dets_out=buf[:, 5:5 + (D + 7) // 8])
    np.testing.assert_array_equal(np.unpackbits(buf[:, 5:5 + (D + 7) // 8], axis=1, bitorder="little")[:, :D], want_d.astype(np.uint8))
    assert (buf[:, :5] == 0xAA).all() and (buf[:, 5 + (D + 7) // 8:] == 0xAA).all()
    # a dense circuit overflows the record stream and falls back to dense rows chunk by chunk
    dense = "X_ERROR(0.5) " + " ".join(str(q) for q in range(64)) + "\nM " + " ".join(str(q) for q in range(64)) + "\n" + "".join(
        f"DETECTOR rec[-{k}]\n" for k in range(64, 0, -1))
    a = stim_b200.Circuit(dense).compile_detector_sampler(seed=6, engine="events").sample(20_000, bit_packed=True)
    monkeypatch.setenv("GSTIM_D2H", "dense")
    b = stim_b200.Circuit(dense).compile_detector_sampler(seed=6, engine="events").sample(20_000, bit_packed=True)
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("seed", range(6))
def test_random_circuits_match_oracle_on_both_engines(seed):
    """Random instruction sequences (tests/test_response_table.py:_random_circuit): the event engine bit for bit against
    its oracle (detectors and measurements), the interpreter bit for bit against the frame oracle, and the two engines against
    each other statistically (their random streams differ)."""
    from oracle import frame_oracle as fo
    from test_response_table import _random_circuit

    text = _random_circuit(np.random.default_rng(2000 + seed))
    circ = stim_b200.Circuit(text)
    ev = circ.compile_detector_sampler(seed=seed, engine="interp")
    if not ev.engine_info()["eligible"]:
        pytest.skip(ev.engine_info()["why_not"])
    check_detectors(text, 400, seed=50 + seed)
    ms = circ.compile_sampler(seed=seed, skip_reference_sample=True, engine="events")
    m = ms.sample(300)
    np.testing.assert_array_equal(m.astype(np.uint8), oracle_rows(ms, seed, 0, 300, circ.num_measurements))
    it = circ.compile_detector_sampler(seed=seed, engine="interp")
    dets, obs = it.sample(400, separate_observables=True)
    od, oo = fo.sample(text, 400, seed, it.last_block_columns(), "detectors")
    np.testing.assert_array_equal(dets.astype(np.uint8), od)
    np.testing.assert_array_equal(obs.astype(np.uint8), oo)
    shots = 1 << 18
    a, _ = circ.compile_detector_sampler(seed=1, engine="events").bit_counts(shots)
    b, _ = circ.compile_detector_sampler(seed=2, engine="interp").bit_counts(shots)
    p = (a + b) / (2.0 * shots)
    z = (a.astype(np.float64) - b.astype(np.float64)) / shots / np.sqrt(np.maximum(p * (1 - p), 1e-12) * 2 / shots)
    assert np.abs(z).max() < 5.5, (z, text)
