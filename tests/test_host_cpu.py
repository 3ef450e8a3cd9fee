"""CPU tests of the host side: C ABI surface, parser / lowering error behaviour (same exception types as the
reference), lowering correctness through the program-level emulator, and the multi-rank sharding logic (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import stim_b200
from conftest import ROOT, gen_circuit
from oracle import frame_oracle as fo
from oracle import program_emulator as pe
from stim_b200 import _native, sharding


def lower(text, mode, slots, chunk=0):
    L = _native.lib()
    d = text.encode()
    n = ctypes.c_size_t(0)
    plan = (ctypes.c_uint32 * 16)()
    _native.check(L.gstim_lower_text(d, len(d), mode, slots, chunk, None, ctypes.byref(n), plan))
    w = np.empty(n.value, dtype=np.uint32)
    _native.check(L.gstim_lower_text(d, len(d), mode, slots, chunk, w.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n), plan))
    return w, list(plan)


# ---------------------------------------------------------------------------------------------------------
# C ABI
# ---------------------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "gstim.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(gstim_[a-z0-9_]+)\s*\(", header))
    declared -= {"gstim_sampler", "gstim_stats"}
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/gstim.h but not exported"
    assert declared == set(_native.exported_symbols()), "ctypes binding and header disagree"
    assert lib.gstim_version() >= 1


def test_no_cpu_fallback():
    """Without a CUDA device creating a sampler must fail loudly (after validating the circuit)."""
    if _native.lib().gstim_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(stim_b200.GstimCudaError):
        stim_b200.Circuit("H 0\nM 0\n").compile_detector_sampler(seed=1)
    with pytest.raises(stim_b200.GstimCudaError):
        stim_b200.Circuit("H 0\nM 0\n").compile_sampler(seed=1, skip_reference_sample=True)


def test_product_path_does_not_import_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "stim_b200")):
        for fn in files:
            if fn.endswith((".py", ".cc", ".cu", ".h", ".cuh")):
                with open(os.path.join(root, fn)) as f:
                    src = f.read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace(
                    "oracle/program_emulator.py", "").replace("oracle/philox.py", "").replace("oracle/log2_table.py", "").replace(
                    "oracle/sparse_oracle.py", ""), fn


# ---------------------------------------------------------------------------------------------------------
# parser / stats / errors (exception types follow the reference: invalid_argument -> ValueError, out_of_range -> IndexError)
# ---------------------------------------------------------------------------------------------------------
def test_circuit_stats_match_reference_numbers():
    # SURVEY Appendix C, computed with the reference
    c3 = stim_b200.Circuit(gen_circuit("surface_code", "rotated_memory_z", 25, 25, 0.001)) if False else None
    with open(os.path.join(ROOT, "tests", "golden", "circuits", "c3_surface_z_d25_r25.stim")) as f:
        c3 = stim_b200.Circuit(f.read())
    assert (c3.num_qubits, c3.num_measurements, c3.num_detectors, c3.num_observables) == (1324, 16225, 15600, 1)
    assert c3._stats.max_lookback == 1248 and c3._stats.active_qubits == 1249
    with open(os.path.join(ROOT, "tests", "golden", "circuits", "c5_surface_x_d51_r51.stim")) as f:
        c5 = stim_b200.Circuit(f.read())
    assert (c5.num_measurements, c5.num_detectors, c5._stats.active_qubits) == (135201, 132600, 5201)
    c = stim_b200.Circuit("MPP X0*Y1 Z2\nMXX 0 1\nMPAD 1\nHERALDED_ERASE(0.1) 3\nOBSERVABLE_INCLUDE(7) rec[-1]\nDETECTOR\n")
    assert (c.num_measurements, c.num_detectors, c.num_observables, c.num_qubits) == (5, 1, 8, 4)


@pytest.mark.parametrize("text", [
    "NOT_A_GATE 0",
    "H rec[-1]",
    "CX 0",
    "CX 0 0",
    "X_ERROR(1.5) 0",
    "X_ERROR 0",
    "PAULI_CHANNEL_1(0.5, 0.4, 0.3) 0",
    "M(0.1, 0.2) 0",
    "DETECTOR 0",
    "MPP X0*",
    "REPEAT 0 {\nH 0\n}",
    "REPEAT 2 {\nH 0\n",
    "}",
    "TICK 0",
    "E(0.1) 5",
    "OBSERVABLE_INCLUDE(1.5) rec[-1]",
    "MPAD 2",
])
def test_invalid_circuits_raise_value_error(text):
    with pytest.raises(ValueError):
        stim_b200.Circuit(text)


@pytest.mark.parametrize("text", ["MPP X0*Z0", "CX 0 rec[-1]", "M 0\nCY 1 rec[-1]"])
def test_errors_the_reference_raises_when_the_circuit_is_run(text):
    """Anti-Hermitian products (gate_decomposition.cc) and bit-as-target (frame_simulator.inl:399-402, 421-424) parse
    fine in the reference and throw std::invalid_argument when simulated: here, when a sampler is compiled."""
    c = stim_b200.Circuit(text)
    with pytest.raises(ValueError):
        c.compile_detector_sampler(seed=1)
    with pytest.raises(ValueError):
        c.compile_sampler(seed=1, skip_reference_sample=True)


@pytest.mark.parametrize("text", ["DETECTOR rec[-1]", "M 0\nDETECTOR rec[-2]", "M 0\nCX rec[-3] 1", "OBSERVABLE_INCLUDE(0) rec[-1]"])
def test_bad_lookback_raises_index_error(text):
    c = stim_b200.Circuit(text)  # parses, like the reference
    with pytest.raises(IndexError):  # measure_record_batch.inl:83-94 throws std::out_of_range
        c.compile_detector_sampler(seed=1)


@pytest.mark.parametrize("text", [
    "M(0.01) 0 1 2 3\nDETECTOR rec[-1]",
    "HERALDED_PAULI_CHANNEL_1(0.01, 0.02, 0.03, 0.04) 1 0",
    "MPP(0.1) X0*X1 Z0*Z1 Y2 Z3 X4\nDETECTOR rec[-2]",
    "MPAD(0.25) 0 1 1 0 1",
])
def test_noisy_results_with_more_targets_than_the_lookback_window(text):
    """The record ring of detector mode must hold every result of one instruction (round-1 bug: it was sized from the
    maximum lookback alone, two record rows of one noisy instruction aliased and lowering threw)."""
    for mode in (0, 1):
        w, plan = lower(text, mode, 32)  # (used to throw "a noise group was cut inside an RNG slice")
        assert w.size > 0


def test_parser_accepts_the_documented_syntax():
    c = stim_b200.Circuit("""
        # comment
        h[tag] 0 1   # lower case names, tags
        CNOT 0 1
        ZCX 2 3
        M !0 1
        MPP !X0*Y1*Z2 X3
        REPEAT 3 {
            MR 0
            DETECTOR(1, 2) rec[-1] rec[-2]
        }
        QUBIT_COORDS(1, 2) 5
        SHIFT_COORDS(0, 0, 1)
        CX sweep[5] 0
        OBSERVABLE_INCLUDE(2) rec[-1] X0 Z1
    """)
    assert c.num_measurements == 2 + 2 + 3 and c.num_detectors == 3 and c.num_observables == 3


# ---------------------------------------------------------------------------------------------------------
# lowering == reference semantics (program emulator vs circuit-level oracle), incl. the stream's race contract
# ---------------------------------------------------------------------------------------------------------
LOWERING_CIRCUITS = [
    ("repetition_code", "memory", 3, 4, 0.05),
    ("surface_code", "rotated_memory_z", 3, 3, 0.02),
    ("surface_code", "unrotated_memory_x", 3, 2, 0.03),
    ("color_code", "memory_xyz", 3, 3, 0.02),
]


@pytest.mark.parametrize("code,task,d,r,p", LOWERING_CIRCUITS)
@pytest.mark.parametrize("mode", [0, 1])
def test_lowered_program_equals_oracle(code, task, d, r, p, mode):
    text = gen_circuit(code, task, d, r, p)
    for slots, chunk in ((32, 0), (5, 256)):
        w, plan = lower(text, mode, slots, chunk)
        em = pe.emulate(w, plan, seed=7, K=2, n_blocks=2, col0=10)
        if mode == 0:
            dd, oo = fo.sample(text, 2 * 2 * 128, 7, 2, "detectors", col0=10)
            ref = np.concatenate([dd, oo], axis=1)
        else:
            ref = fo.sample(text, 2 * 2 * 128, 7, 2, "measurements", col0=10)
        np.testing.assert_array_equal(em, ref)


def test_lowered_program_equals_oracle_on_every_instruction():
    from test_gpu_parity import ALL_OPS

    for mode in (0, 1):
        for slots in (32, 3):
            w, plan = lower(ALL_OPS, mode, slots)
            em = pe.emulate(w, plan, seed=3, K=1, n_blocks=2, col0=4)
            if mode == 0:
                dd, oo = fo.sample(ALL_OPS, 256, 3, 1, "detectors", col0=4)
                ref = np.concatenate([dd, oo], axis=1)
            else:
                ref = fo.sample(ALL_OPS, 256, 3, 1, "measurements", col0=4)
            np.testing.assert_array_equal(em, ref)


def test_batches_are_padded_into_chunks_and_large_circuits_lower_quickly():
    with open(os.path.join(ROOT, "tests", "golden", "circuits", "c3_surface_z_d25_r25.stim")) as f:
        text = f.read()
    w, plan = lower(text, 0, 640)
    p = pe.plan_dict(plan)
    assert p["n_words"] == w.size and w.size % p["chunk_words"] == 0
    assert p["num_qubits"] == 1249 and p["rec_ring"] == 2048 and p["num_det"] == 15600


# ---------------------------------------------------------------------------------------------------------
# multi-rank path: shots shard with no data-path collective; only the flip-count sum is reduced
# ---------------------------------------------------------------------------------------------------------
def test_shard_ranges_partition_the_shot_space():
    for total in (1 << 20, 1000_003, 128, 64):
        for world in (1, 2, 4, 8):
            ranges = [sharding.shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0
            for (a, n), (b, _) in zip(ranges, ranges[1:]):
                assert a + n == b and a % 128 == 0
            assert ranges[-1][0] + ranges[-1][1] == total


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["REPO_ROOT"])
import numpy as np
import torch.distributed as dist
from oracle import frame_oracle as fo
from stim_b200 import sharding

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
text = open(os.environ["CIRCUIT"]).read()
total, K, seed = 2048, 2, 9
first, count = sharding.shard_range(total, rank, world)
dets, obs = fo.sample(text, count, seed, K, "detectors", col0=first // 128)   # this rank's shard of the global shot space
local = np.concatenate([dets, obs], axis=1).sum(axis=0).astype(np.uint64)
reduced = sharding.allreduce_counts(local)
if rank == 0:
    d1, o1 = fo.sample(text, total, seed, K, "detectors", col0=0)             # the same shots in one process
    want = np.concatenate([d1, o1], axis=1).sum(axis=0).astype(np.uint64)
    assert np.array_equal(reduced, want), (reduced, want)
    print("GLOO_OK", int(reduced.sum()))
dist.destroy_process_group()
"""


def test_world_size_2_gloo_flip_count_allreduce(tmp_path):
    """Shard-count invariance: 2 ranks sampling disjoint shot ranges + allreduce == 1 rank sampling all of them."""
    circuit = tmp_path / "c.stim"
    circuit.write_text(gen_circuit("surface_code", "rotated_memory_z", 3, 3, 0.02))
    worker = tmp_path / "worker.py"
    worker.write_text(GLOO_WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, CIRCUIT=str(circuit))
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29517", str(worker)],
        env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GLOO_OK" in r.stdout


def test_stabiliser_comparison_detectors_are_fused_into_the_measurements():
    """lowering.cc fuse_detectors: a detector = (fresh result) xor (one earlier record row) is computed by the MEASURE
    item (GF_DET payload); detectors with other shapes stay XORROWS batches. The emulator executes both forms."""
    text = gen_circuit("surface_code", "rotated_memory_z", 5, 4, 0.01)
    w, plan = lower(text, 0, 64, 0)
    pl = pe.plan_dict(plan)
    pc, chunk, fused_items, xor_items = 0, pl["chunk_words"], 0, 0
    while True:
        h = int(w[pc])
        op = h & 0xFF
        if op == pe.OP_END:
            break
        if op == pe.OP_NEXT:
            pc = (pc // chunk + 1) * chunk
            continue
        n, words = int(w[pc + 1]), int(w[pc + 2])
        if op == pe.OP_MEASURE and (h >> 8) & pe.F_DET:
            assert words == pe.HDR + 3 * n
            fused_items += sum(1 for i in range(n) if int(w[pc + pe.HDR + 3 * i + 1]) != 0xFFFFFFFF)
        if op == pe.OP_XORROWS:
            xor_items += n
        pc += words
    # d=5 rotated memory: 24 stabilisers; rounds 2..4 compare with the previous round (72 fused detectors); the first
    # round (12 single-row detectors), the final data-qubit detectors (12) and the observable stay XORROWS items
    assert fused_items == 72 and xor_items == 12 + 12 + 1 + 1


def _golden_detect_cases():
    from golden_util import load_cases

    return [c for c in load_cases() if c["mode"] == "detect"]


@pytest.mark.parametrize("case", _golden_detect_cases(), ids=lambda c: c["name"])
def test_lowered_golden_circuits_reproduce_the_reference_on_the_emulator(case):
    """The lowered instruction stream (detector fusion, batching, barriers) of every deterministic reference test circuit,
    executed by the program emulator with its race detector on, gives the reference CLI's detection events."""
    from golden_util import arrange, expected_bits

    shots = next(iter(case["outputs"].values()))["shots"]
    K = 1
    n_blocks = (shots + 127) // 128
    for slots in (3,):  # few thread groups: many cross-group hazards, i.e. the barrier analysis is exercised hardest
        w, plan = lower(case["circuit"], 0, slots, 0)
        pl = pe.plan_dict(plan)
        out = pe.emulate(w, plan, seed=9, K=K, n_blocks=n_blocks, col0=0)[:shots]
        dets, obs = out[:, : pl["num_det"]], out[:, pl["num_det"]:]
        got = arrange(dets, obs, case["flags"])
        np.testing.assert_array_equal(got, expected_bits(case, got.shape[1])[:shots])


def test_reference_sample_matches_the_reference_and_its_own_slow_path(monkeypatch):
    """gstim_reference_sample (host inverse-tableau simulator, replaces TableauSimulator::reference_sample_circuit):
    equals the reference's sample on the golden circuits, and its fast paths (fused column pass for random measurements,
    transposed working copy for instructions with many targets, in-place row products) equal the step-by-step path on
    random Clifford circuits."""
    import random

    from golden_util import load_cases
    from stim_b200 import _reference_sample as rs

    for case in load_cases():
        if case["mode"] != "sample":
            continue
        ref = np.array([int(ch) for ch in case["reference_sample"]], dtype=np.uint8)
        got = np.unpackbits(rs.reference_sample_bits(case["circuit"], ref.size), bitorder="little")[: ref.size]
        np.testing.assert_array_equal(got, ref, err_msg=case["name"])

    c = stim_b200.Circuit("X 1\nM 0 1\nH 2\nM 2\nMR !1")  # Circuit.reference_sample mirror
    np.testing.assert_array_equal(c.reference_sample(), [False, True, False, False])
    np.testing.assert_array_equal(c.reference_sample(bit_packed=True), [2])

    g1 = ["H", "S", "S_DAG", "SQRT_X", "SQRT_X_DAG", "SQRT_Y", "SQRT_Y_DAG", "H_XY", "H_YZ", "C_XYZ", "C_ZYX", "X", "Y", "Z"]
    g2 = ["CX", "CY", "CZ", "SWAP", "ISWAP", "ISWAP_DAG", "SQRT_XX", "SQRT_YY", "SQRT_ZZ", "XCX", "XCY", "XCZ", "YCX", "YCY", "YCZ",
          "CXSWAP", "SWAPCX", "CZSWAP"]
    ms = ["M", "MX", "MY", "MR", "MRX", "MRY", "R", "RX", "RY"]
    rng = random.Random(5)
    for _ in range(300):
        n = rng.choice([2, 3, 5, 9, 17, 70, 130])
        lines = []
        for _ in range(rng.choice([10, 40, 150])):
            r = rng.random()
            if r < 0.4:
                lines.append(f"{rng.choice(g1)} {rng.randrange(n)}")
            elif r < 0.8:
                a, b = rng.sample(range(n), 2)
                lines.append(f"{rng.choice(g2)} {a} {b}")
            elif r < 0.9:
                lines.append(f"{rng.choice(ms)} {rng.randrange(n)}")
            elif r < 0.95:  # many targets in one instruction (the transposed working copy of tableau_ref.cc)
                gate = rng.choice(ms)
                lines.append(f"{gate} " + " ".join(("!" if gate[0] == "M" and rng.random() < 0.1 else "") + str(q)
                                                   for q in rng.sample(range(n), rng.randrange(1, n + 1))))
            else:
                lines.append("MPP " + "*".join(rng.choice("XYZ") + str(q) for q in rng.sample(range(n), min(n, 3))))
        lines.append("M " + " ".join(map(str, range(n))))
        text = "\n".join(lines)
        m = stim_b200.Circuit(text).num_measurements
        monkeypatch.setenv("GSTIM_TABLEAU_SLOW", "1")
        slow = rs.reference_sample_bits(text, m)
        monkeypatch.setenv("GSTIM_TABLEAU_SLOW", "0")
        monkeypatch.setenv("GSTIM_TABLEAU_TRANSPOSE", "off")
        fast = rs.reference_sample_bits(text, m)
        np.testing.assert_array_equal(fast, slow)
        monkeypatch.setenv("GSTIM_TABLEAU_TRANSPOSE", "force")
        transposed = rs.reference_sample_bits(text, m)
        np.testing.assert_array_equal(transposed, slow)
        monkeypatch.delenv("GSTIM_TABLEAU_TRANSPOSE")
