"""Generates tests/golden/stats_big/<name>.npz: per-bit flip counts and adjacent-pair (AND) counts of the unmodified
reference (oracle/_ref/stim detect / sample) over N_REF = 2^24 shots for the full-size BASELINE.json configs (c3, c4, the
c4 MPP/dense variant, c5 and the c5 feedback measurement-sampling variant). The statistical-parity tests
(tests/test_gpu_stats_big.py) compare the CUDA sampler's on-device counts against these within 5 sigma, the method of
/root/reference/src/stim/cmd/command_sample.test.cc:39-71 in its two-sample form.

The reference streams ptb64 (64 shots per u64 word, bit-major inside a group), so counts are popcounts of words and of
ANDs of neighbouring words. Needs oracle/_ref/stim; ~10 minutes on 8 cores.

    python tools/gen_stats_big.py [name ...]
"""
import os
import subprocess
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
STIM = os.path.join(ROOT, "oracle", "_ref", "stim")
CDIR = os.path.join(ROOT, "tests", "golden", "circuits")
ODIR = os.path.join(ROOT, "tests", "golden", "stats_big")
N_REF = 1 << 24


def _read(name):
    with open(os.path.join(CDIR, name)) as f:
        return f.read()


def configs():
    import gen_variants

    return {
        "c3_surface_z_d25_r25": ("detect", lambda: _read("c3_surface_z_d25_r25.stim")),
        "c4_color_d15_r15": ("detect", lambda: _read("c4_color_d15_r15.stim")),
        "c4v_color_d15_r15_mpp_dense": ("detect", lambda: _read("c4v_color_d15_r15_mpp_dense.stim")),
        "c5_surface_x_d51_r51": ("detect", lambda: _read("c5_surface_x_d51_r51.stim")),
        "c5f_surface_x_d51_r51_feedback": ("sample", gen_variants.c5_feedback),
    }


def one_chunk(job):
    mode, path, n_bits, shots, seed = job
    args = [STIM, mode, "--shots", str(shots), "--in", path, "--out_format", "ptb64", "--seed", str(seed)]
    if mode == "detect":
        args.append("--append_observables")
    p = subprocess.Popen(args, stdout=subprocess.PIPE)
    single = np.zeros(n_bits, dtype=np.int64)
    pair = np.zeros(n_bits - 1, dtype=np.int64)
    group_bytes = n_bits * 8
    per_read = max(1, (64 << 20) // group_bytes)
    left = shots // 64
    while left:
        g = min(per_read, left)
        buf = p.stdout.read(g * group_bytes)
        assert len(buf) == g * group_bytes, "short read from the reference"
        w = np.frombuffer(buf, dtype=np.uint64).reshape(g, n_bits)
        single += np.bitwise_count(w).sum(axis=0, dtype=np.int64)
        pair += np.bitwise_count(w[:, :-1] & w[:, 1:]).sum(axis=0, dtype=np.int64)
        left -= g
    assert p.stdout.read(1) == b""
    assert p.wait() == 0
    return single, pair


def main():
    import stim_b200

    os.makedirs(ODIR, exist_ok=True)
    want = sys.argv[1:] or list(configs())
    for name in want:
        mode, text_fn = configs()[name]
        text = text_fn()
        path = os.path.join("/tmp", name + ".stim")
        with open(path, "w") as f:
            f.write(text)
        c = stim_b200.Circuit(text)
        n_bits = c.num_detectors + c.num_observables if mode == "detect" else c.num_measurements
        chunk = 1 << (16 if n_bits > 50000 else 18)
        jobs = [(mode, path, n_bits, chunk, 424200 + i) for i in range(N_REF // chunk)]
        single = np.zeros(n_bits, dtype=np.int64)
        pair = np.zeros(n_bits - 1, dtype=np.int64)
        with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
            for s, p in ex.map(one_chunk, jobs):
                single += s
                pair += p
        np.savez_compressed(os.path.join(ODIR, name + ".npz"), mode=mode, n_ref=N_REF, n_bits=n_bits,
                            single=single.astype(np.uint32), pair=pair.astype(np.uint32))
        print(name, mode, n_bits, "bits; mean rate", float(single.mean()) / N_REF, "mean pair rate", float(pair.mean()) / N_REF,
              flush=True)


if __name__ == "__main__":
    main()
