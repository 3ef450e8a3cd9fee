"""Generates the fixtures of the detector-error-model sampler tests from the UNMODIFIED reference (oracle/_ref/stim):

  tests/golden/dem/*.dem[.gz]          `stim analyze_errors` of the committed benchmark circuits (c2 flat, c3 with --fold_loops)
  tests/golden/dem_cases.json          `stim sample_dem` outputs (detectors / observables / errors, several formats) of
                                       deterministic models (every probability in {0, 1}; checked with two seeds)
  tests/golden/stats_big/dem_*.npz     per-detector flip counts, adjacent-pair counts and per-observable counts of
                                       `stim sample_dem` over 2^24 shots, for the 5 sigma tests

  tests/golden/dem_replay_cases.json   errors sampled by the reference from noisy models + the detectors / observables it derives
                                       from them (also through its own --replay_err_in)

    python tools/gen_dem_golden.py [cases] [replay] [stats]"""
import base64
import gzip
import json
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STIM = os.path.join(ROOT, "oracle", "_ref", "stim")
N_REF = 1 << 24

CASES = {
    "readme_example": "error(0) D0\nerror(1) D1 L0\nerror(1) D2 D3\n",
    "separators_and_cancellation": "error(1) D0 D1 ^ D1 D2 L0 ^ L0 L1\nerror(1) D0\nerror(0) D5\ndetector(1, 2) D7\nlogical_observable L3\n",
    "repeat_and_shift": "error(1) D0\nrepeat 5 {\n    error(1) D1 D2 L0\n    error(0) D0\n    shift_detectors(0, 1) 3\n}\nerror[tagged](1) D0 L2\n",
    "nested_repeat": "repeat 3 {\n    repeat 2 {\n        error(1) D0 L1\n        shift_detectors 1\n    }\n    error(1) D0 D1\n    shift_detectors 2\n}\ndetector D4\n",
    "many_errors": "".join(f"error({(k * 7) % 3 == 0 and 1 or 0}) D{k} D{(k * 5 + 1) % 97} L{k % 3}\n" for k in range(97)),
    "no_detectors": "error(1) L0\nerror(1) L0 L1\n",
}
FORMATS = ("01", "b8", "hits", "dets", "r8", "ptb64")


def run_sample_dem(text, shots, seed, fmt):
    with tempfile.TemporaryDirectory() as d:
        p = {k: os.path.join(d, k) for k in ("in", "det", "obs", "err")}
        with open(p["in"], "w") as f:
            f.write(text)
        subprocess.run([STIM, "sample_dem", "--shots", str(shots), "--seed", str(seed), "--in", p["in"], "--out", p["det"],
                        "--out_format", fmt, "--obs_out", p["obs"], "--obs_out_format", fmt, "--err_out", p["err"],
                        "--err_out_format", fmt], check=True)
        return {k: open(p[k], "rb").read() for k in ("det", "obs", "err")}


def gen_replay():
    """Noisy models: errors sampled by the reference, the detectors / observables it derives from them, and the reference's own
    `--replay_err_in` run on those errors (must agree) -> tests/golden/dem_replay_cases.json (b8, base64)."""
    import re

    models = {name: re.sub(r"\((0|1)\)", "(0.3)", text) for name, text in CASES.items()}
    models["c2_surface_d5"] = open(os.path.join(ROOT, "tests", "golden", "dem", "c2_surface_x_d5_r5.dem")).read()
    out = []
    for name, text in models.items():
        shots = 96
        with tempfile.TemporaryDirectory() as d:
            p = {k: os.path.join(d, k) for k in ("in", "det", "obs", "err", "det2", "obs2")}
            with open(p["in"], "w") as f:
                f.write(text)
            subprocess.run([STIM, "sample_dem", "--shots", str(shots), "--seed", "7", "--in", p["in"], "--out", p["det"], "--out_format", "b8",
                            "--obs_out", p["obs"], "--obs_out_format", "b8", "--err_out", p["err"], "--err_out_format", "b8"], check=True)
            subprocess.run([STIM, "sample_dem", "--shots", str(shots), "--in", p["in"], "--out", p["det2"], "--out_format", "b8",
                            "--obs_out", p["obs2"], "--obs_out_format", "b8", "--replay_err_in", p["err"], "--replay_err_in_format", "b8"],
                           check=True)
            files = {k: open(p[k], "rb").read() for k in ("det", "obs", "err", "det2", "obs2")}
        assert files["det"] == files["det2"] and files["obs"] == files["obs2"], name
        out.append({"name": name, "dem": text, "shots": shots, **{k: base64.b64encode(files[k]).decode() for k in ("det", "obs", "err")}})
    with open(os.path.join(ROOT, "tests", "golden", "dem_replay_cases.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out), "replay cases")


def gen_cases():
    out = []
    for name, text in CASES.items():
        entry = {"name": name, "dem": text, "outputs": {}}
        for fmt in FORMATS:
            shots = 64 if fmt == "ptb64" else 5
            a, b = run_sample_dem(text, shots, 1, fmt), run_sample_dem(text, shots, 2, fmt)
            assert a == b, (name, fmt)
            entry["outputs"][fmt] = {"shots": shots, **{k: base64.b64encode(v).decode() for k, v in a.items()}}
        out.append(entry)
    with open(os.path.join(ROOT, "tests", "golden", "dem_cases.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out), "dem cases")


def one_chunk(job):
    path, D, L, shots, seed = job
    with tempfile.TemporaryDirectory() as d:
        obs_path = os.path.join(d, "obs")
        p = subprocess.Popen([STIM, "sample_dem", "--shots", str(shots), "--seed", str(seed), "--in", path, "--out_format", "ptb64",
                              "--obs_out", obs_path, "--obs_out_format", "b8"], stdout=subprocess.PIPE)
        single = np.zeros(D, dtype=np.int64)
        pair = np.zeros(D - 1, dtype=np.int64)
        group_bytes = D * 8
        per_read = max(1, (64 << 20) // group_bytes)
        left = shots // 64
        while left:
            g = min(per_read, left)
            buf = p.stdout.read(g * group_bytes)
            assert len(buf) == g * group_bytes
            w = np.frombuffer(buf, dtype=np.uint64).reshape(g, D)
            single += np.bitwise_count(w).sum(axis=0, dtype=np.int64)
            pair += np.bitwise_count(w[:, :-1] & w[:, 1:]).sum(axis=0, dtype=np.int64)
            left -= g
        assert p.wait() == 0
        ob = np.fromfile(obs_path, dtype=np.uint8).reshape(shots, (L + 7) // 8)
        obs = np.unpackbits(ob, axis=1, bitorder="little")[:, :L].sum(axis=0, dtype=np.int64)
    return single, pair, obs


def gen_stats():
    import stim_b200

    for name, fixture in (("dem_c2_surface_x_d5_r5", "c2_surface_x_d5_r5.dem"), ("dem_c3_surface_z_d25_r25", "c3_surface_z_d25_r25.dem.gz")):
        src = os.path.join(ROOT, "tests", "golden", "dem", fixture)
        text = gzip.open(src, "rt").read() if src.endswith(".gz") else open(src).read()
        path = os.path.join("/tmp", name + ".dem")
        with open(path, "w") as f:
            f.write(text)
        m = stim_b200.DetectorErrorModel(text)
        D, L = m.num_detectors, m.num_observables
        chunk = 1 << 17
        jobs = [(path, D, L, chunk, 777000 + i) for i in range(N_REF // chunk)]
        single = np.zeros(D, dtype=np.int64)
        pair = np.zeros(D - 1, dtype=np.int64)
        obs = np.zeros(L, dtype=np.int64)
        with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
            for s, p, o in ex.map(one_chunk, jobs):
                single += s
                pair += p
                obs += o
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "stats_big", name + ".npz"), n_ref=N_REF, D=D, L=L,
                            single=single.astype(np.uint32), pair=pair.astype(np.uint32), obs=obs.astype(np.uint32))
        print(name, D, L, "mean rate", float(single.mean()) / N_REF, "obs", obs / N_REF, flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["cases", "stats"]
    if "cases" in what:
        gen_cases()
    if "replay" in what:
        gen_replay()
    if "stats" in what:
        gen_stats()
