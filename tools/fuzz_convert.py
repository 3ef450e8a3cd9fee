"""Differential run of `python -m stim_b200 convert` against the reference CLI (`oracle/_ref/stim convert`) on random record
layouts, densities and format pairs (this container only: needs oracle/_ref/stim). usage: python tools/fuzz_convert.py [n] [seed]"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stim_b200.__main__ as cli

STIM = os.path.join(ROOT, "oracle", "_ref", "stim")
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
fmts = ["01", "b8", "r8", "hits", "dets"]
bad = 0
with tempfile.TemporaryDirectory() as tmp:
    for it in range(n_iter):
        kind = rng.integers(0, 3)
        nm, nd, no = [(int(rng.integers(1, 200)), 0, 0), (0, int(rng.integers(1, 200)), int(rng.integers(0, 5))),
                      (int(rng.integers(0, 70)), int(rng.integers(0, 70)), int(rng.integers(0, 4)))][kind]
        n = nm + nd + no
        if n == 0:
            continue
        shots = int(rng.integers(0, 80))
        bits = (rng.random((shots, n)) < rng.choice([0.0, 0.01, 0.1, 0.5, 1.0])).astype(np.uint8)
        counts = ["--num_measurements", str(nm), "--num_detectors", str(nd), "--num_observables", str(no)]
        fin, fout, fobs = rng.choice(fmts), rng.choice(fmts), rng.choice(fmts)
        text = "".join("".join(map(str, r)) + "\n" for r in bits).encode()
        data = subprocess.run([STIM, "convert", "--in_format", "01", "--out_format", fin] + counts, input=text, capture_output=True).stdout
        obs = bool(no) and bool(rng.integers(0, 2))
        paths = {k: os.path.join(tmp, k) for k in ("in", "r_out", "r_obs", "m_out", "m_obs")}
        open(paths["in"], "wb").write(data)
        flags = ["--in_format", fin, "--out_format", fout, "--obs_out_format", fobs] + counts + ["--in", paths["in"]]
        r = subprocess.run([STIM, "convert"] + flags + ["--out", paths["r_out"]] + (["--obs_out", paths["r_obs"]] if obs else []),
                           capture_output=True)
        rc = cli.main(["convert"] + flags + ["--out", paths["m_out"]] + (["--obs_out", paths["m_obs"]] if obs else []))
        same = rc == r.returncode and (rc != 0 or (open(paths["r_out"], "rb").read() == open(paths["m_out"], "rb").read() and (
            not obs or open(paths["r_obs"], "rb").read() == open(paths["m_obs"], "rb").read())))
        if not same:
            bad += 1
            print("MISMATCH", it, fin, fout, fobs, nm, nd, no, shots, obs, rc, r.returncode)
print(n_iter, "cases,", bad, "mismatches")
sys.exit(1 if bad else 0)
