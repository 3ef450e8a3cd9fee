"""Measurement sampling (`compile_sampler().sample`, `stim sample`) on the BASELINE circuits, device-resident rows, both engines,
beside the reference CLI on the host cores. usage: python tools/prof_measure.py [shots_log2] [names...] -> one JSON line"""
import glob
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stim_b200

log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20
names = sys.argv[2:] or ["c3", "c4_"]
REF = os.path.join(ROOT, "oracle", "_ref", "stim")
out = {}
for name in names:
    f = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "circuits", name + "*.stim")))[0]
    text = open(f).read()
    row = {}
    for engine in ["events", "interp"]:
        t0 = time.time()
        try:
            s = stim_b200.Circuit(text).compile_sampler(seed=1, engine=engine)
        except ValueError as e:
            row[engine] = {"unavailable": str(e)}
            continue
        t_compile = time.time() - t0
        M = int(s.stats.num_measurements)
        nb = (M + 7) // 8
        shots = 1 << log2
        buf = torch.empty((shots, nb), dtype=torch.uint8, device="cuda")
        best = 1e9
        for _ in range(3):
            s.sample_device(shots, buf.data_ptr())
            best = min(best, s.last_call_ms())
        row[engine] = {"shots_per_s": shots / best * 1e3, "ms": best, "shots": shots, "bytes_per_shot": nb, "create_s": t_compile,
                       "nonzero_byte_fraction": float(torch.count_nonzero(buf[:4096]).item()) / (4096 * nb)}
    if os.path.exists(REF):
        nproc = len(os.sched_getaffinity(0))
        sp = 16384
        t0 = time.time()
        ps = [subprocess.Popen([REF, "sample", "--shots", str(sp), "--in", f, "--out_format", "b8", "--out", "/dev/null", "--seed",
                                str(i)]) for i in range(nproc)]
        assert all(p.wait() == 0 for p in ps)
        row["reference_cli"] = {"shots_per_s": nproc * sp / (time.time() - t0), "sample": f"{nproc} processes x {sp} shots, stim sample b8"}
    out[os.path.basename(f)] = row
print(json.dumps(out))
