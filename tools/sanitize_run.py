"""Small workloads for compute-sanitizer (memcheck / racecheck): every opcode of both engines plus a d=3 surface code.
usage: compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import stim_b200
from test_gpu_parity import ALL_OPS

d3 = open(os.path.join(ROOT, "tests", "golden", "circuits", "c2_surface_x_d5_r5.stim")).read()
for name, text in (("all_ops", ALL_OPS), ("surface_d5", d3)):
    for engine in ("interp", "events"):
        try:
            s = stim_b200.Circuit(text).compile_detector_sampler(seed=5, engine=engine)
        except ValueError:
            continue
        a = s.sample(1000, bit_packed=True, append_observables=True)
        b = s.bit_counts(777)
        m = stim_b200.Circuit(text).compile_sampler(seed=5, engine=engine).sample(600, bit_packed=True)
        print(name, engine, s.engine_info()["last_engine"], int(a.sum()), int(b[0].sum()), int(m.sum()), flush=True)
