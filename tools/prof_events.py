"""Times both engines on the BASELINE circuits (device-resident results). usage: python tools/prof_events.py [shots_log2] [names...]"""
import glob
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stim_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 22
names = sys.argv[2:] or ["c1", "c2", "c3", "c4_", "c4v_color_d15_r15_mpp_dense", "c5"]
for name in names:
    f = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "circuits", name + "*.stim")))[0]
    text = open(f).read()
    for engine in ["events", "interp"]:
        t0 = time.time()
        try:
            s = stim_b200.Circuit(text).compile_detector_sampler(seed=1, engine=engine)
        except ValueError as e:
            print(os.path.basename(f), engine, "n/a:", e)
            continue
        t_compile = time.time() - t0
        D, L = int(s.stats.num_detectors), int(s.stats.num_observables)
        nb = (D + L + 7) // 8
        shots = 1 << log2
        while shots * nb > (24 << 30):
            shots >>= 1
        out = torch.empty((shots, nb), dtype=torch.uint8, device="cuda")
        best = 1e9
        for _ in range(3):
            s.sample_device(shots, out.data_ptr(), append_observables=True)
            best = min(best, s.last_call_ms())
        info = s.engine_info()
        print(f"{os.path.basename(f):42s} {engine:6s} {shots / best / 1e3:9.1f} M shots/s  {best:8.2f} ms  "
              f"{shots * nb / best / 1e6:7.1f} GB/s  compile {t_compile:.2f}s  "
              f"tile {info['tile_shots']} ev/shot {info['events_per_shot']:.1f} flips {info['flips_per_shot']:.1f} favoured {info['favoured']} "
              f"nonzero-byte frac {float((out[:65536] != 0).float().mean()):.4f}", flush=True)
        del out
