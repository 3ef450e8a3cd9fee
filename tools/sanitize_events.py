"""compute-sanitizer workload for the event engine only: every opcode, a d=5 surface code over enough tiles that every image is
recycled several times (tickets, slice counts, writer hand-over), the same with the folded table, and a dense class.
usage: compute-sanitizer --tool racecheck python tools/sanitize_events.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import stim_b200
from test_gpu_parity import ALL_OPS

d5 = open(os.path.join(ROOT, "tests", "golden", "circuits", "c2_surface_x_d5_r5.stim")).read()
dense = "R 0 1 2\nX_ERROR(0.25) 0 1 2\nDEPOLARIZE2(0.3) 0 1\nM 0 1 2\nDETECTOR rec[-1]\nDETECTOR rec[-2]\nDETECTOR rec[-3]\n"
cases = [("all_ops", ALL_OPS, 1000), ("surface_d5", d5, 1 << 18), ("dense", dense, 1 << 16)]
for fold in ("48", "0"):
    os.environ["GSTIM_TABLE_COMPRESS_MB"] = fold
    for name, text, shots in cases:
        s = stim_b200.Circuit(text).compile_detector_sampler(seed=5, engine="events")
        a = s.sample(shots, bit_packed=True, append_observables=True)
        d, o = s.sample(shots // 2 + 3, bit_packed=True, separate_observables=True)
        b = s.bit_counts(777)
        m = stim_b200.Circuit(text).compile_sampler(seed=5, engine="events").sample(600, bit_packed=True)
        print(name, "fold<=" + fold, s.engine_info()["tile_shots"], int(a.sum()), int(d.sum()), int(o.sum()), int(b[0].sum()), int(m.sum()), flush=True)
