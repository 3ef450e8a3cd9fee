"""Generates tests/golden/c3_det_rows.json and c5_det_rows.json: the (shot-independent) b8 detection-event row the unmodified
reference produces for the full-size d=25 r=25 (and d=51 r=51) benchmark circuits when one family of X_ERROR flips has probability 1 and all
other noise is off (transform: tests/test_gpu_golden.py:_c3_variant). Needs oracle/_ref/stim."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("PYTEST_DISABLE_PLUGIN_AUTOLOAD", "1")
import importlib.util

spec = importlib.util.spec_from_file_location("tgg", os.path.join(ROOT, "tests", "test_gpu_golden.py"))
tgg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tgg)
STIM = os.path.join(ROOT, "oracle", "_ref", "stim")
for cfg, variant, nbytes in (("c3", tgg._c3_variant, 1951), ("c5", tgg._c5_variant, 16576)):
    out = {}
    for knob in ("measure", "reset"):
        text = variant(knob)
        rows = []
        for seed in ("1", "2"):
            r = subprocess.run([STIM, "detect", "--shots", "3", "--out_format", "b8", "--append_observables", "--seed", seed],
                               input=text.encode(), capture_output=True, check=True).stdout
            assert len(r) == 3 * nbytes
            rows += [r[i * nbytes:(i + 1) * nbytes] for i in range(3)]
        assert all(x == rows[0] for x in rows), "not deterministic"
        assert any(rows[0]), "row is all zero"
        out[knob] = rows[0].hex()
        print(cfg, knob, sum(bin(b).count("1") for b in rows[0]), "detection events per shot")
    with open(os.path.join(ROOT, "tests", "golden", cfg + "_det_rows.json"), "w") as f:
        json.dump(out, f)
