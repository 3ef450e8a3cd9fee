"""Runs tests/test_gpu_events.py::test_random_circuits_match_oracle_on_both_engines on more seeds than the suite does
(both engines bit for bit against their oracles, and against each other statistically).  python tools/fuzz_more.py [first] [last]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import pytest
import test_gpu_events as t
ok = skipped = 0
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6, int(sys.argv[2]) if len(sys.argv) > 2 else 70):
    try:
        t.test_random_circuits_match_oracle_on_both_engines(seed)
        ok += 1
    except pytest.skip.Exception:
        skipped += 1
    except Exception as e:
        print("FAIL seed", seed, repr(e)[:400]); raise
print("fuzz ok", ok, "skipped", skipped)
