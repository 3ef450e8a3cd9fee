"""compute-sanitizer workload for the interpreter only: every opcode incl. E / ELSE chains (tools/sanitize_run.py covers both engines)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import stim_b200
from test_gpu_parity import ALL_OPS
s = stim_b200.Circuit(ALL_OPS).compile_detector_sampler(seed=5, engine="interp")
a = s.sample(1000, bit_packed=True, append_observables=True)
m = stim_b200.Circuit(ALL_OPS).compile_sampler(seed=5, engine="interp").sample(600, bit_packed=True)
print("all_ops interp", int(a.sum()), int(m.sum()), flush=True)
