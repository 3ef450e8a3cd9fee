import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import stim_b200
from test_gpu_parity import ALL_OPS
no_else = "\n".join(l for l in ALL_OPS.split("\n") if not l.startswith("ELSE_CORRELATED_ERROR"))
d5 = open("/root/repo/tests/golden/circuits/c2_surface_x_d5_r5.stim").read()
for name, text in (("all_ops_no_else", no_else), ("surface_d5", d5)):
    s = stim_b200.Circuit(text).compile_detector_sampler(seed=5, engine="events")
    a = s.sample(3000, bit_packed=True, append_observables=True)
    d, o = s.sample(1000, separate_observables=True)
    print(name, int(a.sum()), int(d.sum()), flush=True)
