"""Times the default host path: sample(bit_packed=True) into a fresh pageable numpy array.  python tools/prof_pageable.py [log2 shots]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stim_b200
shots = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
c = stim_b200.Circuit(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "circuits", "c3_surface_z_d25_r25.stim")).read())
s = c.compile_detector_sampler(seed=1)
s.sample(1 << 16, bit_packed=True, append_observables=True)
for i in range(4):
    t0 = time.perf_counter()
    a = s.sample(shots, bit_packed=True, append_observables=True)
    dt = time.perf_counter() - t0
    print(f"pageable sample {shots} shots: {shots / dt / 1e6:.2f} M shots/s ({a.nbytes / dt / 1e9:.1f} GB/s)", flush=True)
    del a
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
