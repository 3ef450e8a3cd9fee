"""ncu driver for the event engine: python tools/prof_ev_run.py <circuit> <shots_log2> [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stim_b200

path = sys.argv[1]
shots = 1 << int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
c = stim_b200.Circuit(open(path).read())
s = c.compile_detector_sampler(seed=1, engine="events")
nb = (c.num_detectors + c.num_observables + 7) // 8
out = torch.empty((shots, nb), dtype=torch.uint8, device="cuda")
print(s.engine_info())
for i in range(reps):
    s.sample_device(shots, out.data_ptr(), append_observables=True)
    print(f"rep {i}: call {s.last_call_ms():.3f} ms -> {shots / (s.last_call_ms() * 1e-3) / 1e6:.1f} Mshots/s")
