"""Small driver for ncu / experiments: python tools/prof_run.py <circuit> <shots_log2> [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stim_b200

path = sys.argv[1]
shots = 1 << int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
c = stim_b200.Circuit(open(path).read())
s = c.compile_detector_sampler(seed=1)
nb = (c.num_detectors + c.num_observables + 7) // 8
out = torch.empty((shots, nb), dtype=torch.uint8, device="cuda")
st = s.stats
print("threads", st.threads, "G", st.lanes_per_item, "slots", st.slots, "Kmax", st.max_columns, "batches", st.num_batches,
      "barriers", st.num_barriers, "words", st.program_words, "smem", st.smem_bytes_max)
for i in range(reps):
    t = time.perf_counter()
    s.sample_device(shots, out.data_ptr(), append_observables=True)
    dt = time.perf_counter() - t
    a, b = s.last_kernel_ms()
    print(f"rep {i}: K={s.last_block_columns()} call {s.last_call_ms():.2f} ms interp {a:.2f} transpose {b:.2f} wall {dt*1e3:.2f} -> {shots/ (s.last_call_ms()*1e-3)/1e6:.1f} Mshots/s")
