"""Times the detector-error-model sampler: python tools/prof_dem.py <model.dem[.gz]> <shots_log2>"""
import gzip, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stim_b200
path, shots = sys.argv[1], 1 << int(sys.argv[2])
text = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
m = stim_b200.DetectorErrorModel(text)
s = m.compile_sampler(seed=1)
for rep in range(3):
    t = time.perf_counter()
    single, pair = s.bit_counts(shots)
    dt = time.perf_counter() - t
    print(f"rep {rep}: {m.num_errors} errors, {shots} shots, {dt * 1e3:.1f} ms -> {shots / dt / 1e6:.1f} Mshots/s (device-resident counts), mean rate {single[:m.num_detectors].mean() / shots:.5f}")
