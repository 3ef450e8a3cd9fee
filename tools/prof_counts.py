import sys, time
sys.path.insert(0, ".")
import stim_b200
s = stim_b200.Circuit(open("tests/golden/circuits/c3_surface_z_d25_r25.stim").read()).compile_detector_sampler(seed=1, engine="events")
for pairs in (False, True):
    for _ in range(3):
        t = time.perf_counter(); s.bit_counts(1 << 24, pairs=pairs); dt = time.perf_counter() - t
    print("pairs", pairs, round(dt * 1e3, 1), "ms per 2^24 shots (sampling ~17.5 ms)")
import gzip
m = stim_b200.DetectorErrorModel(gzip.open("tests/golden/dem/c3_surface_z_d25_r25.dem.gz", "rt").read())
d = m.compile_sampler(seed=1)
t = d.response_table()
print("dem classes", len(t["classes"]), "slices", len(t["slices"]), "tile", t["tile_shots"])
