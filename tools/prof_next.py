"""Throughput of the SURVEY 8(f) rows on the headline circuit (c3, d=25 r=25 p=1e-3), with the reference CLI's own implementation
of each timed beside it on the box's host cores (bounded samples):   python tools/prof_next.py > profiles/r2_next_rows.json
  sample_dem  the circuit's detector error model (356 321 mechanisms): device-resident counts, packed rows to host; `stim sample_dem`
  m2d         measurements -> detection events, host rows in and out; `stim m2d`
  FlipSimulator.do(circuit)  one batch of 2^17 instances, tables stay on the device (no CLI equivalent: reported alone)"""
import gzip, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import stim_b200

STIM = os.path.join(ROOT, "oracle", "_ref", "stim")
CIRC = os.path.join(ROOT, "tests", "golden", "circuits", "c3_surface_z_d25_r25.stim")
text = open(CIRC).read()
circ = stim_b200.Circuit(text)
cores = os.cpu_count() or 1
out = {"circuit": "c3 surface_code:rotated_memory_z d=25 r=25 p=1e-3", "host_cores": cores}


def best(f, reps=3):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t)
    return min(ts)


def ref_parallel(cmd_of, shots_per_proc):
    """all cores, one process each (like bench.py --impl reference)"""
    t = time.perf_counter()
    ps = [subprocess.Popen(cmd_of(i), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(cores)]
    for p in ps:
        p.wait()
    return cores * shots_per_proc / (time.perf_counter() - t)


# ---- sample_dem -------------------------------------------------------------------------------------------------------
dem_text = gzip.open(os.path.join(ROOT, "tests", "golden", "dem", "c3_surface_z_d25_r25.dem.gz"), "rt").read()
dem = stim_b200.DetectorErrorModel(dem_text)
ds = dem.compile_sampler(seed=1)
ds.bit_counts(1 << 16)
n = 1 << 22
t = best(lambda: ds.bit_counts(n))
row = {"num_errors": dem.num_errors, "device_resident_counts_shots_per_s": n / t}
n = 1 << 20
ds.sample(1 << 12, bit_packed=True)
t = best(lambda: ds.sample(n, bit_packed=True))
row["packed_rows_to_host_shots_per_s"] = n / t
if os.path.exists(STIM):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "m.dem")
        open(p, "w").write(dem_text)
        per = 1 << 14
        row["reference_cli_shots_per_s"] = ref_parallel(
            lambda i: [STIM, "sample_dem", "--shots", str(per), "--in", p, "--out_format", "b8", "--out", "/dev/null", "--seed", str(i)], per)
        row["reference_sample"] = f"{cores} processes x {per} shots, stim sample_dem --out_format b8"
out["sample_dem"] = row

# ---- m2d ---------------------------------------------------------------------------------------------------------------
conv = circ.compile_m2d_converter()
n = 1 << 20
meas = circ.compile_sampler(seed=3).sample(n, bit_packed=True)
conv.convert(measurements=meas[:4096], append_observables=True, bit_packed=True)
t = best(lambda: conv.convert(measurements=meas, append_observables=True, bit_packed=True))
row = {"host_rows_in_and_out_shots_per_s": n / t, "bytes_in_per_shot": int(meas.shape[1]),
       "bytes_out_per_shot": (circ.num_detectors + circ.num_observables + 7) // 8}
if os.path.exists(STIM):
    with tempfile.TemporaryDirectory() as d:
        per = 1 << 15
        pm = os.path.join(d, "m.b8")
        meas[:per].tofile(pm)
        row["reference_cli_shots_per_s"] = ref_parallel(
            lambda i: [STIM, "m2d", "--circuit", CIRC, "--in", pm, "--in_format", "b8", "--out_format", "b8", "--out", "/dev/null",
                       "--append_observables"], per)
        row["reference_sample"] = f"{cores} processes x {per} shots, stim m2d b8 -> b8 (includes its reference sample of the circuit)"
out["m2d"] = row

# ---- FlipSimulator -----------------------------------------------------------------------------------------------------
batch = 1 << 17
def run_flipsim():
    sim = stim_b200.FlipSimulator(batch_size=batch, num_qubits=circ.num_qubits, seed=5)
    sim.do(text)
    return sim
run_flipsim()
t = best(run_flipsim, reps=2)
out["flip_simulator"] = {"batch_size": batch, "do_whole_circuit_shots_per_s": batch / t,
                         "note": "lowering of the fragment + one kernel per batch over HBM-resident tables; create + do, tables left on the device"}
print(json.dumps(out, indent=1))
