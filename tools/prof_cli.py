"""Times the detect path to a file descriptor (gstim_sample_detectors_to_fd: what `stim detect` maps onto) per output format on c3,
next to the reference CLI on all host cores.   python tools/prof_cli.py"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stim_b200
STIM = os.path.join(ROOT, "oracle", "_ref", "stim")
CIRC = os.path.join(ROOT, "tests", "golden", "circuits", "c3_surface_z_d25_r25.stim")
circ = stim_b200.Circuit(open(CIRC).read())
s = circ.compile_detector_sampler(seed=1)
cores = os.cpu_count() or 1
out = {"host_cores": cores}
for fmt, log2 in (("b8", 22), ("r8", 20), ("hits", 19), ("dets", 19), ("01", 17), ("ptb64", 20)):
    n = 1 << log2
    s.sample_write(4096, filepath="/dev/null", format=fmt, append_observables=True)
    t = time.perf_counter()
    s.sample_write(n, filepath="/dev/null", format=fmt, append_observables=True)
    dt = time.perf_counter() - t
    row = {"shots": n, "shots_per_s": n / dt}
    if os.path.exists(STIM):
        per = max(n // cores // 8, 1024)
        t = time.perf_counter()
        ps = [subprocess.Popen([STIM, "detect", "--shots", str(per), "--in", CIRC, "--out_format", fmt, "--out", "/dev/null",
                                "--append_observables", "--seed", str(i)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(cores)]
        for p in ps:
            p.wait()
        row["reference_cli_shots_per_s"] = cores * per / (time.perf_counter() - t)
    out[fmt] = row
print(json.dumps(out, indent=1))
