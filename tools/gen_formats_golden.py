"""Golden vectors for the six shot-data formats, produced by the unmodified reference CLI (`oracle/_ref/stim convert`):
random bit tables written as 01, converted by the reference into every format. tests/test_formats.py decodes each with
stim_b200._formats and re-encodes with the library's writers.   python tools/gen_formats_golden.py"""
import base64, json, os, subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STIM = os.path.join(ROOT, "oracle", "_ref", "stim")
rng = np.random.default_rng(2024)
cases = []
for (nm, nd, no, shots, density) in [(1, 0, 0, 64, 0.5), (7, 0, 0, 64, 0.3), (10, 0, 0, 128, 0.1), (64, 0, 0, 64, 0.02), (130, 0, 0, 64, 0.5),
                                     (300, 0, 0, 128, 0.004), (0, 5, 3, 64, 0.3), (0, 70, 2, 128, 0.05), (520, 0, 0, 64, 0.0),
                                     (9, 0, 0, 3, 0.5), (0, 12, 1, 5, 0.2)]:
    n = nm + nd + no
    bits = (rng.random((shots, n)) < density).astype(np.uint8)
    text01 = "".join("".join(map(str, r)) + "\n" for r in bits).encode()
    out = {}
    for fmt in ("01", "b8", "r8", "hits", "dets", "ptb64"):
        if fmt == "ptb64":
            # the reference's per-record converter cannot WRITE ptb64; build the bytes here (64 shots per group, one 64-bit
            # little-endian word per bit) and let the reference READ them back into the original 01 text
            if shots % 64:
                continue
            raw = np.packbits(bits.reshape(shots // 64, 64, n).transpose(0, 2, 1), axis=2, bitorder="little").tobytes()
            back = subprocess.run([STIM, "convert", "--in_format", "ptb64", "--out_format", "01", "--num_measurements", str(nm),
                                   "--num_detectors", str(nd), "--num_observables", str(no)], input=raw, capture_output=True)
            assert back.returncode == 0 and back.stdout == text01, back.stderr
            out[fmt] = base64.b64encode(raw).decode()
            continue
        cmd = [STIM, "convert", "--in_format", "01", "--out_format", fmt, "--num_measurements", str(nm), "--num_detectors", str(nd),
               "--num_observables", str(no)]
        r = subprocess.run(cmd, input=text01, capture_output=True)
        assert r.returncode == 0, r.stderr
        out[fmt] = base64.b64encode(r.stdout).decode()
    cases.append({"num_measurements": nm, "num_detectors": nd, "num_observables": no, "shots": shots,
                  "bits": base64.b64encode(np.packbits(bits, axis=1, bitorder="little").tobytes()).decode(), "files": out})
json.dump(cases, open(os.path.join(ROOT, "tests", "golden", "formats_cases.json"), "w"))
print(len(cases), "cases", os.path.getsize(os.path.join(ROOT, "tests", "golden", "formats_cases.json")), "bytes")
