"""Golden cases for `python -m stim_b200 convert`, produced by the unmodified reference CLI (`oracle/_ref/stim convert`,
/root/reference/src/stim/cmd/command_convert.cc): random records in one format, the reference's re-encoding (and its exit
status) for every way of describing the record layout: explicit counts, --bits_per_shot, --circuit + --types, --dem, with and
without --obs_out.   python tools/gen_convert_golden.py"""
import base64, json, os, subprocess, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STIM = os.path.join(ROOT, "oracle", "_ref", "stim")
rng = np.random.default_rng(77)

CIRCUIT = """X_ERROR(0.1) 0 1 2
M 0 1 2
DETECTOR rec[-1] rec[-2]
DETECTOR rec[-2] rec[-3]
M 0 1
DETECTOR rec[-1]
OBSERVABLE_INCLUDE(0) rec[-1]
OBSERVABLE_INCLUDE(1) rec[-2]
"""  # 5 measurements, 3 detectors, 2 observables
DEM = """error(0.1) D0 D1 L0
error(0.2) D2 D3
error(0.1) D4 L1 L2
"""  # 5 detectors, 3 observables


def b64(b):
    return base64.b64encode(b).decode()


def text01(bits):
    return "".join("".join(map(str, r)) + "\n" for r in bits).encode()


cases = []
with tempfile.TemporaryDirectory() as tmp:
    cpath, dpath, opath = (os.path.join(tmp, n) for n in ("c.stim", "m.dem", "obs.out"))
    open(cpath, "w").write(CIRCUIT)
    open(dpath, "w").write(DEM)

    def run(name, flags, data, obs_out=False):
        cmd = [STIM, "convert"] + flags + (["--obs_out", opath] if obs_out else [])
        if os.path.exists(opath):
            os.remove(opath)
        r = subprocess.run(cmd, input=data, capture_output=True)
        flags = [{cpath: "@CIRCUIT", dpath: "@DEM"}.get(f, f) for f in flags]
        case = {"name": name, "flags": flags, "input": b64(data), "rc": r.returncode, "stdout": b64(r.stdout), "obs_out": None}
        if obs_out and r.returncode == 0:
            case["obs_out"] = b64(open(opath, "rb").read())
        cases.append(case)

    def to(fmt, bits, nm, nd, no):
        """bits in format fmt, by the reference itself."""
        r = subprocess.run([STIM, "convert", "--in_format", "01", "--out_format", fmt, "--num_measurements", str(nm),
                            "--num_detectors", str(nd), "--num_observables", str(no)], input=text01(bits), capture_output=True)
        assert r.returncode == 0, r.stderr
        return r.stdout

    for (nm, nd, no, shots, density) in [(6, 0, 0, 10, 0.4), (0, 9, 2, 17, 0.2), (3, 4, 2, 12, 0.3), (70, 0, 0, 9, 0.05),
                                         (0, 130, 3, 8, 0.03), (2, 70, 1, 6, 0.1)]:
        bits = (rng.random((shots, nm + nd + no)) < density).astype(np.uint8)
        counts = ["--num_measurements", str(nm), "--num_detectors", str(nd), "--num_observables", str(no)]
        for fin in ("01", "b8", "r8", "hits", "dets"):
            data = to(fin, bits, nm, nd, no)
            for fout in ("01", "b8", "r8", "hits", "dets"):
                run(f"counts_{nm}_{nd}_{no}_{fin}_to_{fout}", ["--in_format", fin, "--out_format", fout] + counts, data)
            if no:
                for fout, fobs in (("dets", "hits"), ("b8", "01"), ("01", "dets"), ("r8", "b8")):
                    run(f"counts_{nm}_{nd}_{no}_{fin}_to_{fout}_obs_{fobs}",
                        ["--in_format", fin, "--out_format", fout, "--obs_out_format", fobs] + counts, data, obs_out=True)
    # ptb64 on the input side (the reference's per-record converter reads it but cannot write it): 64 shots per group, one
    # 64-bit little-endian word per bit
    for (nm, nd, no, shots) in [(5, 0, 0, 64), (0, 70, 3, 128), (2, 3, 1, 64)]:
        n = nm + nd + no
        bits = (rng.random((shots, n)) < 0.2).astype(np.uint8)
        raw = np.packbits(bits.reshape(shots // 64, 64, n).transpose(0, 2, 1), axis=2, bitorder="little").tobytes()
        counts = ["--num_measurements", str(nm), "--num_detectors", str(nd), "--num_observables", str(no)]
        for fout in ("01", "b8", "r8", "hits", "dets"):
            run(f"ptb64_{nm}_{nd}_{no}_to_{fout}", ["--in_format", "ptb64", "--out_format", fout] + counts, raw)
        if no:
            run(f"ptb64_{nm}_{nd}_{no}_obs_out", ["--in_format", "ptb64", "--out_format", "dets", "--obs_out_format", "01"] + counts,
                raw, obs_out=True)
    bits = (rng.random((11, 13)) < 0.3).astype(np.uint8)
    for fin in ("01", "b8", "r8", "hits"):
        data = to(fin, bits, 13, 0, 0)
        for fout in ("01", "b8", "r8", "hits", "dets"):
            run(f"bits_per_shot_{fin}_to_{fout}", ["--in_format", fin, "--out_format", fout, "--bits_per_shot", "13"], data)
    # malformed input: the reference exits with status 1 (its partial output before the error is not compared)
    m4 = ["--out_format", "01", "--num_measurements", "4"]
    for name, fin, data in [("short_01_line", "01", b"010\n01\n"), ("bad_01_character", "01", b"0120\n"), ("01_without_newline", "01", b"0101"),
                            ("truncated_b8", "b8", b"abc"), ("hit_too_large", "hits", b"1,9\n"), ("hit_not_a_number", "hits", b"1,x\n"),
                            ("r8_past_the_end", "r8", b"\x02\x09"), ("r8_truncated", "r8", b"\x02"), ("dets_too_large", "dets", b"shot M9\n"),
                            ("dets_wrong_type", "dets", b"shot D0\n"), ("dets_without_shot", "dets", b"shut M0\n"), ("empty_input", "01", b""), ("hits_without_newline", "hits", b"1"), ("hits_with_spaces", "hits", b" 1, 2\n"),
                            ("hits_trailing_comma", "hits", b"1,\n"), ("hits_crlf", "hits", b"1\r\n\r\n"), ("hits_listed_twice", "hits", b"1,1\n"),
                            ("dets_listed_twice", "dets", b"shot M0 M0\n"), ("dets_double_space", "dets", b"shot  M0   M3\n"),
                            ("dets_indented_and_blank", "dets", b"  shot M0\n\nshot\n"), ("dets_without_newline", "dets", b"shot M1"),
                            ("01_crlf", "01", b"0101\r\n"), ("01_blank_line", "01", b"0101\n\n"), ("r8_exact", "r8", b"\x04"),
                            ("r8_one_too_far", "r8", b"\x05")]:
        flags = ["--in_format", fin] + (["--out_format", "01", "--num_measurements", "16"] if name == "truncated_b8" else m4)
        run("malformed_" + name, flags, data)
    run("unknown_flag", ["--in_format", "01", "--out_format", "01", "--num_measurements", "4", "--bogus", "1"], b"0101\n")
    run("missing_in_format", ["--out_format", "01", "--num_measurements", "4"], b"0101\n")
    run("unknown_format_name", ["--in_format", "01", "--out_format", "xyz", "--num_measurements", "4"], b"0101\n")
    run("count_not_a_number", ["--in_format", "01", "--out_format", "01", "--num_measurements", "four"], b"0101\n")
    run("nothing_known", ["--in_format", "01", "--out_format", "hits"], text01(bits))
    for types, (nm, nd, no) in (("M", (5, 0, 0)), ("D", (0, 3, 0)), ("L", (0, 0, 2)), ("DL", (0, 3, 2)), ("MDL", (5, 3, 2)),
                                ("LD", (0, 3, 2)), ("MD", (5, 3, 0)), ("ML", (5, 0, 2))):
        bits = (rng.random((9, nm + nd + no)) < 0.35).astype(np.uint8)
        for fin, fout in (("01", "dets"), ("b8", "hits"), ("dets", "01"), ("hits", "r8"), ("r8", "b8")):
            run(f"circuit_{types}_{fin}_to_{fout}", ["--in_format", fin, "--out_format", fout, "--circuit", cpath, "--types", types],
                to(fin, bits, nm, nd, no))
        if "L" in types:
            run(f"circuit_{types}_obs_out", ["--in_format", "01", "--out_format", "dets", "--obs_out_format", "dets", "--circuit",
                                            cpath, "--types", types], text01(bits), obs_out=True)
    run("circuit_without_types", ["--in_format", "01", "--out_format", "01", "--circuit", cpath], b"00000\n")
    run("circuit_unknown_type", ["--in_format", "01", "--out_format", "01", "--circuit", cpath, "--types", "MX"], b"00000\n")
    run("circuit_duplicate_type", ["--in_format", "01", "--out_format", "01", "--circuit", cpath, "--types", "MM"], b"00000\n")
    bits = (rng.random((14, 8)) < 0.3).astype(np.uint8)
    for fin, fout in (("01", "dets"), ("b8", "hits"), ("dets", "b8"), ("hits", "01"), ("r8", "dets")):
        run(f"dem_{fin}_to_{fout}", ["--in_format", fin, "--out_format", fout, "--dem", dpath], to(fin, bits, 0, 5, 3))
    run("dem_obs_out", ["--in_format", "01", "--out_format", "dets", "--obs_out_format", "hits", "--dem", dpath], text01(bits),
        obs_out=True)

out = os.path.join(ROOT, "tests", "golden", "convert_cases.json")
json.dump({"circuit": CIRCUIT, "dem": DEM, "cases": cases}, open(out, "w"))
print(len(cases), "cases", os.path.getsize(out), "bytes", sum(c["rc"] != 0 for c in cases), "error cases")
