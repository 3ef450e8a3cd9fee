"""Hottest SASS lines of one device function in an ncu report: python tools/ncu_hot.py <rep> <lib.so> <function> [min_pct]"""
import csv, io, os, re, subprocess, sys, tempfile
rep, binary, func = sys.argv[1], sys.argv[2], sys.argv[3]
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 1.5
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
variant = re.search(r"\(int\)(\d+)", rows[0][1]).group(1)
hdr = rows[1]
ia, isamp, iinst, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stall = [(j, n) for j, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(binary)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if "interp" in f][0]
sass = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
off2f, cur, inside = {}, None, False
for line in sass.split("\n"):
    if line.startswith("_ZN5gstim19gstim_interp_kernelILi%s" % variant):
        inside, cur = True, "main"
        continue
    if inside and line.startswith("//---"):
        break
    if not inside:
        continue
    m = re.match(r"\$_ZN5gstim19gstim_interp_kernelILi\d+EEEvNS_12InterpParamsE\$_ZN5gstim\d+(\w+?)E[PK]", line)
    if m:
        cur = m.group(1)
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/", line)
    if m:
        off2f[int(m.group(1), 16)] = cur
base = int(rows[2][ia], 16)
sel = [r for r in rows[2:] if len(r) > iinst and off2f.get(int(r[ia], 16) - base) == func]
tot = sum(int(r[isamp]) for r in sel)
print(func, "samples", tot, "lines", len(sel), "warp instr", sum(int(r[iinst]) for r in sel))
for r in sel:
    s_ = int(r[isamp])
    if s_ >= tot * minpct / 100:
        top = sorted(((int(r[j] or 0), n[6:]) for j, n in stall), reverse=True)[:2]
        print(f"{int(r[ia], 16) - base:05x} {100 * s_ / tot:5.1f}% x{r[iinst]:>8} {r[isrc].strip()[:64]:64s} {top}")
