"""Per-device-function breakdown of an ncu report of the interpreter kernel.

usage: python tools/ncu_funcs.py <report.ncu-rep> <libgstim.so | interp.o> [kernel_variant e.g. 768]

Maps the SASS-level source page (ncu --page source --print-source sass) onto the __noinline__ opcode
functions using the function offsets nvdisasm prints for the same build."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, binary = sys.argv[1], sys.argv[2]
variant = sys.argv[3] if len(sys.argv) > 3 else None
csv_text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(csv_text)))
kname = rows[0][1]
if variant is None:
    variant = re.search(r"\(int\)(\d+)", kname).group(1)
tmp = tempfile.mkdtemp()
if binary.endswith(".so"):
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(binary)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if "interp" in f][0]
    sass = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
else:
    sass = subprocess.run(["nvdisasm", "-c", binary], capture_output=True, text=True).stdout
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
    m = re.match(r"\$__internal.*?(div_u64|slowpath)", line)
    if m:
        cur = m.group(1)
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/", line)
    if m:
        off2f[int(m.group(1), 16)] = cur
hdr = rows[1]
ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [(j, n) for j, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
base = int(rows[2][ia], 16)
agg = collections.OrderedDict()
for r in rows[2:]:
    if len(r) <= iinst:
        continue
    f = off2f.get(int(r[ia], 16) - base, "?")
    d = agg.setdefault(f, collections.Counter())
    d["samples"] += int(r[isamp])
    d["inst"] += int(r[iinst])
    for j, n in stall:
        d[n] += int(r[j] or 0)
ts, ti = sum(d["samples"] for d in agg.values()), sum(d["inst"] for d in agg.values())
print(f"{kname}: {ts} samples, {ti} warp instructions")
for f, d in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    top = sorted(((d[n], n) for _, n in stall), reverse=True)[:4]
    print(f"{f:14s} samples {100 * d['samples'] / ts:5.1f}%  inst {100 * d['inst'] / ti:5.1f}%   " +
          ", ".join(f"{n[6:]} {100 * v / max(d['samples'], 1):.0f}%" for v, n in top))
