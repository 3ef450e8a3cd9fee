"""Maps an ncu SASS source page (csv) onto CUDA source lines using nvdisasm -g line info.

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-name-substring> [libgstim.so]
"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, kname = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else "stim_b200/libgstim.so"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
addr2line = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    inside = False
    cur = None
    for ln in sass.split("\n"):
        if ln.startswith("//---") and ".text." in ln:
            inside = kname in ln
            cur = None
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            # keep only the outermost (non-inlined) annotation if "inlined at" follows; simple: last seen
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and cur:
            addr2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
h = next(r for r in rows if "Address" in r and "Instructions Executed" in r)
start = rows.index(h) + 1
ia, ii, iss, it = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
ib = h.index("stall_barrier") if "stall_barrier" in h else None
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rows[start:]:
    if len(r) <= it or not r[ia]:
        continue
    a = int(r[ia], 16) if not r[ia].isdigit() else int(r[ia])
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", 0))
    agg[key][0] += int(r[ii] or 0)
    agg[key][1] += int(r[iss] or 0)
    agg[key][2] += int(r[it] or 0)
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
src = {}
print(f"total warp-inst {ti}  samples {ts}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    f, l = key
    text = ""
    p = os.path.join("stim_b200/csrc", f)
    if os.path.exists(p):
        src.setdefault(p, open(p).read().split("\n"))
        if 0 < l <= len(src[p]):
            text = src[p][l - 1].strip()[:90]
    print(f"{v[1]/ts*100:5.1f}% smp {v[0]/ti*100:5.1f}% inst  thr/inst {v[2]/max(v[0],1):4.1f}  {f}:{l}  {text}")
