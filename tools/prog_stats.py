"""Static statistics of a lowered program: batches, barriers, bank-conflict degree of gate batches.
usage: python tools/prog_stats.py <circuit.stim> [slots] [mode]"""
import ctypes, sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stim_b200 import _native
text = open(sys.argv[1]).read()
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 640
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
L = _native.lib(); d = text.encode(); n = ctypes.c_size_t(0); plan = (ctypes.c_uint32 * 16)()
_native.check(L.gstim_lower_text(d, len(d), mode, slots, 0, None, ctypes.byref(n), plan))
w = np.empty(n.value, dtype=np.uint32)
_native.check(L.gstim_lower_text(d, len(d), mode, slots, 0, w.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n), plan))
names = ["END", "NEXT", "CLIFF1", "CLIFF2", "NOISE1", "NOISE2", "MEASURE", "RECZERO", "XORROWS", "OBS_PAULI", "FEEDBACK", "CORR", "QMAP"]
chunk = plan[7]
pc = 0
cnt = collections.Counter(); items = collections.Counter(); bars = collections.Counter()
wf = collections.Counter(); ideal = collections.Counter()
while True:
    h = int(w[pc]); op = h & 0xFF; flags = (h >> 8) & 0xFF
    if op == 0: break
    if op == 1:
        pc = (pc // chunk + 1) * chunk; continue
    nI = int(w[pc + 1]); words = int(w[pc + 2])
    cnt[names[op]] += 1; items[names[op]] += nI
    if flags & 1: bars[names[op]] += 1
    if op in (2, 3, 4, 5):
        pay = w[pc + 12: pc + words]
        if op == 5 and flags & 16: pay = pay[15:]
        ops_ = [pay & 0xFFFF] + ([pay >> 16] if op in (3, 5) else [])
        for o in ops_:
            for g0 in range(0, nI, 8):
                grp = o[g0:g0 + 8] & 7
                c = collections.Counter(grp.tolist())
                wf[names[op]] += max(c.values()); ideal[names[op]] += 1
    pc += words
print("plan", dict(zip(["Q","pitch","M","D","L","ring","words","chunk","nchunks","slots","mode","max_items","batches","barriers"], list(plan)[:14])))
for k in cnt:
    extra = f" bank-degree {wf[k]/ideal[k]:.2f}" if ideal[k] else ""
    print(f"{k:10s} batches {cnt[k]:6d} items {items[k]:9d} avg {items[k]/cnt[k]:8.1f} flagged-barriers {bars[k]:5d}{extra}")
