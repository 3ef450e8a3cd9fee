"""TEST INFRASTRUCTURE (oracle) — never imported by the product path (stim_b200/).

Sequential numpy emulator of the LOWERED instruction stream (stim_b200/csrc/program.h). It executes
exactly what the CUDA interpreter (stim_b200/csrc/kernels.cu: gstim_interp_kernel) executes, one block at
a time, and additionally checks the stream's concurrency contract (a race detector):

  * items of one batch must touch disjoint resources, and
  * between two barriers no thread group (`slot`) may touch data last written by another slot.

It lets the host lowering be validated on a machine without a GPU: emulator(lowered program) must equal
oracle.frame_oracle (reference semantics on the circuit text) bit for bit.
"""
import numpy as np

from . import philox as px

HDR = 12
(OP_END, OP_NEXT, OP_CLIFF1, OP_CLIFF2, OP_NOISE1, OP_NOISE2, OP_MEASURE, OP_RECZERO, OP_XORROWS, OP_OBS_PAULI,
 OP_FEEDBACK, OP_CORR, OP_QMAP) = range(13)
F_BARRIER, F_REC, F_ACCUM, F_RESET, F_TABLE, F_NOFRAME, F_DET = 1, 2, 4, 8, 16, 32, 64

PLAN_FIELDS = ["num_qubits", "q_pitch", "num_meas", "num_det", "num_obs", "rec_ring", "n_words", "chunk_words", "n_chunks",
               "slots", "mode", "max_items", "n_batches", "n_barriers"]


def plan_dict(plan_words):
    return {k: int(plan_words[i]) for i, k in enumerate(PLAN_FIELDS)}


def read_qmap(w, plan):
    """physical frame row -> logical qubit index, from the GOP_QMAP batches at the head of the program."""
    Q = plan["num_qubits"]
    table = list(range(Q + 1))
    chunk = plan["chunk_words"]
    pc = 0
    while True:
        op = int(w[pc]) & 0xFF
        if op == OP_NEXT:
            pc = (pc // chunk + 1) * chunk
        elif op == OP_QMAP:
            n, words, base = int(w[pc + 1]), int(w[pc + 2]), int(w[pc + 3])
            table[base: base + n] = [int(v) for v in w[pc + HDR: pc + HDR + n]]
            pc += words
        else:
            return table


class RaceError(AssertionError):
    pass


class Emulator:
    def __init__(self, words, plan, seed, K, col0):
        self.w = np.asarray(words, dtype=np.uint32)
        self.plan = plan
        self.Q = plan["num_qubits"]
        self.K = K
        self.B = K * 128
        self.col0 = col0
        self.k0, self.k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
        W = K * 4
        self.W = W
        self.x = np.zeros((self.Q, W), dtype=np.uint32)
        self.z = np.zeros((self.Q, W), dtype=np.uint32)
        self.flag = np.zeros(W, dtype=np.uint32)
        self.mode = plan["mode"]
        self.rec_mask = (plan["rec_ring"] - 1) if self.mode == 0 else 0xFFFFFFFF
        n_rec = plan["rec_ring"] if self.mode == 0 else max(plan["num_meas"], 1)
        self.rec = np.zeros((n_rec, W), dtype=np.uint32)
        self.out = np.zeros((plan["num_det"] + plan["num_obs"], W), dtype=np.uint32)
        self.logical_of = read_qmap(self.w, plan)
        self.group_items = {}  # noise group -> sites of it seen so far (a group may span several batches)
        # race detector state: resource -> (slot that wrote, set of slots that read) since the last barrier
        self.writer = {}
        self.readers = {}
        self.slots = plan["slots"]
        self.batch_w = set()
        self.batch_r = set()

    # ---- race detector ----
    def touch(self, slot, res, write):
        if write:
            if res in self.batch_w or res in self.batch_r:
                raise RaceError(f"batch items share resource {res}")
            self.batch_w.add(res)
        else:
            if res in self.batch_w:
                raise RaceError(f"batch items share resource {res}")
            self.batch_r.add(res)
        w = self.writer.get(res)
        if w is not None and w != slot:
            raise RaceError(f"slot {slot} touches {res} written by slot {w} without a barrier")
        if write:
            rs = self.readers.get(res)
            if rs and (rs - {slot}):
                raise RaceError(f"slot {slot} writes {res} read by slots {rs} without a barrier")
            self.writer[res] = slot
        else:
            self.readers.setdefault(res, set()).add(slot)

    def end_item(self):
        pass

    # ---- helpers ----
    def flip(self, arr, shot):
        arr[shot >> 5] ^= np.uint32(1 << (shot & 31))

    def run_batch(self, group, lam, evs):
        """Noise sites of one batch (evs[i](shot, r) = event callback of item i); see frame_oracle.run_sites."""
        n, S, B = len(evs), 32, self.B  # GSTIM_NOISE_SLICE
        gfirst = self.group_items.get(group, 0)
        self.group_items[group] = gfirst + n
        if lam == 0 or n == 0:
            return
        assert gfirst % S == 0, "a noise group was cut inside an RNG slice"
        c2, hi = self.col0 & 0xFFFFFFFF, self.col0 >> 32
        for i0 in range(0, n, S):
            j = (gfirst + i0) // S
            total = min(S, n - i0) * B
            a = d = 0
            words = None
            while True:
                if d % 2 == 0:  # one Philox call = two draws
                    words = [int(v) for v in px.philox4x32_10(group, 0x80000000 | j, c2, hi | ((d // 2) << 15), self.k0, self.k1)]
                gap_word, pauli_word = words[2 * (d % 2)], words[2 * (d % 2) + 1]
                d += 1
                G = px.exp_draw_fx(gap_word) // lam
                if G >= total - a:
                    break
                a += G
                evs[i0 + a // B](a % B, (0, pauli_word, 0, 0))
                a += 1

    def collapse(self, mgroup, q):
        cols = np.uint64(self.col0) + np.arange(self.K, dtype=np.uint64)
        r = px.philox4x32_10(mgroup, self.logical_of[q], cols & np.uint64(0xFFFFFFFF), np.uint64(px.TAG_COLLAPSE) ^ (cols >> np.uint64(32)),
                             self.k0, self.k1)
        return np.stack(r, axis=1).reshape(-1)

    # ---- main loop ----
    def run(self):
        w = self.w
        chunk = self.plan["chunk_words"]
        pc = 0
        Q = self.Q
        R_CLOCK, R_FLAG, R_REC, = ("clk",), ("flag",), "rec"
        while True:
            h0 = int(w[pc])
            op, flags, aux = h0 & 0xFF, (h0 >> 8) & 0xFF, h0 >> 16
            if op == OP_END:
                break
            if op == OP_NEXT:
                pc = (pc // chunk + 1) * chunk
                continue
            if op == OP_QMAP:
                pc += int(w[pc + 2])
                continue
            n, words, extra = int(w[pc + 1]), int(w[pc + 2]), int(w[pc + 3])
            assert pc // chunk == (pc + words - 1) // chunk, "batch straddles a chunk"
            lam = int(w[pc + 4]) | (int(w[pc + 5]) << 32)
            site0, csite0, rec0 = int(w[pc + 6]), int(w[pc + 7]), int(w[pc + 8])
            t1, t2, t3 = int(w[pc + 9]), int(w[pc + 10]), int(w[pc + 11])
            pay = w[pc + HDR: pc + words]
            if flags & F_BARRIER or op in (OP_NOISE1, OP_NOISE2):  # noise batches are bracketed by block barriers
                self.writer.clear()
                self.readers.clear()
            self.batch_w, self.batch_r = set(), set()
            S = self.slots

            if op == OP_CLIFF1:
                a, b, c, d = [(0xFFFFFFFF if (aux >> i) & 1 else 0) for i in range(4)]
                for i in range(n):
                    q = int(pay[i])
                    self.touch(i % S, q, True)
                    x, z = self.x[q].copy(), self.z[q].copy()
                    self.x[q] = (x & np.uint32(a)) ^ (z & np.uint32(b))
                    self.z[q] = (x & np.uint32(c)) ^ (z & np.uint32(d))
            elif op == OP_CLIFF2:
                m = [np.uint32(0xFFFFFFFF if (aux >> i) & 1 else 0) for i in range(16)]
                for i in range(n):
                    q1, q2 = int(pay[i]) & 0xFFFF, int(pay[i]) >> 16
                    self.touch(i % S, q1, True)
                    self.touch(i % S, q2, True)
                    v = [self.x[q1].copy(), self.z[q1].copy(), self.x[q2].copy(), self.z[q2].copy()]
                    o = [(v[0] & m[4 * k]) ^ (v[1] & m[4 * k + 1]) ^ (v[2] & m[4 * k + 2]) ^ (v[3] & m[4 * k + 3]) for k in range(4)]
                    self.x[q1], self.z[q1], self.x[q2], self.z[q2] = o
            elif op == OP_NOISE1:
                evs = []
                for i in range(n):
                    q = extra - 1 if flags & F_NOFRAME else int(pay[i])
                    self.touch(i % S, ("clk",) if q == Q else q, True)
                    if flags & F_REC:
                        self.touch(i % S, (R_REC, (rec0 + i) & self.rec_mask), True)

                    def ev(shot, r, q=q, i=i):
                        v = r[1]
                        sel = 0 if v < t1 else 2 if v < t2 else 4 if v < t3 else 6
                        cat = (aux >> sel) & 3
                        if cat & 1:
                            self.flip(self.x[q], shot)
                        if cat & 2:
                            self.flip(self.z[q], shot)
                        if flags & F_REC:
                            self.flip(self.rec[(rec0 + i) & self.rec_mask], shot)

                    evs.append(ev)
                self.run_batch(site0, lam, evs)
            elif op == OP_NOISE2:
                table = [int(v) for v in pay[:15]] if flags & F_TABLE else None
                items = pay[15:] if flags & F_TABLE else pay
                evs = []
                for i in range(n):
                    q1, q2 = int(items[i]) & 0xFFFF, int(items[i]) >> 16
                    self.touch(i % S, q1, True)
                    self.touch(i % S, q2, True)

                    def ev(shot, r, q1=q1, q2=q2):
                        v = r[1]
                        if table is None:
                            pr = 1 + ((v * 15) >> 32)
                            f = (pr & 1, (pr >> 1) & 1, (pr >> 2) & 1, (pr >> 3) & 1)
                        else:
                            pr = aux
                            for j in range(15):
                                if v < table[j]:
                                    pr = j + 1
                                    break
                            c1, c2 = pr >> 2, pr & 3
                            f = (((c1 + 1) >> 1) & 1, c1 >> 1, ((c2 + 1) >> 1) & 1, c2 >> 1)
                        for on, arr in zip(f, (self.x[q1], self.z[q1], self.x[q2], self.z[q2])):
                            if on:
                                self.flip(arr, shot)

                    evs.append(ev)
                self.run_batch(site0, lam, evs)
            elif op == OP_MEASURE:
                basis, kind = aux & 3, (aux >> 2) & 3
                stride = 3 if flags & F_DET else 1  # fused detectors: (qubit word, detector row or NONE, record slot)
                for i in range(n):
                    q = int(pay[stride * i]) & 0xFFFF
                    assert int(pay[stride * i]) >> 16 == self.logical_of[q]
                    self.touch(i % S, q, True)
                    rnd = self.collapse(csite0, q)
                    x, z = self.x[q].copy(), self.z[q].copy()
                    if basis == 2:
                        m, nx, nz = x, (x if kind == 0 else np.zeros_like(x)), rnd
                    elif basis == 0:
                        m, nz, nx = z, (z if kind == 0 else np.zeros_like(z)), rnd
                    else:
                        m = x ^ z
                        nz = rnd
                        nx = (m ^ rnd) if kind == 0 else rnd
                    self.x[q], self.z[q] = nx, nz
                    if kind != 2:
                        self.touch(i % S, (R_REC, (rec0 + i) & self.rec_mask), True)
                        self.rec[(rec0 + i) & self.rec_mask] = m
                        if flags & F_DET and int(pay[3 * i + 1]) != 0xFFFFFFFF:
                            d, other = int(pay[3 * i + 1]), int(pay[3 * i + 2])
                            self.touch(i % S, ("out", d), True)
                            self.touch(i % S, (R_REC, other), False)
                            self.out[d] = m ^ self.rec[other]
            elif op == OP_RECZERO:
                for i in range(n):
                    self.touch(i % S, (R_REC, (rec0 + i) & self.rec_mask), True)
                    self.rec[(rec0 + i) & self.rec_mask] = 0
            elif op == OP_XORROWS:
                dst, off, idx = pay[:n], pay[n: 2 * n + 1], pay[2 * n + 1:]
                for i in range(n):
                    acc = np.zeros(self.W, dtype=np.uint32)
                    self.touch(i % S, ("out", int(dst[i])), True)
                    for j in range(int(off[i]), int(off[i + 1])):
                        self.touch(i % S, (R_REC, int(idx[j])), False)
                        acc ^= self.rec[int(idx[j])]
                    if flags & F_ACCUM:
                        acc ^= self.out[int(dst[i])]
                    self.out[int(dst[i])] = acc
            elif op == OP_OBS_PAULI:
                for i in range(n):
                    d, wq = int(pay[2 * i]), int(pay[2 * i + 1])
                    q = wq & 0xFFFFFF
                    self.touch(i % S, ("out", d), True)
                    self.touch(i % S, q, False)
                    if wq & (1 << 30):
                        self.out[d] ^= self.x[q]
                    if wq & (1 << 31):
                        self.out[d] ^= self.z[q]
            elif op == OP_FEEDBACK:
                for i in range(n):
                    ri, wq = int(pay[2 * i]), int(pay[2 * i + 1])
                    q = wq & 0xFFFFFF
                    self.touch(i % S, (R_REC, ri), False)
                    self.touch(i % S, q, True)
                    if wq & (1 << 30):
                        self.x[q] ^= self.rec[ri]
                    if wq & (1 << 31):
                        self.z[q] ^= self.rec[ri]
            elif op == OP_CORR:
                self.touch(0, R_FLAG, True)
                tq = [(int(v) & 0xFFFFFF, bool(int(v) & (1 << 30)), bool(int(v) & (1 << 31))) for v in pay[:n]]
                for q in sorted({t[0] for t in tq}):
                    self.touch(0, q, True)
                if extra == Q:
                    self.touch(0, R_CLOCK, True)
                if flags & F_RESET:
                    self.flag[:] = 0

                def ev(shot, r):
                    bit = np.uint32(1 << (shot & 31))
                    if not (self.flag[shot >> 5] & bit):
                        self.flag[shot >> 5] |= bit
                        for q, fx, fz in tq:
                            if fx:
                                self.flip(self.x[q], shot)
                            if fz:
                                self.flip(self.z[q], shot)

                self.run_batch(site0, lam, [ev])
            else:
                raise ValueError(f"bad opcode {op} at word {pc}")
            if op in (OP_NOISE1, OP_NOISE2):
                self.writer.clear()  # the kernel ends noise batches with a block barrier (event queue)
                self.readers.clear()
            pc += words
        return self


def _unpack(rows):
    if rows.shape[0] == 0:
        return np.zeros((rows.shape[1] * 32, 0), dtype=np.uint8)
    bits = np.unpackbits(rows.view(np.uint8).reshape(rows.shape[0], rows.shape[1], 4), axis=2, bitorder="little")
    return bits.reshape(rows.shape[0], -1).T.copy()


def emulate(words, plan_words, seed, K, n_blocks, col0=0):
    """Returns (dets+obs rows or measurement rows) as uint8 [n_blocks*K*128, n]."""
    plan = plan_dict(plan_words)
    outs = []
    for g in range(n_blocks):
        e = Emulator(words, plan, seed, K, col0 + g * K).run()
        outs.append(_unpack(e.out if plan["mode"] == 0 else e.rec[: plan["num_meas"]]))
    return np.concatenate(outs, axis=0)
