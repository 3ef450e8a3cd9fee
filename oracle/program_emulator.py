"""TEST INFRASTRUCTURE (oracle) — never imported by the product path (stim_b200/).

Sequential numpy emulator of the LOWERED instruction stream (stim_b200/csrc/program.h). It executes
exactly what the CUDA interpreter (stim_b200/csrc/kernels.cu: gstim_interp_kernel) executes, one block at
a time, and additionally checks the stream's concurrency contract (a race detector):

  * items of one batch must touch disjoint resources,
  * between two block barriers no WARP may touch data last written by another warp (item i is executed by thread group
    i % slots, a warp executes 32 >> lanes_log2 consecutive thread groups; noise events are applied by the warp that
    executes the item they hit), and
  * the noise applications form the chain the kernel's staging pipeline follows (GH_*_NEXT, alternating parity).

It lets the host lowering be validated on a machine without a GPU: emulator(lowered program) must equal
oracle.frame_oracle (reference semantics on the circuit text) bit for bit.
"""
import numpy as np

from . import philox as px

HDR = 12
(OP_END, OP_NEXT, OP_CLIFF1, OP_CLIFF2, OP_NOISE1, OP_NOISE2, OP_MEASURE, OP_RECZERO, OP_XORROWS, OP_OBS_PAULI,
 OP_FEEDBACK, OP_CORR, OP_QMAP) = range(13)
F_BARRIER, F_REC, F_ACCUM, F_RESET, F_TABLE, F_NOFRAME, F_DET = 1, 2, 4, 8, 16, 32, 64
(GH_OP, GH_N, GH_WORDS, GH_EXTRA, GH_CSITE0, GH_REC0, GH_PRE, GH_PRE_NEXT, GH_POST, GH_POST_NEXT, GH_PERM, GH_WIDTHS) = range(12)
NO_NOISE = 0xFFFFFFFF
SLICE_WORDS = 8

PLAN_FIELDS = ["num_qubits", "q_pitch", "num_meas", "num_det", "num_obs", "rec_ring", "n_words", "chunk_words", "n_chunks",
               "slots", "lanes_log2", "n_slices", "mode", "max_items", "n_batches", "n_barriers"]


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


def read_schedule(w, plan):
    """The copy of the noise schedule behind the program: (slices [n, 8], rates [n, 2], tables)."""
    o = plan["n_words"]
    assert int(w[o]) == 0x4843534E, "noise schedule marker missing"
    n_sl, n_rates, n_tab = int(w[o + 1]), int(w[o + 2]), int(w[o + 3])
    assert n_sl == plan["n_slices"]
    o += 4
    slices = [[int(v) for v in w[o + SLICE_WORDS * i: o + SLICE_WORDS * (i + 1)]] for i in range(n_sl)]
    o += SLICE_WORDS * n_sl
    rates = [(int(w[o + 2 * i]), int(w[o + 2 * i + 1])) for i in range(n_rates)]
    o += 2 * n_rates
    tables = [int(v) for v in w[o: o + n_tab]]
    return slices, rates, tables


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
        self.x = np.zeros((self.Q + 1, W), dtype=np.uint32)
        self.z = np.zeros((self.Q + 1, W), dtype=np.uint32)
        self.flag = np.zeros(W, dtype=np.uint32)
        self.mode = plan["mode"]
        self.rec_mask = (plan["rec_ring"] - 1) if self.mode == 0 else 0xFFFFFFFF
        n_rec = plan["rec_ring"] if self.mode == 0 else max(plan["num_meas"], 1)
        self.rec = np.zeros((n_rec, W), dtype=np.uint32)
        self.out = np.zeros((plan["num_det"] + plan["num_obs"], W), dtype=np.uint32)
        self.logical_of = read_qmap(self.w, plan)
        self.slices, self.rates, self.tables = read_schedule(self.w, plan)
        self.group_slices = {}  # noise group -> slices of it seen so far (must arrive in order, without gaps)
        self.n_applications = 0
        # what the chain promises for the next application: its first slice | log2(slices per 32 items) << 28
        self.expect_next = NO_NOISE
        if self.slices:
            self.expect_next = None  # (the first application's width is not promised by anyone: the kernel stages slice 0)
        # race detector state: resource -> (warp that wrote, set of warps that read) since the last barrier
        self.writer = {}
        self.readers = {}
        self.slots = plan["slots"]
        self.warp_shift = 5 - plan["lanes_log2"]
        assert self.slots % (1 << self.warp_shift) == 0
        self.batch_w = set()
        self.batch_r = set()

    def warp_of(self, i):
        return (i % self.slots) >> self.warp_shift

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

    def apply_noise(self, att, nxt, n, item_of, rec0, w, corr=None):
        """One noise application: walks its ceil(n / 32) slices exactly like the kernel's producers and applies every event
        to the item it hits. item_of(it) = (row1, row2) of item `it` (noise target order); corr: E / ELSE handler."""
        assert att != NO_NOISE
        slice0, parity = att & 0x7FFFFFFF, att >> 31
        assert parity == self.n_applications & 1, "noise application parity out of step"
        if self.expect_next is None:
            assert slice0 == 0
        else:
            assert slice0 | ((5 - w) << 28) == self.expect_next, "noise chain broken: the previous application promised another slice"
        self.n_applications += 1
        self.expect_next = nxt
        S = 1 << w
        B = self.B
        c2, hi = self.col0 & 0xFFFFFFFF, self.col0 >> 32
        for sl_rel in range((n + S - 1) // S):
            group, j, rs, h0, t1, t2, t3, _ = self.slices[slice0 + sl_rel]
            assert self.group_slices.get(group, 0) == j, "slices of a noise group out of order"
            self.group_slices[group] = j + 1
            rate = self.rates[rs & 0xFFFF]
            sites = rs >> 16
            assert sites == min(S, n - S * sl_rel)
            assert abs(w - px.slice_width_log2(self._prob_of(rate))) <= 1  # (exact up to the rounding of INV at a boundary)
            op, flags, aux = h0 & 0xFF, (h0 >> 8) & 0xFF, h0 >> 16
            total = sites * B
            a = d = 0
            words = None
            while True:
                if d % 2 == 0:  # one Philox call = two draws
                    words = [int(v) for v in px.philox4x32_10(group, 0x80000000 | j, c2, hi | ((d // 2) << 15), self.k0, self.k1)]
                gap_word, v = words[2 * (d % 2)], words[2 * (d % 2) + 1]
                d += 1
                G = px.gap_of(gap_word, rate)
                if G >= total - a:
                    break
                a += G
                it, shot = S * sl_rel + a // B, a % B
                a += 1
                if corr is not None:
                    corr(shot)
                    continue
                q1, q2 = item_of(it)
                if op == OP_NOISE1:
                    sel = 0 if v < t1 else 2 if v < t2 else 4 if v < t3 else 6
                    cat = (aux >> sel) & 3
                    f = (cat & 1, cat >> 1, 0, 0)
                    if flags & F_REC:
                        self.flip(self.rec[(rec0 + it) & self.rec_mask], shot)
                else:
                    assert op == OP_NOISE2
                    if not flags & F_TABLE:
                        pr = 1 + ((v * 15) >> 32)
                        f = (pr & 1, (pr >> 1) & 1, (pr >> 2) & 1, (pr >> 3) & 1)
                    else:
                        table = self.tables[t1: t1 + 15]
                        pr = aux
                        for jj in range(15):
                            if v < table[jj]:
                                pr = jj + 1
                                break
                        c1, c2_ = pr >> 2, pr & 3
                        f = (((c1 + 1) >> 1) & 1, c1 >> 1, ((c2_ + 1) >> 1) & 1, c2_ >> 1)
                for on, arr in zip(f, (self.x[q1], self.z[q1], self.x[q2], self.z[q2])):
                    if on:
                        self.flip(arr, shot)

    @staticmethod
    def _prob_of(rate):
        """probability back from (INV, SH): 1 / lambda = INV * 2^(26 - SH)."""
        import math

        if rate[0] == 0:
            return 1.0  # p >= 1: INV = 0
        return -math.expm1(-1.0 / (rate[0] * 2.0 ** (26 - rate[1])))

    def noise_flags(self, att):
        """flags of the noise of application att (from its first slice)."""
        return (self.slices[att & 0x7FFFFFFF][3] >> 8) & 0xFF

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
            n, words, extra = int(w[pc + GH_N]), int(w[pc + GH_WORDS]), int(w[pc + GH_EXTRA])
            assert pc // chunk == (pc + words - 1) // chunk, "batch straddles a chunk"
            csite0, rec0 = int(w[pc + GH_CSITE0]), int(w[pc + GH_REC0])
            pre, pre_next = int(w[pc + GH_PRE]), int(w[pc + GH_PRE_NEXT])
            post, post_next = int(w[pc + GH_POST]), int(w[pc + GH_POST_NEXT])
            perm_off = int(w[pc + GH_PERM])
            w_pre, w_post = int(w[pc + GH_WIDTHS]) & 15, (int(w[pc + GH_WIDTHS]) >> 4) & 15
            pay = w[pc + HDR: pc + words]
            perm = None
            if perm_off:
                perm = w[pc + perm_off: pc + words].view(np.uint8)
                pay = w[pc + HDR: pc + perm_off]
            if flags & F_BARRIER:
                self.writer.clear()
                self.readers.clear()
            self.batch_w, self.batch_r = set(), set()
            S = self.warp_of

            def pos_of(it, perm=perm):
                return it if perm is None else (it & ~31) + int(perm[it])

            if op == OP_CLIFF1:
                a, b, c, d = [(0xFFFFFFFF if (aux >> i) & 1 else 0) for i in range(4)]
                for i in range(n):
                    q = int(pay[i])
                    self.touch(S(i), q, True)
                    x, z = self.x[q].copy(), self.z[q].copy()
                    self.x[q] = (x & np.uint32(a)) ^ (z & np.uint32(b))
                    self.z[q] = (x & np.uint32(c)) ^ (z & np.uint32(d))
                if post != NO_NOISE:
                    self.apply_noise(post, post_next, n, lambda it: (int(pay[pos_of(it)]), 0), rec0, w_post)
            elif op == OP_CLIFF2:
                m = [np.uint32(0xFFFFFFFF if (aux >> i) & 1 else 0) for i in range(16)]
                for i in range(n):
                    q1, q2 = int(pay[i]) & 0xFFFF, int(pay[i]) >> 16
                    self.touch(S(i), q1, True)
                    self.touch(S(i), q2, True)
                    v = [self.x[q1].copy(), self.z[q1].copy(), self.x[q2].copy(), self.z[q2].copy()]
                    o = [(v[0] & m[4 * k]) ^ (v[1] & m[4 * k + 1]) ^ (v[2] & m[4 * k + 2]) ^ (v[3] & m[4 * k + 3]) for k in range(4)]
                    self.x[q1], self.z[q1], self.x[q2], self.z[q2] = o
                if post != NO_NOISE:
                    self.apply_noise(post, post_next, n, lambda it: (int(pay[pos_of(it)]) & 0xFFFF, int(pay[pos_of(it)]) >> 16), rec0, w_post)
            elif op == OP_NOISE1:
                for i in range(n):
                    if not flags & F_NOFRAME:
                        self.touch(S(i), int(pay[i]), True)
                    if flags & F_REC:
                        self.touch(S(i), (R_REC, (rec0 + i) & self.rec_mask), True)
                if post != NO_NOISE:
                    assert flags & ~F_BARRIER == self.noise_flags(post)
                    self.apply_noise(post, post_next, n, lambda it: (int(pay[it]) & 0xFFFF, 0), rec0, w_post)
            elif op == OP_NOISE2:
                for i in range(n):
                    self.touch(S(i), int(pay[i]) & 0xFFFF, True)
                    self.touch(S(i), int(pay[i]) >> 16, True)
                if post != NO_NOISE:
                    self.apply_noise(post, post_next, n, lambda it: (int(pay[it]) & 0xFFFF, int(pay[it]) >> 16), rec0, w_post)
            elif op == OP_MEASURE:
                basis, kind = aux & 3, (aux >> 2) & 3
                stride = 3 if flags & F_DET else 1  # fused detectors: (qubit word, detector row or NONE, record slot)
                item = lambda it: (int(pay[stride * it]) & 0xFFFF, 0)  # noqa: E731
                if pre != NO_NOISE:
                    assert not self.noise_flags(pre) & F_REC
                    self.apply_noise(pre, pre_next, n, item, rec0, w_pre)
                for i in range(n):
                    q = int(pay[stride * i]) & 0xFFFF
                    assert int(pay[stride * i]) >> 16 == self.logical_of[q]
                    self.touch(S(i), q, True)
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
                        self.touch(S(i), (R_REC, (rec0 + i) & self.rec_mask), True)
                        self.rec[(rec0 + i) & self.rec_mask] = m
                        if flags & F_DET and int(pay[3 * i + 1]) != 0xFFFFFFFF:
                            d, other = int(pay[3 * i + 1]), int(pay[3 * i + 2])
                            self.touch(S(i), ("out", d), True)
                            self.touch(S(i), (R_REC, other), False)
                            self.out[d] = m ^ self.rec[other]
                if post != NO_NOISE:
                    assert kind != 2 or not self.noise_flags(post) & F_REC
                    self.apply_noise(post, post_next, n, item, rec0, w_post)
            elif op == OP_RECZERO:
                for i in range(n):
                    self.touch(S(i), (R_REC, (rec0 + i) & self.rec_mask), True)
                    self.rec[(rec0 + i) & self.rec_mask] = 0
            elif op == OP_XORROWS:
                dst, off, idx = pay[:n], pay[n: 2 * n + 1], pay[2 * n + 1:]
                for i in range(n):
                    acc = np.zeros(self.W, dtype=np.uint32)
                    self.touch(S(i), ("out", int(dst[i])), True)
                    for j in range(int(off[i]), int(off[i + 1])):
                        self.touch(S(i), (R_REC, int(idx[j])), False)
                        acc ^= self.rec[int(idx[j])]
                    if flags & F_ACCUM:
                        acc ^= self.out[int(dst[i])]
                    self.out[int(dst[i])] = acc
            elif op == OP_OBS_PAULI:
                for i in range(n):
                    d, wq = int(pay[2 * i]), int(pay[2 * i + 1])
                    q = wq & 0xFFFFFF
                    self.touch(S(i), ("out", d), True)
                    self.touch(S(i), q, False)
                    if wq & (1 << 30):
                        self.out[d] ^= self.x[q]
                    if wq & (1 << 31):
                        self.out[d] ^= self.z[q]
            elif op == OP_FEEDBACK:
                for i in range(n):
                    ri, wq = int(pay[2 * i]), int(pay[2 * i + 1])
                    q = wq & 0xFFFFFF
                    self.touch(S(i), (R_REC, ri), False)
                    self.touch(S(i), q, True)
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

                if post != NO_NOISE:
                    self.apply_noise(post, post_next, 1, None, rec0, w_post, corr=lambda shot: ev(shot, None))
            else:
                raise ValueError(f"bad opcode {op} at word {pc}")
            pc += words
        assert self.expect_next == NO_NOISE, "noise chain does not end with the last application"
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
