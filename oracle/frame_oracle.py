"""TEST INFRASTRUCTURE (oracle) — never imported by the product path (stim_b200/).

CPU restatement (numpy) of the reference's bulk Pauli-frame sampler, FrameSimulator<W>
(/root/reference/src/stim/simulators/frame_simulator.inl), operating directly on circuit text.
Each handler cites the reference lines it follows. The reference draws its randomness from a
std::mt19937_64 whose stream is explicitly NOT part of the API contract
(/root/reference/src/stim/py/compiled_detector_sampler.pybind.cc:184-196); this restatement draws the
same *distributions* from the counter-based Philox addressing that DESIGN.md ("RNG addressing")
specifies for the device, so that for a given (seed, columns-per-block) the CUDA path must match
this oracle bit for bit — including noisy circuits — while the oracle itself is pinned against the
reference on deterministic circuits (bit-exact, tests/golden + oracle/_ref) and statistically.

Parity status: PINNED (reference golden vectors in tests/golden/, reference binary oracle/_ref/stim).
"""
import math
import re

import numpy as np

from . import philox as px

# ---------------------------------------------------------------------------------------------
# Circuit text -> flat instruction list (independent of the C++ parser in stim_b200/csrc/circuit.cc)
# File format: /root/reference/doc/file_format_stim_circuit.md
# ---------------------------------------------------------------------------------------------
T_INV = 1 << 31
T_X = 1 << 30
T_Z = 1 << 29
T_REC = 1 << 28
T_COMB = 1 << 27
T_SWEEP = 1 << 26
T_VAL = (1 << 24) - 1

ALIASES = {
    "MZ": "M", "MRZ": "MR", "RZ": "R", "ZCX": "CX", "CNOT": "CX", "ZCY": "CY", "ZCZ": "CZ", "H_XZ": "H",
    "CORRELATED_ERROR": "E", "SQRT_Z": "S", "SQRT_Z_DAG": "S_DAG", "SWAPCZ": "CZSWAP",
}


_LINE_RE = re.compile(r"^([A-Za-z_0-9]+)(\[[^\]]*\])?(\(([^)]*)\))?\s*(.*)$")


def _parse_target(tok):
    if tok == "*":
        return T_COMB
    inv = 0
    if tok.startswith("!"):
        inv = T_INV
        tok = tok[1:]
    if tok.startswith("rec[-") and tok.endswith("]"):
        return T_REC | int(tok[5:-1])
    if tok.startswith("sweep[") and tok.endswith("]"):
        return T_SWEEP | int(tok[6:-1])
    c = tok[0].upper()
    if c in "XYZ":
        m = {"X": T_X, "Y": T_X | T_Z, "Z": T_Z}[c]
        return inv | m | int(tok[1:])
    return inv | int(tok)


def parse_circuit(text):
    """Returns a nested list of ops: (name, args, targets) or ("REPEAT", count, body)."""
    stack = [[]]
    counts = []
    for raw in text.split("\n"):
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if line == "}":
            body = stack.pop()
            stack[-1].append(("REPEAT", counts.pop(), body))
            continue
        mt = _LINE_RE.match(line)
        if mt is None:
            raise ValueError("Circuit parse error: " + raw)
        head, rest = mt.group(1), mt.group(5)
        args = [float(a) for a in mt.group(4).split(",") if a.strip()] if mt.group(4) is not None else []
        name = head.upper()
        name = ALIASES.get(name, name)
        if name == "REPEAT":
            rest = rest.strip()
            assert rest.endswith("{")
            counts.append(int(rest[:-1].strip()))
            stack.append([])
            continue
        toks = rest.replace("*", " * ").split()
        stack[-1].append((name, args, [_parse_target(t) for t in toks]))
    assert len(stack) == 1
    return stack[0]


def flatten(ops):
    """REPEAT unrolled in execution order (stim::Circuit::for_each_operation, circuit.h:181-193)."""
    for op in ops:
        if op[0] == "REPEAT":
            for _ in range(op[1]):
                yield from flatten(op[2])
        else:
            yield op


# gate -> frame action, restated from frame_simulator.inl
NOOPS = {"TICK", "QUBIT_COORDS", "SHIFT_COORDS", "I", "X", "Y", "Z", "II", "I_ERROR", "II_ERROR"}  # :1097-1108
C1_SWAP = {"H", "H_NXZ", "SQRT_Y", "SQRT_Y_DAG"}            # do_H_XZ :345-350
C1_Z_XOR_X = {"S", "S_DAG", "H_XY", "H_NXY"}                # do_H_XY :353-358
C1_X_XOR_Z = {"SQRT_X", "SQRT_X_DAG", "H_YZ", "H_NYZ"}      # do_H_YZ :361-366
C1_XYZ = {"C_XYZ", "C_NXYZ", "C_XNYZ", "C_XYNZ"}            # do_C_XYZ :369-375
C1_ZYX = {"C_ZYX", "C_NZYX", "C_ZNYX", "C_ZYNX"}            # do_C_ZYX :378-384
MEASURES = {  # name -> (basis, kind)
    "M": ("Z", "M"), "MX": ("X", "M"), "MY": ("Y", "M"),
    "MR": ("Z", "MR"), "MRX": ("X", "MR"), "MRY": ("Y", "MR"),
    "R": ("Z", "R"), "RX": ("X", "R"), "RY": ("Y", "R"),
}


def _two_qubit(name, x1, z1, x2, z2):
    """Returns new (x1, z1, x2, z2). Same XOR sequences as frame_simulator.inl:387-630."""
    if name == "CX":
        return x1, z1 ^ z2, x2 ^ x1, z2
    if name == "CY":
        nz1 = z1 ^ x2 ^ z2
        return x1, nz1, x2 ^ x1, z2 ^ x1
    if name == "CZ":
        return x1, z1 ^ x2, x2, z2 ^ x1
    if name == "XCZ":  # CX with roles swapped (:592-598)
        a = _two_qubit("CX", x2, z2, x1, z1)
        return a[2], a[3], a[0], a[1]
    if name == "YCZ":  # CY with roles swapped (:624-630)
        a = _two_qubit("CY", x2, z2, x1, z1)
        return a[2], a[3], a[0], a[1]
    if name == "SWAP":
        return x2, z2, x1, z1
    if name in ("ISWAP", "ISWAP_DAG"):
        dx = x1 ^ x2
        return x2, z2 ^ dx, x1, z1 ^ dx
    if name == "CXSWAP":
        z2 = z2 ^ z1
        z1 = z1 ^ z2
        x1 = x1 ^ x2
        x2 = x2 ^ x1
        return x1, z1, x2, z2
    if name == "SWAPCX":
        z1 = z1 ^ z2
        z2 = z2 ^ z1
        x2 = x2 ^ x1
        x1 = x1 ^ x2
        return x1, z1, x2, z2
    if name == "CZSWAP":
        x1, x2 = x2, x1
        z1, z2 = z2, z1
        return x1, z1 ^ x2, x2, z2 ^ x1
    if name in ("SQRT_XX", "SQRT_XX_DAG"):
        dz = z1 ^ z2
        return x1 ^ dz, z1, x2 ^ dz, z2
    if name in ("SQRT_YY", "SQRT_YY_DAG"):
        d = x1 ^ z1 ^ x2 ^ z2
        return x1 ^ d, z1 ^ d, x2 ^ d, z2 ^ d
    if name in ("SQRT_ZZ", "SQRT_ZZ_DAG"):
        dx = x1 ^ x2
        return x1, z1 ^ dx, x2, z2 ^ dx
    if name == "XCX":
        return x1 ^ z2, z1, x2 ^ z1, z2
    if name == "XCY":
        return x1 ^ x2 ^ z2, z1, x2 ^ z1, z2 ^ z1
    if name == "YCX":
        nx2 = x2 ^ x1 ^ z1
        return x1 ^ z2, z1 ^ z2, nx2, z2
    if name == "YCY":
        y1 = x1 ^ z1
        y2 = x2 ^ z2
        return x1 ^ y2, z1 ^ y2, x2 ^ y1, z2 ^ y1
    raise ValueError("unknown two qubit gate " + name)


TWO_QUBIT = {
    "CX", "CY", "CZ", "XCZ", "YCZ", "SWAP", "ISWAP", "ISWAP_DAG", "CXSWAP", "SWAPCX", "CZSWAP", "SQRT_XX",
    "SQRT_XX_DAG", "SQRT_YY", "SQRT_YY_DAG", "SQRT_ZZ", "SQRT_ZZ_DAG", "XCX", "XCY", "YCX", "YCY",
}


rate_of = px.lam_fx  # probability -> per-shot event rate in fixed-point clock units (2**-56 nat)


def thr(frac):
    v = math.floor(frac * 4294967296.0)
    if not v > 0:
        return 0
    return min(int(v), 0xFFFFFFFF)


def used_qubits(ops, used):
    for op in ops:
        if op[0] == "REPEAT":
            used_qubits(op[2], used)
        elif op[0] not in ("QUBIT_COORDS", "MPAD", "TICK", "SHIFT_COORDS"):
            for t in op[2]:
                if t != T_COMB and not (t & (T_REC | T_SWEEP)):
                    used.add(t & T_VAL)


class FrameOracle:
    """Simulates n_blocks blocks of K*128 shots each, starting at global column col0."""

    def __init__(self, text, seed, K, n_blocks, col0=0):
        self.ops = parse_circuit(text)
        used = set()
        used_qubits(self.ops, used)
        self.qmap = {q: i for i, q in enumerate(sorted(used))}
        self.Q = len(self.qmap)
        self.K = K
        self.B = K * 128
        self.nb = n_blocks
        self.k0 = seed & 0xFFFFFFFF
        self.k1 = (seed >> 32) & 0xFFFFFFFF
        self.col0 = np.asarray([col0 + g * K for g in range(n_blocks)], dtype=np.uint64)  # per block
        W = n_blocks * K * 4
        self.W = W
        self.x = np.zeros((self.Q, W), dtype=np.uint32)
        self.z = np.zeros((self.Q, W), dtype=np.uint32)
        self.flag = np.zeros(W, dtype=np.uint32)  # last_correlated_error_occurred (frame_simulator.h:57)
        self.rec = []   # list of uint32[W] rows (flips)
        self.dets = []  # list of uint32[W]
        self.obs = {}
        self.ngroup = 0  # noise group counter   (Philox counter word 0 of event draws)
        self.mgroup = 0  # measure group counter (Philox counter word 0 of collapse draws)

    NOISE_SLICE = 32  # GSTIM_NOISE_SLICE (program.h)
    sweep_rows = None  # see sweep_row
    randomize = True   # False: m2d's frame simulator (see measure)

    def sweep_row(self, k):
        """Row of sweep bit k over the shots, or None when there is no sweep table (sampling)."""
        if self.sweep_rows is None:
            return None
        return self.sweep_rows.get(k, np.zeros(self.W, dtype=np.uint32))

    # -- randomness ---------------------------------------------------------------------------
    def collapse_words(self, mgroup, q):
        """128 fresh random bits per column for the collapse of qubit q in measure group mgroup -> uint32[W]."""
        cols = (self.col0[:, None] + np.arange(self.K, dtype=np.uint64)[None, :]).reshape(-1)
        r = px.philox4x32_10(mgroup, q, cols & np.uint64(0xFFFFFFFF), np.uint64(px.TAG_COLLAPSE) ^ (cols >> np.uint64(32)),
                             self.k0, self.k1)
        return np.stack(r, axis=1).reshape(-1)

    def run_sites(self, clocks, lam, group, on_event):
        """Samples the len(clocks) sites of noise group `group` (in target order) over all blocks.

        The sites are cut into slices of NOISE_SLICE; a slice x a shot block is one Bernoulli sequence (site-major, then
        shot) walked with geometric gaps floor(Exp(1)/lambda) == RareErrorIterator (probability_util.cc:33-43), in
        exact integer arithmetic. Call c of a slice: Philox counter (group, 0x80000000 | slice, col0 lo, col0 hi | c << 15)
        = draws 2c (words 0, 1) and 2c + 1 (words 2, 3); a draw = (gap to the next event, that event's Pauli word). on_event(i, g, shot, r): r[1] = Pauli word."""
        n = len(clocks)
        if n == 0 or lam == 0:
            return
        B, S = self.B, self.NOISE_SLICE
        n_sl = (n + S - 1) // S
        c2 = (self.col0 & np.uint64(0xFFFFFFFF))[None, :]
        c3 = (self.col0 >> np.uint64(32))[None, :]
        assert int(self.col0.max()) < (1 << 47)
        js = (np.arange(n_sl, dtype=np.uint64) | np.uint64(0x80000000))[:, None]
        first = px.philox4x32_10(group, js, c2, c3, self.k0, self.k1)  # call 0 of every (slice, block)
        for j in range(n_sl):
            total = min(S, n - j * S) * B
            for g in range(self.nb):
                words = [int(first[k][j, g]) for k in range(4)]  # one Philox call = draws (w0, w1) and (w2, w3)
                a, d = 0, 0
                while True:
                    if d and d % 2 == 0:
                        r = px.philox4x32_10(group, 0x80000000 | j, int(c2[0, g]), int(c3[0, g]) | ((d // 2) << 15), self.k0, self.k1)
                        words = [int(v) for v in r]
                    gap_word, pauli_word = words[2 * (d % 2)], words[2 * (d % 2) + 1]
                    d += 1
                    G = px.exp_draw_fx(gap_word) // lam
                    if G >= total - a:
                        break
                    a += G
                    on_event(j * S + a // B, g, a % B, (0, pauli_word, 0, 0))
                    a += 1

    def _flip(self, arr, g, shot):
        w = g * self.K * 4 + (shot >> 5)
        arr[w] ^= np.uint32(1 << (shot & 31))

    # -- group bookkeeping (DESIGN.md "RNG addressing") -----------------------------------------
    @staticmethod
    def _runs(keys_per_target):
        """Splits targets into maximal runs in which no qubit repeats. keys_per_target: list of tuples of qubits."""
        runs, start, seen = [], 0, set()
        for i, ks in enumerate(keys_per_target):
            if any(k in seen for k in ks):
                runs.append((start, i))
                start, seen = i, set()
            seen.update(ks)
        if start < len(keys_per_target):
            runs.append((start, len(keys_per_target)))
        return runs

    # -- primitive ops --------------------------------------------------------------------------
    def rec_at(self, lookback, what):
        if lookback == 0 or lookback > len(self.rec):
            raise IndexError("Referred to a measurement record before the beginning of time in %s." % what)
        return self.rec[len(self.rec) - lookback]

    def measure(self, basis, kind, q, mgroup):
        """M/MX/MY :173-208, R/RX/RY :211-219,255-274, MR/MRX/MRY :277-317 (one target)."""
        x, z = self.x[q], self.z[q]
        if not self.randomize:
            # guarantee_anticommutation_via_frame_randomization = false (m2d): the conjugate component is KEPT where the
            # sampler would replace it by fresh random bits (the `if (guarantee...)` branches of :173-317)
            if basis == "Z":
                m = x.copy()
                if kind != "M":
                    x[:] = 0
            elif basis == "X":
                m = z.copy()
                if kind != "M":
                    z[:] = 0
            else:
                m = x ^ z
                if kind != "M":
                    x[:] = z
            if kind != "R":
                self.rec.append(m)
            return
        rnd = self.collapse_words(mgroup, q)
        if basis == "Z":
            m = x.copy()
            if kind != "M":
                x[:] = 0
            z[:] = rnd
        elif basis == "X":
            m = z.copy()
            if kind != "M":
                z[:] = 0
            x[:] = rnd
        else:
            m = x ^ z
            z[:] = rnd
            x[:] = (m ^ rnd) if kind == "M" else rnd
        if kind != "R":
            self.rec.append(m)

    def measure_list(self, basis, kind, qs, args):
        """One measurement-type instruction on qubits qs (may repeat): collapse groups + optional result noise."""
        rec_first = len(self.rec)
        for a, b in self._runs([(q,) for q in qs]):
            g = self.mgroup
            self.mgroup += 1
            for q in qs[a:b]:
                self.measure(basis, kind, q, g)
        if kind != "R" and args:
            self.rec_noise(args[0], qs, rec_first)

    def rec_noise(self, p, clock_qubits, rec_first):
        """Result flips of M(p) etc. (measure_record_batch.inl:49-62). One noise group per run of distinct qubits."""
        lam = rate_of(p)
        for a, b in self._runs([(q,) for q in clock_qubits]):
            g = self.ngroup
            self.ngroup += 1

            def ev(i, blk, shot, r, a=a):
                self._flip(self.rec[rec_first + a + i], blk, shot)

            self.run_sites(clock_qubits[a:b], lam, g, ev)

    def cliff1(self, name, q):
        x, z = self.x[q], self.z[q]
        if name in C1_SWAP:
            t = x.copy()
            x[:] = z
            z[:] = t
        elif name in C1_Z_XOR_X:
            z ^= x
        elif name in C1_X_XOR_Z:
            x ^= z
        elif name in C1_XYZ:
            x ^= z
            z ^= x
        elif name in C1_ZYX:
            z ^= x
            x ^= z
        else:
            raise ValueError(name)

    def cliff2(self, name, a, b):
        nx1, nz1, nx2, nz2 = _two_qubit(name, self.x[a].copy(), self.z[a].copy(), self.x[b].copy(), self.z[b].copy())
        self.x[a], self.z[a], self.x[b], self.z[b] = nx1, nz1, nx2, nz2

    def controlled(self, name, ta, tb):
        """Pair of a controlled gate, either side possibly a classical bit (:140-150, :387-474)."""
        a_bit = bool(ta & (T_REC | T_SWEEP))
        b_bit = bool(tb & (T_REC | T_SWEEP))
        if not a_bit and not b_bit:
            self.cliff2(name, self.qmap[ta & T_VAL], self.qmap[tb & T_VAL])
            return
        if name in ("CX", "CY"):
            if b_bit:
                raise ValueError("Controlled %s had a bit as its target, instead of its control." % name[1])
            bit, qt, comps = ta, tb, ("x" if name == "CX" else "xz")
        elif name in ("XCZ", "YCZ"):
            if a_bit:
                raise ValueError("Controlled %s had a bit as its target, instead of its control." % name[0])
            bit, qt, comps = tb, ta, ("x" if name == "XCZ" else "xz")
        else:
            if a_bit and b_bit:
                return
            bit, qt, comps = (ta, tb, "z") if a_bit else (tb, ta, "z")
        if bit & T_SWEEP:
            # no sweep data when sampling (frame_simulator.inl:146-148: the sweep table is empty); the measurement
            # converter (oracle/m2d_oracle.py) supplies one: sweep_rows[k] = uint32[W] row of sweep bit k
            r = self.sweep_row(bit & T_VAL)
            if r is None:
                return
        else:
            r = self.rec_at(bit & T_VAL, name)
        q = self.qmap[qt & T_VAL]
        if "x" in comps:
            self.x[q] ^= r
        if "z" in comps:
            self.z[q] ^= r

    # -- products (MPP / SPP) -------------------------------------------------------------------
    def products(self, name, targets, allow_bits):
        """accumulate_next_obs_terms_to_pauli_string_helper, gate_decomposition.cc:42-86."""
        out = []
        k = 0
        while k < len(targets):
            end = k + 1
            while end < len(targets) and targets[end] == T_COMB:
                end += 2
            acc = {}
            bits = []
            imag = False
            for j in range(k, end, 2):
                t = targets[j]
                if t & (T_REC | T_SWEEP):
                    if not allow_bits:
                        raise ValueError("Found an unsupported target in " + name)
                    bits.append(t)
                    continue
                q = t & T_VAL
                xz = (1 if t & T_X else 0) | (2 if t & T_Z else 0)
                if q in acc:
                    if acc[q] and xz and acc[q] != xz:
                        imag = not imag
                    acc[q] ^= xz
                else:
                    acc[q] = xz
            if imag:
                raise ValueError("Acted on an anti-Hermitian operator (e.g. X0*Z0 instead of Y0) in " + name)
            terms = [(self.qmap[q], acc[q]) for q in sorted(acc) if acc[q]]
            out.append((terms, bits))
            k = end
        return out

    def do_mpad(self, args, n):
        rec_first = len(self.rec)
        for _ in range(n):
            self.rec.append(np.zeros(self.W, dtype=np.uint32))  # :905-912
        if args:
            lam = rate_of(args[0])
            for i in range(n):  # all share the global clock -> one noise group each, strictly sequential
                g = self.ngroup
                self.ngroup += 1
                self.run_sites([self.Q], lam, g,
                               lambda _i, blk, shot, r, i=i: self._flip(self.rec[rec_first + i], blk, shot))

    def do_mpp(self, args, targets):
        """decompose_mpp_operation, gate_decomposition.cc:88-161."""
        h_xz, h_yz, cx, ms = [], [], [], []
        merged = set()

        def flush():
            if not ms:
                return
            for q in h_xz:
                self.cliff1("H", q)
            for q in h_yz:
                self.cliff1("H_YZ", q)
            for a, b in cx:
                self.cliff2("CX", a, b)
            self.measure_list("Z", "M", list(ms), args)
            for a, b in cx:
                self.cliff2("CX", a, b)
            for q in h_yz:
                self.cliff1("H_YZ", q)
            for q in h_xz:
                self.cliff1("H", q)
            h_xz.clear(); h_yz.clear(); cx.clear(); ms.clear(); merged.clear()

        for terms, _ in self.products("MPP", targets, False):
            if not terms:
                flush()
                self.do_mpad(args, 1)
                continue
            if any(q in merged for q, _ in terms):
                flush()
            first = True
            for q, xz in terms:
                merged.add(q)
                if xz & 1:
                    (h_yz if xz & 2 else h_xz).append(q)
                if first:
                    ms.append(q)
                    first = False
                else:
                    cx.append((q, ms[-1]))
        flush()

    def do_spp(self, name, targets):
        """decompose_spp_or_spp_dag_operation, gate_decomposition.cc:163-243 (frame action of S == S_DAG)."""
        for terms, bits in self.products(name, targets, True):
            if not terms:
                continue
            focus = terms[0][0]
            h_xz = [q for q, xz in terms if xz == 1]
            h_yz = [q for q, xz in terms if xz == 3]

            def cx_layer():
                for q, _ in terms[1:]:
                    self.cliff2("CX", q, focus)
                for b in bits:
                    if b & T_SWEEP:
                        r = self.sweep_row(b & T_VAL)
                        if r is not None:
                            self.x[focus] ^= r
                        continue
                    self.x[focus] ^= self.rec_at(b & T_VAL, name)

            for q in h_xz:
                self.cliff1("H", q)
            for q in h_yz:
                self.cliff1("H_YZ", q)
            cx_layer()
            self.cliff1("S", focus)
            cx_layer()
            for q in h_yz:
                self.cliff1("H_YZ", q)
            for q in h_xz:
                self.cliff1("H", q)

    def do_mpair(self, name, args, targets):
        """do_MXX/MYY/MZZ :842-902 with decompose_pair_instruction_into_disjoint_segments (:245-274)."""
        basis = name[1]
        conj = {"X": "CX", "Y": "CY", "Z": "XCZ"}[basis]
        seg, used = [], set()

        def flush():
            if not seg:
                return
            for a, b in seg:
                self.cliff2(conj, a, b)
            self.measure_list(basis, "M", [a for a, _ in seg], args)
            for a, b in seg:
                self.cliff2(conj, a, b)
            seg.clear()
            used.clear()

        for i in range(0, len(targets), 2):
            a, b = self.qmap[targets[i] & T_VAL], self.qmap[targets[i + 1] & T_VAL]
            if a in used or b in used:
                flush()
            used.add(a)
            used.add(b)
            seg.append((a, b))
        flush()

    # -- noise channels ---------------------------------------------------------------------------
    def noise1(self, lam, cats, t1, t2, t3, qs, rec_first=None):
        """One Pauli-choice site per target (:633-643, :664-695, :779-839). qs may repeat."""
        for a, b in self._runs([(q,) for q in qs]):
            g = self.ngroup
            self.ngroup += 1

            def ev(i, blk, shot, r, a=a):
                v = r[1]
                cat = cats[0] if v < t1 else cats[1] if v < t2 else cats[2] if v < t3 else cats[3]
                q = qs[a + i]
                if cat & 1:
                    self._flip(self.x[q], blk, shot)
                if cat & 2:
                    self._flip(self.z[q], blk, shot)
                if rec_first is not None:
                    self._flip(self.rec[rec_first + a + i], blk, shot)

            self.run_sites(qs[a:b], lam, g, ev)

    def noise2(self, lam, pairs, table=None, last=0):
        """DEPOLARIZE2 :646-661 / PAULI_CHANNEL_2 (tableau_simulator.h:291-324 folded to one draw)."""
        for a0, b0 in self._runs(pairs):
            g = self.ngroup
            self.ngroup += 1

            def ev(i, blk, shot, r, a0=a0):
                a, b = pairs[a0 + i]
                v = r[1]
                if table is None:
                    pr = 1 + ((v * 15) >> 32)
                    f = (pr & 1, (pr >> 1) & 1, (pr >> 2) & 1, (pr >> 3) & 1)
                else:
                    pr = last
                    for j in range(15):
                        if v < table[j]:
                            pr = j + 1
                            break
                    c1, c2 = pr >> 2, pr & 3
                    f = (((c1 + 1) >> 1) & 1, c1 >> 1, ((c2 + 1) >> 1) & 1, c2 >> 1)
                if f[0]:
                    self._flip(self.x[a], blk, shot)
                if f[1]:
                    self._flip(self.z[a], blk, shot)
                if f[2]:
                    self._flip(self.x[b], blk, shot)
                if f[3]:
                    self._flip(self.z[b], blk, shot)

            self.run_sites([p[0] for p in pairs[a0:b0]], lam, g, ev)

    def corr(self, p, targets, reset):
        """do_CORRELATED_ERROR / do_ELSE_CORRELATED_ERROR :747-776."""
        if reset:
            self.flag[:] = 0
        lam = rate_of(p)
        g = self.ngroup
        self.ngroup += 1
        tq = [(self.qmap[t & T_VAL], bool(t & T_X), bool(t & T_Z)) for t in targets]
        clock = tq[0][0] if tq else self.Q

        def ev(i, blk, shot, r):
            w = blk * self.K * 4 + (shot >> 5)
            bit = np.uint32(1 << (shot & 31))
            if not (self.flag[w] & bit):
                self.flag[w] |= bit
                for q, fx, fz in tq:
                    if fx:
                        self.x[q][w] ^= bit
                    if fz:
                        self.z[q][w] ^= bit

        self.run_sites([clock], lam, g, ev)

    # -- driver -------------------------------------------------------------------------------------
    def run(self):
        # reset_all :153-163 — x = 0, z = random for every qubit; measure group 0
        self.measure_list("Z", "R", list(range(self.Q)), [])
        if self.Q == 0:
            self.mgroup = 1
        for name, args, targets in flatten(self.ops):
            self.do_op(name, args, targets)
        return self

    def do_op(self, name, args, targets):
        qm = self.qmap
        if name in NOOPS:
            return
        if name in C1_SWAP or name in C1_Z_XOR_X or name in C1_X_XOR_Z or name in C1_XYZ or name in C1_ZYX:
            for t in targets:
                self.cliff1(name, qm[t])
        elif name in TWO_QUBIT:
            for i in range(0, len(targets), 2):
                self.controlled(name, targets[i], targets[i + 1])
        elif name in MEASURES:
            basis, kind = MEASURES[name]
            self.measure_list(basis, kind, [qm[t & T_VAL] for t in targets], args)  # '!' ignored :176
        elif name == "MPAD":
            self.do_mpad(args, len(targets))
        elif name == "MPP":
            self.do_mpp(args, targets)
        elif name in ("SPP", "SPP_DAG"):
            self.do_spp(name, targets)
        elif name in ("MXX", "MYY", "MZZ"):
            self.do_mpair(name, args, targets)
        elif name in ("X_ERROR", "Y_ERROR", "Z_ERROR"):
            c = {"X": 1, "Z": 2, "Y": 3}[name[0]]
            self.noise1(rate_of(args[0]), (c, c, c, c), 0, 0, 0, [qm[t] for t in targets])
        elif name == "DEPOLARIZE1":
            t1, t2 = thr(1.0 / 3.0), thr(2.0 / 3.0)
            self.noise1(rate_of(args[0]), (1, 2, 3, 3), t1, t2, t2, [qm[t] for t in targets])
        elif name == "DEPOLARIZE2":
            pairs = [(qm[targets[i]], qm[targets[i + 1]]) for i in range(0, len(targets), 2)]
            self.noise2(rate_of(args[0]), pairs)
        elif name == "PAULI_CHANNEL_1":
            pxx, py, pz = args
            tot = pxx + py + pz
            lam = rate_of(min(tot, 1.0))
            if tot > 0:
                last = 2 if pz > 0 else 3 if py > 0 else 1
                t1, t2 = thr(pxx / tot), thr((pxx + py) / tot)
                cats = (1, 3, last, last)
            else:
                t1 = t2 = 0
                cats = (0, 0, 0, 0)
            self.noise1(lam, cats, t1, t2, t2, [qm[t] for t in targets])
        elif name == "PAULI_CHANNEL_2":
            tot = 0.0
            for a in args:
                tot += a
            lam = rate_of(min(tot, 1.0))
            table, cum, last = [], 0.0, 1
            for i in range(15):
                cum += args[i]
                table.append(thr(cum / tot) if tot > 0 else 0)
                if args[i] > 0:
                    last = i + 1
            pairs = [(qm[targets[i]], qm[targets[i + 1]]) for i in range(0, len(targets), 2)]
            self.noise2(lam, pairs, table, last)
        elif name in ("E", "ELSE_CORRELATED_ERROR"):
            self.corr(args[0], targets, name == "E")
        elif name in ("HERALDED_ERASE", "HERALDED_PAULI_CHANNEL_1"):
            if name == "HERALDED_ERASE":
                tot = args[0]
                t1, t2, t3 = 1 << 30, 2 << 30, 3 << 30
            else:
                hi, hx, hy, hz = args
                tot = hi + hx + hy + hz
                t1 = thr(hx / tot) if tot > 0 else 0
                t2 = thr((hx + hz) / tot) if tot > 0 else 0
                t3 = thr((hx + hz + hy) / tot) if tot > 0 else 0
            rec_first = len(self.rec)
            for _ in targets:
                self.rec.append(np.zeros(self.W, dtype=np.uint32))
            self.noise1(rate_of(min(tot, 1.0)), (1, 2, 3, 0), t1, t2, t3, [qm[t] for t in targets], rec_first)
        elif name == "DETECTOR":
            r = np.zeros(self.W, dtype=np.uint32)  # :222-230
            for t in targets:
                r ^= self.rec_at(t & T_VAL, name)
            self.dets.append(r)
        elif name == "OBSERVABLE_INCLUDE":
            k = int(args[0])  # :233-252
            r = self.obs.setdefault(k, np.zeros(self.W, dtype=np.uint32))
            for t in targets:
                if t & T_REC:
                    r ^= self.rec_at(t & T_VAL, name)
                else:
                    q = qm[t & T_VAL]
                    if t & T_X:
                        r ^= self.z[q]
                    if t & T_Z:
                        r ^= self.x[q]
        else:
            raise ValueError("Gate not found: " + name)

    # -- results -------------------------------------------------------------------------------------
    @staticmethod
    def _unpack(rows, W):
        if not rows:
            return np.zeros((W * 32, 0), dtype=np.uint8)
        a = np.stack(rows, axis=0)  # [n, W] uint32
        bits = np.unpackbits(a.view(np.uint8).reshape(len(rows), W, 4), axis=2, bitorder="little")
        return bits.reshape(len(rows), W * 32).T.copy()  # [shots, n]

    def detectors(self):
        return self._unpack(self.dets, self.W)

    def observables(self):
        n = (max(self.obs) + 1) if self.obs else 0
        rows = [self.obs.get(k, np.zeros(self.W, dtype=np.uint32)) for k in range(n)]
        return self._unpack(rows, self.W)

    def measurement_flips(self):
        return self._unpack(self.rec, self.W)


def sample(text, shots, seed, K, mode="detectors", col0=0):
    """Convenience wrapper: returns uint8 [shots, n] arrays (dets, obs) or measurement flips."""
    B = K * 128
    nb = (shots + B - 1) // B
    o = FrameOracle(text, seed, K, nb, col0).run()
    if mode == "detectors":
        return o.detectors()[:shots], o.observables()[:shots]
    return o.measurement_flips()[:shots]
