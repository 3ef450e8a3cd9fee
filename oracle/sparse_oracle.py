"""TEST INFRASTRUCTURE (oracle) — never imported by the product path (stim_b200/).

Checker of the event-driven engine (stim_b200/csrc/response.cc + sparse.cu), in two independent halves:

1. `responses_by_injection` re-derives the response table the other way round from how the library builds it: the
   library propagates sensitivities BACKWARDS through its lowered program; here every (site, outcome) is injected as a
   single deterministic event into one shot of the FORWARD frame oracle (oracle/frame_oracle.py, the restatement of
   /root/reference/src/stim/simulators/frame_simulator.inl that is pinned against the reference CLI's goldens), and the
   output bits that flip in that shot are the response. Collapse sites are injected as a single set bit of the
   collapse randomisation (all other collapse bits zero).

2. `sample` restates the sampling of sparse.cu bit for bit from the table arrays: slices walked with geometric gaps
   (RareErrorIterator over targets x shots, /root/reference/src/stim/util_bot/probability_util.cc:33-43, in the
   32-bit fixed point of oracle/philox.py exp_draw_q26), outcome chooser, XOR of the table entries into dense rows.

Parity status: PINNED through (1) + the reference-pinned frame oracle for the table, through the deterministic goldens
(p in {0, 1}: tests/golden/reference_outputs.json) and the 2^24-shot reference statistics (tests/golden/stats_big) for
the sampler as a whole."""
import numpy as np

from . import frame_oracle as fo
from . import philox as px

NONE = 0xFFFFFFFF
OVERFLOW = 0x80000000
SPARSE_TAG = 0x53500000
RK_SINGLE, RK_UNIFORM, RK_THRESH, RK_CHAIN = 0, 1, 2, 3


def entry_ids(table, e):
    """Output ids of table entry e (ascending)."""
    ids = []
    w = table["entries"][e]
    for j in range(4):
        v = int(w[j])
        if v == NONE:
            break
        if j == 3 and v & OVERFLOW:
            off = v & 0x7FFFFFFF
            cnt = int(table["overflow"][off])
            ids.extend(int(x) for x in table["overflow"][off + 1: off + 1 + cnt])
            break
        ids.append(v)
    return ids


def class_sites(table):
    """Yields (class index, class row, first site index of the class in site_group / site_index, first outcome_word)."""
    s0 = o0 = 0
    for ci, c in enumerate(table["classes"]):
        yield ci, c, s0, o0
        s0 += int(c[21])
        o0 += int(c[5])


class _InjectOracle(fo.FrameOracle):
    """Frame oracle whose only randomness is a list of injected events: noise[group] = [(site index, shot, chooser word)],
    collapse[(measure group, logical qubit)] = [shot, ...]."""

    def __init__(self, text, n_shots, noise, collapse):
        K = max(1, (n_shots + 127) // 128)
        super().__init__(text, 0, K, 1, 0)
        self._noise, self._collapse = noise, collapse

    def collapse_words(self, mgroup, q):
        w = np.zeros(self.W, dtype=np.uint32)
        for shot in self._collapse.get((mgroup, q), ()):
            w[shot >> 5] |= np.uint32(1 << (shot & 31))
        return w

    def run_sites(self, clocks, lam, group, on_event):
        for idx, shot, word in self._noise.get(group, ()):
            assert idx < len(clocks)
            on_event(idx, 0, shot, (0, word, 0, 0))


def responses_by_injection(text, table, mode="detectors", entries=None):
    """{entry index: sorted output ids} for the given entries (default: all), derived by forward injection."""
    D = L = None
    todo = []  # (entry, group, index, word)
    for ci, c, s0, o0 in class_sites(table):
        kind, n_out, n_sites, entry0 = int(c[4]), int(c[5]), int(c[21]), int(c[22])
        for s in range(n_sites):
            for o in range(n_out):
                e = entry0 + s * n_out + o
                if entries is None or e in entries:
                    g, i, w = int(table["site_group"][s0 + s]), int(table["site_index"][s0 + s]), int(table["outcome_word"][o0 + o])
                    if kind == RK_CHAIN:  # outcome o = element o of an E / ELSE chain: its own noise group, fired alone
                        g, w = g + w, 0
                    todo.append((e, g, i, w))
    noise, collapse = {}, {}
    for shot, (e, g, i, w) in enumerate(todo):
        if g & 0x80000000:
            collapse.setdefault((g & 0x7FFFFFFF, i), []).append(shot)
        else:
            noise.setdefault(g, []).append((i, shot, w))
    o = _InjectOracle(text, max(len(todo), 1), noise, collapse).run()
    if mode == "detectors":
        dets, obs = o.detectors(), o.observables()
        rows = np.concatenate([dets, obs], axis=1) if obs.shape[1] else dets
    else:
        rows = o.measurement_flips()
    return {e: [int(v) for v in np.flatnonzero(rows[shot])] for shot, (e, _, _, _) in enumerate(todo)}


def choose(kind, n_out, thr, word):
    if kind == RK_UNIFORM:
        return (word * n_out) >> 32
    if kind in (RK_THRESH, RK_CHAIN):
        return sum(1 for j in range(n_out - 1) if word >= int(thr[j]))
    return 0


def sample(table, slices, tile_shots, seed, first_shot, n_shots, n_outputs):
    """uint8 [n_shots, n_outputs] output bits (ids as in the table: detectors then observables, or measurements) of
    shots [first_shot, first_shot + n_shots); first_shot must be a multiple of tile_shots."""
    S = tile_shots
    log_s = S.bit_length() - 1
    assert 1 << log_s == S and first_shot % S == 0
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    out = np.zeros((n_shots, n_outputs), dtype=np.uint8)
    classes = table["classes"]
    for t in range((n_shots + S - 1) // S):
        gt = first_shot // S + t
        c2, c3 = gt & 0xFFFFFFFF, SPARSE_TAG | (gt >> 32)
        for sl, (ci, total, ebase, _) in enumerate(slices):
            c = classes[int(ci)]
            rate = (int(c[2]), int(c[3]))
            kind, n_out, thr = int(c[4]), int(c[5]), c[6:21]
            total, ebase = int(total), int(ebase)
            dense_thr = int(c[23]) if len(c) > 23 else 0
            if dense_thr:
                # packed Bernoulli words (sparse.cu "Dense class"): P(bit) = dense_thr / 2^32 by the binary expansion of the
                # threshold, lowest set bit first (the bit-sliced part of biased_randomize_bits, probability_util.cc:74-132)
                i0 = (dense_thr & -dense_thr).bit_length() - 1
                for j in range((total + 31) // 32):
                    acc, r = 0, None
                    for i in range(i0, 32):
                        if i == i0 or i % 4 == 0:
                            r = [int(v) for v in px.philox4x32_10(sl, 8 * j + (i >> 2), c2, c3, k0, k1)]
                        acc = (acc | r[i & 3]) if (dense_thr >> i) & 1 else (acc & r[i & 3])
                    left = total - 32 * j
                    if left < 32:
                        acc &= (1 << left) - 1
                    for b in range(32):
                        if not (acc >> b) & 1:
                            continue
                        tr = 32 * j + b
                        site, shot = tr >> log_s, tr & (S - 1)
                        o = 0
                        if n_out > 1:
                            o = choose(kind, n_out, thr, int(px.philox4x32_10(sl, 0x80000000 | tr, c2, c3, k0, k1)[0]))
                        row = t * S + shot
                        if row < n_shots:
                            for v in entry_ids(table, ebase + site * n_out + o):
                                out[row, v] ^= 1
                continue
            a = call = 0
            done = False
            while not done:
                words = [int(v) for v in px.philox4x32_10(sl, call, c2, c3, k0, k1)]
                call += 1
                for h in range(2):
                    G = px.gap_of(words[2 * h], rate)
                    if G >= total - a:
                        done = True
                        break
                    a += G
                    site, shot = a >> log_s, a & (S - 1)
                    a += 1
                    o = choose(kind, n_out, thr, words[2 * h + 1])
                    row = t * S + shot
                    if row < n_shots:
                        for v in entry_ids(table, ebase + site * n_out + o):
                            out[row, v] ^= 1
                    if a >= total:
                        done = True
                        break
    return out
