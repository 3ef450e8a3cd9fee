"""TEST INFRASTRUCTURE (oracle) — never imported by the product path (stim_b200/).

CPU restatement of the reference's detector-error-model sampler, DemSampler<W>::resample
(/root/reference/src/stim/simulators/dem_sampler.inl:52-74): every `error(p)` instruction of the flattened model is an
independent Bernoulli(p) row over the shots (biased_randomize_bits, probability_util.cc:74-132) that is XORed into the rows
of its detector / observable targets. The reference's mt19937 stream is not part of its contract; this restatement draws
the same distribution from the Philox addressing of stim_b200/csrc/dem.cu (header comment), so the CUDA path must match it
bit for bit, while it is pinned against the reference itself on deterministic models (p in {0, 1}: tests/golden/dem_cases.json,
outputs of `stim sample_dem`) and statistically.

Parity status: PINNED (tests/golden/dem_cases.json, tests/golden/stats_big/dem_*.npz)."""
import re

import numpy as np

from . import philox as px

TAG_DEM = 0x44454D53


def parse_dem(text):
    """Flattened model: (num_detectors, num_observables, [(p, [row, ...])]) with rows = detector id or ('L', id).
    Format: /root/reference/doc/file_format_dem_detector_error_model.md."""
    lines = []
    for raw in text.split("\n"):
        ln = raw.split("#", 1)[0].strip()
        if ln:
            lines.append(ln)
    pos = 0
    state = {"off": 0, "D": 0, "L": 0}
    errors = []

    def block(top):
        nonlocal pos
        while pos < len(lines):
            ln = lines[pos]
            pos += 1
            if ln == "}":
                assert not top
                return
            m = re.match(r"^([A-Za-z_]+)(\[[^\]]*\])?(\(([^)]*)\))?\s*(.*)$", ln)
            name, args, rest = m.group(1).lower(), m.group(4), m.group(5).split()
            if name == "repeat":
                assert rest[-1] == "{"
                start = pos
                for _ in range(int(rest[0])):
                    pos = start
                    block(False)
                continue
            if name == "error":
                tg = []
                for t in rest:
                    if t == "^":
                        continue
                    if t[0] in "Dd":
                        d = int(t[1:]) + state["off"]
                        state["D"] = max(state["D"], d + 1)
                        tg.append(d)
                    else:
                        state["L"] = max(state["L"], int(t[1:]) + 1)
                        tg.append(("L", int(t[1:])))
                errors.append((float(args), tg))
            elif name == "detector":
                for t in rest:
                    state["D"] = max(state["D"], int(t[1:]) + state["off"] + 1)
            elif name == "logical_observable":
                for t in rest:
                    state["L"] = max(state["L"], int(t[1:]) + 1)
            elif name == "shift_detectors":
                state["off"] += int(rest[0])
            else:
                raise ValueError("bad dem instruction " + name)
        assert top

    block(True)
    return state["D"], state["L"], errors


def sample(text, shots, seed, K, col0=0):
    """(dets [shots, D], obs [shots, L], errs [shots, E]) uint8, for blocks of K * 128 shots starting at global column col0."""
    D, L, errors = parse_dem(text)
    B = K * 128
    n_blocks = (shots + B - 1) // B
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    dets = np.zeros((n_blocks * B, D), dtype=np.uint8)
    obs = np.zeros((n_blocks * B, L), dtype=np.uint8)
    errs = np.zeros((n_blocks * B, len(errors)), dtype=np.uint8)
    for g in range(n_blocks):
        c0 = col0 + g * K
        for e, (p, tg) in enumerate(errors):
            rate = px.rate_of(p)
            if rate is None:
                continue
            a, call = 0, 0
            done = False
            while not done:
                words = [int(v) for v in px.philox4x32_10(e, TAG_DEM, c0 & 0xFFFFFFFF, (c0 >> 32) | (call << 15), k0, k1)]
                call += 1
                for w in words:
                    G = px.gap_of(w, rate)
                    if G >= B - a:
                        done = True
                        break
                    a += G
                    shot = g * B + a
                    a += 1
                    errs[shot, e] ^= 1
                    for t in tg:
                        if isinstance(t, tuple):
                            obs[shot, t[1]] ^= 1
                        else:
                            dets[shot, t] ^= 1
    return dets[:shots], obs[:shots], errs[:shots]
