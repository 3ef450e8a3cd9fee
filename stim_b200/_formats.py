"""Readers for the reference's six shot-data formats (01, b8, r8, hits, dets, ptb64) into packed shot-major rows:
the input side of `m2d` / `convert_file` and of `read_shot_data_file`
(/root/reference/src/stim/io/measure_record_reader.inl: Format01 :118-230, B8 :232-290, Hits :300-395, R8 :440-480,
Dets :528-595, PTB64 :600-760; doc/result_formats.md). Host-side numpy; the writers are the C side (`writers.cc`)."""
import numpy as np

FORMATS = ("01", "b8", "r8", "hits", "dets", "ptb64")


def _pack(bits: np.ndarray, n_bits: int) -> np.ndarray:
    if n_bits == 0:
        return np.zeros((bits.shape[0], 0), dtype=np.uint8)
    return np.packbits(bits, axis=1, bitorder="little")


def _rows_from_hits(shot_of_hit: np.ndarray, bit_of_hit: np.ndarray, shots: int, n_bits: int, toggle: bool = True) -> np.ndarray:
    """Sparse (shot, bit) pairs -> packed rows. toggle: a bit listed twice toggles back (the reference's per-record hits reader
    XORs, measure_record_reader.inl:315) - its dets reader SETS the bit instead (:484). Sort-based: no per-hit Python work and
    no dense one-byte-per-bit intermediate."""
    nb = (n_bits + 7) // 8
    out = np.zeros((shots, nb), dtype=np.uint8)
    if len(shot_of_hit) == 0 or nb == 0:
        return out
    flat = shot_of_hit.astype(np.int64) * n_bits + bit_of_hit.astype(np.int64)
    u, c = np.unique(flat, return_counts=True)
    if toggle:
        u = u[(c & 1) == 1]
    if len(u) == 0:
        return out
    shot, bit = np.divmod(u, n_bits)
    byte_index = shot * nb + (bit >> 3)               # non-decreasing, since u is sorted
    val = (1 << (bit & 7)).astype(np.uint8)
    first = np.flatnonzero(np.diff(byte_index, prepend=-1))
    out.reshape(-1)[byte_index[first]] = np.bitwise_or.reduceat(val, first)
    return out


def _parse_uints(b: np.ndarray):
    """All maximal digit runs of a byte array at once -> (start index, value) per run, or None if a run is longer than 9 digits."""
    is_d = (b >= 48) & (b <= 57)
    if not is_d.any():
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    edge = np.diff(is_d.astype(np.int8), prepend=0, append=0)
    starts = np.flatnonzero(edge == 1)
    ends = np.flatnonzero(edge == -1) - 1          # last digit of each run
    lens = ends - starts + 1
    if lens.max() > 9:
        return None
    vals = np.zeros(len(starts), dtype=np.int64)
    for k in range(int(lens.max())):
        m = lens > k
        vals[m] += (b[ends[m] - k].astype(np.int64) - 48) * 10 ** k
    return starts, vals


def _fast_hits(data: bytes, n_bits: int):
    """Vectorised reader for well-formed hits data (digits, commas, newlines); None -> let the strict line parser decide."""
    b = np.frombuffer(data, dtype=np.uint8)
    if len(b) == 0:
        return np.zeros((0, (n_bits + 7) // 8), dtype=np.uint8)
    is_d = (b >= 48) & (b <= 57)
    comma, nl = b == 44, b == 10
    if not np.all(is_d | comma | nl) or b[-1] != 10:
        return None
    ci = np.flatnonzero(comma)
    if len(ci) and (ci[0] == 0 or not (np.all(is_d[ci - 1]) and np.all(is_d[ci + 1]))):
        return None
    parsed = _parse_uints(b)
    if parsed is None:
        return None
    starts, vals = parsed
    if len(vals) and vals.max() >= n_bits:
        raise ValueError("hit index is too large.")
    line = np.cumsum(nl)[starts] if len(starts) else np.zeros(0, np.int64)
    return _rows_from_hits(line.astype(np.int64), vals, int(nl.sum()), n_bits)


def _fast_dets(data: bytes, n_bits: int, nm: int, nd: int, no: int):
    """Vectorised reader for well-formed dets data ("shot" + " M3 D0 L1" tokens, one record per line); None -> strict parser."""
    b = np.frombuffer(data, dtype=np.uint8)
    if len(b) == 0:
        return np.zeros((0, (n_bits + 7) // 8), dtype=np.uint8)
    if len(b) < 5 or b[-1] != 10:
        return None
    is_d = (b >= 48) & (b <= 57)
    nl, sp = b == 10, b == 32
    pre = (b == 77) | (b == 68) | (b == 76)
    s4 = np.zeros(len(b), dtype=bool)
    s4[:-3] = (b[:-3] == 115) & (b[1:-2] == 104) & (b[2:-1] == 111) & (b[3:] == 116)
    word = s4.copy()
    for k in (1, 2, 3):
        word[k:] |= s4[:-k]
    if not np.all(is_d | nl | sp | pre | word):
        return None
    shot_pos = np.flatnonzero(s4)
    line_starts = np.concatenate(([0], np.flatnonzero(nl)[:-1] + 1))
    if len(shot_pos) != len(line_starts) or not np.array_equal(shot_pos, line_starts):
        return None  # blank lines, indentation, several records per line: the strict parser handles those
    pi = np.flatnonzero(pre)
    if len(pi) and not (np.all(sp[pi - 1]) and np.all(is_d[np.minimum(pi + 1, len(b) - 1)])):
        return None
    si = np.flatnonzero(sp)
    if len(si) and not np.all(pre[np.minimum(si + 1, len(b) - 1)]):
        return None
    parsed = _parse_uints(b)
    if parsed is None:
        return None
    starts, vals = parsed
    if len(starts) != len(pi) or (len(pi) and not np.array_equal(starts, pi + 1)):
        return None
    letter = b[pi]
    off = np.where(letter == 77, 0, np.where(letter == 68, nm, nm + nd)).astype(np.int64)
    length = np.where(letter == 77, nm, np.where(letter == 68, nd, no)).astype(np.int64)
    if np.any(vals >= length):
        return None  # (the strict parser words the error)
    rec = (np.cumsum(nl)[pi]).astype(np.int64) if len(pi) else np.zeros(0, np.int64)
    return _rows_from_hits(rec, off + vals, len(shot_pos), n_bits, toggle=False)


def read_shots(data: bytes, fmt: str, n_bits: int, *, num_measurements=None, num_detectors: int = 0,
               num_observables: int = 0) -> np.ndarray:
    """Decodes `data` into uint8[shots, ceil(n_bits / 8)] (little-endian bit order). For `dets`, the M / D / L prefixes index
    the three consecutive ranges (num_measurements, num_detectors, num_observables); by default all n_bits are M."""
    if num_measurements is None:
        num_measurements = n_bits - num_detectors - num_observables
    nb = (n_bits + 7) // 8
    if fmt == "b8":
        if nb == 0:
            if len(data):
                raise ValueError("b8 data with zero bits per shot does not say how many shots there are.")
            return np.zeros((0, 0), dtype=np.uint8)
        if len(data) % nb:
            raise ValueError("b8 data ended in middle of record.")
        return np.frombuffer(data, dtype=np.uint8).reshape(-1, nb).copy()
    if fmt == "01":
        if n_bits and len(data) % (n_bits + 1) == 0 and b"\r" not in data[: n_bits + 2]:
            # the common case in one pass: equally long lines of '0' / '1' closed by '\n'
            a = np.frombuffer(data, dtype=np.uint8).reshape(-1, n_bits + 1)
            body = a[:, :n_bits]
            if np.all(a[:, n_bits] == 10) and np.all((body == 48) | (body == 49)):
                return _pack(body & 1, n_bits)
        lines = data.replace(b"\r\n", b"\n").split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
        elif lines and len(lines[-1]) == n_bits and not lines[-1].strip(b"01"):
            # (/root/reference/src/stim/io/measure_record_reader.inl: every record of the 01 format is closed by a newline)
            raise ValueError(f"01 data didn't end with a newline after the expected data length of '{n_bits}'.")
        bits = np.zeros((len(lines), max(n_bits, 1)), dtype=np.uint8)
        for i, ln in enumerate(lines):
            if len(ln) != n_bits or ln.strip(b"01"):
                raise ValueError("01 data didn't have the expected number of 0/1 characters per line.")
            if n_bits:
                bits[i, :n_bits] = np.frombuffer(ln, dtype=np.uint8) - 48
        return _pack(bits[:, :n_bits], n_bits)
    if fmt == "r8":
        b = np.frombuffer(data, dtype=np.uint8).astype(np.int64)
        if len(b) == 0:
            return np.zeros((0, nb), dtype=np.uint8)
        # every record is n_bits + 1 positions long (a virtual 1 terminates it): byte v = v zeros, then a 1 unless v == 255
        ones = b != 255
        start = np.cumsum(b) + np.concatenate(([0], np.cumsum(ones)[:-1]))  # global position of the 1 a byte emits
        g = start[ones]
        rec, pos = np.divmod(g, n_bits + 1)
        # the terminator of every record must sit exactly at its position n_bits: every 1 belongs to the record numbered by the
        # terminators before it, else a run "jumped past" the end of its record
        term = pos == n_bits
        if not np.array_equal(rec, np.cumsum(term) - term):
            raise ValueError(f"r8 data jumped past expected end of encoded data. Expected to decode {n_bits} bits.")
        if not ones[-1] or pos[-1] != n_bits:
            raise ValueError(f"End of file before end of r8 data. Expected to decode {n_bits} bits.")
        shots = int(term.sum())
        keep = ~term
        return _rows_from_hits(rec[keep], pos[keep], shots, n_bits)
    if fmt == "hits":
        fast = _fast_hits(data.replace(b"\r\n", b"\n"), n_bits)
        if fast is not None:
            return fast
        lines = data.replace(b"\r\n", b"\n").split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
        elif lines:  # (the reference's reader wants every record closed by a newline and nothing but digits and commas)
            raise ValueError("HITS data wasn't comma-separated integers terminated by a newline.")
        s_idx, b_idx = [], []
        for i, ln in enumerate(lines):
            if ln == b"":
                continue
            toks = ln.split(b",")
            if not all(tok.isdigit() for tok in toks):
                raise ValueError("HITS data wasn't comma-separated integers terminated by a newline.")
            vals = [int(tok) for tok in toks]
            for v in vals:
                if v < 0 or v >= n_bits:
                    raise ValueError("hit index is too large.")
                s_idx.append(i)
                b_idx.append(v)
        return _rows_from_hits(np.array(s_idx, dtype=np.int64), np.array(b_idx, dtype=np.int64), len(lines), n_bits)
    if fmt == "dets":
        fast = _fast_dets(data.replace(b"\r\n", b"\n"), n_bits, num_measurements, num_detectors, num_observables)
        if fast is not None:
            return fast
        offsets = {ord("M"): (0, num_measurements), ord("D"): (num_measurements, num_detectors),
                   ord("L"): (num_measurements + num_detectors, num_observables)}
        s_idx, b_idx = [], []
        shots = 0
        for ln in data.replace(b"\r\n", b"\n").split(b"\n"):
            ln = ln.lstrip(b" \t")
            if ln == b"":
                continue
            if not ln.startswith(b"shot"):
                raise ValueError("DETS data didn't start with 'shot'")
            rest = ln[4:]
            if rest:
                if rest[:1] != b" " or rest.endswith(b" ") or b"  " in rest:
                    raise ValueError("DETS data wasn't single-space-separated with no trailing spaces.")
                for tok in rest[1:].split(b" "):
                    if tok[:1] not in (b"M", b"D", b"L"):
                        raise ValueError(f"Unrecognized DETS prefix. Expected M or D or L not '{tok[:1].decode(errors='replace')}'")
                    off, length = offsets[tok[0]]
                    if not tok[1:].isdigit():
                        raise ValueError("DETS data had a value prefix (M or D or L) not followed by an integer.")
                    v = int(tok[1:])
                    if v >= length:
                        raise ValueError(f"DETS data had a value larger than expected. Got {chr(tok[0])}{v} but expected length "
                                         f"of {chr(tok[0])} space to be {length}.")
                    s_idx.append(shots)
                    b_idx.append(off + v)
            shots += 1
        return _rows_from_hits(np.array(s_idx, dtype=np.int64), np.array(b_idx, dtype=np.int64), shots, n_bits, toggle=False)
    if fmt == "ptb64":
        if n_bits == 0:
            if len(data):
                raise ValueError("ptb64 data with zero bits per shot does not say how many shots there are.")
            return np.zeros((0, 0), dtype=np.uint8)
        if len(data) % (8 * n_bits):
            raise ValueError("File ended in the middle of a ptb64 record.")
        w = np.frombuffer(data, dtype=np.uint8).reshape(-1, n_bits, 8)  # [group, bit, byte of the 64-shot word]
        bits = np.unpackbits(w, axis=2, bitorder="little")              # [group, bit, shot in group]
        bits = bits.transpose(0, 2, 1).reshape(-1, n_bits)              # [shot, bit]
        return _pack(bits, n_bits)
    raise ValueError(f"Unrecognized format: {fmt!r}")
