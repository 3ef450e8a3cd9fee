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


def _rows_from_hits(shot_of_hit: np.ndarray, bit_of_hit: np.ndarray, shots: int, n_bits: int) -> np.ndarray:
    """Sparse (shot, bit) pairs -> packed rows; a bit listed twice toggles back (the readers XOR, :314)."""
    bits = np.zeros((shots, max(n_bits, 1)), dtype=np.uint8)
    if len(shot_of_hit):
        np.bitwise_xor.at(bits, (shot_of_hit, bit_of_hit), 1)
    return _pack(bits[:, :n_bits], n_bits)


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
        lines = data.replace(b"\r\n", b"\n").split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
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
        # the terminator of every record must sit exactly at its position n_bits; a 1 that skips over one "jumped past"
        last_rec = int(g[-1] // (n_bits + 1)) if len(g) else -1
        need = np.arange(last_rec + (1 if len(g) and pos[-1] == n_bits else 0), dtype=np.int64) * (n_bits + 1) + n_bits
        if not np.all(np.isin(need, g)):
            raise ValueError(f"r8 data jumped past expected end of encoded data. Expected to decode {n_bits} bits.")
        if not ones[-1] or pos[-1] != n_bits:
            raise ValueError(f"End of file before end of r8 data. Expected to decode {n_bits} bits.")
        term = pos == n_bits
        shots = int(term.sum())
        keep = ~term
        return _rows_from_hits(rec[keep], pos[keep], shots, n_bits)
    if fmt == "hits":
        lines = data.replace(b"\r\n", b"\n").split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
        s_idx, b_idx = [], []
        for i, ln in enumerate(lines):
            if ln == b"":
                continue
            try:
                vals = [int(tok) for tok in ln.split(b",")]
            except ValueError:
                raise ValueError("HITS data wasn't comma-separated integers terminated by a newline.") from None
            for v in vals:
                if v < 0 or v >= n_bits:
                    raise ValueError("hit index is too large.")
                s_idx.append(i)
                b_idx.append(v)
        return _rows_from_hits(np.array(s_idx, dtype=np.int64), np.array(b_idx, dtype=np.int64), len(lines), n_bits)
    if fmt == "dets":
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
        return _rows_from_hits(np.array(s_idx, dtype=np.int64), np.array(b_idx, dtype=np.int64), shots, n_bits)
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
