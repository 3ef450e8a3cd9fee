"""Mirror of stim.FlipSimulator (/root/reference/src/stim/simulators/frame_simulator.pybind.cc:507-1561) on the
device-resident simulator of stim_b200/csrc/flipsim.cu (C ABI gstim_flipsim_*): same method names, argument meaning, array
shapes ([row, instance]; bit_packed packs the instance axis little-endian) and exception types. Pauli frames are returned
as strings over "_XYZ" (the reference returns stim.PauliString objects, which are outside the replaced path)."""
import ctypes
import os
from typing import Optional

import numpy as np

from . import _native

_X, _Z, _REC, _DET, _OBS = 0, 1, 2, 3, 4


def _pauli_code(p) -> int:
    """pybind11_object_to_pauli_ixyz (frame_simulator.pybind.cc:158-183)."""
    if isinstance(p, str) and p in ("X", "Y", "Z", "I", "_"):
        return {"I": 0, "_": 0, "X": 1, "Y": 2, "Z": 3}[p]
    if isinstance(p, (int, np.integer)) and not isinstance(p, bool) and 0 <= int(p) < 4:
        return int(p)
    raise ValueError("Need pauli in ['I', 'X', 'Y', 'Z', 0, 1, 2, 3, '_'].")


class FlipSimulator:
    def __init__(self, *, batch_size: int, disable_stabilizer_randomization: bool = False, num_qubits: int = 0, seed=None,
                 device: int = 0):
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        self._handle = ctypes.c_void_p()
        _native.check(_native.lib().gstim_flipsim_create(
            int(batch_size), int(bool(disable_stabilizer_randomization)), int(num_qubits), ctypes.c_uint64(int(seed)), int(device),
            ctypes.byref(self._handle)))

    def copy(self, *, copy_rng: bool = False, seed=None) -> "FlipSimulator":
        """Mirror of FlipSimulator.copy (frame_simulator.pybind.cc:1476-1488): same state; the copy's random stream is fresh
        (seed or OS entropy) unless copy_rng=True."""
        if copy_rng and seed is not None:
            raise ValueError("seed and copy_rng are incompatible")
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        other = FlipSimulator.__new__(FlipSimulator)
        other._handle = ctypes.c_void_p()
        _native.check(_native.lib().gstim_flipsim_copy(self._handle, int(bool(copy_rng)), ctypes.c_uint64(int(seed)),
                                                       ctypes.byref(other._handle)))
        return other

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value and _native is not None:
            _native.lib().gstim_flipsim_destroy(h)
            self._handle = ctypes.c_void_p()

    # -- sizes ---------------------------------------------------------------------------------
    def _sizes(self):
        v = [ctypes.c_uint64(0) for _ in range(6)]
        _native.check(_native.lib().gstim_flipsim_sizes(self._handle, *[ctypes.byref(x) for x in v]))
        return [int(x.value) for x in v]

    batch_size = property(lambda self: self._sizes()[0])
    num_qubits = property(lambda self: self._sizes()[1])
    num_measurements = property(lambda self: self._sizes()[2])
    num_detectors = property(lambda self: self._sizes()[3])
    num_observables = property(lambda self: self._sizes()[4])

    # -- running circuits ------------------------------------------------------------------------
    def do(self, obj) -> None:
        """Applies a circuit, instruction or repeat block (anything whose str() is Stim circuit text)."""
        data = str(obj).encode("utf-8")
        _native.check(_native.lib().gstim_flipsim_do_text(self._handle, data, len(data)))

    def clear(self) -> None:
        _native.check(_native.lib().gstim_flipsim_clear(self._handle))

    # -- tables ----------------------------------------------------------------------------------
    def _rows(self, what: int, first: int, n: int) -> np.ndarray:
        """bool_[n, batch_size] of rows [first, first + n)."""
        batch, _, _, _, _, W = self._sizes()
        words = np.zeros((n, W), dtype=np.uint32)
        if n:
            _native.check(_native.lib().gstim_flipsim_get_rows(self._handle, what, first, n, words.ctypes.data_as(ctypes.c_void_p)))
        bits = np.unpackbits(words.view(np.uint8), axis=1, bitorder="little")[:, :batch]
        return bits.astype(np.bool_)

    def _words(self, bits: np.ndarray) -> np.ndarray:
        """bool_[n, batch_size] -> uint32[n, row_words]."""
        batch, _, _, _, _, W = self._sizes()
        pad = np.zeros((bits.shape[0], W * 32), dtype=np.uint8)
        pad[:, :batch] = bits
        return np.ascontiguousarray(np.packbits(pad, axis=1, bitorder="little").view(np.uint32))

    def _get(self, what: int, count: int, row_index, instance_index, bit_packed: bool, row_name: str, count_name: str):
        batch = self.batch_size
        if row_index is not None:
            if not -count <= row_index < count:
                raise IndexError(f"not (-{count_name} <= {row_name}={row_index} < {count_name}={count})")
            row_index %= count
        if instance_index is not None:
            if not -batch <= instance_index < batch:
                raise IndexError(f"not (-batch_size <= instance_index={instance_index} < batch_size={batch})")
            instance_index %= batch
        rows = self._rows(what, 0 if row_index is None else row_index, count if row_index is None else 1)
        if row_index is not None and instance_index is not None:
            return bool(rows[0, instance_index])
        if row_index is not None:
            out = rows[0]
        elif instance_index is not None:
            out = rows[:, instance_index]
        else:
            out = rows
        if bit_packed:
            return np.packbits(out, axis=-1, bitorder="little")
        return out

    def get_measurement_flips(self, *, record_index: Optional[int] = None, instance_index: Optional[int] = None, bit_packed: bool = False):
        return self._get(_REC, self.num_measurements, record_index, instance_index, bit_packed, "record_index", "num_measurements")

    def get_detector_flips(self, *, detector_index: Optional[int] = None, instance_index: Optional[int] = None, bit_packed: bool = False):
        return self._get(_DET, self.num_detectors, detector_index, instance_index, bit_packed, "detector_index", "num_detectors")

    def get_observable_flips(self, *, observable_index: Optional[int] = None, instance_index: Optional[int] = None, bit_packed: bool = False):
        return self._get(_OBS, self.num_observables, observable_index, instance_index, bit_packed, "observable_index", "num_observables")

    def to_numpy(self, *, bit_packed: bool = False, transpose: bool = False, output_xs=False, output_zs=False,
                 output_measure_flips=False, output_detector_flips=False, output_observable_flips=False):
        """(xs, zs, measure_flips, detector_flips, observable_flips); an entry is None unless its output_* argument is True
        or a preallocated array of the right dtype and shape (frame_simulator.pybind.cc:250-330)."""
        if all(o is False for o in (output_xs, output_zs, output_measure_flips, output_detector_flips, output_observable_flips)):
            raise ValueError("No outputs requested! Specify at least one output_*= argument.")
        res = []
        for what, out, name in ((_X, output_xs, "output_xs"), (_Z, output_zs, "output_zs"), (_REC, output_measure_flips, "output_measure_flips"),
                                (_DET, output_detector_flips, "output_detector_flips"), (_OBS, output_observable_flips, "output_observable_flips")):
            if out is False:
                res.append(None)
                continue
            n = self._sizes()[[1, 1, 2, 3, 4][what]]
            a = self._rows(what, 0, n)
            if transpose:
                a = a.T
            if bit_packed:
                a = np.packbits(a, axis=1, bitorder="little")
            if out is True:
                res.append(np.ascontiguousarray(a))
            elif isinstance(out, np.ndarray) and out.shape == a.shape and out.dtype == a.dtype:
                out[...] = a
                res.append(out)
            else:
                raise ValueError(f"{name} wasn't set to False, True, or a numpy array with dtype={a.dtype} and shape={a.shape}")
        return tuple(res)

    # -- Pauli frames ----------------------------------------------------------------------------
    def peek_pauli_flips(self, *, instance_index: Optional[int] = None):
        batch, Q = self.batch_size, self.num_qubits
        xs, zs = self._rows(_X, 0, Q), self._rows(_Z, 0, Q)
        chars = np.array(list("_XZY"))
        codes = xs.astype(np.uint8) + 2 * zs.astype(np.uint8)  # [qubit, instance]

        def frame(i):
            return "+" + "".join(chars[codes[:, i]])

        if instance_index is not None:
            if not -batch <= instance_index < batch:
                raise IndexError(f"not (-batch_size <= instance_index={instance_index} < batch_size={batch})")
            return frame(instance_index % batch)
        return [frame(i) for i in range(batch)]

    def set_pauli_flip(self, pauli, *, qubit_index: int, instance_index: int) -> None:
        code = _pauli_code(pauli)
        batch = self.batch_size
        if qubit_index < 0:
            raise ValueError("qubit_index")
        if not -batch <= instance_index < batch:
            raise IndexError(f"not (-batch_size <= instance_index={instance_index} < batch_size={batch})")
        instance_index %= batch
        if qubit_index >= self.num_qubits:
            self.do(f"I {qubit_index}")  # (the reference grows the simulator the same way)
        want_x, want_z = code in (1, 2), code in (2, 3)
        for what, want in ((_X, want_x), (_Z, want_z)):
            row = self._rows(what, qubit_index, 1)
            row[0, instance_index] = want
            w = self._words(row)
            _native.check(_native.lib().gstim_flipsim_set_rows(self._handle, what, qubit_index, 1, w.ctypes.data_as(ctypes.c_void_p), 0))

    def broadcast_pauli_errors(self, *, pauli, mask: np.ndarray, p: float = 1) -> None:
        code = _pauli_code(pauli)
        mask = np.asarray(mask)
        if mask.dtype != np.bool_ or mask.ndim != 2:
            raise ValueError("Need a 2d bool_ mask of shape (num_qubits, batch_size).")
        if mask.shape[1] != self.batch_size:
            raise ValueError("mask.shape[1] != flip_sim.batch_size")
        if not 0 <= p <= 1:
            raise ValueError("Need 0 <= p <= 1")
        w = self._words(mask)
        _native.check(_native.lib().gstim_flipsim_broadcast(
            self._handle, code, w.ctypes.data_as(ctypes.c_void_p) if w.size else None, mask.shape[0], float(p)))

    def append_measurement_flips(self, measurement_flip_data: np.ndarray) -> None:
        a = np.asarray(measurement_flip_data)
        batch = self.batch_size
        if a.ndim != 2:
            raise ValueError("measurement_flip_data must be a 2d array.")
        if a.dtype == np.uint8:
            if a.shape[1] != (batch + 7) // 8:
                raise ValueError("measurement_flip_data.shape[1] != ceil(batch_size / 8)")
            a = np.unpackbits(a, axis=1, bitorder="little")[:, :batch].astype(np.bool_)
        elif a.dtype != np.bool_ or a.shape[1] != batch:
            raise ValueError("measurement_flip_data must be bool_[n, batch_size] or uint8[n, ceil(batch_size / 8)].")
        w = self._words(a)
        _native.check(_native.lib().gstim_flipsim_set_rows(
            self._handle, _REC, self.num_measurements, a.shape[0], w.ctypes.data_as(ctypes.c_void_p) if w.size else None, 0))

    def generate_bernoulli_samples(self, num_samples: int, *, p: float, bit_packed: bool = False, out: Optional[np.ndarray] = None):
        if not 0 <= p <= 1:
            raise ValueError("Need 0 <= p <= 1")
        n_words = (int(num_samples) + 31) // 32
        words = np.zeros(n_words, dtype=np.uint32)
        _native.check(_native.lib().gstim_flipsim_bernoulli(self._handle, n_words, float(p), words.ctypes.data_as(ctypes.c_void_p) if n_words else None))
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:num_samples]
        res = np.packbits(bits, bitorder="little") if bit_packed else bits.astype(np.bool_)
        if out is not None:
            if out.shape != res.shape or out.dtype != res.dtype:
                raise ValueError("out has the wrong shape or dtype")
            out[...] = res
            return out
        return res

    def __repr__(self) -> str:
        return f"stim_b200.FlipSimulator(batch_size={self.batch_size}, num_qubits={self.num_qubits})"
