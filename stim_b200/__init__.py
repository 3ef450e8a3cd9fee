"""stim_b200 — B200-native Pauli-frame sampler, drop-in for the sampling hot path of quantumlib/Stim.

Host-side mirror of the reference's Python surface for this path (same names, argument meaning and
error behaviour), on top of the C ABI in include/gstim.h:

    stim.Circuit(...).compile_detector_sampler(seed=...)  -> Circuit.compile_detector_sampler
        /root/reference/src/stim/py/compiled_detector_sampler.pybind.cc
    stim.Circuit(...).compile_sampler(...)                -> Circuit.compile_sampler
        /root/reference/src/stim/py/compiled_measurement_sampler.pybind.cc

All sampling runs in hand-written sm_100a CUDA kernels (stim_b200/csrc/kernels.cu). There is no CPU
fallback: without libgstim.so or without a CUDA device, constructing a sampler raises.
"""
import os
from typing import Optional, Tuple, Union

import ctypes
import numpy as np

from . import _native
from ._native import GstimCudaError, GstimStats

from .dem import CompiledDemSampler, DetectorErrorModel  # noqa: E402,F401
from .flipsim import FlipSimulator  # noqa: E402,F401

__all__ = ["Circuit", "CompiledDetectorSampler", "CompiledMeasurementSampler", "GstimCudaError", "measure_lop3_peak", "response_table",
           "DetectorErrorModel", "CompiledDemSampler", "CompiledMeasurementsToDetectionEventsConverter", "FlipSimulator"]


def measure_lop3_peak(device: int = 0) -> dict:
    """LOP3 microbenchmark (gstim_measure_lop3_peak): the measured integer-ALU roofline of `device`."""
    a, b, c = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
    _native.check(_native.lib().gstim_measure_lop3_peak(int(device), ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
    return {"lane_ops_per_clk_per_sm": a.value, "lane_ops_per_sec": b.value, "sm_mhz": c.value}


def _seed_to_u64(seed) -> int:
    """seed=None -> OS entropy, like make_py_seeded_rng (/root/reference/src/stim/py/base.pybind.cc:23-35)."""
    if seed is None:
        return int.from_bytes(os.urandom(8), "little")
    if not isinstance(seed, (int, np.integer)) or isinstance(seed, bool):
        raise ValueError("Expected seed to be None or a 64 bit unsigned integer.")
    seed = int(seed)
    if seed < 0 or seed >= 1 << 64:
        raise ValueError("Expected seed to be None or a 64 bit unsigned integer.")
    return seed


def _path_str(p, what) -> str:
    if hasattr(p, "__fspath__"):
        p = os.fspath(p)
    if not isinstance(p, str):
        raise ValueError(f"Don't know how to write {what}{p!r}")
    return p


class Circuit:
    """A stabilizer circuit in Stim's text format. Only what the sampling path needs is mirrored."""

    def __init__(self, stim_program_text: str = ""):
        if not isinstance(stim_program_text, str):
            raise TypeError("stim_program_text must be a str")
        self._text = stim_program_text
        st = GstimStats()
        data = self._text.encode("utf-8")
        _native.check(_native.lib().gstim_circuit_stats(data, len(data), ctypes.byref(st)))
        self._stats = st

    @staticmethod
    def from_file(file) -> "Circuit":
        if hasattr(file, "read"):
            return Circuit(file.read())
        with open(os.fspath(file), "r") as f:
            return Circuit(f.read())

    def __str__(self) -> str:
        return self._text

    @property
    def num_qubits(self) -> int:
        return int(self._stats.num_qubits)

    @property
    def num_measurements(self) -> int:
        return int(self._stats.num_measurements)

    @property
    def num_detectors(self) -> int:
        return int(self._stats.num_detectors)

    @property
    def num_observables(self) -> int:
        return int(self._stats.num_observables)

    def reference_sample(self, *, bit_packed: bool = False) -> np.ndarray:
        """Mirror of stim.Circuit.reference_sample (TableauSimulator::reference_sample_circuit, tableau_simulator.inl:1435-1438):
        the noiseless sample with every random measurement result forced to 0; host stabilizer simulation (tableau_ref.cc)."""
        from . import _reference_sample

        m = self.num_measurements
        packed = _reference_sample.reference_sample_bits(self._text, m)
        return packed if bit_packed else np.unpackbits(packed, bitorder="little", count=m).astype(np.bool_)

    def compile_detector_sampler(self, *, seed=None, device: int = 0, engine: str = "auto") -> "CompiledDetectorSampler":
        """`engine` (not in the reference): "auto" | "interp" | "events" — include/gstim.h "sampling engines"."""
        return CompiledDetectorSampler(self, seed=seed, device=device, engine=engine)

    def compile_m2d_converter(self, *, skip_reference_sample: bool = False, device: int = 0) -> "CompiledMeasurementsToDetectionEventsConverter":
        return CompiledMeasurementsToDetectionEventsConverter(self, skip_reference_sample=skip_reference_sample, device=device)

    def compile_sampler(self, *, skip_reference_sample: bool = False, seed=None, reference_sample=None,
                        device: int = 0, engine: str = "auto") -> "CompiledMeasurementSampler":
        return CompiledMeasurementSampler(
            self, skip_reference_sample=skip_reference_sample, seed=seed, reference_sample=reference_sample, device=device,
            engine=engine)


_TABLE_ARRAYS = ["classes", "entries", "overflow", "site_group", "site_index", "outcome_word", "slices"]


def _shape_table(out: dict) -> dict:
    out["classes"] = out["classes"].reshape(-1, 24)
    out["entries"] = out["entries"].reshape(-1, 4)
    if "slices" in out:
        out["slices"] = out["slices"].reshape(-1, 4)
    if "periods" in out:
        out["periods"] = out["periods"].reshape(-1, 4)
    return out


def response_table(circuit, mode: str = "detectors") -> dict:
    """Host-only (no GPU): the event engine's response table of a circuit (gstim_response_table_create) as numpy
    arrays plus an "info" dict; info["eligible"] == 0 means the circuit has none (info["why_not"])."""
    text = str(circuit).encode("utf-8")
    h = ctypes.c_void_p()
    m = _native.MODE_DETECTORS if mode == "detectors" else _native.MODE_MEASUREMENTS
    _native.check(_native.lib().gstim_response_table_create(text, len(text), m, ctypes.byref(h)))
    try:
        info = _native.GstimEngineInfo()
        _native.check(_native.lib().gstim_response_table_info(h, ctypes.byref(info)))
        d = {k: getattr(info, k) for k, _ in info._fields_}
        d["why_not"] = info.why_not.decode()
        out = {"info": d}
        if not info.eligible:
            return out
        for what, name in list(enumerate(_TABLE_ARRAYS[:6])) + [(7, "periods")]:
            n = ctypes.c_size_t(0)
            _native.check(_native.lib().gstim_response_table_get(h, what, None, ctypes.byref(n)))
            a = np.zeros(n.value, dtype=np.uint32)
            if n.value:
                _native.check(_native.lib().gstim_response_table_get(h, what, a.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n)))
            out[name] = a
        return _shape_table(out)
    finally:
        _native.lib().gstim_response_table_destroy(h)


class _Sampler:
    def __init__(self, circuit: Circuit, mode: int, seed, device: int, engine: str = "auto"):
        if engine not in _native.ENGINE_NAMES:
            raise ValueError("engine must be 'auto', 'interp' or 'events'")
        if isinstance(circuit, str):
            circuit = Circuit(circuit)
        self._circuit = circuit
        self._handle = ctypes.c_void_p()
        data = str(circuit).encode("utf-8")
        if isinstance(device, (list, tuple)):
            # several GPUs of this process: host-output sample() calls shard their shots over them (gstim_create_from_text_multi)
            devs = (ctypes.c_int * len(device))(*[int(d) for d in device])
            _native.check(_native.lib().gstim_create_from_text_multi(
                data, len(data), mode, ctypes.c_uint64(_seed_to_u64(seed)), devs, len(device), ctypes.byref(self._handle)))
        else:
            _native.check(_native.lib().gstim_create_from_text(
                data, len(data), mode, ctypes.c_uint64(_seed_to_u64(seed)), int(device), ctypes.byref(self._handle)))
        st = GstimStats()
        _native.check(_native.lib().gstim_get_stats(self._handle, ctypes.byref(st)))
        self.stats = st
        if engine != "auto":
            self.set_engine(engine)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value and _native is not None:  # (module globals are None during interpreter shutdown)
            _native.lib().gstim_destroy(h)
            self._handle = ctypes.c_void_p()

    # -- introspection used by tests / bench ---------------------------------------------------
    def program_words(self) -> np.ndarray:
        n = ctypes.c_size_t(0)
        _native.check(_native.lib().gstim_get_program(self._handle, None, ctypes.byref(n)))
        out = np.empty(n.value, dtype=np.uint32)
        _native.check(_native.lib().gstim_get_program(self._handle, out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n)))
        return out

    @property
    def shot_offset(self) -> int:
        v = ctypes.c_uint64(0)
        _native.check(_native.lib().gstim_get_shot_offset(self._handle, ctypes.byref(v)))
        return int(v.value)

    @shot_offset.setter
    def shot_offset(self, value: int):
        _native.check(_native.lib().gstim_set_shot_offset(self._handle, ctypes.c_uint64(int(value))))

    def set_engine(self, engine: str) -> None:
        """"auto" | "interp" | "events" (gstim_set_engine); "events" raises ValueError when the circuit is not eligible."""
        _native.check(_native.lib().gstim_set_engine(self._handle, _native.ENGINE_NAMES[engine]))

    def engine_info(self) -> dict:
        info = _native.GstimEngineInfo()
        _native.check(_native.lib().gstim_get_engine_info(self._handle, ctypes.byref(info)))
        d = {k: getattr(info, k) for k, _ in info._fields_}
        d["why_not"] = info.why_not.decode()
        d["last_engine"] = {0: "auto", 1: "interp", 2: "events"}[info.last_engine]
        return d

    def response_table(self) -> dict:
        """Arrays of the event engine's response table (gstim_get_response_table), for tests and the oracle."""
        out = {}
        for what, name in enumerate(_TABLE_ARRAYS):
            n = ctypes.c_size_t(0)
            _native.check(_native.lib().gstim_get_response_table(self._handle, what, None, ctypes.byref(n)))
            a = np.zeros(n.value, dtype=np.uint32)
            if n.value:
                _native.check(_native.lib().gstim_get_response_table(self._handle, what, a.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n)))
            out[name] = a
        return _shape_table(out)

    def set_block_columns(self, columns: int) -> None:
        """Pins K (128-shot columns per thread block); 0 = automatic. See gstim_set_block_columns."""
        _native.check(_native.lib().gstim_set_block_columns(self._handle, int(columns)))

    def last_launch_count(self) -> int:
        v = ctypes.c_uint64(0)
        _native.check(_native.lib().gstim_last_launch_count(self._handle, ctypes.byref(v)))
        return int(v.value)

    def last_block_columns(self) -> int:
        v = ctypes.c_uint32(0)
        _native.check(_native.lib().gstim_last_block_columns(self._handle, ctypes.byref(v)))
        return int(v.value)

    def last_call_ms(self) -> float:
        v = ctypes.c_float(0)
        _native.check(_native.lib().gstim_last_call_ms(self._handle, ctypes.byref(v)))
        return float(v.value)

    def last_kernel_ms(self) -> Tuple[float, float]:
        a, b = ctypes.c_float(0), ctypes.c_float(0)
        _native.check(_native.lib().gstim_last_kernel_ms(self._handle, ctypes.byref(a), ctypes.byref(b)))
        return float(a.value), float(b.value)


def _bit_counts(self, shots: int, pairs: bool = True, single_dev_ptr: int = 0, pair_dev_ptr: int = 0):
    """(single[n], pair[n - 1]) uint64 flip counts of every output bit and of adjacent bit pairs over `shots` fresh
    shots, reduced on the device (gstim_bit_counts); optional device copies for an NCCL allreduce."""
    n = int(self.stats.num_detectors + self.stats.num_observables) if isinstance(self, CompiledDetectorSampler) else int(
        self.stats.num_measurements)
    single = np.zeros(n, dtype=np.uint64)
    pair = np.zeros(max(n - 1, 0), dtype=np.uint64)
    _native.check(_native.lib().gstim_bit_counts(
        self._handle, int(shots), single.ctypes.data_as(ctypes.c_void_p),
        pair.ctypes.data_as(ctypes.c_void_p) if pairs and pair.size else None,
        ctypes.c_void_p(single_dev_ptr or None), ctypes.c_void_p(pair_dev_ptr or None)))
    return single, pair


_Sampler.bit_counts = _bit_counts


def _prepare_out(buf, shots: int, n_bits: int, bit_packed: bool):
    """Validates / allocates an output array like numpy.pybind.cc:20-102 does.

    Returns (array_to_return, native_target, copy_back) where native_target is a C-contiguous-in-dim-1 array."""
    width = (n_bits + 7) // 8 if bit_packed else n_bits
    dtype = np.uint8 if bit_packed else np.bool_
    if buf is None:
        arr = np.empty((shots, width), dtype=dtype)
        return arr, arr, False
    if not isinstance(buf, np.ndarray) or buf.dtype != dtype:
        raise ValueError("Output buffer wasn't a numpy.ndarray[np.%s]." % ("uint8" if bit_packed else "bool_"))
    if buf.ndim != 2:
        raise ValueError("Output buffer wasn't two dimensional.")
    if buf.shape != (shots, width):
        raise ValueError(
            f"Expected output buffer to have shape=({shots}, {width}) but its shape is ({buf.shape[0]}, {buf.shape[1]}).")
    if buf.size == 0 or (buf.strides[1] == 1 and buf.strides[0] >= width):
        return buf, buf, False
    tmp = np.empty((shots, width), dtype=dtype)  # exotic strides: sample densely, then scatter
    return buf, tmp, True


def _ptr(a: Optional[np.ndarray]):
    return None if a is None or a.size == 0 else a.ctypes.data_as(ctypes.c_void_p)


def _stride(a: Optional[np.ndarray]) -> int:
    return 0 if a is None or a.size == 0 else int(a.strides[0])


class CompiledDetectorSampler(_Sampler):
    """Mirror of stim.CompiledDetectorSampler (compiled_detector_sampler.pybind.cc:151-421)."""

    def __init__(self, circuit: Circuit, *, seed=None, device: int = 0, engine: str = "auto"):
        super().__init__(circuit, _native.MODE_DETECTORS, seed, device, engine)

    def sample(
        self,
        shots: int,
        *,
        prepend_observables: bool = False,
        append_observables: bool = False,
        separate_observables: bool = False,
        bit_packed: bool = False,
        dets_out: Optional[np.ndarray] = None,
        obs_out: Optional[np.ndarray] = None,
    ) -> Union[np.ndarray, Tuple[np.ndarray, np.ndarray]]:
        if separate_observables and (append_observables or prepend_observables):
            raise ValueError(
                "Can't specify separate_observables=True with append_observables=True or prepend_observables=True")
        shots = int(shots)
        if shots < 0:
            raise ValueError("shots must be non-negative")
        D, L = int(self.stats.num_detectors), int(self.stats.num_observables)
        n_main = D + (L if append_observables else 0) + (L if prepend_observables else 0)
        want_obs = separate_observables or obs_out is not None
        det_ret, det_native, det_copy = _prepare_out(dets_out, shots, n_main, bit_packed)
        obs_ret = obs_native = None
        obs_copy = False
        if want_obs:
            obs_ret, obs_native, obs_copy = _prepare_out(obs_out, shots, L, bit_packed)
        if append_observables and prepend_observables:
            # The reference concatenates obs + dets + obs in this case (compiled_detector_sampler.pybind.cc:64-75).
            both = self.sample(shots, append_observables=True, separate_observables=False, bit_packed=False,
                               obs_out=obs_native if (want_obs and not bit_packed) else None)
            full = np.concatenate([both[:, D:], both], axis=1)
            res = np.packbits(full, axis=1, bitorder="little") if bit_packed else full
            det_native[...] = res
            if want_obs and bit_packed:
                obs_native[...] = np.packbits(both[:, D:], axis=1, bitorder="little")
        else:
            flags = (_native.BIT_PACKED if bit_packed else 0) | (_native.PREPEND_OBS if prepend_observables else 0) | (
                _native.APPEND_OBS if append_observables else 0)
            if want_obs:
                if flags & (_native.PREPEND_OBS | _native.APPEND_OBS):
                    # obs_out together with prepend/append: two passes would resample; emit obs from the main rows.
                    _native.check(_native.lib().gstim_sample_detectors(
                        self._handle, shots, flags, _ptr(det_native), _stride(det_native), None, 0))
                    main_bits = np.unpackbits(det_native, axis=1, bitorder="little")[:, :n_main] if bit_packed else det_native
                    ob = main_bits[:, :L] if prepend_observables else main_bits[:, D:D + L]
                    obs_native[...] = np.packbits(ob, axis=1, bitorder="little") if bit_packed else ob
                else:
                    flags |= _native.SEPARATE_OBS
                    _native.check(_native.lib().gstim_sample_detectors(
                        self._handle, shots, flags, _ptr(det_native), _stride(det_native), _ptr(obs_native),
                        _stride(obs_native)))
            else:
                _native.check(_native.lib().gstim_sample_detectors(
                    self._handle, shots, flags, _ptr(det_native), _stride(det_native), None, 0))
        if det_copy:
            det_ret[...] = det_native
        if obs_copy:
            obs_ret[...] = obs_native
        if separate_observables:
            return det_ret, obs_ret
        return det_ret

    def sample_bit_packed(self, shots: int, *, prepend_observables: bool = False, append_observables: bool = False) -> np.ndarray:
        """[DEPRECATED in the reference] use sample(..., bit_packed=True)."""
        return self.sample(shots, prepend_observables=prepend_observables, append_observables=append_observables, bit_packed=True)

    def sample_write(
        self,
        shots: int,
        *,
        filepath,
        format: str = "01",
        obs_out_filepath=None,
        obs_out_format: str = "01",
        prepend_observables: bool = False,
        append_observables: bool = False,
    ) -> None:
        path = _path_str(filepath, "to ")
        obs_path = None if obs_out_filepath is None else _path_str(obs_out_filepath, "observables to ")
        flags = (_native.PREPEND_OBS if prepend_observables else 0) | (_native.APPEND_OBS if append_observables else 0)
        with _native.open_out(path) as f:
            of = _native.open_out(obs_path) if obs_path is not None else None
            try:
                _native.check(_native.lib().gstim_sample_detectors_to_fd(
                    self._handle, int(shots), flags, f.fileno(), format.encode(), of.fileno() if of else -1,
                    obs_out_format.encode()))
            finally:
                if of:
                    of.close()

    def sample_device(self, shots: int, dets_ptr: int, *, dets_shot_stride: int = 0, obs_ptr: int = 0,
                      obs_shot_stride: int = 0, append_observables: bool = False, prepend_observables: bool = False) -> None:
        """Bit-packed results written straight to DEVICE memory (raw pointers, e.g. tensor.data_ptr())."""
        flags = _native.BIT_PACKED | (_native.PREPEND_OBS if prepend_observables else 0) | (
            _native.APPEND_OBS if append_observables else 0) | (_native.SEPARATE_OBS if obs_ptr else 0)
        _native.check(_native.lib().gstim_sample_detectors_device(
            self._handle, int(shots), flags, ctypes.c_void_p(dets_ptr or None), dets_shot_stride,
            ctypes.c_void_p(obs_ptr or None), obs_shot_stride))

    def sample_torch(self, shots: int, *, separate_observables: bool = True):
        """Device-resident bit-packed results as torch uint8 CUDA tensors: (dets[shots, ceil(D/8)], obs[shots, ceil(L/8)])
        or one [shots, ceil((D+L)/8)] tensor with the observables appended. The consumer hook of SURVEY §8(f) rank 1
        (sinter's sample -> decode loop, glue/sample/src/sinter/_decoding/_stim_then_decode_sampler.py:162-185, without
        the PCIe drain): a decoder or a reduction that lives on the GPU reads the shots where they were produced."""
        import torch

        dev = torch.device("cuda", torch.cuda.current_device())
        D, L = int(self.stats.num_detectors), int(self.stats.num_observables)
        if separate_observables:
            dets = torch.empty((shots, (D + 7) // 8), dtype=torch.uint8, device=dev)
            obs = torch.empty((shots, (L + 7) // 8), dtype=torch.uint8, device=dev)
            self.sample_device(shots, dets.data_ptr() if dets.numel() else 0, obs_ptr=obs.data_ptr() if obs.numel() else 0)
            return dets, obs
        out = torch.empty((shots, (D + L + 7) // 8), dtype=torch.uint8, device=dev)
        self.sample_device(shots, out.data_ptr(), append_observables=True)
        return out

    def sample_pinned(self, shots: int, *, separate_observables: bool = True):
        """Bit-packed results in page-locked host memory (numpy views of pinned torch tensors): the library DMAs straight
        into them (no staging copy), and a consumer can feed them to further async H2D copies."""
        import torch

        D, L = int(self.stats.num_detectors), int(self.stats.num_observables)
        if separate_observables:
            dets = torch.empty((shots, (D + 7) // 8), dtype=torch.uint8, pin_memory=True).numpy()
            obs = torch.empty((shots, (L + 7) // 8), dtype=torch.uint8, pin_memory=True).numpy()
            self.sample(shots, bit_packed=True, dets_out=dets, obs_out=obs)
            return dets, obs
        out = torch.empty((shots, (D + L + 7) // 8), dtype=torch.uint8, pin_memory=True).numpy()
        self.sample(shots, bit_packed=True, append_observables=True, dets_out=out)
        return out

    def flip_counts(self, shots: int, counts_dev_ptr: int = 0) -> np.ndarray:
        """uint64[D+L] flip counts over `shots` fresh shots (device copy optional, for NCCL allreduce)."""
        n = int(self.stats.num_detectors + self.stats.num_observables)
        out = np.zeros(n, dtype=np.uint64)
        _native.check(_native.lib().gstim_detector_flip_counts(
            self._handle, int(shots), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(counts_dev_ptr or None)))
        return out

    def __repr__(self) -> str:
        return f"stim_b200.CompiledDetectorSampler({self._circuit!r})"


class CompiledMeasurementSampler(_Sampler):
    """Mirror of stim.CompiledMeasurementSampler (compiled_measurement_sampler.pybind.cc:26-290)."""

    def __init__(self, circuit: Circuit, *, skip_reference_sample: bool = False, seed=None, reference_sample=None,
                 device: int = 0, engine: str = "auto"):
        if reference_sample is not None and skip_reference_sample:
            raise ValueError("reference_sample is specified but skip_reference_sample=True")
        super().__init__(circuit, _native.MODE_MEASUREMENTS, seed, device, engine)
        M = int(self.stats.num_measurements)
        if reference_sample is not None:
            ref = np.asarray(reference_sample)
            if ref.dtype == np.bool_ and ref.shape == (M,):
                packed = np.packbits(ref, bitorder="little")
            elif ref.dtype == np.uint8 and ref.shape == ((M + 7) // 8,):
                packed = np.ascontiguousarray(ref)
            else:
                raise ValueError(
                    "reference_sample must be a numpy array of dtype bool_ with shape (num_measurements,) or of dtype "
                    "uint8 with shape (ceil(num_measurements / 8),).")
            self._set_reference(packed)
        elif not skip_reference_sample:
            from ._reference_sample import reference_sample_bits
            self._set_reference(reference_sample_bits(str(circuit), M))

    def _set_reference(self, packed: np.ndarray):
        M = int(self.stats.num_measurements)
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        _native.check(_native.lib().gstim_set_reference_sample(self._handle, _ptr(packed) if M else None, M))

    def sample(self, shots: int, *, bit_packed: bool = False) -> np.ndarray:
        shots = int(shots)
        M = int(self.stats.num_measurements)
        ret, native, _ = _prepare_out(None, shots, M, bit_packed)
        flags = _native.BIT_PACKED if bit_packed else 0
        _native.check(_native.lib().gstim_sample_measurements(self._handle, shots, flags, _ptr(native), _stride(native)))
        return ret

    def sample_bit_packed(self, shots: int) -> np.ndarray:
        """[DEPRECATED in the reference] use sample(..., bit_packed=True)."""
        return self.sample(shots, bit_packed=True)

    def sample_write(self, shots: int, *, filepath, format: str = "01") -> None:
        path = _path_str(filepath, "to ")
        with _native.open_out(path) as f:
            _native.check(_native.lib().gstim_sample_measurements_to_fd(self._handle, int(shots), f.fileno(), format.encode()))

    def sample_device(self, shots: int, out_ptr: int, *, shot_stride: int = 0) -> None:
        _native.check(_native.lib().gstim_sample_measurements_device(
            self._handle, int(shots), ctypes.c_void_p(out_ptr or None), shot_stride))

    def __repr__(self) -> str:
        return f"stim_b200.CompiledMeasurementSampler({self._circuit!r})"


class CompiledMeasurementsToDetectionEventsConverter:
    """Mirror of stim.CompiledMeasurementsToDetectionEventsConverter
    (/root/reference/src/stim/simulators/measurements_to_detection_events.pybind.cc:78-137) on gstim_m2d_convert."""

    def __init__(self, circuit: Circuit, *, skip_reference_sample: bool = False, device: int = 0):
        if isinstance(circuit, str):
            circuit = Circuit(circuit)
        self._circuit = circuit
        self._skip = bool(skip_reference_sample)
        self._handle = ctypes.c_void_p()
        data = str(circuit).encode("utf-8")
        _native.check(_native.lib().gstim_m2d_create_from_text(data, len(data), int(self._skip), int(device), ctypes.byref(self._handle)))
        v = [ctypes.c_uint64(0) for _ in range(4)]
        _native.check(_native.lib().gstim_m2d_get_sizes(self._handle, *[ctypes.byref(x) for x in v]))
        self.num_measurements, self.num_detectors, self.num_observables, self.num_sweep_bits = [int(x.value) for x in v]

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value and _native is not None:
            _native.lib().gstim_m2d_destroy(h)
            self._handle = ctypes.c_void_p()

    @staticmethod
    def _packed_rows(a, n_bits: int, what: str):
        """bool_[shots, n_bits] or uint8[shots, ceil(n_bits / 8)] -> (packed C-contiguous uint8 rows, shots), like
        numpy_array_to_transposed_simd_table (/root/reference/src/stim/py/numpy.pybind.cc)."""
        a = np.asarray(a)
        if a.ndim != 2:
            raise ValueError(f"{what} must be a 2-dimensional numpy array.")
        if a.dtype == np.bool_:
            if a.shape[1] != n_bits:
                raise ValueError(f"{what}.dtype == bool_ but {what}.shape[1] != {n_bits}")
            return np.ascontiguousarray(np.packbits(a, axis=1, bitorder="little")) if n_bits else np.zeros((a.shape[0], 0), np.uint8), a.shape[0]
        if a.dtype == np.uint8:
            if a.shape[1] != (n_bits + 7) // 8:
                raise ValueError(f"{what}.dtype == uint8 but {what}.shape[1] != ceil({n_bits} / 8)")
            return np.ascontiguousarray(a), a.shape[0]
        raise ValueError(f"{what} must have dtype bool_ or uint8.")

    def convert(self, *, measurements, sweep_bits=None, separate_observables=None, append_observables=None,
                bit_packed: bool = False, bit_pack_result: bool = False):
        bit_packed = bool(bit_packed or bit_pack_result)
        if separate_observables is None and append_observables is None:
            raise ValueError(
                "To ignore observable flip data, you must explicitly specify either separate_observables=False or "
                "append_observables=False.")
        separate, append = bool(separate_observables), bool(append_observables)
        meas, shots = self._packed_rows(measurements, self.num_measurements, "measurements")
        sweep = None
        if sweep_bits is not None:
            sweep, n2 = self._packed_rows(sweep_bits, self.num_sweep_bits, "sweep_bits")
            if n2 != shots:
                raise ValueError("Need sweep_bits.shape[0] == measurements.shape[0]")
        D, L = self.num_detectors, self.num_observables
        n_main = D + (L if append else 0)
        dets = np.zeros((shots, (n_main + 7) // 8), dtype=np.uint8)
        obs = np.zeros((shots, (L + 7) // 8), dtype=np.uint8) if separate else None
        flags = _native.BIT_PACKED | (_native.APPEND_OBS if append else 0) | (_native.SEPARATE_OBS if separate else 0)
        _native.check(_native.lib().gstim_m2d_convert(
            self._handle, shots, flags, _ptr(meas), _stride(meas), _ptr(sweep), _stride(sweep), _ptr(dets), _stride(dets),
            _ptr(obs), _stride(obs)))
        if not bit_packed:
            dets = np.unpackbits(dets, axis=1, bitorder="little", count=n_main).astype(np.bool_) if n_main else np.zeros((shots, 0), np.bool_)
            if separate:
                obs = np.unpackbits(obs, axis=1, bitorder="little", count=L).astype(np.bool_) if L else np.zeros((shots, 0), np.bool_)
        return (dets, obs) if separate else dets

    def convert_file(self, *, measurements_filepath, measurements_format: str = "01", sweep_bits_filepath=None,
                     sweep_bits_format: str = "01", detection_events_filepath, detection_events_format: str = "01",
                     append_observables: bool = False, obs_out_filepath=None, obs_out_format: str = "01") -> None:
        """File-to-file conversion (measurements_to_detection_events.pybind.cc:convert_file, command_m2d.cc): every one of the
        six formats on either side."""
        from . import _formats

        with open(measurements_filepath, "rb") as f:
            meas = _formats.read_shots(f.read(), measurements_format, self.num_measurements)
        sweep = None
        if sweep_bits_filepath is not None:
            with open(sweep_bits_filepath, "rb") as f:
                sweep = _formats.read_shots(f.read(), sweep_bits_format, self.num_sweep_bits)
        D, L = self.num_detectors, self.num_observables
        separate = obs_out_filepath is not None
        res = self.convert(measurements=meas, sweep_bits=sweep, append_observables=bool(append_observables),
                           separate_observables=separate, bit_packed=True)
        dets, obs = res if separate else (res, None)
        _write_rows(dets, D + (L if append_observables else 0), detection_events_filepath, detection_events_format, b"D", b"L", D)
        if separate:
            _write_rows(obs, L, obs_out_filepath, obs_out_format, b"L", b"L", L)

    def __repr__(self) -> str:
        return f"stim_b200.CompiledMeasurementsToDetectionEventsConverter({self._circuit!r}, skip_reference_sample={self._skip})"


def _write_rows(rows: np.ndarray, n_bits: int, path, fmt: str, prefix1: bytes, prefix2: bytes, transition: int) -> None:
    """Packed shot-major rows -> file in any of the six formats (writers.cc through gstim_write_shots_to_fd)."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    rows = rows if rows.size else rows.reshape(rows.shape[0], 0)
    with _native.open_out(path) as f:
        buf = rows.ctypes.data_as(ctypes.c_void_p) if rows.size else None
        _native.check(_native.lib().gstim_write_shots_to_fd(
            buf, rows.shape[1], rows.shape[0], n_bits, f.fileno(), fmt.encode(), prefix1, prefix2, transition))


def read_shot_data_file(*, path, format: str, bit_packed: bool = False, num_measurements: Optional[int] = None,
                        num_detectors: Optional[int] = None, num_observables: Optional[int] = None,
                        separate_observables: bool = False, bit_pack: bool = False):
    """Mirror of stim.read_shot_data_file (/root/reference/src/stim/io/read_write.pybind.cc:99-142)."""
    from . import _formats

    if num_measurements is None and num_detectors is None and num_observables is None:
        raise ValueError("Must specify num_measurements, num_detectors, num_observables.")
    nm, nd, no = int(num_measurements or 0), int(num_detectors or 0), int(num_observables or 0)
    bit_packed = bool(bit_packed or bit_pack)
    with open(path, "rb") as f:
        rows = _formats.read_shots(f.read(), format, nm + nd + no, num_measurements=nm, num_detectors=nd, num_observables=no)
    bits = np.unpackbits(rows, axis=1, bitorder="little", count=nm + nd + no).astype(np.bool_) if nm + nd + no else \
        np.zeros((rows.shape[0], 0), np.bool_)

    def out(lo, hi):
        part = bits[:, lo:hi]
        if not bit_packed:
            return np.ascontiguousarray(part)
        return np.packbits(part, axis=1, bitorder="little") if hi > lo else np.zeros((bits.shape[0], 0), np.uint8)

    if separate_observables:
        return out(0, nm + nd), out(nm + nd, nm + nd + no)
    return out(0, nm + nd + no)


def write_shot_data_file(*, data, path, format: str, num_measurements: Optional[int] = None,
                         num_detectors: Optional[int] = None, num_observables: Optional[int] = None) -> None:
    """Mirror of stim.write_shot_data_file (/root/reference/src/stim/io/read_write.pybind.cc:144-181)."""
    if num_measurements is None and num_detectors is None and num_observables is None:
        raise ValueError("Must specify num_measurements, num_detectors, num_observables.")
    nm, nd, no = int(num_measurements or 0), int(num_detectors or 0), int(num_observables or 0)
    if nm != 0 and (nd != 0 or no != 0):
        raise ValueError("num_measurements and (num_detectors or num_observables)")
    n_bits = nm + nd + no
    rows, _ = CompiledMeasurementsToDetectionEventsConverter._packed_rows(data, n_bits, "data")
    _write_rows(rows, n_bits, path, format, b"D" if nm == 0 else b"M", b"L" if nm == 0 else b"M", nm + nd)
