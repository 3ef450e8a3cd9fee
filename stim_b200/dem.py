"""Detector-error-model sampling: host-side mirror of stim.DetectorErrorModel.compile_sampler() /
stim.CompiledDemSampler (/root/reference/src/stim/simulators/dem_sampler.pybind.cc) over the C ABI
(include/gstim.h: gstim_dem_*). Sampling runs in stim_b200/csrc/dem.cu on the GPU; there is no CPU fallback."""
import ctypes
import os
from typing import Optional, Tuple

import numpy as np

from . import _native


def _seed_to_u64(seed) -> int:
    if seed is None:
        return int.from_bytes(os.urandom(8), "little")
    if not isinstance(seed, (int, np.integer)) or isinstance(seed, bool) or not 0 <= int(seed) < 1 << 64:
        raise ValueError("Expected seed to be None or a 64 bit unsigned integer.")
    return int(seed)


class DetectorErrorModel:
    """A detector error model in Stim's .dem text format. Only what the sampling path needs is mirrored."""

    def __init__(self, detector_error_model_text: str = ""):
        if not isinstance(detector_error_model_text, str):
            raise TypeError("detector_error_model_text must be a str")
        self._text = detector_error_model_text
        d, l, e = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
        data = self._text.encode("utf-8")
        _native.check(_native.lib().gstim_dem_counts(data, len(data), ctypes.byref(d), ctypes.byref(l), ctypes.byref(e)))
        self.num_detectors, self.num_observables, self.num_errors = int(d.value), int(l.value), int(e.value)

    @staticmethod
    def from_file(file) -> "DetectorErrorModel":
        if hasattr(file, "read"):
            return DetectorErrorModel(file.read())
        with open(os.fspath(file), "r") as f:
            return DetectorErrorModel(f.read())

    def __str__(self) -> str:
        return self._text

    def compile_sampler(self, *, seed=None, device: int = 0) -> "CompiledDemSampler":
        return CompiledDemSampler(self, seed=seed, device=device)


class CompiledDemSampler:
    """Mirror of stim.CompiledDemSampler: sample() returns (detection events, observable flips, errors or None)."""

    def __init__(self, dem: DetectorErrorModel, *, seed=None, device: int = 0):
        if isinstance(dem, str):
            dem = DetectorErrorModel(dem)
        self._dem = dem
        self._handle = ctypes.c_void_p()
        data = str(dem).encode("utf-8")
        _native.check(_native.lib().gstim_dem_create_from_text(
            data, len(data), ctypes.c_uint64(_seed_to_u64(seed)), int(device), ctypes.byref(self._handle)))

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value and _native is not None:
            _native.lib().gstim_dem_destroy(h)
            self._handle = ctypes.c_void_p()

    def set_shot_offset(self, offset: int) -> None:
        _native.check(_native.lib().gstim_dem_set_shot_offset(self._handle, ctypes.c_uint64(int(offset))))

    def sample(self, shots: int, *, bit_packed: bool = False, return_errors: bool = False,
               recorded_errors_to_replay=None) -> Tuple[np.ndarray, np.ndarray, Optional[np.ndarray]]:
        shots = int(shots)
        if shots < 0:
            raise ValueError("shots must be non-negative")
        if recorded_errors_to_replay is not None:
            return self._replay(shots, recorded_errors_to_replay, bit_packed, return_errors)
        dt = np.uint8 if bit_packed else np.bool_

        def alloc(n_bits):
            return np.zeros((shots, (n_bits + 7) // 8 if bit_packed else n_bits), dtype=dt)

        dets, obs = alloc(self._dem.num_detectors), alloc(self._dem.num_observables)
        errs = alloc(self._dem.num_errors) if return_errors else None

        def ptr(a):
            return None if a is None or a.size == 0 else a.ctypes.data_as(ctypes.c_void_p)

        _native.check(_native.lib().gstim_dem_sample(
            self._handle, shots, _native.BIT_PACKED if bit_packed else 0, ptr(dets), 0, ptr(obs), 0, ptr(errs), 0))
        return dets, obs, errs

    def _replay(self, shots: int, recorded, bit_packed: bool, return_errors: bool):
        """recorded_errors_to_replay (dem_sampler.pybind.cc:sample): bool_[shots, num_errors] or uint8[shots, ceil(num_errors / 8)];
        the detectors / observables those errors flip, no sampling."""
        D, L, E = self._dem.num_detectors, self._dem.num_observables, self._dem.num_errors
        a = np.asarray(recorded)
        if a.ndim != 2 or a.shape[0] != shots:
            raise ValueError("recorded_errors_to_replay.shape[0] != shots")
        if a.dtype == np.bool_:
            if a.shape[1] != E:
                raise ValueError(f"recorded_errors_to_replay.dtype == bool_ but shape[1] != num_errors ({E})")
            packed = np.packbits(a, axis=1, bitorder="little") if E else np.zeros((shots, 0), np.uint8)
        elif a.dtype == np.uint8:
            if a.shape[1] != (E + 7) // 8:
                raise ValueError(f"recorded_errors_to_replay.dtype == uint8 but shape[1] != ceil(num_errors / 8) ({(E + 7) // 8})")
            packed = np.ascontiguousarray(a)
        else:
            raise ValueError("recorded_errors_to_replay must have dtype bool_ or uint8")
        dets = np.zeros((shots, (D + 7) // 8), dtype=np.uint8)
        obs = np.zeros((shots, (L + 7) // 8), dtype=np.uint8)

        def ptr(x):
            return None if x.size == 0 else x.ctypes.data_as(ctypes.c_void_p)

        if shots:
            _native.check(_native.lib().gstim_dem_replay(self._handle, shots, ptr(packed), 0, ptr(dets), 0, ptr(obs), 0))
        errs = packed if return_errors else None
        if not bit_packed:
            def unpack(x, n):
                return np.unpackbits(x, axis=1, bitorder="little", count=n).astype(np.bool_) if n else np.zeros((shots, 0), np.bool_)
            dets, obs = unpack(dets, D), unpack(obs, L)
            errs = unpack(errs, E) if return_errors else None
        return dets, obs, errs

    def response_table(self) -> dict:
        """Arrays of the event engine's table for this model (gstim_dem_get_response_table) + "tile_shots", for the oracle."""
        out = {}
        for what, name in [(0, "classes"), (1, "entries"), (2, "overflow"), (3, "site_group"), (6, "slices"), (8, "tile")]:
            n = ctypes.c_size_t(0)
            _native.check(_native.lib().gstim_dem_get_response_table(self._handle, what, None, ctypes.byref(n)))
            a = np.zeros(max(n.value, 1), dtype=np.uint32)
            _native.check(_native.lib().gstim_dem_get_response_table(self._handle, what, a.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n)))
            out[name] = a[:n.value]
        out["classes"] = out["classes"].reshape(-1, 24)
        out["entries"] = out["entries"].reshape(-1, 4)
        out["slices"] = out["slices"].reshape(-1, 4)
        out["tile_shots"] = int(out.pop("tile")[0])
        return out

    def bit_counts(self, shots: int):
        """(single[D + L], pair[D + L - 1]) uint64 flip counts over `shots` fresh shots, reduced on the device."""
        n = self._dem.num_detectors + self._dem.num_observables
        single = np.zeros(n, dtype=np.uint64)
        pair = np.zeros(max(n - 1, 0), dtype=np.uint64)
        _native.check(_native.lib().gstim_dem_bit_counts(
            self._handle, int(shots), single.ctypes.data_as(ctypes.c_void_p), pair.ctypes.data_as(ctypes.c_void_p) if pair.size else None))
        return single, pair

    def sample_write(self, shots: int, *, det_out_file=None, det_out_format: str = "01", obs_out_file=None,
                     obs_out_format: str = "01", err_out_file=None, err_out_format: str = "01",
                     replay_err_in_file=None, replay_err_in_format: str = "01") -> None:
        if replay_err_in_file is not None:
            from . import _formats
            from . import _write_rows

            with open(os.fspath(replay_err_in_file), "rb") as f:
                errors = _formats.read_shots(f.read(), replay_err_in_format, self._dem.num_errors)
            if errors.shape[0] < int(shots):
                raise ValueError("The replay file held fewer shots than were requested.")
            errors = errors[: int(shots)]
            dets, obs, _ = self._replay(int(shots), errors, True, False)
            D, L, E = self._dem.num_detectors, self._dem.num_observables, self._dem.num_errors
            for rows, n, path, fmt, prefix in ((dets, D, det_out_file, det_out_format, b"D"), (obs, L, obs_out_file, obs_out_format, b"L"),
                                               (errors, E, err_out_file, err_out_format, b"M")):
                if path is not None:
                    _write_rows(rows, n, os.fspath(path), fmt, prefix, prefix, n)
            return
        files = []
        try:
            fds = []
            for path in (det_out_file, obs_out_file, err_out_file):
                if path is None:
                    fds.append(-1)
                else:
                    f = _native.open_out(path)
                    files.append(f)
                    fds.append(f.fileno())
            _native.check(_native.lib().gstim_dem_sample_to_fd(
                self._handle, int(shots), fds[0], det_out_format.encode(), fds[1], obs_out_format.encode(), fds[2],
                err_out_format.encode()))
        finally:
            for f in files:
                f.close()
