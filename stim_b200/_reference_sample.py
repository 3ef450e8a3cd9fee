"""Noiseless reference sample for compile_sampler(). Host-side stabilizer simulation (the reference keeps
this on the CPU too: TableauSimulator::reference_sample_circuit, tableau_simulator.inl:1435-1438)."""
import ctypes

import numpy as np

from . import _native


def reference_sample_bits(circuit_text: str, num_measurements: int) -> np.ndarray:
    """Returns the reference sample as little-endian packed uint8[ceil(M/8)]."""
    out = np.zeros((num_measurements + 7) // 8, dtype=np.uint8)
    fn = getattr(_native.lib(), "gstim_reference_sample", None)
    if fn is None:
        raise NotImplementedError(
            "Computing the reference sample is not available in this build; pass skip_reference_sample=True or "
            "an explicit reference_sample.")
    data = circuit_text.encode("utf-8")
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
    _native.check(fn(data, len(data), out.ctypes.data_as(ctypes.c_void_p), num_measurements))
    return out
