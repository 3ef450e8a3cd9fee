"""Noiseless reference sample for compile_sampler(): host-side stabilizer simulation (the reference keeps this on
the CPU too: TableauSimulator::reference_sample_circuit, tableau_simulator.inl:1435-1438)."""
import ctypes

import numpy as np

from . import _native


def reference_sample_bits(circuit_text: str, num_measurements: int) -> np.ndarray:
    """Returns the reference sample as little-endian packed uint8[ceil(M/8)]."""
    out = np.zeros((num_measurements + 7) // 8, dtype=np.uint8)
    data = circuit_text.encode("utf-8")
    _native.check(_native.lib().gstim_reference_sample(
        data, len(data), out.ctypes.data_as(ctypes.c_void_p) if num_measurements else None, num_measurements))
    return out
