"""ctypes binding of libgstim.so (C ABI: include/gstim.h). Fails loudly if the library is missing —
there is no Python/CPU fallback for the sampling path."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgstim.so")

GSTIM_OK = 0
ERR_INVALID_ARGUMENT = 1
ERR_OUT_OF_RANGE = 2
ERR_CUDA = 3
ERR_OOM = 4
ERR_IO = 5
ERR_INTERNAL = 6

MODE_DETECTORS = 0
MODE_MEASUREMENTS = 1

BIT_PACKED = 0x01
PREPEND_OBS = 0x02
APPEND_OBS = 0x04
SEPARATE_OBS = 0x08


class GstimStats(ctypes.Structure):
    _fields_ = [
        ("num_qubits", ctypes.c_uint64),
        ("num_measurements", ctypes.c_uint64),
        ("num_detectors", ctypes.c_uint64),
        ("num_observables", ctypes.c_uint64),
        ("max_lookback", ctypes.c_uint64),
        ("active_qubits", ctypes.c_uint64),
        ("program_words", ctypes.c_uint64),
        ("num_batches", ctypes.c_uint64),
        ("num_barriers", ctypes.c_uint64),
        ("num_noise_sites", ctypes.c_uint64),
        ("num_collapse_sites", ctypes.c_uint64),
        ("threads", ctypes.c_uint32),
        ("lanes_per_item", ctypes.c_uint32),
        ("slots", ctypes.c_uint32),
        ("max_columns", ctypes.c_uint32),
        ("chunk_words", ctypes.c_uint32),
        ("smem_bytes_max", ctypes.c_uint32),
    ]


ENGINE_AUTO, ENGINE_INTERPRETER, ENGINE_EVENTS = 0, 1, 2
ENGINE_NAMES = {"auto": ENGINE_AUTO, "interp": ENGINE_INTERPRETER, "interpreter": ENGINE_INTERPRETER, "events": ENGINE_EVENTS}


class GstimEngineInfo(ctypes.Structure):
    _fields_ = [
        ("eligible", ctypes.c_int32),
        ("favoured", ctypes.c_int32),
        ("last_engine", ctypes.c_int32),
        ("tile_shots", ctypes.c_uint32),
        ("blocks_per_sm", ctypes.c_uint32),
        ("num_classes", ctypes.c_uint32),
        ("num_slices", ctypes.c_uint32),
        ("max_response", ctypes.c_uint32),
        ("num_sites", ctypes.c_uint64),
        ("num_entries", ctypes.c_uint64),
        ("device_entries", ctypes.c_uint64),
        ("overflow_words", ctypes.c_uint64),
        ("events_per_shot", ctypes.c_double),
        ("flips_per_shot", ctypes.c_double),
        ("why_not", ctypes.c_char * 96),
    ]


class GstimCudaError(RuntimeError):
    """CUDA failure or no usable device (the library has no CPU fallback)."""


# Every symbol include/gstim.h declares: (name, restype, argtypes)
_P = ctypes.c_void_p
_SIGNATURES = [
    ("gstim_version", ctypes.c_int, []),
    ("gstim_last_error", ctypes.c_char_p, []),
    ("gstim_device_count", ctypes.c_int, []),
    ("gstim_circuit_stats", ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(GstimStats)]),
    ("gstim_reference_sample", ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, _P, ctypes.c_size_t]),
    ("gstim_lower_text", ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32,
                                        _P, ctypes.POINTER(ctypes.c_size_t), _P]),
    ("gstim_create_from_text", ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64, ctypes.c_int,
                                              ctypes.POINTER(_P)]),
    ("gstim_create_from_text_multi", ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64,
                                                    ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(_P)]),
    ("gstim_destroy", None, [_P]),
    ("gstim_get_stats", ctypes.c_int, [_P, ctypes.POINTER(GstimStats)]),
    ("gstim_get_program", ctypes.c_int, [_P, _P, ctypes.POINTER(ctypes.c_size_t)]),
    ("gstim_set_reference_sample", ctypes.c_int, [_P, _P, ctypes.c_size_t]),
    ("gstim_get_shot_offset", ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_uint64)]),
    ("gstim_set_shot_offset", ctypes.c_int, [_P, ctypes.c_uint64]),
    ("gstim_sample_detectors", ctypes.c_int, [_P, ctypes.c_uint64, ctypes.c_uint32, _P, ctypes.c_int64, _P, ctypes.c_int64]),
    ("gstim_sample_measurements", ctypes.c_int, [_P, ctypes.c_uint64, ctypes.c_uint32, _P, ctypes.c_int64]),
    ("gstim_sample_detectors_device", ctypes.c_int, [_P, ctypes.c_uint64, ctypes.c_uint32, _P, ctypes.c_int64, _P, ctypes.c_int64]),
    ("gstim_sample_measurements_device", ctypes.c_int, [_P, ctypes.c_uint64, _P, ctypes.c_int64]),
    ("gstim_sample_detectors_to_fd", ctypes.c_int, [_P, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_char_p,
                                                    ctypes.c_int, ctypes.c_char_p]),
    ("gstim_sample_measurements_to_fd", ctypes.c_int, [_P, ctypes.c_uint64, ctypes.c_int, ctypes.c_char_p]),
    ("gstim_write_shots_to_fd", ctypes.c_int, [_P, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int,
                                               ctypes.c_char_p, ctypes.c_char, ctypes.c_char, ctypes.c_uint64]),
    ("gstim_detector_flip_counts", ctypes.c_int, [_P, ctypes.c_uint64, _P, _P]),
    ("gstim_bit_counts", ctypes.c_int, [_P, ctypes.c_uint64, _P, _P, _P, _P]),
    ("gstim_dem_counts", ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64),
                                        ctypes.POINTER(ctypes.c_uint64)]),
    ("gstim_dem_create_from_text", ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(_P)]),
    ("gstim_dem_destroy", None, [_P]),
    ("gstim_dem_set_shot_offset", ctypes.c_int, [_P, ctypes.c_uint64]),
    ("gstim_dem_sample", ctypes.c_int, [_P, ctypes.c_uint64, ctypes.c_uint32, _P, ctypes.c_int64, _P, ctypes.c_int64, _P, ctypes.c_int64]),
    ("gstim_dem_replay", ctypes.c_int, [_P, ctypes.c_uint64, _P, ctypes.c_int64, _P, ctypes.c_int64, _P, ctypes.c_int64]),
    ("gstim_dem_get_response_table", ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.POINTER(ctypes.c_size_t)]),
    ("gstim_dem_bit_counts", ctypes.c_int, [_P, ctypes.c_uint64, _P, _P]),
    ("gstim_dem_sample_to_fd", ctypes.c_int, [_P, ctypes.c_uint64, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p,
                                              ctypes.c_int, ctypes.c_char_p]),
    ("gstim_set_engine", ctypes.c_int, [_P, ctypes.c_int]),
    ("gstim_get_engine_info", ctypes.c_int, [_P, ctypes.POINTER(GstimEngineInfo)]),
    ("gstim_get_response_table", ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.POINTER(ctypes.c_size_t)]),
    ("gstim_response_table_create", ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(_P)]),
    ("gstim_response_table_destroy", None, [_P]),
    ("gstim_response_table_info", ctypes.c_int, [_P, ctypes.POINTER(GstimEngineInfo)]),
    ("gstim_response_table_get", ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.POINTER(ctypes.c_size_t)]),
    ("gstim_m2d_create_from_text", ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_P)]),
    ("gstim_m2d_destroy", None, [_P]),
    ("gstim_m2d_get_sizes", ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64),
                                           ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]),
    ("gstim_m2d_convert", ctypes.c_int, [_P, ctypes.c_uint64, ctypes.c_uint32, _P, ctypes.c_int64, _P, ctypes.c_int64, _P, ctypes.c_int64,
                                         _P, ctypes.c_int64]),
    ("gstim_flipsim_create", ctypes.c_int, [ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(_P)]),
    ("gstim_flipsim_destroy", None, [_P]),
    ("gstim_flipsim_copy", ctypes.c_int, [_P, ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(_P)]),
    ("gstim_flipsim_sizes", ctypes.c_int, [_P] + [ctypes.POINTER(ctypes.c_uint64)] * 6),
    ("gstim_flipsim_do_text", ctypes.c_int, [_P, ctypes.c_char_p, ctypes.c_size_t]),
    ("gstim_flipsim_get_rows", ctypes.c_int, [_P, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, _P]),
    ("gstim_flipsim_set_rows", ctypes.c_int, [_P, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, _P, ctypes.c_int]),
    ("gstim_flipsim_broadcast", ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.c_uint64, ctypes.c_double]),
    ("gstim_flipsim_bernoulli", ctypes.c_int, [_P, ctypes.c_uint64, ctypes.c_double, _P]),
    ("gstim_flipsim_clear", ctypes.c_int, [_P]),
    ("gstim_set_block_columns", ctypes.c_int, [_P, ctypes.c_uint32]),
    ("gstim_measure_lop3_peak", ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                               ctypes.POINTER(ctypes.c_double)]),
    ("gstim_last_launch_count", ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_uint64)]),
    ("gstim_last_block_columns", ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_uint32)]),
    ("gstim_last_call_ms", ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_float)]),
    ("gstim_last_kernel_ms", ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
]

_lib = None


def lib():
    """Loads libgstim.so (building nothing: run `python -m stim_b200.build` or __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing. Build it with `python -m stim_b200.build`; "
                "stim_b200 has no CPU fallback for the sampling path.")
        l = ctypes.CDLL(LIB_PATH)
        for name, restype, argtypes in _SIGNATURES:
            fn = getattr(l, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = l
    return _lib


def exported_symbols():
    return [s[0] for s in _SIGNATURES]


def check(code):
    """Maps ABI error codes to the exception types the reference raises through pybind."""
    if code == GSTIM_OK:
        return
    msg = lib().gstim_last_error().decode("utf-8", "replace")
    if code == ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    if code == ERR_OUT_OF_RANGE:
        raise IndexError(msg)
    if code == ERR_OOM:
        raise MemoryError(msg)
    if code == ERR_IO:
        raise OSError(msg)
    if code == ERR_CUDA:
        raise GstimCudaError(msg)
    raise RuntimeError(msg)


def open_out(path):
    """Opens a result file for writing. "/dev/stdout" means this process's standard output as it is (a duplicate of
    descriptor 1: no truncation, the offset shared with whoever else writes there) like the reference writing to its `stdout`
    FILE* (/root/reference/src/stim/util_bot/arg_parse.cc find_open_file_argument) - opening the path again would truncate a
    regular file that stdout was redirected or appended to."""
    import os
    import sys

    path = os.fspath(path)
    if path == "/dev/stdout":
        sys.stdout.flush()
        return os.fdopen(os.dup(1), "wb")
    return open(path, "wb")
