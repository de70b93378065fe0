"""ctypes binding of include/twstft.h.

The product library is amaranth_twstft_b200/libtwstft.so (built in-tree by
``__graft_entry__.build()`` / ``make -C amaranth_twstft_b200/csrc``).  There is no
CPU fallback: if the library is missing, or no CUDA device is present, loading /
plan creation raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libtwstft.so")

VARIANT_G, VARIANT_K = 0, 1
BPSK, QPSK = 0, 1
LAYOUT_INTERLEAVED2, LAYOUT_SINGLE = 0, 1
AXIS_G, AXIS_K = 0, 1

EXPORTS = [
    "twstft_default_config", "twstft_plan_create", "twstft_plan_destroy", "twstft_last_error",
    "twstft_backend", "twstft_set_code", "twstft_gen_code", "twstft_prn_generate", "twstft_process",
    "twstft_process_device", "twstft_process_windows", "twstft_estimate_carrier", "twstft_host_alloc",
    "twstft_host_free", "twstft_device_alloc", "twstft_device_free", "twstft_copy_to_device",
    "twstft_copy_to_host", "twstft_synchronize", "twstft_kernel_launches", "twstft_debug_fetch",
    "twstft_window_bins", "twstft_doppler_bank", "twstft_search_df", "twstft_estimate_carrier_block", "twstft_estimate_bins",
    "twstft_search_df_block", "twstft_estimate_decimated", "twstft_processing_cf64", "twstft_processing_cf64_df",
    "twstft_search_df_cf64", "twstft_shard_create", "twstft_shard_destroy", "twstft_shard_ngpus", "twstft_shard_plan",
    "twstft_shard_last_error", "twstft_shard_set_code", "twstft_shard_process", "twstft_process_cycled",
    "twstft_shard_process_cycled", "twstft_synth_records", "twstft_synth_stream", "twstft_fast_suspended", "twstft_acquire", "twstft_fast_counters", "twstft_timer_start", "twstft_timer_stop", "twstft_profile_enable", "twstft_profile_read",
]


class Config(C.Structure):
    _fields_ = [
        ("fs", C.c_double), ("nsamples", C.c_int64), ("samples_per_chip", C.c_int), ("variant", C.c_int),
        ("modulation", C.c_int), ("nint", C.c_int), ("layout", C.c_int), ("freq_axis", C.c_int),
        ("win_lo_hz", C.c_double), ("win_hi_hz", C.c_double), ("est_sub_mean", C.c_int), ("ddof", C.c_int),
        ("max_batch", C.c_int), ("device_id", C.c_int), ("fine_freq", C.c_int), ("code_window", C.c_int), ("fast_peak", C.c_int),
    ]


RESULT_FIELDS = ["indice", "correction", "xval_re", "xval_im", "xvalm1_re", "xvalm1_im", "xvalp1_re",
                 "xvalp1_im", "SNRr", "SNRi", "df", "puissance", "puissancecode", "puissancenoise",
                 "carrier_bin", "status"]


class Result(C.Structure):
    _fields_ = [(n, C.c_double) for n in RESULT_FIELDS]


BANK_FIELDS = ["b0", "prnsig", "metric", "status"]


class LibraryMissing(RuntimeError):
    pass


_cache = {}


def load(path=None):
    """Load the C-ABI library (default: the in-tree CUDA build) and declare its signatures."""
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise LibraryMissing(
            f"{path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C amaranth_twstft_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(path)
    P = C.c_void_p
    lib.twstft_default_config.argtypes = [C.POINTER(Config)]
    lib.twstft_default_config.restype = None
    lib.twstft_plan_create.argtypes = [C.POINTER(Config), C.POINTER(P)]
    lib.twstft_plan_destroy.argtypes = [P]
    lib.twstft_plan_destroy.restype = None
    lib.twstft_last_error.argtypes = [P]
    lib.twstft_last_error.restype = C.c_char_p
    lib.twstft_backend.restype = C.c_char_p
    lib.twstft_set_code.argtypes = [P, P, C.c_size_t, C.c_int]
    lib.twstft_gen_code.argtypes = [P, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t, P]
    lib.twstft_prn_generate.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t, P]
    lib.twstft_process.argtypes = [P, P, C.c_size_t, P]
    lib.twstft_process_device.argtypes = [P, P, C.c_size_t, P]
    lib.twstft_process_windows.argtypes = [P, P, C.c_int, C.c_size_t, P, P, C.c_size_t, P]
    lib.twstft_estimate_carrier.argtypes = [P, P, C.c_int, C.c_size_t, P, P]
    lib.twstft_doppler_bank.argtypes = [P, P, C.c_int, C.c_size_t, C.c_int64, P, C.c_size_t, C.c_int, P]
    lib.twstft_acquire.argtypes = [P, P, C.c_int, C.c_size_t, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                   C.c_int, C.c_int, P, P, P]
    lib.twstft_search_df.argtypes = [P, P, C.c_int, C.c_size_t, C.c_int64, C.c_double, P, P, P]
    lib.twstft_estimate_carrier_block.argtypes = [P, P, C.c_int, C.c_size_t, C.c_int, P, P]
    lib.twstft_search_df_block.argtypes = [P, P, P, C.c_int, C.c_int, C.c_double, P, P, P]
    lib.twstft_processing_cf64.argtypes = [P, P, C.c_size_t, C.c_int64, C.c_int64, P]
    lib.twstft_processing_cf64_df.argtypes = [P, P, C.c_size_t, P, P]
    lib.twstft_search_df_cf64.argtypes = [P, P, P, C.c_int, C.c_int64, C.c_int64, C.c_double, P, P, P]
    lib.twstft_shard_create.argtypes = [C.POINTER(Config), P, C.c_int, C.POINTER(P)]
    lib.twstft_shard_destroy.argtypes = [P]
    lib.twstft_shard_destroy.restype = None
    lib.twstft_shard_ngpus.argtypes = [P]
    lib.twstft_shard_plan.argtypes = [P, C.c_int]
    lib.twstft_shard_plan.restype = P
    lib.twstft_shard_last_error.argtypes = [P]
    lib.twstft_shard_last_error.restype = C.c_char_p
    lib.twstft_shard_set_code.argtypes = [P, P, C.c_size_t, C.c_int]
    lib.twstft_shard_process.argtypes = [P, P, C.c_size_t, P]
    lib.twstft_process_cycled.argtypes = [P, P, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, P]
    lib.twstft_shard_process_cycled.argtypes = [P, P, C.c_size_t, C.c_size_t, C.c_int, P, P, P]
    lib.twstft_synth_records.argtypes = [P, C.c_uint64, C.c_size_t, C.c_double, C.c_double, P]
    lib.twstft_synth_stream.argtypes = [P, C.c_uint64, C.c_size_t, C.c_double, C.c_double, C.c_double, P]
    lib.twstft_estimate_bins.argtypes = [P, P, C.c_int, C.c_size_t, C.c_int64, C.c_int, P]
    lib.twstft_fast_counters.argtypes = [P, P, P]
    lib.twstft_fast_suspended.argtypes = [P]
    lib.twstft_fast_suspended.restype = C.c_int64
    lib.twstft_estimate_decimated.argtypes = [P, P, C.c_size_t, C.c_int, C.c_double, C.c_int, P, P]
    lib.twstft_host_alloc.argtypes = [C.c_size_t]
    lib.twstft_host_alloc.restype = P
    lib.twstft_host_free.argtypes = [P]
    lib.twstft_host_free.restype = None
    lib.twstft_device_alloc.argtypes = [P, C.c_size_t, C.POINTER(P)]
    lib.twstft_device_free.argtypes = [P, P]
    lib.twstft_copy_to_device.argtypes = [P, P, P, C.c_size_t]
    lib.twstft_copy_to_host.argtypes = [P, P, P, C.c_size_t]
    lib.twstft_synchronize.argtypes = [P]
    lib.twstft_kernel_launches.argtypes = [P]
    lib.twstft_kernel_launches.restype = C.c_int64
    lib.twstft_debug_fetch.argtypes = [P, C.c_int, C.c_int, P, C.c_size_t]
    lib.twstft_window_bins.argtypes = [P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.twstft_timer_start.argtypes = [P]
    lib.twstft_timer_stop.argtypes = [P, C.POINTER(C.c_float)]
    lib.twstft_profile_enable.argtypes = [P, C.c_int]
    lib.twstft_profile_read.argtypes = [P, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    _cache[path] = lib
    return lib
