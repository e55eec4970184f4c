"""ctypes binding of libcphnsw_b200.so (include/cphnsw_b200.h).

The library is the product: if it is missing or cannot be loaded this module raises -- there is
no Python or CPU fallback for the query path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

HERE = Path(__file__).resolve().parent
# CPHNSW_B200_LIB: another build of the same library (A/B runs of kernel variants, build.py --variant); still native code
LIB_PATH = Path(os.environ["CPHNSW_B200_LIB"]).resolve() if os.environ.get("CPHNSW_B200_LIB") else HERE / "libcphnsw_b200.so"

OK, EINVAL, ERUNTIME, ECUDA, ENOMEM = 0, -1, -2, -3, -4


class Info(C.Structure):
    _fields_ = [
        ("D", C.c_uint32), ("bits", C.c_uint32), ("dim", C.c_uint32), ("n", C.c_uint64),
        ("max_level", C.c_int32), ("entry_point", C.c_uint32), ("n_layers", C.c_uint32),
        ("block_stride", C.c_uint32), ("device_bytes", C.c_uint64),
        ("affine_a", C.c_float), ("affine_b", C.c_float), ("ip_qo_floor", C.c_float),
        ("search_gamma", C.c_float), ("gamma_max", C.c_float), ("gamma_beta", C.c_float),
        ("gamma_warmup", C.c_uint64), ("num_slack_levels", C.c_int32), ("slack_levels", C.c_float * 32),
    ]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "pops", "expansions", "exact_calls", "beam_pushes", "max_beam", "nn_pushes", "lb_skips",
        "gamma_terms", "msb_skipped", "estimated", "descent_dists", "overflow_retries", "kernel_launches")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class HostIndex(C.Structure):
    _u32pp = C.POINTER(C.POINTER(C.c_uint32))
    _fields_ = [
        ("D", C.c_uint32), ("bits", C.c_uint32), ("dim", C.c_uint32), ("n", C.c_uint64),
        ("search_data", C.c_void_p), ("rec_size", C.c_uint64), ("nb_off", C.c_uint32),
        ("raw", C.c_void_p), ("norm_sq", C.c_void_p), ("centroid", C.c_void_p), ("calibration", C.c_void_p),
        ("max_level", C.c_int32), ("entry_point", C.c_uint32), ("graph_entry_point", C.c_uint32),
        ("rotation_seed", C.c_uint64), ("n_layers", C.c_uint32),
        ("layer_nodes", _u32pp), ("layer_offs", _u32pp), ("layer_nbrs", _u32pp),
        ("layer_sizes", C.POINTER(C.c_uint32)),
    ]


# every symbol include/cphnsw_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "cphnsw_b200_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "cphnsw_b200_destroy": (None, [_P]),
    "cphnsw_b200_last_error": (C.c_char_p, [_P]),
    "cphnsw_b200_load": (C.c_int, [_P, C.c_char_p]),
    "cphnsw_b200_upload": (C.c_int, [_P, C.POINTER(HostIndex)]),
    "cphnsw_b200_get_info": (C.c_int, [_P, C.POINTER(Info)]),
    "cphnsw_b200_search_batch": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P, _P]),
    "cphnsw_b200_search_batch_device": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P, _P, _P]),
    "cphnsw_b200_search_batch_submit": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P, _P, C.POINTER(C.c_uint64)]),
    "cphnsw_b200_search_batch_wait": (C.c_int, [_P, C.c_uint64]),
    "cphnsw_b200_synchronize": (C.c_int, [_P]),
    "cphnsw_b200_last_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "cphnsw_b200_last_timings": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "cphnsw_b200_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "cphnsw_b200_prepare_queries": (C.c_int, [_P, _P, C.c_uint64, C.c_int, _P, _P, _P, _P, _P]),
    "cphnsw_b200_fastscan_blocks": (C.c_int, [_P, _P, _P, C.c_uint64, _P, _P, C.c_uint64, C.c_uint64, _P, _P,
                                              _P, _P, _P, _P, _P, _P, _P]),
    "cphnsw_b200_exact_l2": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_uint64, _P, _P]),
    "cphnsw_b200_greedy_descent": (C.c_int, [_P, _P, C.c_uint64, _P, _P]),
    "cphnsw_b200_exhaustive_search": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                                _P, _P, _P]),
    "cphnsw_b200_exhaustive_candidates": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _P, _P, _P, _P, _P, _P]),
    "cphnsw_b200_merge_candidates": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _P, _P, _P, _P]),
    "cphnsw_b200_exhaustive_estimates": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.c_uint64, _P, _P, _P]),
    "cphnsw_b200_unique_topk": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint64, C.c_uint64, _P, C.c_uint64, _P, _P, _P]),
    "cphnsw_b200_calibration_samples": (C.c_int, [_P, _P, _P, C.c_uint64, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "cphnsw_b200_neighbor_codes": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint64, _P, C.c_uint64, C.c_uint64, _P, _P,
                                             C.c_uint64, _P, _P, _P, C.c_uint64, _P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the native library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python rabitq-ann-search_b200/build.py` "
                "(cphnsw_b200 has no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(handle, rc: int):
    """Map a status code to the exception the reference's pybind layer would raise."""
    if rc == OK:
        return
    msg = lib().cphnsw_b200_last_error(handle)
    msg = msg.decode() if msg else f"cphnsw_b200 error {rc}"
    if rc == EINVAL:
        raise ValueError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
