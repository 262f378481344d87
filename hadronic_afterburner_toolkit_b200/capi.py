"""ctypes binding of libhbt_b200.so — the C ABI declared in include/hbt_b200.h.

The library is the product; this module only declares its entry points to Python (tests,
bench.py).  It fails loudly when the library is missing or cannot be loaded: there is no
Python or CPU implementation of the pair loops behind it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

from .params import CParams

HERE = os.path.dirname(os.path.abspath(__file__))
# HBT_B200_LIB: load another build of the same library (A/B runs of kernel variants)
LIB_PATH = os.environ.get("HBT_B200_LIB") or os.path.join(HERE, "libhbt_b200.so")

# every symbol include/hbt_b200.h declares
EXPORTS = (
    "hbt_create", "hbt_destroy", "hbt_last_error", "hbt_num_bins", "hbt_num_slabs", "hbt_reset",
    "hbt_gather_rapidity", "hbt_psi_ref", "hbt_rng_create", "hbt_rng_destroy", "hbt_rng_int_uniform",
    "hbt_rng_uniform", "hbt_rng_mixed_plan", "hbt_accumulate_same", "hbt_accumulate_mixed",
    "hbt_accumulate_batch", "hbt_accumulate_same_dev", "hbt_accumulate_mixed_dev", "hbt_accumulate_batch_dev",
    "hbt_synchronize",
    "hbt_read", "hbt_read_qinv", "hbt_get_stage_counters", "hbt_get_timers", "hbt_get_deferred_pairs",
    "hbt_get_launch_count", "hbt_measure_fp64_peak", "hbt_timer_start", "hbt_timer_stop",
    "hbt_comm_unique_id", "hbt_comm_init_rank", "hbt_comm_init_all", "hbt_allreduce", "hbt_allreduce_all",
    "hbt_version", "hbt_device_count", "hbt_set_option",
    "hbt_reader_open", "hbt_reader_next", "hbt_reader_error", "hbt_reader_bytes", "hbt_reader_close",
    "hbt_group_create", "hbt_group_destroy", "hbt_group_last_error", "hbt_group_size", "hbt_group_ctx",
    "hbt_group_accumulate_batch", "hbt_group_reduce", "hbt_group_ordered_batches",
    "hbt_cap_channels", "hbt_cap_get_counts", "hbt_cap_set_foreign",
    "hbt_bf_create", "hbt_bf_destroy", "hbt_bf_last_error", "hbt_bf_accumulate", "hbt_bf_read", "hbt_bf_get_timers",
)

HBT_OK = 0
ERRORS = {-1: "HBT_ERR_INVALID", -2: "HBT_ERR_CUDA", -3: "HBT_ERR_NO_DEVICE", -4: "HBT_ERR_CAP",
          -5: "HBT_ERR_OVERFLOW", -6: "HBT_ERR_NCCL", -7: "HBT_ERR_STATE"}


class HBTError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERRORS.get(code, code)}: {message}")
        self.code = code


def build(verbose: bool = False) -> None:
    """Compile libhbt_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libhbt_b200.so failed")


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `make -C {HERE}` (or __graft_entry__.build()); "
                          "there is no fallback implementation")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    pp = ctypes.POINTER(CParams)
    sig = {
        "hbt_create": (ctypes.c_int, [pp, i32, ctypes.POINTER(vp)]),
        "hbt_destroy": (None, [vp]),
        "hbt_last_error": (ctypes.c_char_p, [vp]),
        "hbt_num_bins": (i64, [vp]),
        "hbt_num_slabs": (i32, [vp]),
        "hbt_reset": (ctypes.c_int, [vp]),
        "hbt_gather_rapidity": (i64, [pp, vp, i64, vp]),
        "hbt_psi_ref": (dbl, [vp, i64, i32]),
        "hbt_rng_create": (ctypes.c_int, [i32, ctypes.POINTER(vp)]),
        "hbt_rng_destroy": (None, [vp]),
        "hbt_rng_int_uniform": (i32, [vp]),
        "hbt_rng_uniform": (dbl, [vp]),
        "hbt_rng_mixed_plan": (i32, [vp, i32, i32, vp, vp, vp]),
        "hbt_accumulate_same": (ctypes.c_int, [vp, vp, i64, dbl]),
        "hbt_accumulate_mixed": (ctypes.c_int, [vp, vp, vp, i32, vp, vp, i32, vp, vp, i32, dbl]),
        "hbt_accumulate_batch": (ctypes.c_int, [vp, vp, vp, i32, vp, vp, i32, vp, vp, i32, dbl, i32, i32]),
        "hbt_accumulate_same_dev": (ctypes.c_int, [vp, vp, i64, dbl]),
        "hbt_accumulate_mixed_dev": (ctypes.c_int, [vp, vp, vp, i32, vp, vp, i32, vp, vp, i32, dbl]),
        "hbt_accumulate_batch_dev": (ctypes.c_int, [vp, vp, vp, i32, vp, vp, i32, dbl]),
        "hbt_synchronize": (ctypes.c_int, [vp]),
        "hbt_read": (ctypes.c_int, [vp] * 9),
        "hbt_read_qinv": (ctypes.c_int, [vp] * 7),
        "hbt_get_stage_counters": (ctypes.c_int, [vp, vp, vp]),
        "hbt_get_timers": (ctypes.c_int, [vp, vp, vp, vp, vp]),
        "hbt_get_deferred_pairs": (ctypes.c_int, [vp, vp]),
        "hbt_get_launch_count": (ctypes.c_int, [vp, vp]),
        "hbt_timer_start": (ctypes.c_int, [vp]),
        "hbt_timer_stop": (ctypes.c_int, [vp, vp]),
        "hbt_measure_fp64_peak": (ctypes.c_int, [i32, dbl, vp]),
        "hbt_comm_unique_id": (ctypes.c_int, [vp]),
        "hbt_comm_init_rank": (ctypes.c_int, [vp, i32, i32, vp]),
        "hbt_comm_init_all": (ctypes.c_int, [vp, i32]),
        "hbt_allreduce": (ctypes.c_int, [vp]),
        "hbt_allreduce_all": (ctypes.c_int, [vp, i32]),
        "hbt_version": (ctypes.c_char_p, []),
        "hbt_device_count": (i32, []),
        "hbt_set_option": (ctypes.c_int, [vp, i32, i32]),
        "hbt_reader_open": (ctypes.c_int, [ctypes.c_char_p, i32, i32, i64, dbl, vp, vp]),
        "hbt_reader_next": (i32, [vp, vp, vp, vp]),
        "hbt_reader_error": (ctypes.c_char_p, [vp]),
        "hbt_reader_bytes": (ctypes.c_uint64, [vp]),
        "hbt_reader_close": (None, [vp]),
        "hbt_group_create": (ctypes.c_int, [pp, i32, vp, ctypes.POINTER(vp)]),
        "hbt_group_destroy": (None, [vp]),
        "hbt_group_last_error": (ctypes.c_char_p, [vp]),
        "hbt_group_size": (i32, [vp]),
        "hbt_group_ctx": (vp, [vp, i32]),
        "hbt_group_accumulate_batch": (ctypes.c_int, [vp, vp, vp, i32, vp, vp, i32, vp, vp, i32, dbl, i32, i32]),
        "hbt_group_reduce": (ctypes.c_int, [vp]),
        "hbt_group_ordered_batches": (ctypes.c_int, [vp, vp]),
        "hbt_cap_channels": (i32, [vp]),
        "hbt_cap_get_counts": (ctypes.c_int, [vp, vp, vp]),
        "hbt_cap_set_foreign": (ctypes.c_int, [vp, vp, vp]),
        "hbt_bf_create": (ctypes.c_int, [i32, dbl, i32, vp]),
        "hbt_bf_destroy": (None, [vp]),
        "hbt_bf_last_error": (ctypes.c_char_p, [vp]),
        "hbt_bf_accumulate": (ctypes.c_int, [vp, i32, vp, vp, i32, vp, vp, i32, vp, vp]),
        "hbt_bf_read": (ctypes.c_int, [vp, vp]),
        "hbt_bf_get_timers": (ctypes.c_int, [vp, vp, vp]),
    }
    for name in EXPORTS:
        f = getattr(L, name)  # AttributeError if the library lacks a declared symbol
        f.restype, f.argtypes = sig[name]
    _lib = L
    return L
