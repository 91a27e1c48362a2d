"""ctypes binding of the C-ABI declared in include/lps_b200.h.

There is no CPU fallback: if the CUDA library is missing or no CUDA device is
visible, the first use of the device path raises `NativeUnavailable` with the
reason.  The library is built in-tree by `__graft_entry__.build()` (or
`make -C linear_program_solver_b200/csrc`).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_longlong, c_size_t, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "liblps_b200.so")

OK, OPTIMAL, UNBOUNDED, ZERO_PIVOT, MAX_PIVOTS, BAD_PRIME, STALLED = 0, 1, 2, 3, 4, 5, 6
INFEASIBLE, REDUNDANT_ROW, NEED_ARTIFICIALS, NEED_PRIMES = 7, 8, 9, 10
E_INVALID, E_CUDA, E_NOMEM, E_STATE, E_PEER = -1, -2, -3, -4, -5
RULE_DANTZIG, RULE_BLAND = 0, 1


class NativeUnavailable(RuntimeError):
    """The CUDA extension cannot be used (not built, or no device).  Never silently ignored."""


class NativeError(RuntimeError):
    """A C-ABI call returned a negative status."""


_lib = None

# name -> (restype, argtypes); this table is also what tests compare with include/lps_b200.h
SIGNATURES = {
    "lps_version": (c_int, []),
    "lps_last_error": (c_char_p, []),
    "lps_device_count": (c_int, [POINTER(c_int)]),
    "lps_grid_ld": (c_size_t, [c_int]),
    "lps_create_f64": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, POINTER(c_void_p)]),
    "lps_destroy": (c_int, [c_void_p]),
    "lps_sync": (c_int, [c_void_p]),
    "lps_cand_bytes": (c_size_t, [c_int]),
    "lps_set_shard": (c_int, [c_void_p, c_longlong, c_longlong, c_int, c_int, c_void_p, c_void_p]),
    "lps_upload_f64": (c_int, [c_void_p, c_void_p, c_size_t]),
    "lps_download_f64": (c_int, [c_void_p, c_void_p, c_size_t]),
    "lps_read_row_f64": (c_int, [c_void_p, c_int, c_void_p]),
    "lps_read_col_f64": (c_int, [c_void_p, c_int, c_void_p]),
    "lps_grid_ptr": (c_void_p, [c_void_p]),
    "lps_fill_dense_f64": (c_int, [c_void_p, c_ulonglong, c_int, c_int]),
    "lps_rebind_grid_f64": (c_int, [c_void_p, c_void_p]),
    "lps_pivot": (c_int, [c_void_p, c_int, c_int]),
    "lps_row_scale_f64": (c_int, [c_void_p, c_int, c_double]),
    "lps_row_axpy_f64": (c_int, [c_void_p, c_int, c_int, c_double]),
    "lps_phase2_f64": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "lps_find_pivot": (c_int, [c_void_p, c_int, POINTER(c_int), POINTER(c_int)]),
    "lps_find_pivot_max_increase": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int)]),
    "lps_find_pivot_all": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, POINTER(c_longlong)]),
    "lps_is_canonical": (c_int, [c_void_p, c_void_p, POINTER(c_int)]),
    "lps_is_optimal": (c_int, [c_void_p, POINTER(c_int)]),
    "lps_is_infeasible": (c_int, [c_void_p, POINTER(c_int)]),
    "lps_solve": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_longlong, POINTER(c_longlong)]),
    "lps_solve_begin": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p, c_longlong]),
    "lps_step_select": (c_int, [c_void_p]),
    "lps_step_update": (c_int, [c_void_p]),
    "lps_step_block": (c_int, [c_void_p]),
    "lps_block_size": (c_int, [c_void_p]),
    "lps_set_block_size": (c_int, [c_void_p, c_int]),
    "lps_set_panel_ctas": (c_int, [c_void_p, c_int]),
    "lps_p2p_prepare": (c_int, [c_void_p, c_int]),
    "lps_p2p_export": (c_int, [c_void_p, c_void_p]),
    "lps_p2p_attach": (c_int, [c_void_p, c_longlong, c_longlong, c_int, c_int, c_void_p]),
    "lps_p2p_block": (c_int, [c_void_p, POINTER(c_void_p)]),
    "lps_p2p_attach_local": (c_int, [c_void_p, c_longlong, c_longlong, c_int, c_int, c_void_p]),
    "lps_solve_poll": (c_int, [c_void_p, POINTER(c_int), POINTER(c_longlong)]),
    "lps_solve_end": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, POINTER(c_longlong)]),
    "lps_exact_ld": (c_size_t, [c_int]),
    "lps_create_exact": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_void_p, POINTER(c_void_p)]),
    "lps_upload_exact": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_int]),
    "lps_upload_exact_weighted": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_int, c_void_p, c_int]),
    "lps_read_exact": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "lps_read_alive": (c_int, [c_void_p, c_void_p]),
    "lps_read_exact_cells": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lps_exact_col_ratios": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_longlong]),
    "lps_exact_info": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_size_t)]),
    "lps_exact_set_weights": (c_int, [c_void_p, c_void_p, c_int]),
    "lps_exact_working": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "lps_row_combine_exact": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int]),
    "lps_phase2_exact": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                 c_void_p, c_int]),
    "lps_exact_loop_state": (c_int, [c_void_p, c_int, POINTER(c_int), POINTER(c_longlong), POINTER(c_int), POINTER(c_int)]),
    "lps_batch_smem_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "lps_batch_plan": (c_int, [c_int, c_int, c_int, c_void_p, c_int, POINTER(c_int), POINTER(c_double)]),
    "lps_batch_create": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, POINTER(c_void_p)]),
    "lps_batch_destroy": (c_int, [c_void_p]),
    "lps_batch_upload": (c_int, [c_void_p, c_void_p]),
    "lps_batch_grids_ptr": (c_void_p, [c_void_p]),
    "lps_batch_run": (c_int, [c_void_p, c_longlong]),
    "lps_batch_sync": (c_int, [c_void_p]),
    "lps_batch_set_stuck_limit": (c_int, [c_void_p, c_longlong]),
    "lps_batch_set_capacity_check": (c_int, [c_void_p, c_int]),
    "lps_batch_download": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lps_batch_timing": (c_int, [c_void_p, c_void_p]),
    "lps_batch_info": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_size_t), POINTER(c_int), POINTER(c_double),
                               POINTER(c_longlong)]),
    "lps_prof_reset": (c_int, [c_void_p]),
    "lps_prof_read": (c_int, [c_void_p, POINTER(c_double), POINTER(c_longlong), POINTER(c_longlong)]),
}


def load():
    """dlopen the library and declare every signature.  Does not touch the GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeUnavailable(
            f"CUDA extension not built: {LIB_PATH} is missing. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc); there is no CPU fallback.")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # e.g. libcudart not found
        raise NativeUnavailable(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def lib():
    """The loaded library, after checking that a CUDA device is usable."""
    L = load()
    n = c_int(0)
    rc = L.lps_device_count(byref(n))
    if rc != OK or n.value <= 0:
        raise NativeUnavailable(
            "no usable CUDA device: " + (L.lps_last_error() or b"").decode(errors="replace")
            + " -- the simplex pivot engine has no CPU fallback")
    return L


def check(rc: int) -> int:
    """Raise on negative status codes; pass non-negative ones (they carry meaning) through."""
    if rc < 0:
        msg = (load().lps_last_error() or b"").decode(errors="replace")
        if rc == E_INVALID:
            raise ValueError(msg)
        if rc == E_NOMEM:
            raise MemoryError(msg)
        raise NativeError(f"lps status {rc}: {msg}")
    return rc


def ptr(t) -> int:
    """Device / host pointer of a torch tensor or numpy array as an int (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data
