"""ctypes loader for oracle/pivot_oracle.c -- TEST INFRASTRUCTURE ONLY (see that file's header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpivot_oracle.so")
_lib = None

OK, OPTIMAL, UNBOUNDED, ZERO_PIVOT, MAX_PIVOTS = 0, 1, 2, 3, 4
OBJ_INCREASED = 6


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pivot_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_set_threads.argtypes = [ctypes.c_int]
        L.orc_pivot_f64.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_int, ctypes.c_int]
        L.orc_price_f64.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_int]
        L.orc_ratio_f64.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_int]
        L.orc_solve_f64.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ip, ip, ctypes.c_long,
                                    ctypes.POINTER(ctypes.c_long)]
        L.orc_pivot_modp.argtypes = [ctypes.POINTER(ctypes.c_uint32), ctypes.c_int, ctypes.c_int, ctypes.c_size_t,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32]
        L.orc_pivot_modp.restype = None
        L.orc_fill_dense_f64.argtypes = [dp, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_size_t,
                                         ctypes.c_ulonglong, ctypes.c_int, ctypes.c_int]
        L.orc_fill_dense_f64.restype = None
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def num_threads() -> int:
    return lib().orc_num_threads()


def set_threads(k: int) -> None:
    lib().orc_set_threads(int(k))


def pivot(T: np.ndarray, r: int, c: int) -> int:
    """In-place pivot of a C-contiguous float64 (m+1, ld) grid (ld >= n+1 = T.shape[1])."""
    m, n = T.shape[0] - 1, T.shape[1] - 1
    return lib().orc_pivot_f64(_dp(T), m, n, T.strides[0] // 8, r, c)


def price(T: np.ndarray, rule: int) -> int:
    m, n = T.shape[0] - 1, T.shape[1] - 1
    return lib().orc_price_f64(_dp(T), m, n, T.strides[0] // 8, rule)


def ratio(T: np.ndarray, c: int) -> int:
    m, n = T.shape[0] - 1, T.shape[1] - 1
    return lib().orc_ratio_f64(_dp(T), m, n, T.strides[0] // 8, c)


def solve(T: np.ndarray, basis: np.ndarray, max_pivots: int = -1, want_trace: bool = True):
    """Returns (status, npiv, trace[npiv,3])."""
    m, n = T.shape[0] - 1, T.shape[1] - 1
    cap = max_pivots if max_pivots >= 0 else 50 * (m + n) + 1000
    trace = np.zeros((cap, 3), dtype=np.int32) if want_trace else None
    npiv = ctypes.c_long(0)
    st = lib().orc_solve_f64(_dp(T), m, n, T.strides[0] // 8,
                             basis.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                             trace.ctypes.data_as(ctypes.POINTER(ctypes.c_int)) if want_trace else None,
                             cap if want_trace else max_pivots, ctypes.byref(npiv))
    return st, npiv.value, (trace[: npiv.value] if want_trace else None)


def pivot_modp(N: np.ndarray, r: int, c: int, q: int, dinv: int) -> None:
    m, n = N.shape[0] - 1, N.shape[1] - 1
    lib().orc_pivot_modp(N.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), m, n, N.strides[0] // 4, r, c, q, dinv)


def fill_dense(T: np.ndarray, m_total: int, row0: int = 0, seed: int = 0, lo: int = 1, hi: int = 9) -> None:
    """Fill T ((m_local+1, ld) float64, ld >= 2*m_total+1) with rows [row0, row0+m_local) of the hashed dense family."""
    m_local = T.shape[0] - 1
    lib().orc_fill_dense_f64(_dp(T), m_total, row0, m_local, 2 * m_total, T.strides[0] // 8, seed, lo, hi)
