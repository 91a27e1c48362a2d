"""Device engines: thin Python objects around one C-ABI handle each.

`F64Engine` owns the fp64 device grid of one tableau (or of one row shard of
it).  PyTorch is used only to allocate device / pinned buffers and to name the
stream; every arithmetic step of the path happens inside the CUDA library.
"""
from __future__ import annotations

from ctypes import byref, c_int, c_longlong, c_void_p, c_double
from typing import Optional, Tuple

import numpy as np

from . import _native as N

STATUS_NAME = {N.OPTIMAL: "optimal", N.UNBOUNDED: "unbounded", N.MAX_PIVOTS: "max_pivots", N.OK: "running",
               N.ZERO_PIVOT: "zero_pivot", N.BAD_PRIME: "bad_prime"}


def _torch():
    import torch
    return torch


# Closed engines (grid, stream, handle) wait here for the next tableau of the same shape: creating and destroying a
# handle costs milliseconds of driver calls (tens with gigabytes of pinned host memory mapped), more than a small solve,
# and an upload overwrites the whole grid anyway.  Bounded in count; grids above 256 MB are released and the handle is
# re-bound to a new one (`lps_rebind_grid_f64`); sharded, re-shaped (`phase2`) or profiled engines are destroyed.
_ENGINE_POOL: dict = {}
_ENGINE_POOL_MAX = 8
_ENGINE_POOL_BYTES = 256 << 20


def drain_engine_pool():
    """Destroy every pooled engine (tests, memory-sensitive callers)."""
    L = N.lib()
    for lst in _ENGINE_POOL.values():
        for grid, stream, h in lst:
            L.lps_destroy(h)
    _ENGINE_POOL.clear()


class F64Engine:
    """fp64 tableau on one GPU.  Grid layout: see include/lps_b200.h."""

    mode = "fp64"

    def __init__(self, m: int, n: int, device: Optional[int] = None):
        torch = _torch()
        self.L = N.lib()
        if device is None:
            device = torch.cuda.current_device()
        self.device = int(device)
        self.m, self.n = int(m), int(n)
        self.ld = int(self.L.lps_grid_ld(self.n))
        self.row0, self.m_total, self.rank, self.nranks = 0, self.m, 0, 1
        self.cand_send = self.cand_all = None
        self._pool_key = (self.device, self.m, self.n)
        pooled = _ENGINE_POOL.get(self._pool_key)
        if pooled:
            self.grid, self.stream, self.h = pooled.pop()
            if self.grid is None:                         # a large engine was pooled without its grid
                with torch.cuda.device(self.device):
                    self.grid = torch.zeros((self.m + 1, self.ld), dtype=torch.float64, device=f"cuda:{self.device}")
                    self.stream.wait_stream(torch.cuda.current_stream(self.device))
                N.check(self.L.lps_rebind_grid_f64(self.h, self.grid.data_ptr()))
            else:
                with torch.cuda.stream(self.stream):
                    self.grid.zero_()                     # a new engine starts from a zero grid
            return
        with torch.cuda.device(self.device):
            self.grid = torch.zeros((self.m + 1, self.ld), dtype=torch.float64, device=f"cuda:{self.device}")
            # A stream of our own, handed to the library by handle: torch's default stream has handle 0,
            # which the C-ABI reads as "make a private stream" -- collectives and CUDA events issued from
            # Python must be ordered on the very stream the kernels run on (use `with eng.on_stream():`).
            self.stream = torch.cuda.Stream(self.device)
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        h = c_void_p()
        N.check(self.L.lps_create_f64(self.m, self.n, self.device, self.grid.data_ptr(), self.stream.cuda_stream, byref(h)))
        self.h = h

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            key = getattr(self, "_pool_key", None)
            small = (self.m + 1) * self.ld * 8 <= _ENGINE_POOL_BYTES
            if key is not None and sum(len(v) for v in _ENGINE_POOL.values()) < _ENGINE_POOL_MAX:
                # large tableaux: the handle (small buffers, stream, state) is kept, the grid goes back to the allocator
                _ENGINE_POOL.setdefault(key, []).append((self.grid if small else None, self.stream, self.h))
            else:
                self.L.lps_destroy(self.h)
            self.h = None
            self.grid = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        N.check(self.L.lps_sync(self.h))

    def on_stream(self):
        """Context manager: make the engine's stream torch's current stream (for collectives, events)."""
        return _torch().cuda.stream(self.stream)

    # -- sharding ------------------------------------------------------------------------------
    def set_shard(self, row0: int, m_total: int, rank: int, nranks: int):
        self._pool_key = None
        torch = _torch()
        nb = int(self.L.lps_cand_bytes(self.n))
        dev = f"cuda:{self.device}"
        with self.on_stream():
            self.cand_send = torch.zeros(nb, dtype=torch.uint8, device=dev)
            self.cand_all = torch.zeros(nb * nranks, dtype=torch.uint8, device=dev) if nranks > 1 else self.cand_send
        N.check(self.L.lps_set_shard(self.h, row0, m_total, rank, nranks, self.cand_send.data_ptr(), self.cand_all.data_ptr()))
        self.row0, self.m_total, self.rank, self.nranks = row0, m_total, rank, nranks

    def p2p_export(self, nranks: int) -> np.ndarray:
        """64-byte CUDA IPC handle of this engine's exchange block, sized for `nranks` ranks (peer-store exchange,
        include/lps_b200.h)."""
        self._pool_key = None
        N.check(self.L.lps_p2p_prepare(self.h, int(nranks)))
        out = np.zeros(64, dtype=np.uint8)
        N.check(self.L.lps_p2p_export(self.h, out.ctypes.data))
        return out

    def p2p_attach(self, row0: int, m_total: int, rank: int, nranks: int, handles: np.ndarray):
        """Map every rank's exchange block (handles: uint8 [nranks, 64] in rank order) and switch to peer stores."""
        self._pool_key = None
        handles = np.ascontiguousarray(handles, dtype=np.uint8)
        assert handles.size == 64 * nranks
        N.check(self.L.lps_p2p_attach(self.h, row0, m_total, rank, nranks, handles.ctypes.data))
        self.row0, self.m_total, self.rank, self.nranks = row0, m_total, rank, nranks

    def p2p_block(self, nranks: int) -> int:
        """Device address of this engine's exchange block, sized for `nranks` ranks (ranks of one process:
        `p2p_attach_local`)."""
        from ctypes import c_void_p
        self._pool_key = None
        N.check(self.L.lps_p2p_prepare(self.h, int(nranks)))
        out = c_void_p()
        N.check(self.L.lps_p2p_block(self.h, byref(out)))
        return int(out.value)

    def p2p_attach_local(self, row0: int, m_total: int, rank: int, nranks: int, blocks):
        """Peer-store exchange between engines driven by one process; blocks: `p2p_block()` of every rank."""
        self._pool_key = None
        from ctypes import c_void_p
        arr = (c_void_p * nranks)(*[int(b) for b in blocks])
        N.check(self.L.lps_p2p_attach_local(self.h, row0, m_total, rank, nranks, arr))
        self.row0, self.m_total, self.rank, self.nranks = row0, m_total, rank, nranks

    # -- transfers -----------------------------------------------------------------------------
    def upload(self, host: np.ndarray):
        """host: float64 (m+1, >= n+1) array in grid layout (C-contiguous rows)."""
        assert host.dtype == np.float64 and host.shape[0] == self.m + 1 and host.shape[1] >= self.n + 1
        assert host.strides[1] == 8
        N.check(self.L.lps_upload_f64(self.h, host.ctypes.data, host.strides[0] // 8))
        self.sync()

    def download(self) -> np.ndarray:
        out = np.empty((self.m + 1, self.n + 1), dtype=np.float64)
        N.check(self.L.lps_download_f64(self.h, out.ctypes.data, self.n + 1))
        return out

    def read_row(self, i: int) -> np.ndarray:
        out = np.empty(self.n + 1, dtype=np.float64)
        N.check(self.L.lps_read_row_f64(self.h, i, out.ctypes.data))
        return out

    def read_col(self, j: int) -> np.ndarray:
        out = np.empty(self.m + 1, dtype=np.float64)
        N.check(self.L.lps_read_col_f64(self.h, j, out.ctypes.data))
        return out

    def fill_dense(self, seed: int = 0, lo: int = 1, hi: int = 9):
        """Generate this handle's rows of the hashed dense family (problems.hashed_dense_grid) in place."""
        N.check(self.L.lps_fill_dense_f64(self.h, seed, lo, hi))

    # -- hot path ------------------------------------------------------------------------------
    def pivot(self, r: int, c: int) -> int:
        return N.check(self.L.lps_pivot(self.h, r, c))

    def row_scale(self, r: int, mult: float):
        N.check(self.L.lps_row_scale_f64(self.h, r, float(mult)))

    def row_axpy(self, rd: int, rs: int, mult: float):
        N.check(self.L.lps_row_axpy_f64(self.h, rd, rs, float(mult)))

    def phase2(self, n_new: int, cost: np.ndarray, basis):
        """Simplex._find_bfs simplex.py:95-105 on the device (`lps_phase2_f64`): cut the columns from n_new on, install
        `cost` (n_new reduced costs, then the _z cell) priced out against `basis`."""
        cost = np.ascontiguousarray(cost, dtype=np.float64)
        bas = np.ascontiguousarray(basis, dtype=np.int32)
        assert cost.size == n_new + 1 and bas.size == self.m
        N.check(self.L.lps_phase2_f64(self.h, n_new, cost.ctypes.data, bas.ctypes.data))
        self.n = int(n_new)
        self._pool_key = None                            # the handle's width changed: not reusable for the old shape

    def find_pivot(self, rule: int):
        r, c = c_int(-1), c_int(-1)
        rc = N.check(self.L.lps_find_pivot(self.h, rule, byref(r), byref(c)))
        if rc == N.OPTIMAL:
            return "optimal"
        if rc == N.UNBOUNDED:
            return "unbounded"
        return r.value, c.value

    def find_pivot_max_increase(self):
        """Simplex.findPivotMaxIncrease (simplex.py:286-328) over the device tableau."""
        r, c = c_int(-1), c_int(-1)
        rc = N.check(self.L.lps_find_pivot_max_increase(self.h, byref(r), byref(c)))
        if rc == N.OPTIMAL:
            return "optimal"
        if rc == N.UNBOUNDED:
            return "unbounded"
        return r.value, c.value

    def has_unbounded_column(self) -> bool:
        """Tableau.isUnbounded (tableau.py:504-508): a column with c_j < 0 and no positive entry -- exactly when the
        max-increase pass reports 'unbounded' (simplex.py:319-320)."""
        return self.find_pivot_max_increase() == "unbounded"

    def find_pivot_all(self):
        """Simplex.findPivotAll (simplex.py:330-360) over the device tableau: list of (row, column)."""
        cap = self.n + self.m
        while True:
            rows, cols = np.empty(cap, dtype=np.int32), np.empty(cap, dtype=np.int32)
            count = c_longlong(0)
            N.check(self.L.lps_find_pivot_all(self.h, rows.ctypes.data, cols.ctypes.data, cap, byref(count)))
            if count.value <= cap:
                k = count.value
                return list(zip(rows[:k].tolist(), cols[:k].tolist()))
            cap = count.value

    def is_canonical(self, want_cols: bool = True, cols: Optional[np.ndarray] = None):
        """(canonical?, bcols).  The library leaves bcols untouched when some b_i < 0 (tableau.py:476-477): pass
        `cols` pre-filled to tell."""
        if cols is None:
            cols = np.full(self.m, -1, dtype=np.int32)
        flag = c_int(0)
        N.check(self.L.lps_is_canonical(self.h, cols.ctypes.data if want_cols else None, byref(flag)))
        return bool(flag.value), cols

    def is_optimal(self) -> bool:
        flag = c_int(0)
        N.check(self.L.lps_is_optimal(self.h, byref(flag)))
        return bool(flag.value)

    def is_infeasible(self) -> bool:
        """Tableau.isInfeasible (tableau.py:510-514) on the device."""
        flag = c_int(0)
        N.check(self.L.lps_is_infeasible(self.h, byref(flag)))
        return bool(flag.value)

    def solve(self, basis: np.ndarray, marks: Optional[np.ndarray] = None, max_pivots: int = -1,
              trace_cap: int = 0) -> Tuple[int, int, Optional[np.ndarray]]:
        """Run `Simplex.solve` on the device.  basis / marks are int32 arrays updated in place."""
        assert basis.dtype == np.int32 and basis.size == self.m_total
        trace = np.zeros((trace_cap, 3), dtype=np.int32) if trace_cap > 0 else None
        npiv = c_longlong(0)
        rc = N.check(self.L.lps_solve(self.h, max_pivots, basis.ctypes.data,
                                      marks.ctypes.data if marks is not None else None,
                                      trace.ctypes.data if trace is not None else None, trace_cap, byref(npiv)))
        k = npiv.value
        return rc, k, (trace[: min(k, trace_cap)] if trace is not None else None)

    # -- stepping (sharded driver) -----------------------------------------------------------------
    def solve_begin(self, basis: Optional[np.ndarray], marks: Optional[np.ndarray], max_pivots: int, trace_cap: int):
        self._pool_key = None                            # a stepped solve may be left open: not reusable
        N.check(self.L.lps_solve_begin(self.h, max_pivots, basis.ctypes.data if basis is not None else None,
                                       marks.ctypes.data if marks is not None else None, trace_cap))

    def step_select(self):
        N.check(self.L.lps_step_select(self.h))

    def step_update(self):
        N.check(self.L.lps_step_update(self.h))

    def step_block(self):
        """One block of the blocked loop (`lps_step_block`): up to `block_size()` pivots decided by the look-ahead
        panel, applied to the grid in one pass."""
        N.check(self.L.lps_step_block(self.h))

    def block_size(self) -> int:
        return int(self.L.lps_block_size(self.h))

    def set_block_size(self, k: int):
        self._pool_key = None                            # a tuned handle does not go back to the pool
        N.check(self.L.lps_set_block_size(self.h, int(k)))

    def set_panel_ctas(self, ctas: int):
        self._pool_key = None
        N.check(self.L.lps_set_panel_ctas(self.h, int(ctas)))

    def poll(self) -> Tuple[int, int]:
        st, k = c_int(0), c_longlong(0)
        N.check(self.L.lps_solve_poll(self.h, byref(st), byref(k)))
        return st.value, k.value

    def solve_end(self, basis: np.ndarray, marks: Optional[np.ndarray], trace_cap: int = 0):
        trace = np.zeros((trace_cap, 3), dtype=np.int32) if trace_cap > 0 else None
        npiv = c_longlong(0)
        rc = N.check(self.L.lps_solve_end(self.h, basis.ctypes.data, marks.ctypes.data if marks is not None else None,
                                          trace.ctypes.data if trace is not None else None, trace_cap, byref(npiv)))
        k = npiv.value
        return rc, k, (trace[: min(k, trace_cap)] if trace is not None else None)

    # -- profiling -----------------------------------------------------------------------------
    def prof_reset(self):
        self._pool_key = None                            # per-launch profiling stays on for the handle's life
        N.check(self.L.lps_prof_reset(self.h))

    def prof_read(self):
        ms, nu, na = c_double(0), c_longlong(0), c_longlong(0)
        N.check(self.L.lps_prof_read(self.h, byref(ms), byref(nu), byref(na)))
        return ms.value, nu.value, na.value
