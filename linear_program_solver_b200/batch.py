"""Batched exact solves of small LPs: `Simplex(tab); solve()` for many tableaux in one kernel launch.

This is the node-LP engine of the branch-and-bound driver in `linprog.py` (BASELINE.json config 5).
The reference has no such thing -- its `LinProg` is an empty stub (linprog.py:383-393) -- but what a
B&B driver does per node is exactly the reference's `Simplex(tab)` constructor (phase 1,
simplex.py:36-108) followed by `solve()` (simplex.py:110-148); `lps_batch_*` (include/lps_b200.h) runs
that whole sequence for every LP of a batch on the GPU, one LP per thread-block cluster with the
multi-modular tableau in shared memory.

The host side here only SIZES the batch (how many artificial columns to reserve, how many primes the
Hadamard bound asks for) and DECODES results (mixed-radix digits -> Python ints -> Fractions).  No
pivot, pricing, ratio test or phase-1 arithmetic runs on the host; there is no CPU fallback.
"""
from __future__ import annotations

import math
from ctypes import byref, c_double, c_int, c_longlong, c_size_t, c_void_p
from fractions import Fraction
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _native as N
from .exact import MixedRadix, primes_below_2_31, primes_for_bits

MAX_PRIMES = 16

STATUS_NAME = {N.OPTIMAL: "optimal", N.UNBOUNDED: "unbounded", N.MAX_PIVOTS: "max_pivots", N.BAD_PRIME: "bad_prime",
               N.STALLED: "objective_increased", N.INFEASIBLE: "infeasible", N.REDUNDANT_ROW: "redundant_row",
               N.NEED_ARTIFICIALS: "need_artificials"}


# ------------------------------------------------------------------------------------------------
# sizing (no arithmetic of the path: only counts and norms)
# ------------------------------------------------------------------------------------------------
def missing_basic_rows(G: np.ndarray) -> np.ndarray:
    """[count, m] bool: rows that lack a basic column after the sign normalisation of simplex.py:43-45,
    i.e. the rows that get an artificial column (tableau.py:466-496 decides what a basic column is)."""
    cnt, m, n = G.shape[0], G.shape[1] - 1, G.shape[2] - 1
    body = G[:, :m, :n]
    # candidates: zero cost and exactly one nonzero in the column; only those are looked at more closely
    cand = (G[:, m, :n] == 0) & (np.count_nonzero(body, axis=1) == 1)
    b_idx, j_idx = np.nonzero(cand)
    cols = body[b_idx, :, j_idx]                                     # [candidates, m]
    hit = np.abs(cols).argmax(axis=1)                                # row of the single nonzero
    val = cols[np.arange(cols.shape[0]), hit]
    sign = np.where(G[b_idx, hit, n] < 0, -1, 1)                     # the row is negated when b_i < 0
    ok = val * sign == 1
    has = np.zeros((cnt, m), dtype=bool)
    has[b_idx[ok], hit[ok]] = True
    return ~has


def bound_bits(G: np.ndarray, missing: np.ndarray) -> int:
    """Bits of a Hadamard bound on every minor of every LP's extended matrix, the phase-1 objective
    (minus the sum of the rows that get an artificial, simplex.py:56-60) included."""
    m = G.shape[1] - 1
    A = np.abs(G).astype(np.float32) if np.abs(G).max(initial=0) < (1 << 20) else np.abs(G).astype(np.float64)
    body = A[:, :m, :]
    obj = A[:, m, :].astype(np.float64)
    if missing.any():                                                # |phase-1 objective| <= sum of |rows| with an artificial
        obj = np.maximum(obj, np.einsum("bij,bi->bj", body, missing.astype(body.dtype)).astype(np.float64))
    row_sq = np.concatenate([np.einsum("bij,bij->bi", body, body).astype(np.float64), (obj * obj).sum(axis=1)[:, None]], axis=1)
    col_sq = np.einsum("bij,bij->bj", body, body).astype(np.float64) + obj * obj
    row_sq = np.maximum(row_sq, 1.0)
    col_sq = np.maximum(col_sq, 1.0)
    order = min(row_sq.shape[1], col_sq.shape[1])

    def top_log_sum(sq):
        if sq.shape[1] > order:
            sq = -np.partition(-sq, order - 1, axis=1)[:, :order]
        return np.log2(sq).sum(axis=1) / 2

    bits = np.minimum(top_log_sum(row_sq), top_log_sum(col_sq)).max()
    return int(math.ceil(float(bits) * (1 + 1e-6))) + 3             # float32 norms: a relative 1e-6 covers their rounding


class BatchResult:
    """Results of one batch; numbers are decoded lazily into Fractions."""

    def __init__(self, m: int, n: int, status, npiv, basis, sign, digits, trace, codec: MixedRadix):
        self.m, self.n = m, n
        self.status = status                  # [count] lps status codes
        self.npiv = npiv                      # [count, 4]: phase-1, drive-out, phase-2 pivots, artificial columns
        self.basis = basis                    # [count, m]
        self._sign, self._digits, self._codec = sign, digits, codec
        self.trace = trace                    # [count, trace_cap, 3] or None
        self._ints: Dict[int, List[int]] = {}
        self.no_row_no_zero_cost = None       # [count] bool from `solve_batch`: all rows zero and no cost equal to 0

    def __len__(self) -> int:
        return len(self.status)

    def status_name(self, k: int) -> str:
        return STATUS_NAME.get(int(self.status[k]), str(int(self.status[k])))

    def integers(self, k: int) -> List[int]:
        """[N_0n, ..., N_(m-1)n, the z cell, D] of LP k as Python ints."""
        if k not in self._ints:
            self._ints[k] = self._codec.decode(self._sign[k], self._digits[k])
        return self._ints[k]

    def _one(self, k: int, idx: int) -> int:
        """Number `idx` of LP k alone (Horner over its mixed-radix digits, highest nonzero digit first)."""
        if k in self._ints:
            return self._ints[k][idx]
        dig = self._digits[k, :, idx].tolist()
        primes = self._codec.primes
        top = len(dig) - 1
        while top >= 0 and dig[top] == 0:
            top -= 1
        v = 0
        for j in range(top, -1, -1):
            v = v * primes[j] + dig[j]
        return -v if self._sign[k, idx] < 0 else v

    def rhs(self, k: int) -> List[Fraction]:
        v = self.integers(k)
        D = v[self.m + 1]
        return [Fraction(x, D) for x in v[: self.m]]

    def objective(self, k: int) -> Fraction:
        """`Simplex.getObjValue` (simplex.py:188-190): minus the corner cell."""
        return Fraction(-self._one(k, self.m), self._one(k, self.m + 1))

    def objectives(self) -> List[Fraction]:
        """`objective(k)` of every LP at once: Horner over the digit planes with numpy object arrays (one big-integer
        multiply-add per LP and digit inside numpy instead of a Python loop per LP)."""
        primes = self._codec.primes
        out = []
        for idx in (self.m, self.m + 1):
            dig = self._digits[:, :, idx]                         # [count, P]
            nz = np.nonzero(dig.any(axis=0))[0]
            acc = np.zeros(dig.shape[0], dtype=object)
            if nz.size:
                for j in range(int(nz[-1]), -1, -1):
                    acc = acc * primes[j] + dig[:, j].astype(object)
            neg = self._sign[:, idx] < 0
            acc[neg] = -acc[neg]
            out.append(acc)
        num, den = out
        return [Fraction(-int(a), int(b)) if b else Fraction(0) for a, b in zip(num, den)]

    def bfs(self, k: int) -> Dict[int, Fraction]:
        """`Simplex.getBFS` (simplex.py:165-172)."""
        b = self.rhs(k)
        return {int(j): b[i] for i, j in enumerate(self.basis[k])}

    def pivots(self, k: int) -> List[tuple]:
        """(row, column, rule) of every pivot of LP k in order; rule 2 marks a phase-1 drive-out pivot."""
        if self.trace is None:
            raise ValueError("the batch was run without a pivot trace")
        cnt = int(self.npiv[k, :3].sum())
        return [tuple(int(v) for v in row) for row in self.trace[k, : min(cnt, self.trace.shape[1])]]

    def raise_for_status(self, k: int):
        """The exception the reference's `Simplex(tab); solve()` raises for this outcome."""
        st = int(self.status[k])
        if st == N.OPTIMAL:
            return
        if st == N.INFEASIBLE:
            raise ValueError(f"infeasible problem, artificial opt = {self.objective(k)}")
        if st == N.REDUNDANT_ROW:
            # simplex.py:93 leaves the row count stale after dropping dependent rows; the closing isCanonical then runs
            # past the last row (IndexError) as soon as it scans a zero-cost column -- every kept row has one (its basis
            # column).  Only an LP whose rows are ALL zero keeps no row: then a cost of 0 still gives the IndexError,
            # and with none the scan touches nothing and the constructor's assert fires (simplex.py:106).
            if self.no_row_no_zero_cost is not None and self.no_row_no_zero_cost[k]:
                raise AssertionError("tableau not canonical (internal error)")
            raise IndexError("list index out of range")
        if st == N.UNBOUNDED:
            raise AssertionError("unbounded artificial problem (internal error)")
        if st == N.STALLED:
            raise AssertionError("objective value increased (internal error)")
        raise N.NativeError(f"batched solve ended with status {self.status_name(k)}")


# Closed batch handles wait here for the next batch of the same shape and prime set (a branch-and-bound driver and the
# benchmark ask for the same shapes again and again): creating and destroying one costs about 23 ms of driver calls
# next to a 16 ms kernel.  Every upload / run resets the per-LP state; the two knobs are set again by `solve_batch`.
_BATCH_POOL: dict = {}
_BATCH_POOL_MAX = 4


def drain_batch_pool():
    """Destroy every pooled batch handle."""
    L = N.lib()
    for lst in _BATCH_POOL.values():
        for h in lst:
            L.lps_batch_destroy(h)
    _BATCH_POOL.clear()


class BatchExactSolver:
    """`count` LPs of one shape on one GPU.  `solve(grids)` = `[Simplex(Tableau(grid)).solve() for grid in grids]`."""

    def __init__(self, count: int, m: int, n: int, primes: Sequence[int], art_cap: int = 0, cluster: int = 0,
                 trace_cap: int = 0, device: Optional[int] = None):
        import torch
        self.L = N.lib()
        self.device = int(torch.cuda.current_device() if device is None else device)
        self.count, self.m, self.n = int(count), int(m), int(n)
        self.art_cap, self.trace_cap = int(art_cap), int(trace_cap)
        self.primes = sorted(int(p) for p in primes)       # ascending: digit j of the mixed radix is < every later prime
        self.codec = MixedRadix(self.primes)
        self._pool_key = (self.device, self.count, self.m, self.n, tuple(self.primes), self.art_cap, int(cluster), self.trace_cap)
        pooled = _BATCH_POOL.get(self._pool_key)
        if pooled:
            self.h = pooled.pop()
            self.set_stuck_limit(-1)                      # a new solver starts from the defaults
            self.set_capacity_check(False)
            return
        arr = np.asarray(self.primes, dtype=np.uint32)
        h = c_void_p()
        N.check(self.L.lps_batch_create(self.count, self.m, self.n, self.art_cap, len(self.primes), arr.ctypes.data,
                                        int(cluster), self.trace_cap, self.device, None, byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            key = getattr(self, "_pool_key", None)
            if key is not None and sum(len(v) for v in _BATCH_POOL.values()) < _BATCH_POOL_MAX:
                self.L.lps_batch_sync(self.h)             # nothing of this batch may still be running on it
                _BATCH_POOL.setdefault(key, []).append(self.h)
            else:
                self.L.lps_batch_destroy(self.h)
        self.h = None
        self._keep = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        c, pl, sm, rc, ms, nl = c_int(0), c_int(0), c_size_t(0), c_int(0), c_double(0), c_longlong(0)
        N.check(self.L.lps_batch_info(self.h, byref(c), byref(pl), byref(sm), byref(rc), byref(ms), byref(nl)))
        return {"cluster": c.value, "primes": len(self.primes), "rows_per_cta": pl.value, "smem_bytes": sm.value,
                "resident_clusters": rc.value, "last_run_ms": ms.value, "launches": nl.value}

    def upload(self, grids):
        """grids: int32 [count, m+1, n+1], a C-contiguous numpy array or a torch tensor already on the device."""
        assert tuple(grids.shape) == (self.count, self.m + 1, self.n + 1)
        if isinstance(grids, np.ndarray):
            assert grids.dtype == np.int32 and grids.flags.c_contiguous
            N.check(self.L.lps_batch_upload(self.h, grids.ctypes.data))
        else:
            assert grids.is_contiguous() and grids.element_size() == 4
            N.check(self.L.lps_batch_upload(self.h, grids.data_ptr()))
        self._keep = grids                    # the copy is asynchronous

    def run(self, max_pivots: int = -1):
        N.check(self.L.lps_batch_run(self.h, max_pivots))

    def sync(self):
        N.check(self.L.lps_batch_sync(self.h))

    PHASES = ["price local", "barrier 1", "price fold", "ratio garner", "ratio local", "barrier 2", "ratio global",
              "pivot consts+row bcast", "barrier 3", "update", "inverse", "loop", "setup", "phase-1 glue", "results"]

    def timing(self, read: bool = True):
        """Profiling aid: arm the per-phase cycle counters (read=False) or fetch [count, 16] cycles of the last run."""
        if not read:
            self._pool_key = None                         # the counters stay armed for the handle's life
            N.check(self.L.lps_batch_timing(self.h, None))
            return None
        out = np.zeros((self.count, 16), dtype=np.int64)
        N.check(self.L.lps_batch_timing(self.h, out.ctypes.data))
        return out

    def set_capacity_check(self, on: bool = True):
        """The batch holds fewer primes than the static bound asks for: LPs whose basis grows too heavy for them end
        with NEED_PRIMES before the pivot that might not fit (include/lps_b200.h)."""
        N.check(self.L.lps_batch_set_capacity_check(self.h, 1 if on else 0))

    def set_stuck_limit(self, limit: int = -1):
        """Degenerate Dantzig pivots before the switch to Bland; -1 = the reference's m + n (simplex.py:123)."""
        N.check(self.L.lps_batch_set_stuck_limit(self.h, int(limit)))

    def download(self, want_numbers: bool = True) -> BatchResult:
        c, m, P = self.count, self.m, len(self.primes)
        status = np.empty(c, dtype=np.int32)
        npiv = np.empty((c, 4), dtype=np.int32)
        basis = np.empty((c, m), dtype=np.int32)
        sign = np.empty((c, m + 2), dtype=np.int8) if want_numbers else None
        digits = np.empty((c, P, m + 2), dtype=np.uint32) if want_numbers else None
        trace = np.zeros((c, self.trace_cap, 3), dtype=np.int32) if self.trace_cap > 0 else None
        N.check(self.L.lps_batch_download(self.h, status.ctypes.data, npiv.ctypes.data, basis.ctypes.data,
                                          sign.ctypes.data if want_numbers else None,
                                          digits.ctypes.data if want_numbers else None,
                                          trace.ctypes.data if trace is not None else None))
        return BatchResult(m, self.n, status, npiv, basis, sign, digits, trace, self.codec)

    def solve(self, grids: np.ndarray, max_pivots: int = -1) -> BatchResult:
        self.upload(grids)
        self.run(max_pivots)
        return self.download()


def plan_batch(grids: np.ndarray, banned: Sequence[int] = (), device: Optional[int] = None, on_device: bool = True,
               dev_ptr: Optional[int] = None):
    """(art_cap, primes, bound bits) for a batch of integer tableaux, or raises ValueError when the Hadamard bound
    needs more primes than the batched kernel holds (solve those LPs one at a time).  The counting runs on
    the device (`lps_batch_plan`); `on_device=False` uses the numpy statement of the same rule (tests compare them)."""
    G = np.ascontiguousarray(grids, dtype=np.int32)
    if on_device:
        import torch
        L = N.lib()
        dev = int(torch.cuda.current_device() if device is None else device)
        k, b = c_int(0), c_double(0.0)
        # dev_ptr: the same grids already on the device (solve_batch uploads them once for planning and solving)
        N.check(L.lps_batch_plan(G.shape[0], G.shape[1] - 1, G.shape[2] - 1, dev_ptr if dev_ptr else G.ctypes.data, dev,
                                 byref(k), byref(b)))
        art_cap, bits = k.value, int(math.ceil(b.value * (1 + 1e-9))) + 3
    else:
        missing = missing_basic_rows(G)
        art_cap = int(missing.sum(axis=1).max(initial=0))
        bits = bound_bits(G, missing)
    primes = primes_for_bits(bits, banned)
    if len(primes) & 1:                      # the kernel is instantiated for even prime counts
        primes = primes_below_2_31(len(primes) + 1, banned)
    if len(primes) > MAX_PRIMES:
        raise ValueError(f"Hadamard bound of {bits} bits needs {len(primes)} primes; the batched kernel holds {MAX_PRIMES}")
    return art_cap, primes, bits


def solve_batch(grids, trace_cap: int = 0, cluster: int = 0, device: Optional[int] = None, max_pivots: int = -1,
                stuck_limit: int = -1, two_tier: bool = True) -> BatchResult:
    """`Simplex(tab); solve()` of the reference for every integer tableau of `grids`
    ([count, m+1, n+1], device layout: row m = costs, column n = rhs, cell (m, n) = minus the objective)."""
    G = np.asarray(grids)
    if G.ndim != 3:
        raise ValueError("grids must be [count, m+1, n+1]")
    if G.dtype != np.int32:
        if G.size and (G.max() >= 2 ** 31 or G.min() < -(2 ** 31)):
            raise ValueError("batched tableaux take 32-bit integer coefficients")
    G32 = np.ascontiguousarray(G, dtype=np.int32)
    import torch
    dev = torch.device("cuda", int(torch.cuda.current_device() if device is None else device))
    Gd = torch.from_numpy(G32).to(dev)       # ONE host-to-device copy serves the planner and every tier
    torch.cuda.current_stream(dev).synchronize()          # the library works on its own streams
    art_cap, primes, _ = plan_batch(G32, device=device, dev_ptr=Gd.data_ptr())
    primes = sorted(primes)
    if max_pivots < 0:                       # Bland's rule terminates; the cap only bounds a run gone wrong
        max_pivots = 64 * (G.shape[1] + G.shape[2]) + 256
    m, n = G.shape[1] - 1, G.shape[2] - 1

    def run(grids32, prime_set, check):
        solver = BatchExactSolver(grids32.shape[0], m, n, prime_set, art_cap=art_cap, cluster=cluster,
                                  trace_cap=trace_cap, device=device)
        try:
            solver.set_stuck_limit(stuck_limit)
            solver.set_capacity_check(check)
            return solver.solve(grids32, max_pivots)
        finally:
            solver.close()

    # Two tiers.  The static bound covers a basis made of the heaviest columns; node LPs rarely get there.  A first
    # pass on HALF the primes (a 64 x 128 LP then sits in one CTA instead of a cluster of two, twice as many LPs are
    # resident) solves every LP whose basis stays light -- each LP checks that rigorously before every pivot -- and
    # the few that end with NEED_PRIMES are solved again on the full set.
    short = len(primes) // 2
    short += short & 1                       # the kernel takes an even number of primes
    if two_tier and short >= 2 and short < len(primes) and G.shape[0] >= 2:
        first = run(Gd, primes[:short], True)
        redo = np.nonzero(first.status == N.NEED_PRIMES)[0]
        second = None
        if redo.size:
            Gr = Gd[torch.from_numpy(redo).to(dev)].contiguous()
            torch.cuda.current_stream(dev).synchronize()
            second = run(Gr, primes, False)
        out = _merge_tiers(first, second, redo, MixedRadix(primes))
    else:
        out = run(Gd, primes, False)
    if (out.status == N.REDUNDANT_ROW).any():             # which exception `raise_for_status` owes the caller
        out.no_row_no_zero_cost = ~G32[:, :m, :].any(axis=(1, 2)) & G32[:, m, :n].all(axis=1)
    return out


def _merge_tiers(first: "BatchResult", second: Optional["BatchResult"], redo: np.ndarray, codec: MixedRadix) -> "BatchResult":
    """Results of the short-prime pass with the re-run LPs patched in; digits padded to the full prime set (the short
    set is a prefix of the ascending prime list, so the mixed-radix digits agree and the upper ones are 0)."""
    P = len(codec.primes)
    digits = np.zeros((first._digits.shape[0], P, first._digits.shape[2]), dtype=np.uint32)
    digits[:, : first._digits.shape[1]] = first._digits
    status, npiv, basis, sign = first.status.copy(), first.npiv.copy(), first.basis.copy(), first._sign.copy()
    trace = None if first.trace is None else first.trace.copy()
    if second is not None:
        status[redo], npiv[redo], basis[redo], sign[redo] = second.status, second.npiv, second.basis, second._sign
        digits[redo] = second._digits
        if trace is not None:
            trace[redo] = second.trace
    out = BatchResult(first.m, first.n, status, npiv, basis, sign, digits, trace, codec)
    out.tiers = (int(first.status.shape[0]), int(redo.size))     # LPs in the first pass, LPs solved again
    return out

# ------------------------------------------------------------------------------------------------
# several GPUs: node LPs are independent, so they shard by LP with no data-path collective
# ------------------------------------------------------------------------------------------------
def shard_lps(count: int, rank: int, nranks: int) -> tuple:
    """[lo, hi) of the LPs rank `rank` of `nranks` solves: contiguous, balanced to within one LP."""
    return count * rank // nranks, count * (rank + 1) // nranks


def solve_batch_sharded(grids, rank: int, nranks: int, device: Optional[int] = None, **kw):
    """This rank's share of `solve_batch(grids)`: returns (lo, hi, BatchResult of LPs lo..hi-1).  One process per GPU;
    the caller gathers whatever it needs of the results (statuses, objectives) with its own transport."""
    G = np.asarray(grids)
    lo, hi = shard_lps(G.shape[0], rank, nranks)
    if hi <= lo:
        return lo, hi, None
    return lo, hi, solve_batch(G[lo:hi], device=device, **kw)
