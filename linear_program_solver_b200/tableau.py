"""`Tableau` of the `lpsol` API, with the data resident on a B200.

Public surface and error behaviour follow /root/reference/lpsol/tableau.py (cited per
method); the implementation is new.  The numbers live in one dense grid

    rows 0..m-1 constraints | row m reduced costs ;  columns 0..n-1 variables | column n rhs

whose corner (m, n) is the reference's `_z` field (minus the objective value).  A host copy
and a device copy exist; whichever was written last is authoritative and the other is
refreshed lazily, so a solve never round-trips the tableau through the host.

  * setters, resizing, permutation, JSON and printing are host-side data-structure edits;
  * `pivot`, the row operations, `isCanonical`, `isOptimal` and everything `Simplex` does run
    on the GPU through the C-ABI (include/lps_b200.h).  There is no CPU fallback: without the
    CUDA library or a device these calls raise `NativeUnavailable`.

Two arithmetic modes (BASELINE.json north_star):
  mode='exact' (default, the reference's semantics): entries are `fractions.Fraction`; the
      device holds a fraction-free integer grid as residues modulo 31-bit primes.
  mode='fp64': entries are Python floats; the device holds doubles and rounds exactly like
      the reference's control flow run on floats.

One deliberate deviation ("aliasing"): `getC/getB/getA` return snapshots,
not live references into the storage.
"""
from __future__ import annotations

import csv
import io
import json
from fractions import Fraction
from typing import Any, List, Optional

import numpy as np

MODES = ("exact", "fp64")


class Tableau:
    # ------------------------------------------------------------------------------------
    # construction (tableau.py:36-52)
    # ------------------------------------------------------------------------------------
    def __init__(self, m: int, n: int, mode: str = "exact", device: Optional[int] = None):
        if m <= 0:
            raise ValueError(f"need m > 0, provided m = {m}")
        if n <= 0:
            raise ValueError(f"need n > 0, provided n = {n}")
        if mode not in MODES:
            raise ValueError(f"mode must be one of {MODES}, provided {mode!r}")
        self._m, self._n = m, n
        self._mode = mode
        self._device = device
        self._grid = self._blank(m + 1, n + 1)
        self._cl: List[str] = [""] * n
        self._cm: List[bool] = [False] * n
        self._eng = None
        self._host_ok, self._dev_ok = True, False
        self._peek: dict = {}
        self._int_grid = None        # exact mode, from_grid(integer ndarray): uploaded as is, Fractions only on demand
        self._cost_hint = None       # exact mode, phase 1: the objective `_leave_phase1` will install (sizes the prime set)
        self.phase2_on_device = 0    # how many times `_leave_phase1` ran as a device call (tests look at it)
        self.row_ops_on_device = 0   # exact mode: row operations applied to the device residues (tests look at it)

    # -- number handling -----------------------------------------------------------------
    def _num(self, x):
        q = x if isinstance(x, Fraction) else Fraction(x)
        return float(q) if self._mode == "fp64" else q

    def _blank(self, rows: int, cols: int):
        if self._mode == "fp64":
            return np.zeros((rows, cols), dtype=np.float64)
        zero = Fraction(0)
        return [[zero] * cols for _ in range(rows)]

    def _out(self, v):
        return float(v) if self._mode == "fp64" else v

    @property
    def mode(self) -> str:
        return self._mode

    # -- host / device coherence ---------------------------------------------------------------
    def _engine(self):
        if self._eng is None:
            if self._mode == "fp64":
                from .engine import F64Engine
                self._eng = F64Engine(self._m, self._n, self._device)
            else:
                from .exact import ExactEngine
                self._eng = ExactEngine(self._m, self._n, self._device)
        return self._eng

    def _drop_engine(self):
        if self._eng is not None:
            self._eng.close()
        self._eng = None
        self._dev_ok = False

    def _need_host(self):
        if not self._host_ok:
            if self._int_grid is not None:           # still the tableau: every edit, host or device side, drops it
                self._grid = [[Fraction(x) for x in row] for row in self._int_grid.tolist()]
            else:
                self._grid = self._need_dev().download()
            self._int_grid = None
            self._host_ok = True

    def _touch_host(self):
        """About to edit the host copy: it becomes the only valid one."""
        self._need_host()
        self._dev_ok = False
        self._int_grid = None
        self._peek.clear()

    def _need_dev(self):
        eng = self._engine()
        if not self._dev_ok:
            src = self._grid if self._host_ok else self._int_grid
            if self._mode == "exact" and self._cost_hint is not None:
                eng.upload(src, cost_hint=self._cost_hint)
            else:
                eng.upload(src)
            self._dev_ok = True
        return eng

    def _touch_dev(self):
        """About to edit the device copy."""
        eng = self._need_dev()
        self._host_ok = False
        self._int_grid = None
        self._peek.clear()
        return eng

    _INT_LIMIT = 1 << 62

    def _int_host(self, *values):
        """The integer host copy (exact mode, `from_grid` with integers) when it is the tableau and every value of an
        edit is a small integer: the edit is applied to the array in place -- no Fraction is built for any entry -- and
        the device copy becomes stale.  None: the caller takes the general (Fraction) path."""
        if self._int_grid is None or self._host_ok:
            return None
        for v in values:
            if v.denominator != 1 or not -self._INT_LIMIT < v.numerator < self._INT_LIMIT:
                return None
        self._dev_ok = False
        self._peek.clear()
        return self._int_grid

    def _row(self, i: int):
        """Row i of the grid (n+1 values) without downloading the whole tableau."""
        if self._host_ok:
            return self._grid[i]
        if self._int_grid is not None:
            return [Fraction(x) for x in self._int_grid[i].tolist()]
        key = ("r", i)
        if key not in self._peek:
            self._peek[key] = self._need_dev().read_row(i)
        return self._peek[key]

    def _col(self, j: int):
        """Column j of the grid (m+1 values)."""
        if self._host_ok:
            if self._mode == "fp64":
                return self._grid[:, j]
            return [row[j] for row in self._grid]
        if self._int_grid is not None:
            return [Fraction(x) for x in self._int_grid[:, j].tolist()]
        key = ("c", j)
        if key not in self._peek:
            self._peek[key] = self._need_dev().read_col(j)
        return self._peek[key]

    # ------------------------------------------------------------------------------------
    # comparison (tableau.py:56-66)
    # ------------------------------------------------------------------------------------
    def __eq__(self, t) -> bool:
        if not isinstance(t, Tableau):
            raise TypeError(f"cannot compare to type {type(t)}")
        if (self._m, self._n) != (t._m, t._n) or self._cl != t._cl or self._cm != t._cm:
            return False
        self._need_host()
        t._need_host()
        for i in range(self._m + 1):
            for j in range(self._n + 1):
                if self._grid[i][j] != t._grid[i][j]:
                    return False
        return True

    __hash__ = None

    # ------------------------------------------------------------------------------------
    # getters (tableau.py:70-124); IndexError when out of range
    # ------------------------------------------------------------------------------------
    def getNumCons(self) -> int:
        return self._m

    def getNumVars(self) -> int:
        return self._n

    def getTableauSize(self):
        return self._m, self._n

    def _chk_i(self, i: int) -> int:
        if not -self._m <= i < self._m:
            raise IndexError("list index out of range")
        return i % self._m

    def _chk_j(self, j: int) -> int:
        if not -self._n <= j < self._n:
            raise IndexError("list index out of range")
        return j % self._n

    def getZ(self):
        """objective value: the corner cell stores -z (tableau.py:82-84)"""
        return self._out(-self._row(self._m)[self._n])

    def getC(self) -> list:
        row = self._row(self._m)
        return [self._out(row[j]) for j in range(self._n)]

    def getCj(self, j: int):
        return self._out(self._row(self._m)[self._chk_j(j)])

    def getB(self) -> list:
        col = self._col(self._n)
        return [self._out(col[i]) for i in range(self._m)]

    def getBi(self, i: int):
        return self._out(self._col(self._n)[self._chk_i(i)])

    def getA(self) -> list:
        self._need_host()
        return [[self._out(self._grid[i][j]) for j in range(self._n)] for i in range(self._m)]

    def getAij(self, i: int, j: int):
        i, j = self._chk_i(i), self._chk_j(j)
        return self._out(self._row(i)[j])

    def getVarNames(self) -> List[str]:
        return self._cl

    def getVarName(self, j: int) -> str:
        return self._cl[j]

    def getVarMarks(self) -> List[bool]:
        return self._cm

    def getVarMark(self, j: int) -> bool:
        return self._cm[j]

    # reference-private fields some callers peek at (read-only snapshots here)
    @property
    def _z(self):
        return self._out(self._row(self._m)[self._n])

    @property
    def _c(self):
        return self.getC()

    @property
    def _b(self):
        return self.getB()

    @property
    def _a(self):
        return self.getA()

    # ------------------------------------------------------------------------------------
    # setters (tableau.py:128-180): host-side edits
    # ------------------------------------------------------------------------------------
    def setZ(self, z):
        v = -self._num(z)
        if self._mode == "exact" and (G := self._int_host(v)) is not None:
            G[self._m, self._n] = v.numerator
            return
        self._touch_host()
        self._grid[self._m][self._n] = v

    @staticmethod
    def _converted_prefix(convert, count: int):
        """The reference's bulk setters assign element by element (tableau.py:131-153), so a bad or missing element
        leaves everything before it stored.  -> (values converted before the failure, the exception or None)."""
        vals = []
        try:
            for k in range(count):
                vals.append(convert(k))
        except Exception as e:
            return vals, e
        return vals, None

    def setC(self, c: List[Any]):
        vals, err = self._converted_prefix(lambda j: self._num(c[j]), self._n)
        if not vals:
            pass
        elif self._mode == "exact" and (G := self._int_host(*vals)) is not None:
            G[self._m, :len(vals)] = [v.numerator for v in vals]
        else:
            self._touch_host()
            for j, v in enumerate(vals):
                self._grid[self._m][j] = v
        if err is not None:
            raise err

    def setCj(self, j: int, cj):
        v = self._num(cj)
        j = self._chk_j(j)
        if self._mode == "exact" and (G := self._int_host(v)) is not None:
            G[self._m, j] = v.numerator
            return
        self._touch_host()
        self._grid[self._m][j] = v

    def setB(self, b: List[Any]):
        vals, err = self._converted_prefix(lambda i: self._num(b[i]), self._m)
        if vals:
            self._touch_host()
            for i, v in enumerate(vals):
                self._grid[i][self._n] = v
        if err is not None:
            raise err

    def setBi(self, i: int, bi):
        v = self._num(bi)
        i = self._chk_i(i)
        self._touch_host()
        self._grid[i][self._n] = v

    def setA(self, a: List[List[Any]]):
        m, n = self._m, self._n
        try:                                   # the usual case: every element converts, row by row
            rows = [[self._num(a[i][j]) for j in range(n)] for i in range(m)]
        except Exception:
            rows = None
        if rows is not None:
            if m and n:
                self._touch_host()
                for i in range(m):
                    self._grid[i][:n] = rows[i]
            return
        vals, err = self._converted_prefix(lambda k: self._num(a[k // n][k % n]), m * n)
        if vals:
            self._touch_host()
            for k, v in enumerate(vals):
                self._grid[k // n][k % n] = v
        raise err

    def setAij(self, i: int, j: int, aij):
        v = self._num(aij)
        i, j = self._chk_i(i), self._chk_j(j)
        if self._mode == "exact" and (G := self._int_host(v)) is not None:
            G[i, j] = v.numerator
            return
        self._touch_host()
        self._grid[i][j] = v

    def setVarNames(self, cl: List[str]):
        for j in range(self._n):
            self._cl[j] = cl[j]

    def setVarName(self, j: int, l: str):
        self._cl[j] = l

    def setVarMarks(self, cm: List[bool]):
        for j in range(self._n):
            self._cm[j] = cm[j]

    def setVarMark(self, j: int, m: bool):
        self._cm[j] = m

    def toggleVarMark(self, j: int):
        self._cm[j] = not self._cm[j]

    @classmethod
    def from_grid(cls, grid, mode: str = "exact", names: Optional[List[str]] = None, device: Optional[int] = None):
        """Build a tableau from a whole (m+1) x (n+1) grid in device layout (fast path for
        large inputs: a numpy float64 array in fp64 mode, integers/Fractions in exact mode)."""
        rows, cols = len(grid), len(grid[0])
        t = cls(rows - 1, cols - 1, mode, device)
        if mode == "fp64":
            t._grid = np.ascontiguousarray(grid, dtype=np.float64)
        elif isinstance(grid, np.ndarray) and grid.dtype.kind in "iu":
            # integer input goes to the device as it is; Fractions are only built if a host-side method asks
            t._int_grid = np.array(grid, dtype=np.int64, order="C")      # own copy: integer edits are applied to it in place
            t._grid = None
            t._host_ok = False
        else:
            # integer input (ints only, all below 2^62) takes the array route: nothing is converted entry by entry
            ints = None
            if all(type(x) is int for row in grid for x in row):
                try:
                    ints = np.array(grid, dtype=np.int64)
                except (OverflowError, ValueError):
                    ints = None
            if ints is not None and ints.ndim == 2 and int(np.abs(ints).max(initial=0)) < (1 << 62):
                t._int_grid = ints
                t._grid = None
                t._host_ok = False
            else:
                t._grid = [[x if type(x) is Fraction else Fraction(x) for x in row] for row in grid]
        if names:
            t._cl = list(names)
        return t

    # ------------------------------------------------------------------------------------
    # resizing / permutation / copy (tableau.py:184-250): host-side re-layout
    # ------------------------------------------------------------------------------------
    def _reshape_host(self, new_grid, m: int, n: int):
        self._cost_hint = None
        self._int_grid = None
        self._grid = new_grid
        self._m, self._n = m, n
        self._drop_engine()
        self._host_ok = True
        self._peek.clear()

    def addVar(self, v: str = ""):
        self.addVars([v])

    def addVars(self, vs: List[str]):
        k = len(vs)
        if self._mode == "exact" and (G := self._int_host()) is not None:
            m, n = self._m, self._n
            wide = np.zeros((m + 1, n + k + 1), dtype=np.int64)
            wide[:, :n] = G[:, :n]
            wide[:, n + k] = G[:, n]
            self._cl += list(vs)
            self._cm += [False] * k
            self._m, self._n = m, n + k
            self._drop_engine()
            self._cost_hint = None
            self._int_grid = wide
            return
        self._need_host()
        m, n = self._m, self._n
        new = self._blank(m + 1, n + k + 1)
        for i in range(m + 1):
            for j in range(n):
                new[i][j] = self._grid[i][j]
            new[i][n + k] = self._grid[i][n]
        self._cl += list(vs)
        self._cm += [False] * k
        self._reshape_host(new, m, n + k)

    def addCon(self):
        self._add_cons(1)

    def addCons(self, count: int):
        if count <= 0:
            raise ValueError(f"need count > 0, provided count = {count}")
        self._add_cons(count)

    def _add_cons(self, count: int):
        self._need_host()
        m, n = self._m, self._n
        new = self._blank(m + count + 1, n + 1)
        for i in range(m):
            for j in range(n + 1):
                new[i][j] = self._grid[i][j]
        for j in range(n + 1):
            new[m + count][j] = self._grid[m][j]
        self._reshape_host(new, m + count, n)

    def permuteRows(self, perm: List[int]):
        m = self._m
        if len(perm) != m or set(perm) != set(range(m)):
            raise ValueError(f"not a permutation of 0..{m-1}")
        self._touch_host()
        if self._mode == "fp64":
            self._grid[:m] = self._grid[list(perm)]
        else:
            self._grid[:m] = [self._grid[i] for i in perm]

    def permuteCols(self, perm: List[int]):
        n = self._n
        if len(perm) != n or set(perm) != set(range(n)):
            raise ValueError(f"not a permutation of 0..{n-1}")
        self._touch_host()
        if self._mode == "fp64":
            self._grid[:, :n] = self._grid[:, list(perm)]
        else:
            self._grid = [[row[j] for j in perm] + [row[n]] for row in self._grid]
        self._cl = [self._cl[j] for j in perm]
        self._cm = [self._cm[j] for j in perm]

    def copy(self) -> "Tableau":
        self._need_host()
        dup = Tableau(self._m, self._n, self._mode, self._device)
        dup._grid = self._grid.copy() if self._mode == "fp64" else [row[:] for row in self._grid]
        dup._cl = self._cl[:]
        dup._cm = self._cm[:]
        return dup

    def _expect_phase2(self, cost: List[Any], z):
        """Phase 1 is about to run on this tableau (simplex.py:50-61): remember the objective that `_leave_phase1`
        installs afterwards, so that the exact mode sizes its prime set for it at the upload."""
        self._cost_hint = [self._num(v) for v in cost] + [-self._num(z)] if self._mode == "exact" else None

    def _leave_phase1(self, ncols: int, cost: List[Any], z, basis: List[int]) -> bool:
        """What simplex.py:95-105 does when no row is dropped -- cut the columns from `ncols` on, setZ(z), setC(cost),
        rowSubFromObj(i, c_{basis[i]}) for every row -- as one device call, without moving the tableau through the
        host.  False (nothing done) when the device copy is not the valid one or the engine cannot do it."""
        hint, self._cost_hint = self._cost_hint, None
        if not self._dev_ok or self._eng is None or len(cost) != ncols:
            return False
        if self._mode == "fp64":
            vec = np.array([self._num(v) for v in cost] + [-self._num(z)], dtype=np.float64)
            self._eng.phase2(ncols, vec, basis)
        elif hint is None or not self._eng.phase2(ncols, basis):
            return False
        self._n = ncols
        self._cl = self._cl[:ncols]
        self._cm = self._cm[:ncols]
        self._host_ok = False
        self._int_grid = None
        self._peek.clear()
        self.phase2_on_device += 1
        return True

    def _delete(self, keep_rows: List[bool], ncols: int):
        """Drop constraint rows / trailing columns (what simplex.py:88-100 does through private fields)."""
        self._need_host()
        m, n = self._m, self._n
        rows = [i for i in range(m) if keep_rows[i]] + [m]
        new = self._blank(len(rows), ncols + 1)
        for ii, i in enumerate(rows):
            for j in range(ncols):
                new[ii][j] = self._grid[i][j]
            new[ii][ncols] = self._grid[i][n]
        self._cl = self._cl[:ncols]
        self._cm = self._cm[:ncols]
        self._reshape_host(new, len(rows) - 1, ncols)

    # ------------------------------------------------------------------------------------
    # math operations (tableau.py:254-308): on the device
    # ------------------------------------------------------------------------------------
    def _dev_row_op(self, rd: int, rs, mult) -> bool:
        """exact mode, tableau resident on the device: the row operation on the residues (`lps_row_combine_exact`).
        False when the host copy is the valid one, or the handle cannot hold the grown numbers -- the caller then
        edits the host copy, which re-encodes at the next device operation like a setter."""
        if self._host_ok or not self._dev_ok or self._eng is None:
            return False
        if not self._eng.row_combine(rd, rs, mult):
            return False
        self._peek.clear()
        self._int_grid = None
        self.row_ops_on_device += 1
        return True

    def _int_row_room(self, G, rd: int, rs, mult: int) -> bool:
        """Does row rd (+)= mult * row rs stay inside the integer host copy's range?"""
        top = abs(mult) * int(np.abs(G[rd] if rs is None else G[rs]).max(initial=0)) + (0 if rs is None else int(np.abs(G[rd]).max(initial=0)))
        return abs(mult) < self._INT_LIMIT and top < self._INT_LIMIT

    def _host_row_axpy(self, rd: int, rs: int, mult):
        """exact mode: row rd += mult * row rs.  On the device when the tableau lives there; else on the host copy."""
        if self._dev_row_op(rd, rs, mult):
            return
        if self._int_grid is not None and not self._host_ok and mult.denominator == 1 and \
                self._int_row_room(self._int_grid, rd, rs, mult.numerator):
            G = self._int_host()
            G[rd] += mult.numerator * G[rs]
            return
        self._touch_host()
        dst, src = self._grid[rd], self._grid[rs]
        for j in range(self._n + 1):
            dst[j] = dst[j] + mult * src[j]

    def rowMult(self, r: int, m):
        mult = self._num(m)
        if mult == 1:
            return
        r = self._chk_i(r)
        if self._mode == "fp64":
            self._touch_dev().row_scale(r, mult)
            return
        if self._dev_row_op(r, None, mult):
            return
        if self._int_grid is not None and not self._host_ok and mult.denominator == 1 and \
                self._int_row_room(self._int_grid, r, None, mult.numerator):
            G = self._int_host()
            G[r] *= mult.numerator
            return
        self._touch_host()
        row = self._grid[r]
        for j in range(self._n + 1):
            row[j] = row[j] * mult

    def rowDiv(self, r: int, d):
        d = self._num(d)
        if d == 0:
            raise ZeroDivisionError("cannot divide row by zero")
        self.rowMult(r, self._num(1) / d)

    def rowAdd(self, rd: int, rs: int, m: Any = 1):
        mult = self._num(m)
        if mult == 0:
            return
        rd, rs = self._chk_i(rd), self._chk_i(rs)
        if self._mode == "fp64":
            self._touch_dev().row_axpy(rd, rs, mult)
        else:
            self._host_row_axpy(rd, rs, mult)

    def rowSub(self, rd: int, rs: int, m: Any = 1):
        self.rowAdd(rd, rs, -m)            # negated before coercion, like tableau.py:278-280 (a str raises TypeError)

    def rowAddToObj(self, r: int, m: Any = 1):
        mult = self._num(m)
        if mult == 0:
            return
        r = self._chk_i(r)
        if self._mode == "fp64":
            self._touch_dev().row_axpy(self._m, r, mult)
        else:
            self._host_row_axpy(self._m, r, mult)

    def rowSubFromObj(self, r: int, m: Any = 1):
        self.rowAddToObj(r, -m)            # tableau.py:291-293

    def pivot(self, r: int, c: int):
        """simplex pivot on (r, c), assuming nothing about the tableau (tableau.py:295-308)"""
        r, c = self._chk_i(r), self._chk_j(c)
        from . import _native as N
        if self._touch_dev().pivot(r, c) == N.ZERO_PIVOT:
            raise ZeroDivisionError(f"zero pivot {r},{c}")

    # ------------------------------------------------------------------------------------
    # input / output (tableau.py:312-462): host-side
    # ------------------------------------------------------------------------------------
    def _s(self, v) -> str:
        return repr(float(v)) if self._mode == "fp64" else str(v)

    def loadFile(self, file: str):
        with open(file, "r") as f:
            self.loadJson(json.loads(f.read()))

    def saveFile(self, file: str):
        with open(file, "w") as f:
            f.write(json.dumps(self.saveJson(), separators=(",", ":")))

    def loadJson(self, data: dict):
        assert isinstance(data["m"], int) and data["m"] > 0
        assert isinstance(data["n"], int) and data["n"] > 0
        m, n = data["m"], data["n"]
        grid = self._blank(m + 1, n + 1)
        grid[m][n] = self._num(data["z"])          # the file stores the raw `_z` cell (tableau.py:354)
        for j in range(n):
            grid[m][j] = self._num(data["c"][j])
        for i in range(m):
            grid[i][n] = self._num(data["b"][i])
            for j in range(n):
                grid[i][j] = self._num(data["a"][i][j])
        self._cl = [str(data["cl"][j]) for j in range(n)]
        self._cm = [bool(data["cm"][j]) for j in range(n)]
        self._reshape_host(grid, m, n)

    def saveJson(self) -> dict:
        self._need_host()
        m, n, g = self._m, self._n, self._grid
        return {
            "m": m, "n": n,
            "z": self._s(g[m][n]),
            "c": [self._s(g[m][j]) for j in range(n)],
            "b": [self._s(g[i][n]) for i in range(m)],
            "a": [[self._s(g[i][j]) for j in range(n)] for i in range(m)],
            "cl": self._cl, "cm": self._cm,
        }

    def printGrid(self, labels: bool = True, rownums: bool = True, mpre: str = "(", msuf: str = ")"):
        self._need_host()
        m, n, g = self._m, self._n, self._grid
        lead = 2 if rownums else 1
        out = []
        if labels:
            out.append([""] * lead + [f"{mpre}{l}{msuf}" if self._cm[j] else l for j, l in enumerate(self._cl)])
        out.append(([""] if rownums else []) + [self._s(g[m][n])] + [self._s(g[m][j]) for j in range(n)])
        for i in range(m):
            out.append(([f"{i}"] if rownums else []) + [self._s(g[i][n])] + [self._s(g[i][j]) for j in range(n)])
        return out

    def printText(self, labels: bool = True, rownums: bool = False, spacing: int = 2, left: bool = False,
                  mpre: str = "(", msuf: str = ")") -> str:
        if spacing < 1:
            raise ValueError(f"spacing must be positive, provided {spacing}")
        grid = self.printGrid(labels, rownums, mpre, msuf)
        width = [max(len(row[k]) for row in grid) for k in range(len(grid[0]))]
        pad = (lambda s, w: s.ljust(w)) if left else (lambda s, w: s.rjust(w))
        grid = [[pad(s, width[k]) for k, s in enumerate(row)] for row in grid]
        gap = " " * spacing
        rule = "-" * (spacing * (len(grid[0]) + 2) + sum(width) + 3)
        head = 2 if labels else 1
        lead = 2 if rownums else 1
        lines = [rule]
        for k, row in enumerate(grid):
            if k == head:
                lines.append(rule)
            lines.append(f"|{gap}" + gap.join(row[:lead] + ["|"] + row[lead:]) + f"{gap}|")
        lines.append(rule)
        return "\n".join(lines) + "\n"

    def printLatex(self, labels: bool = True, rownums: bool = False, mpre: str = "(", msuf: str = ")") -> str:
        grid = self.printGrid(labels, rownums, mpre, msuf)
        head = 2 if labels else 1
        lines = ["\\begin{tabular}{" + ("|c" * head) + "|" + ("c" * self._n) + "|} \\hline"]
        for k, row in enumerate(grid):
            line = " & ".join(f"${s}$" if s else s for s in row) + " \\\\"
            if k < head or k == len(grid) - 1:
                line += "\\hline"
            lines.append(line)
        lines.append("\\end{tabular}")
        return "\n".join(lines) + "\n"

    def printCSV(self, labels: bool = True, rownums: bool = False, mpre: str = "(", msuf: str = ")") -> str:
        buf = io.StringIO()
        csv.writer(buf).writerows(self.printGrid(labels, rownums, mpre, msuf))
        return buf.getvalue()

    def __str__(self) -> str:
        return self.printText()

    def __repr__(self) -> str:
        return f"<{type(self).__name__} object at {hex(id(self))}, m = {self._m}, n = {self._n}>"

    # ------------------------------------------------------------------------------------
    # form checks (tableau.py:466-521)
    # ------------------------------------------------------------------------------------
    def isCanonical(self, bcols: Optional[List[int]] = None) -> bool:
        """canonical form?  bcols[i] receives the lowest basic column of row i or -1; it is left
        untouched when some b_i < 0 (tableau.py:476-477).  Runs on the device."""
        cols = np.full(self._m, -2, dtype=np.int32)          # -2: not written by the library
        ok, cols = self._need_dev().is_canonical(cols=cols)
        if bcols is not None and not (cols == -2).all():
            for i in range(self._m):
                bcols[i] = int(cols[i])
        return ok

    def isOptimal(self) -> bool:
        return self._need_dev().is_optimal()

    # The three predicates below are not used by Simplex.  isUnbounded reuses the device's
    # per-column ratio pass, isInfeasible is a row pass on the device, isDegenerate reads the right-hand side column.
    def isUnbounded(self) -> bool:
        """some column with c_j < 0 has no positive entry (tableau.py:504-508): the per-column ratio pass on the device"""
        return self._need_dev().has_unbounded_column()

    def isInfeasible(self) -> bool:
        """some row with b_i > 0 has no positive entry (tableau.py:510-514): one pass over the device tableau"""
        return self._need_dev().is_infeasible()

    def isDegenerate(self) -> bool:
        col = self._col(self._n)
        return any(col[i] == 0 for i in range(self._m))
