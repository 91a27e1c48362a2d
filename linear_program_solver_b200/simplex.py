"""`Simplex` of the `lpsol` API, driving the B200 pivot engine.

Public surface, return values and error behaviour follow /root/reference/lpsol/simplex.py
(cited per method); the implementation is new.  The optimisation loop (`solve`) is ONE call
into the CUDA library: pricing, ratio test, rank-1 update, the stuck counter and the rule
switch all run on the device and only the status, basis, marks and pivot trace come back
(the boundary sits at `Simplex.solve`, not at `Tableau.pivot`).
"""
from __future__ import annotations

from typing import Dict, List, Tuple, Union

import numpy as np

from . import _native as N
from .tableau import Tableau

PivotResult = Union[Tuple[int, int], str]


class Simplex:
    """state of the simplex algorithm on a canonical-form tableau (simplex.py:16-34)

    The constructor finds an initial basic feasible solution (phase 1) and raises ValueError
    when the problem is infeasible, leaving the tableau as the optimum of the artificial
    problem.  The tableau is not copied.
    """

    #: `pivot_log` keeps at most this many pivots of ONE `solve()` call (the device trace buffer is sized up front);
    #: `last_pivots` is always exact.  Raise it before a solve whose whole pivot sequence is wanted.
    log_cap: int = 1 << 16

    def __init__(self, tab: Tableau, drop_redundant: bool = False):
        #: not in the reference: when phase 1 finds dependent constraint rows, drop them and go on instead of
        #: reproducing the reference's crash (simplex.py:93 re-installs the stale row count).  `LinProg` sets it.
        self._drop_redundant = bool(drop_redundant)
        self._tab: Tableau = tab
        self._bfs: List[int] = [-1] * tab.getNumCons()
        #: (row, column, rule) of every pivot made through this object; rule 0 = Dantzig,
        #: 1 = Bland, 2 = explicit.  Not in the reference (its `num_piv` is a dead local).
        self.pivot_log: List[Tuple[int, int, int]] = []
        #: pivots made by the latest `solve()` (which returns None like simplex.py:110)
        self.last_pivots: int = 0
        self._find_bfs()

    # ------------------------------------------------------------------------------------
    # phase 1 (simplex.py:36-108)
    # ------------------------------------------------------------------------------------
    def _find_bfs(self):
        tab = self._tab
        m, n = tab.getTableauSize()
        b = tab.getB()
        for i in range(m):                       # all b_i >= 0 (simplex.py:43-45)
            if b[i] < 0:
                tab.rowMult(i, -1)
        if tab.isCanonical(self._bfs):
            return
        # method of artificial variables, only for the rows that lack a basic column
        orig_c = tab.getC()[:]
        orig_z = tab.getZ()
        tab.setC([0] * n)
        tab.setZ(0)
        missing = [i for i, j in enumerate(self._bfs) if j == -1]
        tab.addVars([f"$a{i}" for i in missing])
        for ind, i in enumerate(missing):
            tab.setCj(n + ind, 1)
            tab.setAij(i, n + ind, 1)
            tab.rowSubFromObj(i, 1)
            self._bfs[i] = n + ind
        tab._expect_phase2(orig_c, orig_z)
        self.solve()
        z = tab.getZ()
        if z != 0:
            tab._cost_hint = None
            raise ValueError(f"infeasible problem, artificial opt = {z}")
        rowkeep = [True] * m
        for i in range(m):
            if self._bfs[i] < n:
                continue
            assert tab.getBi(i) == 0, "invalid artificial solution (internal error)"
            row = tab._row(i)
            j1 = next((j2 for j2 in range(n) if row[j2] != 0), -1)
            if j1 == -1:                          # dependent constraint
                rowkeep[i] = False
            else:                                 # the pivot element may be negative here
                self._pivot(i, j1)
        self._bfs = [j for i, j in enumerate(self._bfs) if rowkeep[i]]
        assert all(0 <= j < n for j in self._bfs), "invalid basic feasible solution (internal error)"
        # simplex.py:95-105.  With every row kept this is one device call (the tableau stays on the GPU); dropping a
        # dependent row goes through the host-side re-layout, like the reference's private-field surgery.
        if not (all(rowkeep) and tab._leave_phase1(n, orig_c, orig_z, self._bfs)):
            tab._delete(rowkeep, n)
            tab.setZ(orig_z)
            tab.setC(orig_c)
            for i, j in enumerate(self._bfs):
                tab.rowSubFromObj(i, tab.getCj(j))
        if not all(rowkeep) and not self._drop_redundant:
            self._fail_like_stale_row_count()
        assert tab.isCanonical(), "tableau not canonical (internal error)"
        for j in self._bfs:
            tab.setVarMark(j, True)

    def _fail_like_stale_row_count(self):
        """The reference re-installs the OLD row count after dropping dependent rows (simplex.py:93), so its closing
        `isCanonical` (simplex.py:106, tableau.py:466-496) walks that many rows over the fewer that are left: a
        zero-cost column whose scan for its 1 (tableau.py:482-485) or whose all-others-zero check (:487-488) runs past
        the last row dies with IndexError; when no column does, the dropped rows have no basic column and the assert
        fires.  Same exception here, on a tableau that is left sane (rows really dropped).  Always raises."""
        tab = self._tab
        a, c = tab.getA(), tab.getC()
        rows = range(len(a))
        if not any(bi < 0 for bi in tab.getB()):
            for j in range(len(c)):
                if c[j] != 0:
                    continue
                one = next((i for i in rows if a[i][j] == 1), -1)
                if one == -1 or all(i == one or a[i][j] == 0 for i in rows):
                    raise IndexError("list index out of range")
        raise AssertionError("tableau not canonical (internal error)")

    # ------------------------------------------------------------------------------------
    # the optimisation loop (simplex.py:110-148): one device call
    # ------------------------------------------------------------------------------------
    def solve(self, max_pivots: int = -1):
        """pivot to an optimal solution: most-negative-cost pivots, switching to min-index
        pivots when the objective seems stuck (simplex.py:110-148).

        `max_pivots` (not in the reference) stops after a prefix of pivots; the default runs to
        optimality.  Returns None like the reference; the number of pivots made is `self.last_pivots`
        (`pivot_log` records the first `log_cap` of them)."""
        tab = self._tab
        eng = tab._touch_dev()
        m, n = tab.getTableauSize()
        basis = np.asarray(self._bfs, dtype=np.int32)
        marks = np.asarray(tab._cm, dtype=np.int32)
        cap = max_pivots if max_pivots >= 0 else int(self.log_cap)
        status, npiv, trace = eng.solve(basis, marks, max_pivots=max_pivots, trace_cap=cap)
        self.last_pivots = int(npiv)
        self._bfs[:] = [int(j) for j in basis]
        tab._cm[:] = [bool(x) for x in marks]
        if trace is not None:
            self.pivot_log.extend((int(r), int(c), int(k)) for r, c, k in trace)
        assert status != N.UNBOUNDED, "unbounded artificial problem (internal error)"
        if status == N.STALLED:
            raise AssertionError("objective value increased (internal error)")
        assert status in (N.OPTIMAL, N.MAX_PIVOTS), "solver failed (internal error)"

    # ------------------------------------------------------------------------------------
    # results (simplex.py:150-190)
    # ------------------------------------------------------------------------------------
    def getBasicSequence(self) -> List[int]:
        return self._bfs

    def getBasicSequenceNames(self) -> List[str]:
        return [self._tab.getVarName(j) for j in self._bfs]

    def getBFS(self) -> Dict[int, object]:
        b = self._tab.getB()
        return {j: b[i] for i, j in enumerate(self._bfs)}

    def getObjValue(self):
        return self._tab.getZ()

    def getBFSNames(self) -> Dict[str, object]:
        b = self._tab.getB()
        return {name: b[i] for i, name in enumerate(self.getBasicSequenceNames())}

    # ------------------------------------------------------------------------------------
    # pivots (simplex.py:192-216)
    # ------------------------------------------------------------------------------------
    def _pivot(self, r: int, c: int, rule: int = 2):
        tab = self._tab
        tab.pivot(r, c)
        tab.setVarMark(self._bfs[r], False)
        self._bfs[r] = c
        tab.setVarMark(c, True)
        self.pivot_log.append((r, c, rule))

    def pivot(self, r: int, c: int):
        """pivot after checking that (r, c) passes the minimum ratio test (simplex.py:199-216)"""
        tab = self._tab
        col = [tab._out(v) for v in tab._col(c)]          # Python numbers: a zero a_rc raises ZeroDivisionError like
        b = [tab._out(v) for v in tab._col(tab.getNumVars())]   # the reference's `getBi(r)/getAij(r,c)` (simplex.py:213)
        minratio = None
        for i in range(tab.getNumCons()):
            if col[i] <= 0:
                continue
            ratio = b[i] / col[i]
            if minratio is None or ratio < minratio:
                minratio = ratio
        ratio = b[r] / col[r]
        if ratio != minratio:
            raise ValueError(f"bad pivot by min ratio test, r = {r}, c = {c}")
        self._pivot(r, c)

    # ------------------------------------------------------------------------------------
    # pivot rules
    # ------------------------------------------------------------------------------------
    def _find(self, rule: int, do_pivot: bool) -> PivotResult:
        found = self._tab._need_dev().find_pivot(rule)
        if isinstance(found, str):
            return found
        i, j = found
        if do_pivot:
            self._pivot(i, j, rule)
        return i, j

    def findPivotMinIndex(self, do_pivot: bool = False) -> PivotResult:
        """Bland's rule: first negative reduced cost, first row of minimum ratio (simplex.py:218-249)"""
        return self._find(N.RULE_BLAND, do_pivot)

    def findPivotStandard(self, do_pivot: bool = False) -> PivotResult:
        """most negative reduced cost (lowest index among equals), first row of minimum ratio
        (simplex.py:251-284)"""
        return self._find(N.RULE_DANTZIG, do_pivot)

    def findPivotMaxIncrease(self, do_pivot: bool = False) -> PivotResult:
        """pivot of maximum objective decrease (simplex.py:286-328).  Not on the solve path (SURVEY.md section 8f
        rank 3).  The ratio test of every column with c_j < 0 is one pass over the device tableau
        (`lps_find_pivot_max_increase`; exact mode: `lps_exact_col_ratios`, then one exact product per column)."""
        found = self._tab._need_dev().find_pivot_max_increase()
        if isinstance(found, str):
            return found
        if do_pivot:
            self._pivot(*found)
        return found

    def findPivotAll(self) -> List[Tuple[int, int]]:
        """every pivot that keeps canonical form (simplex.py:330-360): the ratio test of every column on the device
        (`lps_find_pivot_all` / `lps_exact_col_ratios`)"""
        return self._tab._need_dev().find_pivot_all()

    # ------------------------------------------------------------------------------------
    # text (simplex.py:362-379)
    # ------------------------------------------------------------------------------------
    def __str__(self) -> str:
        tab = self._tab
        b = tab.getB()
        names = [tab.getVarName(j) for j in self._bfs]
        vals = [tab._s(b[i]) for i in range(len(self._bfs))]
        width = [max(len(a), len(v)) for a, v in zip(names, vals)]
        names = [a.rjust(w) for a, w in zip(names, width)]
        vals = [v.rjust(w) for v, w in zip(vals, width)]
        return str(tab) + f"BFS: ({','.join(names)})\n" + f"   = ({','.join(vals)})\n"

    def __repr__(self) -> str:
        # the reference's repr has no closing bracket (simplex.py:377-379)
        return f"<{type(self).__name__} object at {hex(id(self))}, " \
               f"m = {self._tab._m}, n = {self._tab._n}"
