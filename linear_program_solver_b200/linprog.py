"""`LinProg`: a linear (mixed-integer) program in general form, solved by branch and bound on the GPU.

The reference's `LinProg` is an empty stub (/root/reference/lpsol/linprog.py:383-393: a docstring and an
empty `__init__`), so there is no reference behaviour to reproduce below the constructor.
What is kept from the reference: the class name and its zero-argument constructor, and the model-layer
primitives the lowering is written with -- `LinExpr.substitute` (linprog.py:198-222) for the bound
shifts, `LinCon.simplify` (linprog.py:261-268) to bring a constraint to "variables left, constant
right", `LinVar.boundAbove/boundBelow` (linprog.py:338-352, integer bounds are floored / ceiled) for
branching and `LinVar.isFeasible` (linprog.py:354-356) to drop empty boxes.

Everything new is design:
  * lowering to the tableau the reference's `Simplex` accepts (non-negative columns, `<=` rows with a unit
    slack each, `==` rows without, integer coefficients so that the batched exact kernel can take it);
  * a deterministic breadth-first branch and bound whose node LPs -- each one a full `Simplex(tab); solve()`
    with phase 1 -- are solved MANY AT A TIME by `batch.solve_batch` (one LP per thread-block cluster);
    node LPs too large for the batched kernel go through `Tableau`/`Simplex` in exact mode one by one.
No arithmetic of a node LP runs on the host; there is no CPU fallback.  Parity: `oracle/bnb_oracle.py`
restates this driver on the CPU with every node LP solved by the pinned simplex oracle ("parity
unpinned by the reference" for the MILP layer as a whole).
"""
from __future__ import annotations

import math
from fractions import Fraction
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .model import LinCon, LinExpr, LinVar  # noqa: F401  (re-exported like the reference module)

Bounds = Dict[str, Tuple[Optional[Fraction], Optional[Fraction]]]


def _lcm_den(values) -> int:
    s = 1
    for v in values:
        d = Fraction(v).denominator
        s = s * d // math.gcd(s, d)
    return s


class StandardForm:
    """One node LP lowered to an integer tableau.

    grid        (m+1) x (n+1) Python ints in the device layout (row m = costs, column n = rhs)
    columns     per tableau column < n_struct: (variable name, kind) with kind '+': x = lb + col,
                '-': x = ub - col, 'p'/'n': positive / negative part of a free variable
    obj_scale   positive integer the cost row was multiplied by
    obj_const   constant of the objective (after the bound shifts), in the user's sense
    sign        +1 for a minimisation, -1 for a maximisation (the tableau always minimises)
    """

    def __init__(self, grid, columns, shifts, obj_scale, obj_const, sign, names):
        self.grid, self.columns, self.shifts = grid, columns, shifts
        self.obj_scale, self.obj_const, self.sign, self.names = obj_scale, obj_const, sign, names
        self.m, self.n = len(grid) - 1, len(grid[0]) - 1

    def values(self, bfs: Dict[int, Fraction]) -> Dict[str, Fraction]:
        """Variable values from a basic solution {column: value} of the tableau."""
        out: Dict[str, Fraction] = {}
        for j, (name, kind) in enumerate(self.columns):
            v = bfs.get(j, Fraction(0))
            lb, ub = self.shifts[name]
            if kind == "+":
                out[name] = lb + v
            elif kind == "-":
                out[name] = ub - v
            elif kind == "p":
                out[name] = out.get(name, Fraction(0)) + v
            else:
                out[name] = out.get(name, Fraction(0)) - v
        return out

    def objective(self, tableau_obj: Fraction) -> Fraction:
        """The user's objective from the tableau's (`Simplex.getObjValue`)."""
        return self.sign * (tableau_obj / self.obj_scale) + self.obj_const


class LinProgResult:
    def __init__(self, status: str, objective=None, values=None, nodes: int = 0, lps: int = 0, batches: int = 0):
        self.status = status            # 'optimal' | 'infeasible'
        self.objective = objective      # Fraction or None
        self.values = values or {}      # variable name -> Fraction
        self.nodes = nodes              # branch-and-bound nodes processed
        self.lps = lps                  # node LPs solved on the device
        self.batches = batches          # kernel batches launched
        self.trace: List[tuple] = []    # (depth, status, objective) per processed node, in order

    def __repr__(self) -> str:
        return f"LinProgResult({self.status!r}, objective={self.objective}, nodes={self.nodes}, lps={self.lps})"


class LinProg:
    """representation of a linear program in possibly non standard form (linprog.py:383-393)"""

    def __init__(self):
        self.vars: Dict[str, LinVar] = {}
        self.cons: List[LinCon] = []
        self.obj: LinExpr = LinExpr()
        self.sense: str = "min"

    # -- building ------------------------------------------------------------------------------
    def addVar(self, var, integral: bool = False, lb=0, ub=None) -> LinVar:
        """Declare a variable: a `LinVar`, or a name with bounds (default: continuous, x >= 0)."""
        if not isinstance(var, LinVar):
            var = LinVar(var, integral, lb, ub)
        if var.x in self.vars:
            raise ValueError(f"variable {var.x} declared twice")
        self.vars[var.x] = var
        return var

    def addConstraint(self, con: LinCon) -> LinCon:
        if not isinstance(con, LinCon):
            raise TypeError("constraint must be a LinCon (LinExpr.constraintLeq / constraintGeq / constrantEq)")
        if con.comp not in ("<=", ">=", "=="):
            raise ValueError(f"unknown comparison {con.comp!r}")
        self.cons.append(con)
        return con

    def setObjective(self, expr, sense: str = "min"):
        if sense not in ("min", "max"):
            raise ValueError("sense must be 'min' or 'max'")
        self.obj = expr if isinstance(expr, LinExpr) else LinExpr(expr)
        self.sense = sense

    def _declare_missing(self):
        """Variables that only appear in expressions are continuous and non-negative."""
        seen = list(self.obj.m)
        for con in self.cons:
            seen += list(con.left.m) + list(con.right.m)
        for name in seen:
            if name not in self.vars:
                self.vars[name] = LinVar(name, False, 0, None)

    def bounds(self) -> Bounds:
        return {name: v.getBounds() for name, v in self.vars.items()}

    # -- lowering ------------------------------------------------------------------------------
    def lower(self, bounds: Optional[Bounds] = None) -> Optional[StandardForm]:
        """The LP relaxation over the box `bounds` (default: the declared bounds) as an integer tableau,
        or None when the box is empty."""
        self._declare_missing()
        bounds = bounds or self.bounds()
        names = list(self.vars)
        columns: List[Tuple[str, str]] = []
        subst: Dict[str, LinExpr] = {}
        shifts: Dict[str, Tuple[Fraction, Fraction]] = {}
        ub_rows: List[Tuple[int, Fraction]] = []
        for name in names:
            lb, ub = bounds[name]
            if lb is not None and ub is not None and lb > ub:        # LinVar.isFeasible (linprog.py:354-356)
                return None
            col = f"c{len(columns)}"
            if lb is not None:
                subst[name] = LinExpr(1, col, lb)                    # x = lb + x'
                shifts[name] = (lb, ub if ub is not None else Fraction(0))
                columns.append((name, "+"))
                if ub is not None:
                    ub_rows.append((len(columns) - 1, ub - lb))      # x' <= ub - lb
            elif ub is not None:
                subst[name] = LinExpr(-1, col, ub)                   # x = ub - x'
                shifts[name] = (Fraction(0), ub)
                columns.append((name, "-"))
            else:
                col2 = f"c{len(columns) + 1}"
                subst[name] = LinExpr(1, col, -1, col2)              # x = x+ - x-
                shifts[name] = (Fraction(0), Fraction(0))
                columns.append((name, "p"))
                columns.append((name, "n"))
        ns = len(columns)
        rows: List[Tuple[List[Fraction], Fraction, bool]] = []       # (coefficients, rhs, has_slack)
        for con in self.cons:
            c2 = LinCon(con.left.substitute(subst), con.comp, con.right.substitute(subst)).simplify()
            coefs = [c2.left.getCoefficient(f"c{j}") for j in range(ns)]
            rhs = c2.right.getConstant() - c2.left.getConstant()
            if con.comp == ">=":                                     # a.x >= b  <=>  -a.x <= -b
                coefs, rhs = [-a for a in coefs], -rhs
            rows.append((coefs, rhs, con.comp != "=="))
        for j, width in ub_rows:
            rows.append(([Fraction(1) if k == j else Fraction(0) for k in range(ns)], width, True))
        if not rows:
            raise ValueError("a LinProg needs at least one constraint or one two-sided bound")
        nslack = sum(1 for r in rows if r[2])
        m, n = len(rows), ns + nslack
        grid = []
        k = 0
        for coefs, rhs, has_slack in rows:
            s = _lcm_den(coefs + [rhs])                              # integer row; the slack joins afterwards with a unit entry
            line = [int(a * s) for a in coefs] + [0] * nslack + [int(rhs * s)]
            if has_slack:
                line[ns + k] = 1
                k += 1
            grid.append(line)
        obj = self.obj.substitute(subst)
        sign = 1 if self.sense == "min" else -1
        cost = [sign * obj.getCoefficient(f"c{j}") for j in range(ns)]
        s = _lcm_den(cost)
        grid.append([int(a * s) for a in cost] + [0] * nslack + [0])
        return StandardForm(grid, columns, shifts, s, obj.getConstant(), sign, names)

    # -- node LPs on the device -------------------------------------------------------------------
    @staticmethod
    def _solve_nodes(forms: Sequence[StandardForm], device=None, stats=None):
        """`Simplex(tab); solve()` for every lowered node: [(status, tableau objective, bfs)], batched by shape."""
        from . import _native as N
        from .batch import solve_batch
        out: List[Optional[tuple]] = [None] * len(forms)
        by_shape: Dict[Tuple[int, int], List[int]] = {}
        for idx, f in enumerate(forms):
            by_shape.setdefault((f.m, f.n), []).append(idx)
        for (m, n), idxs in by_shape.items():
            grids = np.asarray([forms[i].grid for i in idxs], dtype=object)
            small = all(abs(int(v)) < 2 ** 31 for v in grids.reshape(-1))
            res = None
            if small:
                g32 = np.ascontiguousarray(grids.astype(np.int64), dtype=np.int32)
                try:
                    res = solve_batch(g32, device=device)
                except ValueError:
                    res = None                                       # too many primes / too large for shared memory
            if res is not None:
                if stats is not None:
                    stats["batches"] += 1
                for k, i in enumerate(idxs):
                    st = int(res.status[k])
                    if st == N.OPTIMAL:
                        out[i] = ("optimal", res.objective(k), res.bfs(k))
                    elif st == N.INFEASIBLE:
                        out[i] = ("infeasible", None, None)
                    elif st == N.UNBOUNDED:
                        out[i] = ("unbounded", None, None)
                    elif st in (N.BAD_PRIME, N.REDUNDANT_ROW, N.MAX_PIVOTS, N.NEED_PRIMES, N.NEED_ARTIFICIALS):
                        # solved alone below: the single-LP engine re-encodes around a bad prime, has no pivot cap,
                        # and drops dependent equality rows instead of reproducing the reference constructor's crash
                        # (this layer has no reference behaviour to mirror)
                        out[i] = None
                    else:
                        res.raise_for_status(k)
            for i in idxs:
                if out[i] is None:
                    out[i] = LinProg._solve_one(forms[i], device)
                    if stats is not None:
                        stats["batches"] += 1
        return out

    @staticmethod
    def _solve_one(form: StandardForm, device=None):
        """One node LP through `Tableau`/`Simplex` in exact mode (the single-LP engine)."""
        from .simplex import Simplex
        from .tableau import Tableau
        t = Tableau.from_grid(form.grid, mode="exact", device=device)
        try:
            s = Simplex(t, drop_redundant=True)
        except ValueError as e:
            if str(e).startswith("infeasible problem"):
                return ("infeasible", None, None)
            raise
        try:
            s.solve()
        except AssertionError as e:
            if "unbounded" in str(e):
                return ("unbounded", None, None)
            raise
        return ("optimal", s.getObjValue(), dict(s.getBFS()))

    # -- branch and bound ------------------------------------------------------------------------
    def solve(self, batch_size: int = 4096, max_nodes: int = 1_000_000, device=None) -> LinProgResult:
        """Minimise / maximise the objective; integer variables by breadth-first branch and bound.

        The frontier is processed in rounds of up to `batch_size` nodes whose LP relaxations are solved in one
        batched kernel launch per tableau shape; within a round the nodes are examined in frontier order, the
        incumbent prunes as it improves, and children are appended down-branch first (x <= floor(v), then
        x >= ceil(v)) on the first integer variable, in declaration order, with a fractional value."""
        self._declare_missing()
        intvars = [name for name, v in self.vars.items() if v.isint]
        frontier: List[Tuple[int, Dict[str, LinVar]]] = [(0, {name: v.copy() for name, v in self.vars.items()})]
        best_obj: Optional[Fraction] = None
        best_val: Dict[str, Fraction] = {}
        stats = {"batches": 0}
        result = LinProgResult("infeasible")
        better = (lambda a, b: a < b) if self.sense == "min" else (lambda a, b: a > b)
        while frontier:
            batch, frontier = frontier[:batch_size], frontier[batch_size:]
            forms = [self.lower({name: v.getBounds() for name, v in box.items()}) for _, box in batch]
            live = [k for k, f in enumerate(forms) if f is not None]
            solved = self._solve_nodes([forms[k] for k in live], device=device, stats=stats)
            result.lps += len(live)
            answers: Dict[int, tuple] = dict(zip(live, solved))
            for k, (depth, box) in enumerate(batch):
                result.nodes += 1
                if result.nodes > max_nodes:
                    raise RuntimeError(f"branch and bound stopped after {max_nodes} nodes")
                if k not in answers:
                    result.trace.append((depth, "empty", None))
                    continue
                status, tab_obj, bfs = answers[k]
                if status == "unbounded":
                    raise ValueError("unbounded problem")
                if status == "infeasible":
                    result.trace.append((depth, "infeasible", None))
                    continue
                form = forms[k]
                obj = form.objective(tab_obj)
                result.trace.append((depth, "optimal", obj))
                if best_obj is not None and not better(obj, best_obj):
                    continue                                         # bound: cannot beat the incumbent
                val = form.values(bfs)
                branch = next((name for name in intvars if val[name].denominator != 1), None)
                if branch is None:
                    best_obj, best_val = obj, val
                    continue
                v = val[branch]
                down = {name: lv.copy() for name, lv in box.items()}
                down[branch].boundAbove(math.floor(v))               # linprog.py:338-344
                up = {name: lv.copy() for name, lv in box.items()}
                up[branch].boundBelow(math.ceil(v))                  # linprog.py:346-352
                frontier.append((depth + 1, down))
                frontier.append((depth + 1, up))
        result.batches = stats["batches"]
        if best_obj is not None:
            result.status, result.objective, result.values = "optimal", best_obj, best_val
        return result
