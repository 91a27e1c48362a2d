"""CPU oracle for the MILP layer -- TEST INFRASTRUCTURE ONLY.

The reference has no branch and bound (its `LinProg` is an empty stub, /root/reference/lpsol/linprog.py:383-393;
SURVEY.md F1), so this is not a restatement of reference code: it is the CPU counterpart of
`linear_program_solver_b200/linprog.py`, written independently on plain data, in which EVERY NODE LP is
solved by the reference's algorithm -- by default the pinned oracle `lpsol_oracle.simplex` (the restatement of
`Simplex(tab); solve()`, simplex.py:24-148), or, in tools/make_golden.py, by the unmodified reference `Simplex`
itself.  Parity status of the MILP layer: unpinned by the reference (nothing to pin against); what IS pinned
is each node LP, and `tests/test_bnb_oracle.py` additionally checks this driver against brute-force
enumeration on small bounded integer programs.

Problem description (plain data, no product classes):
    variables    [(name, is_integer, lb | None, ub | None), ...]        in declaration order
    constraints  [({name: coef}, '<=' | '>=' | '==', rhs), ...]
    objective    ({name: coef}, constant, 'min' | 'max')

Only `tests/`, `tools/make_golden.py`, `__graft_entry__.smoke()` and bench scripts' CPU legs may import this.
"""
from __future__ import annotations

import math
from fractions import Fraction as F
from typing import Callable, Dict

from . import lpsol_oracle as O


def _lp_via_oracle(c, A, b):
    """('optimal', objective, {column: value}) | ('infeasible',) | ('unbounded',) with the pinned simplex oracle."""
    try:
        res = O.simplex(c, A, b)
    except O.InfeasibleError:
        return ("infeasible",)
    except O.UnboundedError:
        return ("unbounded",)
    return ("optimal", res["objective"], res["bfs"])


def lower(variables, constraints, objective, box):
    """Integer tableau data (c, A, b) of the relaxation over `box` {name: (lb, ub)}, or None for an empty box.

    Column plan, in declaration order: a variable with a lower bound becomes lb + y; with only an upper bound
    ub - y; a free variable y1 - y2.  Rows: the constraints in order ('>=' rows negated into '<='), then one
    row y <= ub - lb per two-sided variable; every '<=' row is made integral and then gets a unit slack, '=='
    rows get none.  Slack columns follow the structural ones in row order.  The cost row is made integral by a
    positive scale.  Returns also what is needed to map a basic solution back."""
    plan = []          # per structural column: (name, a, sign) meaning x = a + sign * y   (free: two entries, a = 0)
    for name, _isint, _lb0, _ub0 in variables:
        lb, ub = box[name]
        if lb is not None and ub is not None and lb > ub:
            return None
        if lb is not None:
            plan.append((name, F(lb), 1))
        elif ub is not None:
            plan.append((name, F(ub), -1))
        else:
            plan.append((name, F(0), 1))
            plan.append((name, F(0), -1))
    ns = len(plan)

    def expand(coefs: Dict[str, F]):
        """sum coef * x  ->  (row over structural columns, constant)"""
        row, const = [F(0)] * ns, F(0)
        for j, (name, a, sgn) in enumerate(plan):
            cf = F(coefs.get(name, 0))
            if cf == 0:
                continue
            row[j] += cf * sgn
        seen = set()
        for name, a, sgn in plan:
            if name in seen:
                continue
            seen.add(name)
            const += F(coefs.get(name, 0)) * a
        return row, const

    rows = []
    for coefs, comp, rhs in constraints:
        row, const = expand(coefs)
        rhs = F(rhs) - const
        if comp == ">=":
            row, rhs = [-v for v in row], -rhs
        rows.append((row, rhs, comp != "=="))
    col = 0
    for name, _isint, _l, _u in variables:
        lb, ub = box[name]
        if lb is not None:
            if ub is not None:
                rows.append(([F(1) if k == col else F(0) for k in range(ns)], F(ub) - F(lb), True))
            col += 1
        elif ub is not None:
            col += 1
        else:
            col += 2
    nslack = sum(1 for r in rows if r[2])
    A, b, k = [], [], 0
    for row, rhs, slack in rows:
        scale = 1
        for v in row + [rhs]:
            scale = scale * v.denominator // math.gcd(scale, v.denominator)
        line = [int(v * scale) for v in row] + [0] * nslack
        if slack:
            line[ns + k] = 1
            k += 1
        A.append(line)
        b.append(int(rhs * scale))
    ocoefs, oconst, sense = objective
    sgn = 1 if sense == "min" else -1
    crow, cconst = expand({n: F(v) for n, v in ocoefs.items()})
    crow = [sgn * v for v in crow]
    cscale = 1
    for v in crow:
        cscale = cscale * v.denominator // math.gcd(cscale, v.denominator)
    c = [int(v * cscale) for v in crow] + [0] * nslack
    return {"c": c, "A": A, "b": b, "plan": plan, "cscale": cscale, "const": F(oconst) + cconst, "sgn": sgn}


def branch_and_bound(variables, constraints, objective, batch_size: int = 4096,
                     solve_lp: Callable = _lp_via_oracle, max_nodes: int = 100000):
    """Breadth-first rounds of `batch_size` nodes; within a round nodes are examined in frontier order, the
    incumbent prunes (a node must be strictly better), children are appended down branch first on the first
    integer variable (declaration order) with a fractional value; integer bounds are floored / ceiled like
    `LinVar.boundAbove/boundBelow` (linprog.py:338-352)."""
    sense = objective[2]
    better = (lambda a, b: a < b) if sense == "min" else (lambda a, b: a > b)
    intnames = [name for name, isint, _l, _u in variables if isint]
    root = {}
    for name, isint, lb, ub in variables:
        lb = None if lb is None else (F(math.ceil(F(lb))) if isint else F(lb))
        ub = None if ub is None else (F(math.floor(F(ub))) if isint else F(ub))
        root[name] = (lb, ub)
    frontier = [(0, root)]
    best, best_x = None, None
    trace, nodes, lps = [], 0, 0
    while frontier:
        batch, frontier = frontier[:batch_size], frontier[batch_size:]
        for depth, box in batch:
            nodes += 1
            if nodes > max_nodes:
                raise RuntimeError("too many nodes")
            low = lower(variables, constraints, objective, box)
            if low is None:
                trace.append((depth, "empty", None))
                continue
            lps += 1
            ans = solve_lp(low["c"], low["A"], low["b"])
            if ans[0] == "unbounded":
                raise ValueError("unbounded problem")
            if ans[0] == "infeasible":
                trace.append((depth, "infeasible", None))
                continue
            _, tab_obj, bfs = ans
            obj = low["sgn"] * (F(tab_obj) / low["cscale"]) + low["const"]
            trace.append((depth, "optimal", obj))
            if best is not None and not better(obj, best):
                continue
            x: Dict[str, F] = {}
            for j, (name, a, sgn) in enumerate(low["plan"]):
                y = F(bfs.get(j, 0))
                x[name] = (x[name] if name in x else a) + sgn * y
            frac = next((name for name in intnames if x[name].denominator != 1), None)
            if frac is None:
                best, best_x = obj, x
                continue
            v = x[frac]
            lb, ub = box[frac]
            down, up = dict(box), dict(box)
            fl, ce = F(math.floor(v)), F(math.ceil(v))
            down[frac] = (lb, fl if ub is None or fl < ub else ub)
            up[frac] = (ce if lb is None or ce > lb else lb, ub)
            frontier.append((depth + 1, down))
            frontier.append((depth + 1, up))
    return {"status": "optimal" if best is not None else "infeasible", "objective": best, "values": best_x or {},
            "nodes": nodes, "lps": lps, "trace": trace}


def brute_force(variables, constraints, objective):
    """Exhaustive optimum of a PURE integer program whose variables all have finite bounds (independent check)."""
    import itertools
    names = [v[0] for v in variables]
    ranges = [range(math.ceil(F(v[2])), math.floor(F(v[3])) + 1) for v in variables]
    ocoefs, oconst, sense = objective
    best, best_x = None, None
    for point in itertools.product(*ranges):
        x = dict(zip(names, point))
        ok = True
        for coefs, comp, rhs in constraints:
            lhs = sum(F(cf) * x[nm] for nm, cf in coefs.items())
            if (comp == "<=" and lhs > rhs) or (comp == ">=" and lhs < rhs) or (comp == "==" and lhs != rhs):
                ok = False
                break
        if not ok:
            continue
        val = sum(F(cf) * x[nm] for nm, cf in ocoefs.items()) + F(oconst)
        if best is None or (val < best if sense == "min" else val > best):
            best, best_x = val, x
    return best, best_x
