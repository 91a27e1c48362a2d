"""MILP layer, CPU side: the branch-and-bound oracle against brute force, the product's lowering against the
oracle's, and the golden MILP fixtures (node LPs solved by the UNMODIFIED reference `Simplex`, see
tools/make_golden.py) against the oracle driver with the pinned simplex oracle as node solver."""
import random
from fractions import Fraction as F

import pytest

from conftest import load_golden
from oracle import bnb_oracle as B


def random_ip(rng, nv, nc, pure=True):
    """Small bounded program: knapsack-like rows plus a few general ones."""
    variables = []
    for j in range(nv):
        isint = pure or rng.random() < 0.6
        lb = rng.choice([0, 0, 0, -2, 1])
        ub = lb + rng.randint(1, 4)
        variables.append((f"x{j}", isint, lb, ub))
    cons = []
    for _ in range(nc):
        coefs = {f"x{j}": rng.randint(-3, 6) for j in range(nv) if rng.random() < 0.8}
        if not coefs:
            coefs = {"x0": 1}
        comp = rng.choice(["<=", "<=", ">=", "=="]) if rng.random() < 0.3 else "<="
        rhs = rng.randint(2, 12) if comp != "==" else rng.randint(0, 6)
        if rng.random() < 0.2:
            rhs = F(rhs * 2 + 1, 2)
        cons.append((coefs, comp, rhs))
    obj = ({f"x{j}": rng.randint(-5, 5) for j in range(nv)}, rng.randint(-3, 3), rng.choice(["min", "max"]))
    return variables, cons, obj


def to_linprog(variables, cons, obj):
    from linear_program_solver_b200 import LinExpr, LinProg, LinVar
    lp = LinProg()
    for name, isint, lb, ub in variables:
        lp.addVar(LinVar(name, isint, lb, ub))
    for coefs, comp, rhs in cons:
        e = LinExpr()
        for nm, cf in coefs.items():
            e += LinExpr(cf, nm)
        make = {"<=": e.constraintLeq, ">=": e.constraintGeq, "==": e.constrantEq}[comp]
        lp.addConstraint(make(rhs))
    oc, const, sense = obj
    e = LinExpr(const)
    for nm, cf in oc.items():
        e += LinExpr(cf, nm)
    lp.setObjective(e, sense)
    return lp


@pytest.mark.parametrize("seed", range(12))
def test_oracle_bnb_matches_brute_force(seed):
    rng = random.Random(seed)
    variables, cons, obj = random_ip(rng, rng.randint(2, 4), rng.randint(1, 3), pure=True)
    want, want_x = B.brute_force(variables, cons, obj)
    got = B.branch_and_bound(variables, cons, obj)
    if want is None:
        assert got["status"] == "infeasible"
    else:
        assert got["status"] == "optimal" and got["objective"] == want
        # the optimum may be attained at several points: check feasibility and value of the one returned
        x = got["values"]
        assert all(v.denominator == 1 for v in x.values())
        assert sum(F(cf) * x[nm] for nm, cf in obj[0].items()) + obj[1] == want


@pytest.mark.parametrize("seed", range(20))
def test_product_lowering_equals_oracle_lowering(seed):
    """`LinProg.lower` is host-only bookkeeping: the integer tableau must be the oracle's, cell for cell."""
    rng = random.Random(100 + seed)
    variables, cons, obj = random_ip(rng, rng.randint(2, 5), rng.randint(1, 4), pure=False)
    if seed % 4 == 0:                           # a free variable and a variable bounded only from above
        variables[0] = (variables[0][0], False, None, None)
        if len(variables) > 1:
            variables[1] = (variables[1][0], False, None, 5)
    lp = to_linprog(variables, cons, obj)
    form = lp.lower()
    box = {name: (None if lb is None else F(lb), None if ub is None else F(ub)) for name, _i, lb, ub in variables}
    low = B.lower(variables, cons, obj, box)
    m, n = len(low["A"]), len(low["c"])
    assert (form.m, form.n) == (m, n)
    assert [row[:n] for row in form.grid[:m]] == low["A"]
    assert [row[n] for row in form.grid[:m]] == low["b"]
    assert form.grid[m][:n] == low["c"] and form.grid[m][n] == 0
    assert form.obj_scale == low["cscale"] and form.obj_const == low["const"] and form.sign == low["sgn"]


def test_empty_box_and_bound_rounding():
    from linear_program_solver_b200 import LinExpr, LinProg, LinVar
    lp = LinProg()
    lp.addVar(LinVar("x", True, F(1, 2), F(7, 2)))          # integer bounds are ceiled / floored: [1, 3]
    lp.addVar("y")
    lp.addConstraint(LinExpr(1, "x", 1, "y").constraintLeq(4))
    assert lp.vars["x"].getBounds() == (F(1), F(3))
    assert lp.lower({"x": (F(3), F(2)), "y": (F(0), None)}) is None
    form = lp.lower()
    assert form.columns == [("x", "+"), ("y", "+")]
    assert form.grid == [[1, 1, 1, 0, 3], [1, 0, 0, 1, 2], [0, 0, 0, 0, 0]]


def test_golden_milp_fixtures():
    """tests/golden/milp.json: this B&B driver with the unmodified reference `Simplex` solving every node."""
    recs = load_golden("milp")
    assert len(recs) >= 10
    for rec in recs:
        variables = [(n, i, None if lb is None else F(lb), None if ub is None else F(ub)) for n, i, lb, ub in rec["variables"]]
        cons = [({k: F(v) for k, v in c.items()}, comp, F(rhs)) for c, comp, rhs in rec["constraints"]]
        obj = ({k: F(v) for k, v in rec["objective"][0].items()}, F(rec["objective"][1]), rec["objective"][2])
        got = B.branch_and_bound(variables, cons, obj)
        assert got["status"] == rec["status"]
        assert (None if got["objective"] is None else str(got["objective"])) == rec["objective_value"]
        assert {k: str(v) for k, v in got["values"].items()} == rec["values"]
        assert got["nodes"] == rec["nodes"]
        assert [[d, s, None if o is None else str(o)] for d, s, o in got["trace"]] == rec["trace"]


@pytest.mark.parametrize("seed", range(24))
def test_oracle_bnb_matches_an_independent_milp_solver(seed):
    """The MILP layer has no reference implementation, so its oracle is pinned from outside as well: the
    optimum of random mixed-integer programs -- free variables, variables bounded only from above, negative lower
    bounds, equality and >= rows, rational right-hand sides, both senses -- against scipy's HiGHS (`optimize.milp`) on
    the ORIGINAL general-form model, never on our lowering.  Objective within 1e-7; infeasibility must agree."""
    import numpy as np
    from scipy.optimize import Bounds, LinearConstraint, milp
    rng = random.Random(9000 + seed)
    variables, cons, obj = random_ip(rng, rng.randint(2, 5), rng.randint(1, 4), pure=seed % 3 == 0)
    if seed % 4 == 1:                            # a free variable and one bounded only from above, kept bounded by rows
        variables[0] = (variables[0][0], False, None, None)
        cons.append(({variables[0][0]: 1}, ">=", -7))
        cons.append(({variables[0][0]: 1}, "<=", F(15, 2)))
        if len(variables) > 1:
            variables[1] = (variables[1][0], variables[1][1], None, 5)
            cons.append(({variables[1][0]: 2}, ">=", -9))
    names = [v[0] for v in variables]
    oc, const, sense = obj
    sgn = 1.0 if sense == "min" else -1.0
    c = np.array([sgn * float(oc.get(nm, 0)) for nm in names])
    rows, lo, hi = [], [], []
    for coefs, comp, rhs in cons:
        rows.append([float(coefs.get(nm, 0)) for nm in names])
        lo.append(-np.inf if comp == "<=" else float(rhs))
        hi.append(np.inf if comp == ">=" else float(rhs))
    res = milp(c, constraints=LinearConstraint(np.array(rows), lo, hi),
               integrality=np.array([1 if v[1] else 0 for v in variables]),
               bounds=Bounds([-np.inf if v[2] is None else float(v[2]) for v in variables],
                             [np.inf if v[3] is None else float(v[3]) for v in variables]))
    assert res.status in (0, 2), res.message     # optimal or infeasible: the rows above keep every program bounded
    got = B.branch_and_bound(variables, cons, obj)
    if res.status == 2:
        assert got["status"] == "infeasible"
        return
    assert got["status"] == "optimal"
    assert abs(float(got["objective"]) - (sgn * res.fun + const)) <= 1e-7 * max(1.0, abs(res.fun))
    x = got["values"]                            # our optimum is feasible for the original model, exactly
    for nm, isint, lb, ub in variables:
        assert (lb is None or x[nm] >= lb) and (ub is None or x[nm] <= ub) and (not isint or x[nm].denominator == 1)
    for coefs, comp, rhs in cons:
        lhs = sum(F(cf) * x[nm] for nm, cf in coefs.items())
        assert {"<=": lhs <= rhs, ">=": lhs >= rhs, "==": lhs == rhs}[comp]
