"""GPU parity of `LinProg.solve` (branch and bound with batched node LPs) against the golden MILP fixtures --
produced with the UNMODIFIED reference `Simplex` solving every node (tools/make_golden.py) -- and against the
CPU oracle driver on random mixed-integer programs: status, objective, solution, node count and the per-node
trace (depth, outcome, LP objective) must be identical."""
import random
from fractions import Fraction as F

import pytest

from conftest import load_golden
from test_bnb_oracle import random_ip, to_linprog

pytestmark = pytest.mark.gpu


def _same(res, want):
    assert res.status == want["status"]
    assert res.objective == want["objective"]
    assert res.values == want["values"]
    assert res.nodes == want["nodes"] and res.lps == want["lps"]
    assert res.trace == want["trace"]


def test_golden_milp_fixtures():
    recs = load_golden("milp")
    for rec in recs:
        variables = [(n, i, None if lb is None else F(lb), None if ub is None else F(ub)) for n, i, lb, ub in rec["variables"]]
        cons = [({k: F(v) for k, v in c.items()}, comp, F(rhs)) for c, comp, rhs in rec["constraints"]]
        obj = ({k: F(v) for k, v in rec["objective"][0].items()}, F(rec["objective"][1]), rec["objective"][2])
        res = to_linprog(variables, cons, obj).solve()
        assert res.status == rec["status"]
        assert (None if res.objective is None else str(res.objective)) == rec["objective_value"]
        assert {k: str(v) for k, v in res.values.items()} == rec["values"]
        assert res.nodes == rec["nodes"] and res.lps == rec["lps"]
        assert [[d, s, None if o is None else str(o)] for d, s, o in res.trace] == rec["trace"]


@pytest.mark.parametrize("seed", range(10))
def test_random_mixed_integer_programs_against_the_oracle(seed):
    from oracle import bnb_oracle as B
    rng = random.Random(9000 + seed)
    variables, cons, obj = random_ip(rng, rng.randint(3, 7), rng.randint(2, 6), pure=(seed % 3 == 0))
    if seed % 5 == 1:                            # a free continuous variable and one bounded only from above
        variables[0] = (variables[0][0], False, None, None)
        variables[1] = (variables[1][0], False, None, 4)
        cons.append(({variables[0][0]: 1}, ">=", -3))
        cons.append(({variables[0][0]: 1}, "<=", 6))
    want = B.branch_and_bound(variables, cons, obj)
    _same(to_linprog(variables, cons, obj).solve(), want)


def test_small_rounds_follow_the_same_order():
    """batch_size = 2 and 1: more rounds, same deterministic node order as the oracle with that batch size."""
    from oracle import bnb_oracle as B
    rng = random.Random(77)
    variables, cons, obj = random_ip(rng, 6, 4, pure=True)
    for bs in (1, 2, 7):
        want = B.branch_and_bound(variables, cons, obj, batch_size=bs)
        _same(to_linprog(variables, cons, obj).solve(batch_size=bs), want)


def test_knapsack_and_relaxation_only():
    from linear_program_solver_b200 import LinExpr, LinProg, LinVar
    lp = LinProg()
    for nm in "abcd":
        lp.addVar(LinVar(nm, True, 0, 1))
    lp.addConstraint(LinExpr(5, "a", 7, "b", 4, "c", 3, "d").constraintLeq(14))
    lp.setObjective(LinExpr(8, "a", 11, "b", 6, "c", 4, "d"), "max")
    res = lp.solve()
    assert res.status == "optimal" and res.objective == 21
    assert res.values == {"a": 0, "b": 1, "c": 1, "d": 1}
    # no integer variables: one LP, the reference's textbook problem (test_tableau.py:9-29) as a LinProg
    lp = LinProg()
    lp.addConstraint(LinExpr(1, "x1", 1, "x2").constraintLeq(12))
    lp.addConstraint(LinExpr(2, "x1", 1, "x2").constraintLeq(16))
    lp.setObjective(LinExpr(40, "x1", 30, "x2"), "max")
    res = lp.solve()
    assert (res.objective, res.values, res.nodes) == (400, {"x1": 4, "x2": 8}, 1)


def test_unbounded_relaxation_raises():
    from linear_program_solver_b200 import LinExpr, LinProg
    lp = LinProg()
    lp.addVar("x", integral=True)
    lp.addConstraint(LinExpr(-1, "x", 1, "y").constraintLeq(1))
    lp.setObjective(LinExpr(-1, "x"))
    with pytest.raises(ValueError):
        lp.solve()


def test_dependent_equality_rows_are_dropped_instead_of_crashing():
    """x + y = 4 stated twice (once scaled): phase 1 ends with an artificial variable on a zero row.  The reference
    constructor crashes there (simplex.py:93, reproduced by `Simplex`); the MILP layer -- new design, SURVEY F1 --
    drops the dependent row and goes on (round-1 advisor finding)."""
    from linear_program_solver_b200 import LinExpr, LinProg, LinVar, Simplex, Tableau
    lp = LinProg()
    lp.addVar(LinVar("x", True, 0, 3))
    lp.addVar(LinVar("y", True, 0, None))
    lp.addConstraint(LinExpr(1, "x", 1, "y").constrantEq(4))
    lp.addConstraint(LinExpr(2, "x", 2, "y").constrantEq(8))
    lp.addConstraint(LinExpr(1, "x", -1, "y").constraintLeq(1))
    lp.setObjective(LinExpr(3, "x", 1, "y"), "max")
    res = lp.solve()
    assert res.status == "optimal" and res.objective == 8 and res.values == {"x": 2, "y": 2}
    # the class surface keeps the reference's behaviour unless asked otherwise
    grid = [[1, 1, 0, 2], [2, 2, 0, 4], [0, 1, 1, 3], [1, 1, 1, 0]]
    with pytest.raises((IndexError, AssertionError)):
        Simplex(Tableau.from_grid(grid, mode="exact"))
    s = Simplex(Tableau.from_grid(grid, mode="exact"), drop_redundant=True)
    s.solve()
    assert s._tab.getNumCons() == 2 and s.getObjValue() == 3      # x2 = 2, x3 = 1
