"""Pin the CPU oracle (oracle/lpsol_oracle.py) to the reference.

Sources of truth, in order: the reference's only known-answer test on the path
(test_tableau.py:220-227, restated in tests/golden/textbook.json), the known
answers SURVEY.md section 8c records, and the fixtures tools/make_golden.py
produced by running the unmodified reference.  No GPU needed.
"""
from fractions import Fraction

import pytest

from conftest import load_golden
from linear_program_solver_b200 import problems
from oracle import lpsol_oracle as O


def test_textbook_pivots_match_reference_known_answer():
    g = load_golden("textbook")
    grid = O.make_grid(g["c"], g["A"], g["b"])
    for step in g["steps"]:
        O.pivot(grid, *step["pivot"])
        m, n = O.shape(grid)
        assert [str(x) for x in grid[m][:n]] == step["c"]
        assert str(grid[m][n]) == step["z_cell"]
        assert [str(grid[i][n]) for i in range(m)] == step["b"]
        assert [[str(x) for x in grid[i][:n]] for i in range(m)] == step["a"]
    # the fixture's final state is the reference's tab1c: -z cell 400, x1=4, x2=8
    assert grid[2][4] == 400 and grid[0][4] == 8 and grid[1][4] == 4
    # third assertion of test_pivot: pivoting (1,0) then (0,1) from tab1a reaches tab1c too
    grid2 = O.make_grid(g["c"], g["A"], g["b"])
    O.pivot(grid2, 1, 0)
    O.pivot(grid2, 0, 1)
    assert grid2 == grid


def test_getz_sign_convention():
    # test_tableau.py:69-72: corner cell 320 <=> getZ() == -320
    grid = O.make_grid(*problems.TEXTBOOK)
    O.pivot(grid, 1, 0)
    assert grid[2][4] == 320 and O.objective(grid) == -320


def test_zero_pivot_raises_like_reference():
    grid = O.make_grid(*problems.TEXTBOOK)
    with pytest.raises(ZeroDivisionError, match="zero pivot 0,3"):
        O.pivot(grid, 0, 3)


def _check(rec, res, num):
    fmt = (lambda x: x.hex()) if num is float else (lambda x: str(Fraction(x)))
    assert res["initial_basis"] == rec["initial_basis"]
    assert [list(p) for p in res["phase1_pivots"]] == rec["phase1_pivots"]
    assert [list(p) for p in res["pivots"]] == rec["pivots"]
    assert res["basis"] == rec["basis"]
    assert fmt(res["objective"]) == rec["objective"]
    assert {str(j): fmt(v) for j, v in res["bfs"].items()} == rec["bfs"]


@pytest.mark.parametrize("rec", load_golden("dense_exact"), ids=lambda r: f"m{r['m']}s{r['seed']}h{r['hi']}")
def test_dense_exact_matches_reference(rec):
    if rec["m"] >= 100:
        pytest.skip("large cases are covered by the GPU parity tests; the oracle takes seconds here")
    c, A, b = problems.dense_lp(rec["m"], rec["seed"], rec["lo"], rec["hi"])
    res = O.simplex(c, A, b, max_pivots=rec.get("max_pivots"))
    _check(rec, res, Fraction)


@pytest.mark.parametrize("rec", load_golden("dense_float"), ids=lambda r: f"m{r['m']}s{r['seed']}")
def test_dense_float_bit_identical_to_rebound_reference(rec):
    if rec["m"] > 100:
        pytest.skip("slow in pure Python")
    c, A, b = problems.dense_lp(rec["m"], rec["seed"], rec["lo"], rec["hi"])
    res = O.simplex(c, A, b, num=float)
    _check(rec, res, float)


@pytest.mark.parametrize("rec", load_golden("phase1"), ids=lambda r: f"m{r['m']}n{r['n']}s{r['seed']}")
def test_phase1_matches_reference(rec):
    c, A, b = problems.equality_lp(rec["m"], rec["n"], rec["seed"])
    res = O.simplex(c, A, b)
    _check(rec, res, Fraction)


def test_beale_rule_switch_and_stale_objective():
    g = load_golden("special")["beale"]
    c, A, b = problems.BEALE
    res = O.simplex(c, A, b)
    _check(g, res, Fraction)
    # exactly m+n = 10 Dantzig pivots at objective 0, then 8 Bland pivots
    assert [p[2] for p in res["pivots"]] == [0] * 10 + [1] * 8
    assert res["objective"] == Fraction(-5, 4)
    gf = load_golden("special")["beale_float"]
    resf = O.simplex([float(x) for x in c], [[float(x) for x in r] for r in A], [float(x) for x in b], num=float)
    _check(gf, resf, float)


def test_special_cases_raise_like_reference():
    s = load_golden("special")
    with pytest.raises(IndexError):
        O.simplex([1, 1, 1], [[1, 1, 0], [2, 2, 0], [0, 1, 1]], [2, 4, 3])
    assert s["redundant_row"]["ctor_error"][0] == "IndexError"
    with pytest.raises(ValueError, match="infeasible problem, artificial opt = 1"):
        O.simplex([0, 0], [[1, 1], [1, 1]], [1, 2])
    assert s["infeasible"]["ctor_error"] == ["ValueError", "infeasible problem, artificial opt = 1"]
    with pytest.raises(AssertionError, match="unbounded artificial problem"):
        O.simplex([-1, 0], [[-1, 1]], [1])
    grid = O.make_grid([-1, 0], [[-1, 1]], [1])
    assert O.find_pivot(grid, O.RULE_DANTZIG) == s["unbounded"]["findPivotStandard"] == "unbounded"
    assert O.find_pivot(grid, O.RULE_BLAND) == s["unbounded"]["findPivotMinIndex"] == "unbounded"


def test_survey_known_answers():
    # SURVEY.md section 8c(6), seed 0
    for m, npiv, obj in [(8, 5, Fraction(-789, 14)), (50, 33, Fraction(-56571599, 254337))]:
        res = O.simplex(*problems.dense_lp(m, 0))
        assert len(res["pivots"]) == npiv and res["objective"] == obj


def test_bareiss_form_equals_fraction_form_at_every_pivot():
    """The integer form the exact GPU mode stores: entry == N/D after every pivot."""
    for m, seed in [(6, 0), (12, 1), (20, 2)]:
        c, A, b = problems.dense_lp(m, seed)
        res = O.simplex(c, A, b)
        grid = O.make_grid(c, A, b)
        N = [[int(x) for x in row] for row in grid]
        D = 1
        for (r, cc, _rule) in res["pivots"]:
            O.pivot(grid, r, cc)
            D = O.bareiss_pivot(N, D, r, cc)
            assert D > 0
            for i in range(m + 1):
                for j in range(2 * m + 1):
                    assert Fraction(N[i][j], D) == grid[i][j]


@pytest.mark.parametrize("rec", load_golden("rules"), ids=lambda r: f"{r['family']}m{r['m']}s{r['seed']}")
def test_whole_tableau_rules_match_reference(rec):
    """findPivotMaxIncrease / findPivotAll (simplex.py:286-360) against the reference's answers at every state of
    the walk tools/make_golden_rules.py recorded (alternating max-increase and Dantzig pivots)."""
    if rec["family"] == "dense":
        c, A, b = problems.dense_lp(rec["m"], rec["seed"], rec["lo"], rec["hi"])
    else:
        c, A, b = problems.equality_lp(rec["m"], rec["n"], rec["seed"])
    for kind, num in (("exact", Fraction), ("float", float)):
        if kind not in rec:
            continue
        grid = O.make_grid(c, A, b, num=num)
        grid, basis = O.find_bfs(grid)
        for step, want in enumerate(rec[kind]):
            got = O.find_pivot_max_increase(grid)
            assert (list(got) if isinstance(got, tuple) else got) == want["max_increase"], (kind, step)
            assert [list(p) for p in O.find_pivot_all(grid)] == want["all"], (kind, step)
            if isinstance(got, str):
                break
            r, cc = got if step % 2 else O.find_pivot(grid, O.RULE_DANTZIG)
            O.pivot(grid, r, cc)
