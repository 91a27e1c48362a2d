"""GPU parity of the public API (Tableau / Simplex) in fp64 mode against the CPU oracle and the
float-rebound reference fixtures.  fp64 mode rounds exactly like the reference's control flow run on
floats, so pivot sequences and results are compared bit-for-bit (hex floats); the north star's 1e-9
relative tolerance against the exact answer is checked on top."""
from fractions import Fraction

import numpy as np
import pytest

from conftest import load_golden
from linear_program_solver_b200 import problems

pytestmark = pytest.mark.gpu
TOL = 1e-9      # north star: fp64 mode within 1e-9 relative of the exact objective / primal values


def _tab(c, A, b, mode="fp64"):
    from linear_program_solver_b200 import Tableau
    m, n = len(A), len(c)
    t = Tableau(m, n, mode=mode)
    t.setC(list(c))
    t.setB(list(b))
    t.setA([list(r) for r in A])
    return t


def test_reference_known_answer_test_pivot():
    """test_tableau.py:220-227 restated on the new Tableau (fp64 mode)."""
    g = load_golden("textbook")
    t = _tab(g["c"], g["A"], g["b"])
    for step in g["steps"]:
        t.pivot(*step["pivot"])
        assert t.getC() == [float(Fraction(x)) for x in step["c"]]
        assert t._z == float(Fraction(step["z_cell"]))
        assert t.getB() == [float(Fraction(x)) for x in step["b"]]
        assert t.getA() == [[float(Fraction(x)) for x in r] for r in step["a"]]
    assert t.getZ() == -400.0
    with pytest.raises(ZeroDivisionError, match="zero pivot 0,3"):
        _tab(g["c"], g["A"], g["b"]).pivot(0, 3)


@pytest.mark.parametrize("rec", load_golden("dense_float"), ids=lambda r: f"m{r['m']}s{r['seed']}")
def test_simplex_matches_float_rebound_reference(rec):
    from linear_program_solver_b200 import Simplex
    c, A, b = problems.dense_lp(rec["m"], rec["seed"], rec["lo"], rec["hi"])
    t = _tab(c, A, b)
    s = Simplex(t)
    assert s.getBasicSequence() == rec["initial_basis"]
    s.solve()
    assert [list(p) for p in s.pivot_log] == rec["pivots"]
    assert s.getBasicSequence() == rec["basis"]
    assert float(s.getObjValue()).hex() == rec["objective"]
    assert {str(j): float(v).hex() for j, v in s.getBFS().items()} == rec["bfs"]
    assert [bool(x) for x in t.getVarMarks()] == rec["marks"]


@pytest.mark.parametrize("rec", [r for r in load_golden("dense_exact") if r["m"] in (8, 50, 100) and r["hi"] == 9][:8],
                         ids=lambda r: f"m{r['m']}s{r['seed']}")
def test_fp64_within_tolerance_of_exact_reference(rec):
    from linear_program_solver_b200 import Simplex
    c, A, b = problems.dense_lp(rec["m"], rec["seed"], rec["lo"], rec["hi"])
    s = Simplex(_tab(c, A, b))
    s.solve()
    want = Fraction(rec["objective"])
    assert abs(Fraction(s.getObjValue()) - want) <= TOL * abs(want)
    # primal values: basic variables of the exact optimum (other variables are 0 in both)
    got = {j: v for j, v in s.getBFS().items()}
    for j, v in rec["bfs"].items():
        v = Fraction(v)
        assert abs(Fraction(got.get(int(j), 0.0)) - v) <= TOL * max(abs(v), 1)


@pytest.mark.parametrize("m,n,seed", [(3, 6, 0), (5, 9, 1), (8, 14, 2), (12, 20, 0), (20, 36, 3)])
def test_phase1_matches_float_oracle(m, n, seed):
    from linear_program_solver_b200 import Simplex
    from oracle import lpsol_oracle as O
    c, A, b = problems.equality_lp(m, n, seed)
    try:
        want = O.simplex(c, A, b, num=float)
    except (ValueError, IndexError) as e:
        with pytest.raises(ValueError if isinstance(e, ValueError) else IndexError):
            Simplex(_tab(c, A, b))
        return
    t = _tab(c, A, b)
    s = Simplex(t)
    assert t.phase2_on_device == 1                       # simplex.py:95-105 ran as `lps_phase2_f64`, not through the host
    n1 = len(want["phase1_pivots"])
    assert [list(p) for p in s.pivot_log] == [list(p) for p in want["phase1_pivots"]]
    assert s.getBasicSequence() == want["initial_basis"]
    s.solve()
    assert [list(p) for p in s.pivot_log[n1:]] == [list(p) for p in want["pivots"]]
    assert s.getBasicSequence() == want["basis"]
    assert float(s.getObjValue()).hex() == float(want["objective"]).hex()
    assert t.getTableauSize() == (m, n)


@pytest.mark.parametrize("m,n,k,seed", [(6, 10, 3, 0), (150, 260, 90, 1), (700, 1500, 400, 2)])
def test_leave_phase1_kernels_match_the_row_operations(m, n, k, seed):
    """`lps_phase2_f64` (cut of the trailing columns + objective restore) against the oracle's restatement of
    simplex.py:95-105 -- `_c = _c[:n]`, setZ / setC, rowSubFromObj per basic row -- bit for bit, on a tableau that a
    few pivots brought to a basis inside the first n columns.  (The float-rebound reference itself leaves phase 1 only
    on tiny inputs: its `z != 0` test is exact.)"""
    from linear_program_solver_b200 import Tableau
    from oracle import lpsol_oracle as O
    rng = np.random.default_rng(seed)
    grid = rng.integers(-9, 10, size=(m + 1, n + k + 1)).astype(np.float64)
    grid[:m, n - m:n] = np.eye(m)                          # a unit basis inside the first n columns
    grid[m, n - m:n] = 0.0
    t = Tableau.from_grid(grid.copy(), mode="fp64")
    ref = [[float(x) for x in row] for row in grid]
    basis = list(range(n - m, n))
    for step in range(min(m, 5)):                          # a few pivots move the basis into the structural columns
        r, c = (step * 7) % m, step
        if ref[r][c] == 0.0:
            continue
        t.pivot(r, c)
        O.pivot(ref, r, c)
        basis[r] = c
    # In floating point the basis columns are not exact unit vectors (rowDiv multiplies by 1/a_rc): plant a few
    # rounding-sized entries in them, above and below the diagonal, so that the multiplier of a row depends on the
    # steps before it -- the reference reads the CURRENT reduced cost (simplex.py:104-105)
    with t._eng.on_stream():
        for q in range(min(m, 40)):
            i, kb = (q * 5 + 1) % m, (q * 3) % m
            if i == kb:
                continue
            val = float((q % 7 - 3) * 2.0 ** -48)
            ref[i][basis[kb]] = val
            t._eng.grid[i, basis[kb]] = val
    t._eng.sync()
    cost = [float(v) for v in rng.integers(-20, 21, size=n)]
    cost[basis[0]] = 0.0                                   # a zero multiplier: the row is skipped (tableau.py:285)
    z = 12.5
    assert t._leave_phase1(n, cost, z, basis) and t.phase2_on_device == 1
    want = [row[:n] + [row[n + k]] for row in ref]
    want[m] = cost + [-z]
    for i, j in enumerate(basis):
        O.row_axpy(want, m, i, -want[m][j])
    got = t._eng.download()
    assert t.getTableauSize() == (m, n) and got.shape == (m + 1, n + 1)
    assert got.tobytes() == np.array(want, dtype=np.float64).tobytes()
    # the handle keeps working at its new width: one more pivot, compared again
    r = 1 % m
    c = next(j for j in range(n) if want[r][j] != 0.0 and j not in basis)
    t.pivot(r, c)
    O.pivot(want, r, c)
    assert t._eng.download().tobytes() == np.array(want, dtype=np.float64).tobytes()
    assert t.getZ() == -want[m][n]


def test_special_cases_error_behaviour():
    from linear_program_solver_b200 import Simplex
    sp = load_golden("special")
    # infeasible: ValueError with the reference's message; tableau keeps the artificial columns
    t = _tab([0, 0], [[1, 1], [1, 1]], [1, 2])
    with pytest.raises(ValueError, match="infeasible problem, artificial opt = 1"):
        Simplex(t)
    assert sp["infeasible"]["ctor_error"][0] == "ValueError" and t.getNumVars() == 4
    # unbounded: find* return the literal, solve() asserts
    t = _tab([-1, 0], [[-1, 1]], [1])
    s = Simplex(t)
    assert s.findPivotStandard() == "unbounded" == sp["unbounded"]["findPivotStandard"]
    assert s.findPivotMinIndex() == "unbounded"
    with pytest.raises(AssertionError, match="unbounded artificial problem"):
        s.solve()
    # redundant equality row: the reference fails with IndexError
    with pytest.raises(IndexError):
        Simplex(_tab([1, 1, 1], [[1, 1, 0], [2, 2, 0], [0, 1, 1]], [2, 4, 3]))
    # Beale cycling LP: stale stuck counter then the switch to Bland
    c, A, b = problems.BEALE
    s = Simplex(_tab(c, A, b))
    s.solve()
    assert [list(p) for p in s.pivot_log] == sp["beale_float"]["pivots"]
    assert float(s.getObjValue()).hex() == sp["beale_float"]["objective"]


def test_find_pivot_rules_and_checked_pivot():
    from linear_program_solver_b200 import Simplex
    from oracle import lpsol_oracle as O
    c, A, b = problems.dense_lp(12, 5)
    s = Simplex(_tab(c, A, b))
    grid = O.make_grid(c, A, b, num=float)
    assert s.findPivotStandard() == O.find_pivot(grid, O.RULE_DANTZIG)
    assert s.findPivotMinIndex() == O.find_pivot(grid, O.RULE_BLAND)
    r, cc = s.findPivotStandard(True)
    O.pivot(grid, r, cc)
    assert s.getBasicSequence()[r] == cc
    assert s._tab.getA() == [row[:-1] for row in grid[:-1]]
    r2, c2 = s.findPivotStandard()
    col = [row[c2] for row in grid[:-1]]
    lowest = min(grid[i][-1] / col[i] for i in range(12) if col[i] > 0)
    bad = next(i for i in range(12) if col[i] > 0 and grid[i][-1] / col[i] > lowest)     # positive entry, not the minimum ratio
    with pytest.raises(ValueError, match=f"bad pivot by min ratio test, r = {bad}, c = {c2}"):
        s.pivot(bad, c2)
    zero = next(i for i in range(12) if col[i] == 0.0)             # the reference divides first (simplex.py:213)
    with pytest.raises(ZeroDivisionError):
        s.pivot(zero, c2)
    assert all(isinstance(p, tuple) for p in s.findPivotAll())
    assert s.findPivotMaxIncrease() not in ("optimal", "unbounded")


@pytest.mark.parametrize("m,row0,rows", [(40, 0, 40), (64, 16, 24), (300, 123, 100)])
def test_device_fill_matches_hashed_family(m, row0, rows):
    from linear_program_solver_b200.engine import F64Engine
    from oracle import c_oracle as C
    e = F64Engine(rows, 2 * m)
    e.set_shard(row0, m, 1, 2) if rows != m else None
    e.fill_dense(7, 1, 9)
    got = e.download()
    want = problems.hashed_dense_grid(m, 7, 1, 9, row0, rows).astype(np.float64)
    assert got.tobytes() == want.tobytes()
    T = np.zeros_like(want)
    C.fill_dense(T, m, row0, 7, 1, 9)
    assert T.tobytes() == want.tobytes()
    e.close()


def test_stepping_api_equals_solve():
    from linear_program_solver_b200.sharded import ShardedF64Solver
    from oracle import c_oracle as C
    m = 96
    T = problems.hashed_dense_grid(m, 3).astype(np.float64)
    sv = ShardedF64Solver(m, 2 * m, burst=5)
    sv.eng.fill_dense(3, 1, 9)
    basis = np.arange(m, 2 * m, dtype=np.int32)
    st, k, trace = sv.solve(basis, None, trace_cap=4096)
    ob = np.arange(m, 2 * m, dtype=np.int32)
    ost, ok, otrace = C.solve(T, ob)
    assert (st, k) == (1, ok) and trace.tolist() == otrace.tolist() and basis.tolist() == ob.tolist()
    assert sv.eng.download().tobytes() == T.tobytes()
    sv.close()


@pytest.mark.parametrize("m,seed,hi", [(12, 5, 9), (40, 1, 3), (150, 2, 9)])
def test_whole_tableau_rules_on_the_device_match_the_snapshot_evaluation(m, seed, hi):
    """findPivotMaxIncrease / findPivotAll (simplex.py:286-360): the device pass against the host evaluation that
    follows the reference line by line, at every state along a solve (small integer data: many ratio ties)."""
    from linear_program_solver_b200 import Simplex
    from oracle import lpsol_oracle as O
    c, A, b = problems.dense_lp(m, seed, 1, hi)
    s = Simplex(_tab(c, A, b))
    for step in range(3 * m):
        snap = [[float(x) for x in row] for row in s._tab._eng.download()]     # the oracle reads a snapshot
        want_max, want_all = O.find_pivot_max_increase(snap), O.find_pivot_all(snap)
        assert s.findPivotMaxIncrease() == want_max, step
        assert s.findPivotAll() == want_all, step
        if isinstance(want_max, str):
            assert want_max == "optimal"
            break
        nxt = s.findPivotMaxIncrease(True) if step % 2 else s.findPivotStandard(True)
        assert not isinstance(nxt, str)
    else:
        raise AssertionError("did not reach the optimum")


@pytest.mark.parametrize("rec", [r for r in load_golden("rules") if "float" in r], ids=lambda r: f"m{r['m']}s{r['seed']}")
def test_whole_tableau_rules_match_float_rebound_reference(rec):
    """The same rules against the reference's own answers (float-rebound run, tests/golden/rules.json)."""
    from linear_program_solver_b200 import Simplex
    c, A, b = problems.dense_lp(rec["m"], rec["seed"], rec["lo"], rec["hi"])
    s = Simplex(_tab(c, A, b))
    for step, want in enumerate(rec["float"]):
        got = s.findPivotMaxIncrease()
        assert (list(got) if isinstance(got, tuple) else got) == want["max_increase"], step
        assert [list(p) for p in s.findPivotAll()] == want["all"], step
        if isinstance(got, str):
            break
        nxt = s.findPivotMaxIncrease(True) if step % 2 else s.findPivotStandard(True)
        assert not isinstance(nxt, str)


def test_max_increase_reports_unbounded_and_optimal():
    from linear_program_solver_b200 import Simplex
    unb = _tab([-1, -2, 0], [[1, -1, 1]], [1])                  # column 1 is attractive and has no positive entry
    from oracle import lpsol_oracle as O
    s = Simplex(unb)
    assert s.findPivotMaxIncrease() == "unbounded" == O.find_pivot_max_increase(O.make_grid([-1, -2, 0], [[1, -1, 1]], [1], num=float))
    opt = _tab([1, 2, 0], [[1, 1, 1]], [1])
    s = Simplex(opt)
    g = O.make_grid([1, 2, 0], [[1, 1, 1]], [1], num=float)
    assert s.findPivotMaxIncrease() == "optimal" == O.find_pivot_max_increase(g)
    assert s.findPivotAll() == O.find_pivot_all(g)
