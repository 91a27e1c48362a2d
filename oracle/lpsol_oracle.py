"""CPU oracle for the simplex pivot path of `lpsol` -- TEST INFRASTRUCTURE ONLY.

This file restates, from scratch, the algorithm of the reference's hot path
(`/root/reference/lpsol/tableau.py` and `simplex.py`) on a single dense grid:

    grid[i][j]   i in 0..m-1, j in 0..n-1   constraint matrix  (reference `_a`)
    grid[i][n]                              right-hand side    (reference `_b`)
    grid[m][j]                              reduced costs      (reference `_c`)
    grid[m][n]                              the `_z` cell = minus the objective

It is generic over the number type: `fractions.Fraction` gives the exact path,
`float` gives the path the reference takes when its module globals
`Frac/ZERO/ONE` are rebound to float.  The order and
rounding of every arithmetic operation follows the reference so that the float
instantiation is bit-identical to that rebound run.

Parity status: PINNED.  `tests/test_oracle.py` checks this restatement against
(a) the reference's only known-answer test on the path
(`test_tableau.py:220-227`), (b) the known answers SURVEY.md section 8c records,
(c) `tests/golden/*.json`, which `tools/make_golden.py` produced by running
the unmodified reference in the build container, and (d) the irregular corpus
`tests/golden/wild.json` (`tools/make_golden_wild.py`: 400 small LPs with every
constructor / solve outcome, exact and float-rebound; `tests/test_wild.py`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu-baseline /
`--impl reference` legs may import this module.  The product package
(`linear_program_solver_b200`) never does.
"""
from __future__ import annotations

from fractions import Fraction
from typing import Callable, List, Optional, Sequence, Tuple, Union

Number = Union[Fraction, float]
Grid = List[List[Number]]

OPTIMAL = "optimal"
UNBOUNDED = "unbounded"

RULE_DANTZIG = 0  # reference `findPivotStandard`
RULE_BLAND = 1    # reference `findPivotMinIndex`


def make_grid(c: Sequence, A: Sequence[Sequence], b: Sequence, z=0, num: Callable = Fraction) -> Grid:
    """Build the grid from (c, A, b, objective z); `num` is Fraction or float.

    The corner stores -z exactly like `Tableau.setZ` (tableau.py:128-130).
    """
    m = len(A)
    grid = [[num(x) for x in A[i]] + [num(b[i])] for i in range(m)]
    grid.append([num(x) for x in c] + [-num(z)])
    return grid


def shape(grid: Grid) -> Tuple[int, int]:
    return len(grid) - 1, len(grid[0]) - 1


def objective(grid: Grid) -> Number:
    """`Tableau.getZ` (tableau.py:82-84): the negated corner cell."""
    return -grid[-1][-1]


# --------------------------------------------------------------------------
# a2-a5: the pivot and the row operations it is made of
# --------------------------------------------------------------------------

def row_scale(grid: Grid, r: int, mult: Number) -> None:
    """`Tableau.rowMult` (tableau.py:254-260): row r (and b_r) times mult; skipped when mult == 1."""
    if mult == 1:
        return
    row = grid[r]
    n = len(row) - 1
    row[n] = row[n] * mult                 # b_r first, like the reference
    for j in range(n):
        row[j] = row[j] * mult


def row_axpy(grid: Grid, rd: int, rs: int, mult: Number) -> None:
    """`Tableau.rowAdd` / `rowAddToObj` (tableau.py:269-289): row rd += mult * row rs.

    Skipped entirely when mult == 0 (tableau.py:272, :285).  With rd == m this
    is the objective-row form, whose RHS update is the `_z` cell.
    Each element is product-then-sum (two roundings in float), never fused.
    """
    if mult == 0:
        return
    dst, src = grid[rd], grid[rs]
    n = len(dst) - 1
    dst[n] = dst[n] + mult * src[n]
    for j in range(n):
        dst[j] = dst[j] + mult * src[j]


def pivot(grid: Grid, r: int, c: int) -> None:
    """`Tableau.pivot` (tableau.py:295-308).

    1. zero pivot -> ZeroDivisionError('zero pivot r,c')
    2. row r times the reciprocal of the pivot (rowDiv -> rowMult(1/p), :262-267)
    3. objective row += (-c_c) * row r                (:303)
    4. every other row rr += (-a_rr,c) * row r, the multiplier being read from
       the not-yet-modified row rr                   (:305-308)
    """
    m, n = shape(grid)
    p = grid[r][c]
    if p == 0:
        raise ZeroDivisionError(f"zero pivot {r},{c}")
    one = type(p)(1)
    row_scale(grid, r, one / p)
    row_axpy(grid, m, r, -grid[m][c])
    for rr in range(m):
        if rr != r:
            row_axpy(grid, rr, r, -grid[rr][c])


# --------------------------------------------------------------------------
# a6, a7, a10: pricing and the ratio test
# --------------------------------------------------------------------------

def price(grid: Grid, rule: int) -> int:
    """Entering column, or -1 when every reduced cost is >= 0.

    Dantzig (simplex.py:262-267): most negative cost, strict `<` so the lowest
    index wins among equal minima.  Bland (simplex.py:228-232): first negative.
    """
    m, n = shape(grid)
    cost = grid[m]
    best = -1
    for j in range(n):
        if cost[j] >= 0:
            continue
        if rule == RULE_BLAND:
            return j
        if best == -1 or cost[j] < cost[best]:
            best = j
    return best


def ratio_test(grid: Grid, c: int) -> int:
    """Leaving row for column c, or -1 when no entry is positive.

    simplex.py:272-279 (identical at :237-244): rows with a_ic > 0 only, ratio
    b_i / a_ic by true division, strict `<` so the lowest row index wins ties.
    """
    m, n = shape(grid)
    best, best_ratio = -1, None
    for i in range(m):
        a = grid[i][c]
        if a <= 0:
            continue
        ratio = grid[i][n] / a
        if best_ratio is None or ratio < best_ratio:
            best, best_ratio = i, ratio
    return best


def find_pivot(grid: Grid, rule: int):
    """`findPivotStandard` / `findPivotMinIndex` with do_pivot=False: (i, j) or a status string."""
    c = price(grid, rule)
    if c < 0:
        return OPTIMAL
    r = ratio_test(grid, c)
    if r < 0:
        return UNBOUNDED
    return r, c


def find_pivot_max_increase(grid: Grid):
    """`Simplex.findPivotMaxIncrease` with do_pivot=False (simplex.py:286-328): among the columns with c_j < 0,
    the pivot of largest objective decrease -c_j * min ratio; within a column the first row of minimum ratio
    (the tie clause of :316-318 never fires: equal ratios give equal decreases), across columns the first
    column among equals (strict ">", :323); 'unbounded' as soon as a column with c_j < 0 has no positive entry."""
    m, n = shape(grid)
    best = None
    si = sj = -1
    any_negative = False
    for j in range(n):
        cj = grid[m][j]
        if cj >= 0:
            continue
        any_negative = True
        ratio = colinc = None
        isel = -1
        for i in range(m):
            aij = grid[i][j]
            if aij <= 0:
                continue
            r = grid[i][n] / aij
            if ratio is None or r < ratio:
                ratio, colinc, isel = r, -cj * r, i
            elif r == ratio and colinc is not None and -cj * r > colinc:
                colinc, isel = -cj * r, i
        if colinc is None:
            return UNBOUNDED
        if best is None or colinc > best:
            best, si, sj = colinc, isel, j
    if not any_negative:
        return OPTIMAL
    return si, sj


def find_pivot_all(grid: Grid) -> List[Tuple[int, int]]:
    """`Simplex.findPivotAll` (simplex.py:330-360): for every column, all rows of minimum ratio b_i / a_ij over
    a_ij > 0, rows ascending, column after column."""
    m, n = shape(grid)
    out: List[Tuple[int, int]] = []
    for j in range(n):
        ratio = None
        picks: List[Tuple[int, int]] = []
        for i in range(m):
            aij = grid[i][j]
            if aij <= 0:
                continue
            r = grid[i][n] / aij
            if ratio is None or r < ratio:
                ratio, picks = r, [(i, j)]
            elif r == ratio:
                picks.append((i, j))
        out += picks
    return out


def is_optimal(grid: Grid) -> bool:
    """`Tableau.isOptimal` (tableau.py:500-502)."""
    m, n = shape(grid)
    return all(grid[m][j] >= 0 for j in range(n))


def is_canonical(grid: Grid, bcols: Optional[List[int]] = None) -> bool:
    """`Tableau.isCanonical` (tableau.py:466-496).

    False immediately (bcols untouched) when some b_i < 0.  A column counts as
    basic for row i when its cost is exactly 0, its only nonzero is an exact 1,
    and that 1 sits in row i; bcols[i] receives the lowest such column or -1.
    """
    m, n = shape(grid)
    if any(grid[i][n] < 0 for i in range(m)):
        return False
    found = [-1] * m
    for j in range(n):
        if grid[m][j] != 0:
            continue
        hit, clean = -1, True
        for i in range(m):
            v = grid[i][j]
            if v == 0:
                continue
            if v == 1 and hit == -1:
                hit = i
            else:
                clean = False
                break
        if clean and hit != -1 and found[hit] == -1:
            found[hit] = j
    if bcols is not None:
        for i in range(m):
            bcols[i] = found[i]
    return all(j != -1 for j in found)


# --------------------------------------------------------------------------
# a8, a9: the solve loop
# --------------------------------------------------------------------------

class UnboundedError(AssertionError):
    """The reference asserts on an unbounded LP (simplex.py:125, :140)."""


def solve(grid: Grid, basis: List[int], trace: Optional[list] = None, max_pivots: Optional[int] = None,
          stuck_limit: Optional[int] = None) -> int:
    """`Simplex.solve` (simplex.py:110-148).  Mutates grid and basis; returns the pivot count.

    Loop A prices with Dantzig while the stuck counter is below m+n; the counter
    goes up when the objective equals its value AT ENTRY (the reference never
    refreshes `obj_val`, simplex.py:118/:134 -- SURVEY F5a) and resets
    otherwise.  Loop B prices with Bland until optimal and never goes back.
    `trace` receives (row, column, rule) per pivot.  `max_pivots` (not in the
    reference) stops after a prefix, for timing large cases.  `stuck_limit` (not in the reference, whose
    limit is m+n) lets tests reach loop B on LPs that do not cycle.
    """
    m, n = shape(grid)
    z_entry = objective(grid)
    stuck = 0
    done = 0
    if is_optimal(grid):
        return 0
    rule = RULE_DANTZIG
    while True:
        if rule == RULE_DANTZIG and not (stuck < (m + n if stuck_limit is None else stuck_limit)):
            rule = RULE_BLAND
        if max_pivots is not None and done >= max_pivots:
            return done
        found = find_pivot(grid, rule)
        if found == UNBOUNDED:
            raise UnboundedError("unbounded artificial problem (internal error)")
        if found == OPTIMAL:
            break
        r, c = found
        pivot(grid, r, c)
        basis[r] = c
        done += 1
        if trace is not None:
            trace.append((r, c, rule))
        if rule == RULE_DANTZIG:
            z = objective(grid)
            assert z <= z_entry, "objective value increased (internal error)"
            stuck = stuck + 1 if z == z_entry else 0
    assert is_optimal(grid), "solver failed (internal error)"
    return done


# --------------------------------------------------------------------------
# a12: phase 1 (constructor of the reference's Simplex)
# --------------------------------------------------------------------------

class InfeasibleError(ValueError):
    pass


def find_bfs(grid: Grid, trace: Optional[list] = None, stuck_limit: Optional[int] = None) -> Tuple[Grid, List[int]]:
    """`Simplex._find_bfs` (simplex.py:36-108).  Returns (grid, basis); the grid may be re-shaped.

    Rows with b_i < 0 are negated; if the tableau is then canonical the basis
    `isCanonical` reports is used.  Otherwise artificial columns (cost 1) are
    appended for exactly the rows that lack a basic column, the artificial
    objective is priced out, `solve` runs, a nonzero optimum raises
    ValueError('infeasible problem, artificial opt = z'), basic artificials are
    pivoted out on the first nonzero original column (the pivot may be
    negative) or their row is dropped, artificial columns are cut and the
    original objective is restored and priced out against the basis.

    The reference resets `_m` to the OLD row count after dropping a dependent
    row (simplex.py:93) and then fails in its closing `isCanonical` (SURVEY
    F5b): `stale_row_count_failure` reproduces which exception that is.
    """
    m, n = shape(grid)
    num = type(grid[0][0])
    for i in range(m):
        if grid[i][n] < 0:
            row_scale(grid, i, -num(1))
    basis = [-1] * m
    if is_canonical(grid, basis):
        return grid, basis
    saved_cost = grid[m][:n]
    saved_z = objective(grid)
    missing = [i for i in range(m) if basis[i] == -1]
    k = len(missing)
    # widen: artificial columns sit at n..n+k-1, the RHS moves to the new last column
    for i in range(m + 1):
        row = grid[i]
        rhs = row[n]
        row[n:] = [num(0)] * k + [rhs]
    for j in range(n):
        grid[m][j] = num(0)
    grid[m][n + k] = -num(0)
    for idx, i in enumerate(missing):
        grid[m][n + idx] = num(1)
        grid[i][n + idx] = num(1)
        row_axpy(grid, m, i, -num(1))
        basis[i] = n + idx
    solve(grid, basis, trace, stuck_limit=stuck_limit)
    z = objective(grid)
    if z != 0:
        raise InfeasibleError(f"infeasible problem, artificial opt = {z}")
    keep = [True] * m
    for i in range(m):
        if basis[i] < n:
            continue
        assert grid[i][n + k] == 0, "invalid artificial solution (internal error)"
        col = next((j for j in range(n) if grid[i][j] != 0), -1)
        if col == -1:
            keep[i] = False
        else:
            pivot(grid, i, col)
            basis[i] = col
            if trace is not None:
                trace.append((i, col, 2))
    m_old = m
    if not all(keep):
        # simplex.py:89-93 filters `_bfs`, `_b`, `_a` but re-installs the OLD row count
        grid[:] = [row for i, row in enumerate(grid) if i == m or keep[i]]
        basis[:] = [j for i, j in enumerate(basis) if keep[i]]
        m = len(basis)
    # cut the artificial columns, restore and price out the original objective
    for i in range(m + 1):
        row = grid[i]
        rhs = row[n + k]
        row[n:] = [rhs]
    grid[m][n] = -saved_z
    for j in range(n):
        grid[m][j] = saved_cost[j]
    for i, j in enumerate(basis):
        row_axpy(grid, m, i, -grid[m][j])
    if m != m_old:
        stale_row_count_failure(grid, m_old)
    assert is_canonical(grid), "tableau not canonical (internal error)"
    return grid, basis


def stale_row_count_failure(grid: Grid, m_old: int) -> None:
    """What simplex.py:106 does after dependent rows were dropped: `isCanonical` (tableau.py:466-496)
    walks `range(m_old)` over the FEWER rows that are left.  A zero-cost column whose scan for its 1 (tableau.py:482-485)
    or whose all-others-zero check (:487-488) runs past the last row dies with IndexError; when no column does (every
    zero-cost column holds a 1 and another nonzero, or no reduced cost is 0) the dropped rows have no basic column and
    the constructor's assert fires instead.  Always raises."""
    m, n = shape(grid)
    assert m < m_old
    if not any(grid[i][n] < 0 for i in range(m)):
        for j in range(n):
            if grid[m][j] != 0:
                continue
            one = next((i for i in range(m) if grid[i][j] == 1), -1)
            if one == -1 or all(i == one or grid[i][j] == 0 for i in range(m)):
                raise IndexError("list index out of range")
    raise AssertionError("tableau not canonical (internal error)")


def simplex(c, A, b, z=0, num: Callable = Fraction, max_pivots: Optional[int] = None, stuck_limit: Optional[int] = None):
    """`Simplex(tab); solve()` end to end.  Returns a dict of everything parity tests compare."""
    grid = make_grid(c, A, b, z, num)
    phase1: list = []
    grid, basis = find_bfs(grid, phase1, stuck_limit)
    initial_basis = list(basis)
    trace: list = []
    solve(grid, basis, trace, max_pivots, stuck_limit)
    m, n = shape(grid)
    return {
        "grid": grid,
        "initial_basis": initial_basis,
        "basis": basis,
        "phase1_pivots": phase1,
        "pivots": trace,
        "objective": objective(grid),
        "bfs": {basis[i]: grid[i][n] for i in range(m)},
    }


# --------------------------------------------------------------------------
# fraction-free (Bareiss) integer form -- the arithmetic the exact GPU mode uses
# --------------------------------------------------------------------------

def bareiss_pivot(N: List[List[int]], D: int, r: int, c: int) -> int:
    """One fraction-free pivot on the integer grid N with common denominator D.

    SURVEY.md section 8a note (verified there against the reference at every
    pivot): for i != r, N_ij <- (p*N_ij - N_ic*N_rj) / D with p = N_rc, the
    division being exact; row r is unchanged; the new denominator is p.  The
    reference's entry is then Fraction(N_ij, D) everywhere.
    """
    p = N[r][c]
    if p == 0:
        raise ZeroDivisionError(f"zero pivot {r},{c}")
    prow = N[r]
    for i, row in enumerate(N):
        if i == r:
            continue
        f = row[c]
        if f == 0:
            if p != D:
                for j in range(len(row)):
                    q, rem = divmod(p * row[j], D)
                    assert rem == 0
                    row[j] = q
            continue
        for j in range(len(row)):
            q, rem = divmod(p * row[j] - f * prow[j], D)
            assert rem == 0
            row[j] = q
    return p
