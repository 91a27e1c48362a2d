"""Seeded synthetic LP families used by the parity tests and by bench.py.

The dense family is the one SURVEY.md section 8(d) defines (and BASELINE.md timed the
reference on): m structural + m slack columns (n = 2m), strictly positive integer
constraint matrix, so x = 0 is feasible, the LP is bounded and the tableau is
already canonical with the slack basis (phase 1 is a no-op).

Everything here is plain Python / numpy: it only *describes* inputs.  Nothing in
this module computes a pivot.
"""
from __future__ import annotations

import random
from fractions import Fraction
from typing import List, Sequence, Tuple

Grid = Tuple[List[int], List[List[int]], List[int]]


def dense_lp(m: int, seed: int = 0, lo: int = 1, hi: int = 9) -> Grid:
    """Return (c, A, b) of the seeded dense feasible family.

    Draw order (fixed, part of the fixture contract): all of c, then A row by
    row, then b.  c has m negative structural costs followed by m zeros; A is
    [random m x m | I_m]; b[i] in [hi*m//4, hi*m//2].
    """
    rng = random.Random(seed)
    c = [-rng.randint(lo, hi) for _ in range(m)] + [0] * m
    A = []
    for i in range(m):
        row = [rng.randint(lo, hi) for _ in range(m)] + [0] * m
        row[m + i] = 1
        A.append(row)
    b = [rng.randint(hi * m // 4, hi * m // 2) for _ in range(m)]
    return c, A, b


def dense_lp_grid(m: int, seed: int = 0, lo: int = 1, hi: int = 9) -> List[List[int]]:
    """Same family as one (m+1) x (n+1) grid in device layout.

    Rows 0..m-1 are the constraints, row m is the objective row; columns
    0..n-1 are the variables, column n is the right-hand side.  Cell (m, n) is
    the reference's `_z` cell (= minus the objective value, 0 here).
    """
    c, A, b = dense_lp(m, seed, lo, hi)
    grid = [A[i] + [b[i]] for i in range(m)]
    grid.append(c + [0])
    return grid


def equality_lp(m: int, n: int, seed: int = 0, lo: int = -5, hi: int = 9) -> Grid:
    """Non-canonical family (equality rows, no slack columns) for phase 1.

    A feasible point x0 >= 0 is drawn first and b = A x0, so the LP is feasible;
    costs are positive so it is bounded.  Rows are independent with
    probability ~1 for n > m; b may come out negative, which exercises the
    reference's sign normalisation (simplex.py:43-45).
    """
    rng = random.Random(seed)
    A = [[rng.randint(lo, hi) for _ in range(n)] for _ in range(m)]
    x0 = [rng.randint(0, 4) for _ in range(n)]
    b = [sum(A[i][j] * x0[j] for j in range(n)) for i in range(m)]
    c = [rng.randint(1, 9) for _ in range(n)]
    return c, A, b


def tie_heavy_grid(m: int, parts: int, seed: int = 0):
    """Canonical LP whose ratio tests tie ACROSS shards: the structural part of row i repeats every m // parts rows
    with the same right-hand side (small integers, so the ratios are exactly equal doubles); the twins of a leaving row
    drop to b = 0, so degenerate pivots follow.  The leaving row must be the lowest GLOBAL row among the ties
    (simplex.py:277), whichever shard holds it."""
    import numpy as np
    rng = random.Random(seed)
    period = m // parts
    base = [[rng.randint(1, 3) for _ in range(m)] for _ in range(period)]
    rhs = [rng.randint(1, 4) * 6 for _ in range(period)]
    T = np.zeros((m + 1, 2 * m + 1))
    for i in range(m):
        T[i, :m] = base[i % period]
        T[i, m + i] = 1.0
        T[i, 2 * m] = rhs[i % period]
    T[m, :m] = [-rng.randint(1, 3) for _ in range(m)]
    return T


def irregular_lp(rng: random.Random, m: int = 0, n: int = 0):
    """One LP of the irregular family (tests/golden/wild.json, tools/make_golden_wild.py): mixed signs, zeros, sometimes
    rational data, negative / zero right-hand sides, partial slack bases, a bounding row, a dependent row -- feasible,
    infeasible, unbounded or rank deficient as it comes.  The draw order is part of the fixture contract; `m`, `n`
    override the drawn size (1..7 x 1..10) without disturbing it.  Returns (c, A, b) of Fractions; the final size may
    exceed (m, n) by the slack columns and the bounding row."""
    m0, n0 = rng.randint(1, 7), rng.randint(1, 10)
    m, n = m or m0, n or n0
    density = rng.choice([0.3, 0.6, 0.9, 1.0])
    rational = rng.random() < 0.3
    span = rng.choice([2, 3, 9])

    def num(nonneg=False):
        if rng.random() > density:
            return Fraction(0)
        v = Fraction(rng.randint(0 if nonneg else -span, span))
        if rational and rng.random() < 0.5:
            v /= rng.choice([2, 3, 4, 6])
        return v

    A = [[num() for _ in range(n)] for _ in range(m)]
    style = rng.choice([0, 0, 0, 1, 2, 2, 3])
    if style == 0:                                   # right-hand side of a known non-negative point: feasible
        x = [Fraction(rng.randint(0, 3)) for _ in range(n)]
        b = [sum(a * v for a, v in zip(row, x)) for row in A]
    elif style == 1:
        b = [Fraction(rng.randint(-6, 6)) for _ in range(m)]
    elif style == 2:
        b = [Fraction(rng.randint(0, 8)) for _ in range(m)]
    else:
        b = [Fraction(0) if rng.random() < 0.5 else Fraction(rng.randint(1, 5)) for _ in range(m)]
    if rng.random() < 0.5:                           # slack columns on some rows (a partial or complete slack basis)
        for i in range(m):
            if rng.random() < 0.7:
                for r in range(m):
                    A[r].append(Fraction(1 if r == i else 0))
                n += 1
    if rng.random() < 0.4:                           # a bounding row  sum_j w_j x_j + s = K,  w > 0: no unbounded ray
        for r in range(m):
            A[r].append(Fraction(0))
        A.append([Fraction(rng.randint(1, 3)) for _ in range(n)] + [Fraction(1)])
        b.append(Fraction(rng.randint(0, 40)))
        m, n = m + 1, n + 1
    if m >= 2 and rng.random() < 0.2:                # a dependent row: copy or multiple of another one
        src, dst = rng.sample(range(m), 2)
        k = rng.choice([1, 2, -1, Fraction(1, 2)])
        A[dst] = [k * v for v in A[src]]
        b[dst] = k * b[src] + (0 if rng.random() < 0.7 else 1)
    sign = rng.choice([-1, -1, 1, 0])
    c = [Fraction(rng.randint(-span, span)) if sign == 0 else sign * Fraction(rng.randint(0, span)) for _ in range(n)]
    return c, A, b


def _mix(x):
    import numpy as np
    x = x + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def hashed_dense_grid(m: int, seed: int = 0, lo: int = 1, hi: int = 9, row0: int = 0, m_local: int = None):
    """The dense family from a counter-based hash instead of a sequential generator, as an
    int64 (m_local+1, 2m+1) grid in device layout: rows [row0, row0+m_local) plus the cost row.

    This is the function the CUDA library's `lps_fill_dense_f64` evaluates on the device (any
    shard can be generated independently); same shape and ranges as `dense_lp`.
    """
    import numpy as np
    if m_local is None:
        m_local = m - row0
    special = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        s2 = _mix(np.uint64(seed))
        rows = np.concatenate([np.arange(row0, row0 + m_local, dtype=np.uint64), [special]])
        key = s2 + rows[:, None] * np.uint64(0x100000001B3)
        draw = (_mix(key + np.arange(m, dtype=np.uint64)[None, :]) >> np.uint64(33)) % np.uint64(hi - lo + 1)
        rhs = (_mix(key[:, 0] + special) >> np.uint64(33)) % np.uint64(hi * m // 2 - hi * m // 4 + 1)
    grid = np.zeros((m_local + 1, 2 * m + 1), dtype=np.int64)
    grid[:, :m] = lo + draw.astype(np.int64)
    grid[m_local, :m] *= -1
    idx = np.arange(m_local)
    grid[idx, m + row0 + idx] = 1
    grid[:m_local, 2 * m] = hi * m // 4 + rhs[:m_local].astype(np.int64)
    return grid


#: Beale's cycling example: exercises the stale
#: objective "stuck" counter and the switch from Dantzig to Bland pricing.
BEALE = (
    [Fraction(0), Fraction(0), Fraction(0), Fraction(-3, 4), Fraction(20), Fraction(-1, 2), Fraction(6)],
    [
        [Fraction(1), Fraction(0), Fraction(0), Fraction(1, 4), Fraction(-8), Fraction(-1), Fraction(9)],
        [Fraction(0), Fraction(1), Fraction(0), Fraction(1, 2), Fraction(-12), Fraction(-1, 2), Fraction(3)],
        [Fraction(0), Fraction(0), Fraction(1), Fraction(0), Fraction(0), Fraction(1), Fraction(0)],
    ],
    [Fraction(0), Fraction(0), Fraction(1)],
)

#: The textbook LP of the reference's own test fixture (test_tableau.py:9-29).
TEXTBOOK = (
    [-40, -30, 0, 0],
    [[1, 1, 1, 0], [2, 1, 0, 1]],
    [12, 16],
)


def var_names(m: int) -> List[str]:
    """x0..x{m-1}, s0..s{m-1} for the dense family."""
    return [f"x{j}" for j in range(m)] + [f"s{i}" for i in range(m)]


def to_grid(c: Sequence, A: Sequence[Sequence], b: Sequence, z=0):
    """(c, A, b[, z]) -> device-layout grid (objective last row, RHS last column).

    `z` is the objective value as the user sets it (`setZ`); the grid's corner
    cell stores -z like the reference's `_z` field (tableau.py:128-130).
    """
    grid = [list(A[i]) + [b[i]] for i in range(len(A))]
    grid.append(list(c) + [-z])
    return grid
