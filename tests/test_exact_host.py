"""Host-side pieces of the exact mode (no GPU): prime set, Hadamard bound, limb encoding, mixed-radix
decoding, and the fraction-free integer form the kernels implement, checked against the oracle."""
import math
import random
from fractions import Fraction

import numpy as np
import pytest

from linear_program_solver_b200 import exact as X
from linear_program_solver_b200 import problems
from oracle import lpsol_oracle as O


def test_prime_set_is_descending_31_bit_primes():
    sympy = pytest.importorskip("sympy")
    ps = X.primes_below_2_31(40)
    assert ps[0] == 2147483647 and ps == sorted(ps, reverse=True) and len(set(ps)) == 40
    assert all(sympy.isprime(p) and (1 << 30) < p < (1 << 31) for p in ps)
    # no prime skipped between the first and the last
    assert sum(1 for q in range(ps[-1], ps[0] + 1, 2) if sympy.isprime(q)) == 40
    assert X.primes_below_2_31(3, skip=[ps[1]]) == [ps[0], ps[2], ps[3]]


def test_primes_for_bits_cover_the_symmetric_range():
    for bits in (1, 30, 31, 62, 400, 533, 15970):
        ps = X.primes_for_bits(bits)
        assert math.prod(ps) > (1 << (bits + 2))
        assert math.prod(ps[:-1]) <= (1 << (bits + 2)) or len(ps) == 2


@pytest.mark.parametrize("m,seed,hi", [(6, 0, 9), (12, 1, 9), (20, 2, 127)])
def test_hadamard_bound_covers_every_bareiss_integer_of_a_solve(m, seed, hi):
    """|N_ij| and |D| along a whole reference solve stay below the bound that sizes the prime set."""
    c, A, b = problems.dense_lp(m, seed, 1, hi)
    grid = O.make_grid(c, A, b)
    ints = [[int(x) for x in row] for row in grid]
    bits = X.hadamard_bits(ints)
    assert abs(X.hadamard_bits_np(np.array(ints, dtype=np.int64)) - bits) <= 1
    N, D = [row[:] for row in ints], 1
    basis = list(range(m, 2 * m))
    trace = []
    O.solve(grid, basis, trace)
    for r, cc, _ in trace:
        D = O.bareiss_pivot(N, D, r, cc)
        assert abs(D).bit_length() <= bits
        assert max(abs(x).bit_length() for row in N for x in row) <= bits
    # and the integer form is the reference's tableau: entry == N_ij / D everywhere
    assert all(Fraction(N[i][j], D) == grid[i][j] for i in range(m + 1) for j in range(2 * m + 1))


def test_clear_denominators_and_limbs_roundtrip():
    rng = random.Random(5)
    grid = [[Fraction(rng.randint(-50, 50), rng.randint(1, 12)) for _ in range(7)] for _ in range(4)]
    ints, sigma = X.clear_denominators(grid)
    assert all(Fraction(ints[i][j], sigma[i]) == grid[i][j] for i in range(4) for j in range(7))
    assert all(s == math.lcm(*[x.denominator for x in row]) for s, row in zip(sigma, grid))
    vals = [0, 1, -1, (1 << 32) - 1, -(1 << 32), (1 << 64) - 1, -(1 << 200) + 12345, 1 << 1000]
    limbs, neg, nl = X.to_limbs(vals)
    back = [int.from_bytes(limbs[i].astype("<u4").tobytes(), "little") * (-1 if neg[i] else 1) for i in range(len(vals))]
    assert back == vals and nl == 32
    small, neg2, nl2 = X.to_limbs([5, -7, 0, (1 << 40) + 3])
    assert nl2 == 2 and small[3].tolist() == [3, 256] and neg2.tolist() == [0, 1, 0, 0]


def test_mixed_radix_decode():
    ps = X.primes_below_2_31(12)
    codec = X.MixedRadix(ps)
    rng = random.Random(1)
    M = math.prod(ps)
    vals = [0, 1, -1, ps[0], -ps[0] * ps[1] - 5, M // 2, -(M // 2)] + [rng.randrange(-M // 2, M // 2) for _ in range(20)]
    digits = np.zeros((len(ps), len(vals)), dtype=np.uint32)
    sign = np.zeros(len(vals), dtype=np.int8)
    for e, v in enumerate(vals):
        sign[e] = (v > 0) - (v < 0)
        a = abs(v)
        for k, p in enumerate(ps):
            digits[k, e] = a % p
            a //= p
    assert codec.decode(sign, digits) == vals


def _phase1_state(c, A, b):
    """Run the oracle's phase 1 (simplex.py:36-94) up to, but not including, the cut of the artificial columns, keeping
    the fraction-free integer state (N, D, alive) the device would hold next to the oracle's Fraction grid."""
    grid = O.make_grid(c, A, b)
    m, n = O.shape(grid)
    for i in range(m):
        if grid[i][n] < 0:
            O.row_scale(grid, i, Fraction(-1))
    basis = [-1] * m
    assert not O.is_canonical(grid, basis)
    cost, zcell = grid[m][:n], grid[m][n]
    missing = [i for i in range(m) if basis[i] == -1]
    k = len(missing)
    for i in range(m + 1):
        row = grid[i]
        row[n:] = [Fraction(0)] * k + [row[n]]
    for j in range(n):
        grid[m][j] = Fraction(0)
    grid[m][n + k] = Fraction(0)
    for idx, i in enumerate(missing):
        grid[m][n + idx] = Fraction(1)
        grid[i][n + idx] = Fraction(1)
        O.row_axpy(grid, m, i, Fraction(-1))
        basis[i] = n + idx
    ints, sigma = X.clear_denominators(grid)
    N, D, alive = [row[:] for row in ints], 1, [1] * (m + 1)
    trace = []
    O.solve(grid, basis, trace)
    assert O.objective(grid) == 0
    for i in range(m):
        if basis[i] >= n:
            col = next(j for j in range(n) if grid[i][j] != 0)
            O.pivot(grid, i, col)
            basis[i] = col
            trace.append((i, col, 2))
    for r, cc, _ in trace:
        D = O.bareiss_pivot(N, D, r, cc)
        alive[r] = 0
    den = lambda i: D * sigma[i] if alive[i] else D
    assert all(Fraction(N[i][j], den(i)) == grid[i][j] for i in range(m + 1) for j in range(n + k + 1))
    return grid, ints, sigma, N, D, alive, basis, cost, zcell, n, k


@pytest.mark.parametrize("m,n,seed,rational", [(5, 9, 0, False), (8, 14, 1, False), (12, 20, 2, False), (6, 11, 3, True), (10, 17, 4, True)])
def test_phase2_objective_formula_and_bounds(m, n, seed, rational):
    """What `lps_phase2_exact` computes -- N_mj = D C_j - sum_i G_i N_ij with the integers `ExactEngine.phase2`
    prepares -- is the reference's restored objective row (simplex.py:101-105) in the fraction-free form, it fits the
    column-weight bound handed to the device, and the following pivots stay integral."""
    c, A, b = problems.equality_lp(m, n, seed)
    if rational:
        rng = random.Random(seed)
        A = [[Fraction(v, rng.choice([1, 2, 3])) for v in row] for row in A]
        c = [Fraction(v, rng.choice([1, 5])) for v in c]
        b = [sum(A[i][j] * ((j * 7 + seed) % 5) for j in range(n)) for i in range(m)]
    grid, ints, sigma, N, D, alive, basis, cost, zcell, n, k = _phase1_state(c, A, b)
    eng = object.__new__(X.ExactEngine)
    eng.m, eng.n = m, n + k
    bits, info = eng._prepare_phase2(cost + [zcell], ints, sigma)
    assert bits >= X.hadamard_bits(ints) - 1
    C, sigma_m = info["C"], info["sigma_m"]
    assert all(Fraction(C[j], sigma_m) == (cost + [zcell])[j] for j in range(n + 1))
    G = []
    for i in range(m):
        q, rem = divmod(C[basis[i]], sigma[i] if alive[i] else 1)
        assert rem == 0
        G.append(q)
    cols = list(range(n)) + [n + k]                       # the columns that survive the cut
    new_obj = [D * C[jj] - sum(G[i] * N[i][j] for i in range(m)) for jj, j in enumerate(cols)]
    # the reference: cut, setZ / setC, rowSubFromObj per basic row
    want = [row[:n] + [row[n + k]] for row in grid]
    want[m] = cost + [zcell]
    for i, j in enumerate(basis):
        O.row_axpy(want, m, i, -want[m][j])
    assert all(Fraction(new_obj[jj], D * sigma_m) == want[m][jj] for jj in range(n + 1))
    assert all(want[m][j] == 0 for j in basis)
    # magnitude: below the product of the basis columns' norms times the entry's own column norm
    sq = [info["col_sq_a"][j] + C[jj] * C[jj] for jj, j in enumerate(cols)]
    w = X.weights_from_squares(sq)
    wsum = sum(w[j] for j in basis)
    Nc = [[row[j] for j in cols] for row in N[:m]] + [new_obj]
    for i in range(m + 1):
        for jj in range(n + 1):
            assert abs(Nc[i][jj]).bit_length() * 256 <= wsum + w[jj] + 256
    assert abs(D).bit_length() * 256 <= wsum + 256
    # phase 2 keeps dividing exactly (bareiss_pivot asserts it) and stays the reference's tableau
    sig2 = sigma[:m] + [sigma_m]
    trace = []
    O.solve(want, basis[:], trace)
    for r, cc, _ in trace:
        D = O.bareiss_pivot(Nc, D, r, cc)
        alive[r] = 0
    den = lambda i: D * sig2[i] if alive[i] else D
    assert all(Fraction(Nc[i][j], den(i)) == want[i][j] for i in range(m + 1) for j in range(n + 1))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_row_operations_in_fraction_free_form(seed):
    """What `lps_row_combine_exact` computes with the integers `row_combine_coefficients` prepares is the reference's
    rowMult / rowAdd / rowAddToObj (tableau.py:254-293) on N / (D s_i), at any point of a pivot sequence, and the
    pivots that follow still divide exactly (bareiss_pivot asserts it)."""
    rng = random.Random(seed)
    m, n = 6, 10
    grid = [[Fraction(rng.randint(-9, 9), rng.choice([1, 1, 2, 3])) for _ in range(n + 1)] for _ in range(m + 1)]
    ints, sigma = X.clear_denominators(grid)
    N, D, alive = [row[:] for row in ints], 1, [1] * (m + 1)
    den = lambda i: D * sigma[i] if alive[i] else D

    def check():
        assert all(Fraction(N[i][j], den(i)) == grid[i][j] for i in range(m + 1) for j in range(n + 1))

    def pivot(r, c):
        nonlocal D
        O.pivot(grid, r, c)
        D = O.bareiss_pivot(N, D, r, c)
        alive[r] = 0

    def row_op(rd, rs, mult):
        s_d = sigma[rd] if alive[rd] else 1
        s_s = None if rs is None else (sigma[rs] if alive[rs] else 1)
        a, b, scale = X.row_combine_coefficients(s_d, s_s, mult)
        N[rd] = [a * N[rd][j] + (b * N[rs][j] if rs is not None else 0) for j in range(n + 1)]
        sigma[rd], alive[rd] = scale, 1
        if rs is None:
            O.row_scale(grid, rd, mult)
        else:
            O.row_axpy(grid, rd, rs, mult)

    steps = 0
    for rnd in range(12):
        kind = rng.choice(["pivot", "pivot", "scale", "add", "add", "obj"])
        if kind == "pivot":
            r, c = rng.randrange(m), rng.randrange(n)
            if grid[r][c] == 0:
                continue
            pivot(r, c)
        elif kind == "scale":
            row_op(rng.randrange(m), None, Fraction(rng.choice([-1, 2, -3, 5]), rng.choice([1, 2, 7])))
        elif kind == "add":
            rd, rs = rng.randrange(m), rng.randrange(m)
            row_op(rd, rs, Fraction(rng.randint(-4, 4) or 1, rng.choice([1, 1, 3])))
        else:
            row_op(m, rng.randrange(m), Fraction(rng.randint(-5, 5) or 2, rng.choice([1, 2])))
        check()
        steps += 1
    assert steps >= 6


def test_prepare_phase2_agrees_between_array_and_list_inputs():
    """The integer-array path of `ExactEngine._prepare_phase2` (what `Tableau.from_grid` with integers takes) gives the
    same objective integers, scale and column norms as the entry-by-entry path, and a bound that covers both the
    phase-1 matrix and the matrix with the original objective."""
    rng = random.Random(9)
    m, n0, k = 7, 11, 4
    n = n0 + k
    rows = [[rng.randint(-9, 9) for _ in range(n0)] + [1 if j == i else 0 for j in range(k)] + [rng.randint(0, 40)] for i in range(m)]
    obj = [-sum(rows[i][j] for i in range(k)) for j in range(n + 1)]
    grid = rows + [obj]
    cost = [Fraction(rng.randint(1, 9)) for _ in range(n0)] + [Fraction(-5)]
    eng = object.__new__(X.ExactEngine)
    eng.m, eng.n = m, n
    sigma = [1] * (m + 1)
    bits_l, info_l = eng._prepare_phase2(cost, grid, sigma)
    bits_a, info_a = eng._prepare_phase2(cost, np.array(grid, dtype=np.int64), sigma)
    assert info_l["C"] == info_a["C"] and info_l["sigma_m"] == info_a["sigma_m"] == 1
    assert info_l["col_sq_a"] == info_a["col_sq_a"] and abs(bits_l - bits_a) <= 1
    with_cost = rows + [[int(v) for v in cost[:n0]] + [0] * k + [int(cost[n0])]]
    assert bits_l >= max(X.hadamard_bits(grid), X.hadamard_bits(with_cost)) - 1
    w = X.weights_from_squares([0, 1, 2, 1 << 40, (1 << 200) + 1])
    assert w[0] == 0 and w[1] >= 0 and w[3] >= 256 * 20 and w[4] >= 256 * 100 and w == sorted(w)


def _det(M):
    """Exact determinant of a small integer matrix (fraction-free elimination)."""
    M = [row[:] for row in M]
    n, sign, prev = len(M), 1, 1
    for k in range(n):
        piv = next((i for i in range(k, n) if M[i][k] != 0), None)
        if piv is None:
            return 0
        if piv != k:
            M[k], M[piv] = M[piv], M[k]
            sign = -sign
        for i in range(k + 1, n):
            for j in range(k + 1, n):
                M[i][j] = (M[i][j] * M[k][k] - M[i][k] * M[k][j]) // prev
        prev = M[k][k]
    return sign * M[n - 1][n - 1]


@pytest.mark.parametrize("seed", range(6))
def test_static_bound_covers_every_minor_of_irregular_matrices(seed):
    """`hadamard_bits` (what sizes the ALLOCATED prime set) against brute force: every square minor of small matrices
    with mixed signs, zero rows / columns and entries up to 8 bits stays below 2^bits."""
    import itertools
    rng = random.Random(70 + seed)
    nr, nc = rng.randint(2, 4), rng.randint(2, 5)
    span = rng.choice([1, 3, 9, 127])
    G = [[0 if rng.random() < 0.3 else rng.randint(-span, span) for _ in range(nc)] for _ in range(nr)]
    if seed % 2:
        G[rng.randrange(nr)] = [0] * nc
    bits = X.hadamard_bits(G)
    assert X.hadamard_bits_np(np.array(G, dtype=np.int64)) >= bits
    for k in range(1, min(nr, nc) + 1):
        for rows in itertools.combinations(range(nr), k):
            for cols in itertools.combinations(range(nc), k):
                d = _det([[G[i][j] for j in cols] for i in rows])
                assert abs(d).bit_length() <= bits, (rows, cols, d, bits)


@pytest.mark.parametrize("seed", range(8))
def test_working_set_bound_covers_every_bareiss_integer_on_irregular_lps(seed):
    """The WORKING prime set rests on: every tableau integer and D is, up to sign, a minor of the
    uploaded matrix over the last pivot columns of the rows pivoted so far plus at most one more column, hence
    log2|x| <= sum of those columns' weights (`column_weights`, 1/256 bit) + the heaviest other column.  Checked along
    whole oracle solves of the irregular family (mixed signs, zero columns, slack columns, dependent rows),
    phase 1 included, with the artificial problem's matrix as the uploaded one."""
    rng = random.Random(500 + seed)
    while True:                                  # integer data whose constructor and solve go through
        c, A, b = problems.irregular_lp(rng, rng.randint(4, 9), rng.randint(5, 12))
        if any(v.denominator != 1 for row in A + [b, c] for v in row):
            continue
        try:
            O.simplex(c, A, b)
        except (ValueError, IndexError, AssertionError):
            continue
        break
    grid = O.make_grid(c, A, b)
    m, n = O.shape(grid)
    for i in range(m):
        if grid[i][n] < 0:
            O.row_scale(grid, i, Fraction(-1))
    basis = [-1] * m
    if not O.is_canonical(grid, basis):          # the artificial problem of simplex.py:47-60, as one uploaded matrix
        missing = [i for i in range(m) if basis[i] == -1]
        k = len(missing)
        for i in range(m + 1):
            row = grid[i]
            row[n:] = [Fraction(0)] * k + [row[n]]
        for j in range(n + k + 1):
            grid[m][j] = Fraction(0)
        for idx, i in enumerate(missing):
            grid[i][n + idx] = Fraction(1)
            basis[i] = n + idx
            for j in range(n + k + 1):           # cost 1 on the artificial, priced out at once
                if j != n + idx:
                    grid[m][j] -= grid[i][j]
        n += k
    N = [[int(x) for x in row] for row in grid]
    w = X.column_weights(N)
    assert X.column_weights(np.array(N, dtype=np.int64)) == w or all(
        abs(a - b_) <= 1 for a, b_ in zip(X.column_weights(np.array(N, dtype=np.int64)), w))
    D, last, trace = 1, {}, []
    O.solve(grid, basis, trace)
    assert trace, "the walk made no pivot"
    for r, cc, _ in trace:
        D = O.bareiss_pivot(N, D, r, cc)
        last[r] = cc
        wsum = sum(w[j] for j in last.values())
        assert abs(D).bit_length() - 1 <= wsum / 256 + 1e-9
        heaviest = max(w[j] for j in range(n + 1))
        top = max(abs(x).bit_length() for row in N for x in row)
        assert top - 1 <= (wsum + heaviest) / 256 + 1e-9, (top, wsum, heaviest)
