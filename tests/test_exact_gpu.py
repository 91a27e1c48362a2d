"""GPU parity of the exact (multi-modular Bareiss) mode against the reference fixtures and the oracle.

Exact mode must reproduce the reference's pivot sequence, basis and rational results bit-for-bit, so
every comparison here is `==` on ints / Fractions / pivot lists."""
import math
import random
from fractions import Fraction

import numpy as np
import pytest

from conftest import load_golden
from linear_program_solver_b200 import problems

pytestmark = pytest.mark.gpu


def _tab(c, A, b, mode="exact"):
    from linear_program_solver_b200 import Tableau
    t = Tableau(len(A), len(c), mode=mode)
    t.setC(list(c))
    t.setB(list(b))
    t.setA([list(r) for r in A])
    return t


def _frac_grid(c, A, b):
    from oracle import lpsol_oracle as O
    return O.make_grid(c, A, b)


@pytest.mark.parametrize("bits", [8, 31, 32, 63, 64, 200, 1000])
def test_encode_decode_roundtrip(bits):
    """Residue encoding + on-device Garner reconstruction give back the uploaded integers."""
    from linear_program_solver_b200.exact import ExactEngine
    rng = random.Random(bits)
    m, n = 5, 9
    grid = [[rng.choice([-1, 1]) * rng.getrandbits(bits) if rng.random() < 0.9 else 0 for _ in range(n + 1)] for _ in range(m + 1)]
    grid[0][0], grid[1][1], grid[2][2] = (1 << bits) - 1, -((1 << bits) - 1), 0
    e = ExactEngine(m, n)
    e.upload(grid)
    assert e.denominator() == 1
    N, D = e.integer_state()
    assert N == grid and D == 1
    assert e.download() == [[Fraction(x) for x in row] for row in grid]
    assert e.read_col(3) == [Fraction(grid[i][3]) for i in range(m + 1)]
    e.close()


def test_rational_rows_keep_their_scale_until_pivoted():
    from linear_program_solver_b200.exact import ExactEngine
    from oracle import lpsol_oracle as O
    c, A, b = problems.BEALE
    grid = O.make_grid(c, A, b)
    e = ExactEngine(3, 7)
    e.upload(grid)
    assert e.download() == grid
    O.pivot(grid, 0, 3)
    assert e.pivot(0, 3) == 0
    assert e.download() == grid
    O.pivot(grid, 1, 4)
    assert e.pivot(1, 4) == 0
    assert e.download() == grid
    e.close()


def test_reference_known_answer_test_pivot():
    """test_tableau.py:220-227 on the new Tableau in exact mode (whole-tableau equality, marks included)."""
    g = load_golden("textbook")
    t = _tab(g["c"], g["A"], g["b"])
    for step in g["steps"]:
        t.pivot(*step["pivot"])
        assert [str(x) for x in t.getC()] == step["c"]
        assert str(t._z) == step["z_cell"]
        assert [str(x) for x in t.getB()] == step["b"]
        assert [[str(x) for x in r] for r in t.getA()] == step["a"]
    assert t.getZ() == -400
    t2 = _tab(g["c"], g["A"], g["b"])
    t2.pivot(0, 1)
    t3 = _tab(g["c"], g["A"], g["b"])
    t3.pivot(1, 0)
    t3.pivot(0, 1)
    t2.pivot(1, 0)
    assert t3 == t and t2 == t
    with pytest.raises(ZeroDivisionError, match="zero pivot 0,3"):
        _tab(g["c"], g["A"], g["b"]).pivot(0, 3)


@pytest.mark.parametrize("m,seed", [(6, 0), (20, 1), (50, 2)])
def test_pivots_match_oracle_fractions_and_bareiss_integers(m, seed):
    from linear_program_solver_b200.exact import ExactEngine
    from oracle import lpsol_oracle as O
    c, A, b = problems.dense_lp(m, seed)
    grid = O.make_grid(c, A, b)
    Nint = [[int(x) for x in row] for row in grid]
    D = 1
    e = ExactEngine(m, 2 * m)
    e.upload(grid)
    rng = random.Random(seed)
    for _ in range(5):
        r, cc = rng.randrange(m), rng.randrange(2 * m)
        while grid[r][cc] == 0:
            cc = rng.randrange(2 * m)
        O.pivot(grid, r, cc)
        D = O.bareiss_pivot(Nint, D, r, cc)
        assert e.pivot(r, cc) == 0
        gotN, gotD = e.integer_state()
        assert gotD == D and gotN == Nint
        assert e.download() == grid
    e.close()


@pytest.mark.parametrize("rec", load_golden("dense_exact"), ids=lambda r: f"m{r['m']}s{r['seed']}h{r['hi']}")
def test_simplex_matches_reference_exactly(rec):
    from linear_program_solver_b200 import Simplex
    c, A, b = problems.dense_lp(rec["m"], rec["seed"], rec["lo"], rec["hi"])
    t = _tab(c, A, b)
    s = Simplex(t)
    assert s.getBasicSequence() == rec["initial_basis"]
    mp = rec.get("max_pivots")
    s.solve(max_pivots=-1 if mp is None else mp)
    assert [list(p) for p in s.pivot_log] == rec["pivots"]
    assert s.getBasicSequence() == rec["basis"]
    assert str(s.getObjValue()) == rec["objective"]
    assert {str(j): str(v) for j, v in s.getBFS().items()} == rec["bfs"]
    assert [bool(x) for x in t.getVarMarks()] == rec["marks"]


@pytest.mark.parametrize("rec", load_golden("phase1"), ids=lambda r: f"m{r['m']}n{r['n']}s{r['seed']}")
def test_phase1_matches_reference_exactly(rec):
    from linear_program_solver_b200 import Simplex
    c, A, b = problems.equality_lp(rec["m"], rec["n"], rec["seed"])
    t = _tab(c, A, b)
    if "ctor_error" in rec:
        with pytest.raises({"ValueError": ValueError, "IndexError": IndexError}[rec["ctor_error"][0]]) as ei:
            Simplex(t)
        if rec["ctor_error"][0] == "ValueError":
            assert str(ei.value) == rec["ctor_error"][1]
        return
    s = Simplex(t)
    assert t.phase2_on_device == 1                       # simplex.py:95-105 ran as `lps_phase2_exact`, not through the host
    n1 = len(rec["phase1_pivots"])
    assert [list(p) for p in s.pivot_log] == rec["phase1_pivots"]
    assert s.getBasicSequence() == rec["initial_basis"]
    s.solve()
    assert [list(p) for p in s.pivot_log[n1:]] == rec["pivots"]
    assert s.getBasicSequence() == rec["basis"]
    assert str(s.getObjValue()) == rec["objective"]
    assert {str(j): str(v) for j, v in s.getBFS().items()} == rec["bfs"]
    assert [bool(x) for x in t.getVarMarks()] == rec["marks"]
    assert list(t.getTableauSize()) == rec["m_n"]


@pytest.mark.parametrize("m,n,seed,rational", [(40, 72, 0, False), (60, 100, 1, False), (9, 16, 2, True), (14, 24, 3, True)])
def test_phase1_leaves_on_the_device(m, n, seed, rational):
    """Larger equality-form LPs (and rational data, whose rows carry scales): the whole constructor -- artificial
    problem, drive-out pivots, cut of the artificial columns, objective restore (`lps_phase2_exact`) -- and the
    phase-2 solve against the oracle; the tableau after the constructor is the oracle's, entry by entry; the prime
    set stays the working set of the ORIGINAL matrix (no re-encoding of the pivoted tableau)."""
    from linear_program_solver_b200 import Simplex
    from oracle import lpsol_oracle as O
    c, A, b = problems.equality_lp(m, n, seed)
    if rational:
        rng = random.Random(seed)
        A = [[Fraction(v, rng.choice([1, 2, 3])) for v in row] for row in A]
        c = [Fraction(v, rng.choice([1, 5])) for v in c]
        b = [sum(A[i][j] * ((j * 7 + seed) % 5) for j in range(n)) for i in range(m)]
    want = O.simplex(c, A, b)
    t = _tab(c, A, b)
    s = Simplex(t)
    assert t.phase2_on_device == 1 and t.getTableauSize() == (m, n)
    info = t._eng.info()
    assert info["primes"] <= info["allocated_primes"] and t._eng.reencodes == 0
    assert [tuple(p) for p in s.pivot_log] == [tuple(p) for p in want["phase1_pivots"]]
    assert s.getBasicSequence() == want["initial_basis"]
    if m <= 20:                                          # the tableau phase 2 starts from, entry by entry
        g0 = O.make_grid(c, A, b)
        g0, _ = O.find_bfs(g0)
        assert t._eng.download() == g0
    n1 = len(s.pivot_log)
    s.solve()
    assert [tuple(p) for p in s.pivot_log[n1:]] == [tuple(p) for p in want["pivots"]]
    assert s.getBasicSequence() == want["basis"]
    assert s.getObjValue() == want["objective"]
    assert s.getBFS() == want["bfs"]
    assert t._eng.reencodes == 0


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_phase1_with_slack_rows_and_rational_data(seed):
    """Half the rows come with a slack column (they keep their basis through phase 1 unless pivoted, and their row
    scale with it), the other half are equalities that need artificials; entries are rationals.  The rows that are
    never pivoted still carry a scale when the device leaves phase 1 (`lps_phase2_exact`, G_i = C_{B_i} / s_i)."""
    from linear_program_solver_b200 import Simplex
    from oracle import lpsol_oracle as O
    rng = random.Random(100 + seed)
    ms, me, nx = 5, 5, 9
    m, n = ms + me, nx + ms
    x0 = [rng.randint(0, 3) for _ in range(nx)]
    A, b = [], []
    for i in range(m):
        den = rng.choice([1, 2, 3, 4])
        row = [Fraction(rng.randint(-4, 9), den) for _ in range(nx)] + [Fraction(0)] * ms
        val = sum(row[j] * x0[j] for j in range(nx))
        if i < ms:
            row[nx + i] = Fraction(1)
            val += rng.randint(0, 5)                       # slack row: feasible with room to spare
        A.append(row)
        b.append(val)
    c = [Fraction(rng.randint(1, 9), rng.choice([1, 2])) for _ in range(nx)] + [Fraction(0)] * ms
    try:
        want = O.simplex(c, A, b)
    except (ValueError, IndexError) as e:
        with pytest.raises(type(e).__mro__[1] if type(e).__name__ == "InfeasibleError" else type(e)):
            Simplex(_tab(c, A, b))
        return
    t = _tab(c, A, b)
    s = Simplex(t)
    assert [tuple(p) for p in s.pivot_log] == [tuple(p) for p in want["phase1_pivots"]]
    assert s.getBasicSequence() == want["initial_basis"]
    if want["phase1_pivots"]:
        assert t.phase2_on_device == 1
    g0, _ = O.find_bfs(O.make_grid(c, A, b))
    assert t._eng.download() == g0
    n1 = len(s.pivot_log)
    s.solve()
    assert [tuple(p) for p in s.pivot_log[n1:]] == [tuple(p) for p in want["pivots"]]
    assert s.getBasicSequence() == want["basis"] and s.getObjValue() == want["objective"] and s.getBFS() == want["bfs"]


def test_special_cases():
    from linear_program_solver_b200 import Simplex
    sp = load_golden("special")
    c, A, b = problems.BEALE
    s = Simplex(_tab(c, A, b))
    s.solve()
    assert [list(p) for p in s.pivot_log] == sp["beale"]["pivots"]       # 10 Dantzig then 8 Bland pivots
    assert str(s.getObjValue()) == sp["beale"]["objective"] == "-5/4"
    assert {str(j): str(v) for j, v in s.getBFS().items()} == sp["beale"]["bfs"]
    t = _tab([0, 0], [[1, 1], [1, 1]], [1, 2])
    with pytest.raises(ValueError, match="infeasible problem, artificial opt = 1"):
        Simplex(t)
    assert t.getNumVars() == 4
    s = Simplex(_tab([-1, 0], [[-1, 1]], [1]))
    assert s.findPivotStandard() == "unbounded" and s.findPivotMinIndex() == "unbounded"
    with pytest.raises(AssertionError, match="unbounded artificial problem"):
        s.solve()
    with pytest.raises(IndexError):
        Simplex(_tab([1, 1, 1], [[1, 1, 0], [2, 2, 0], [0, 1, 1]], [2, 4, 3]))


def test_find_pivot_is_canonical_is_optimal_match_oracle():
    from linear_program_solver_b200.exact import ExactEngine
    from oracle import lpsol_oracle as O
    c, A, b = problems.dense_lp(16, 4)
    grid = O.make_grid(c, A, b)
    e = ExactEngine(16, 32)
    e.upload(grid)
    want = [-1] * 16
    assert O.is_canonical(grid, want)
    ok, cols = e.is_canonical()
    assert ok and cols.tolist() == want
    for _ in range(6):
        for rule in (0, 1):
            assert e.find_pivot(rule) == O.find_pivot(grid, rule)
        found = O.find_pivot(grid, 0)
        if isinstance(found, str):
            break
        r, cc = found
        O.pivot(grid, r, cc)
        assert e.pivot(r, cc) == 0
        want = [-1] * 16
        O.is_canonical(grid, want)
        ok, cols = e.is_canonical()
        assert cols.tolist() == want
        assert e.is_optimal() == O.is_optimal(grid)
    e.close()


def test_prime_dividing_a_pivot_forces_a_reencode_and_keeps_the_answer():
    """Small primes make 'p_k divides the pivot element' certain; results must not change."""
    from linear_program_solver_b200.exact import ExactEngine, primes_for_bits
    from oracle import lpsol_oracle as O
    c, A, b = problems.dense_lp(8, 0)
    want = O.simplex(c, A, b)
    small = [3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 101, 103,
             107, 109, 113, 127, 131, 137, 139, 149, 151, 157, 163, 167, 173, 179, 181, 191, 193, 197, 199, 211, 223,
             227, 229, 233, 239, 241, 251, 257, 263, 269, 271, 277, 281, 283, 293]
    e = ExactEngine(8, 16, primes=small + primes_for_bits(64))
    e.upload(O.make_grid(c, A, b))
    basis = np.arange(8, 16, dtype=np.int32)
    st, k, trace = e.solve(basis, None, trace_cap=256)
    assert st == 1 and e.reencodes > 0
    assert [tuple(p) for p in trace.tolist()] == [tuple(p) for p in want["pivots"]]
    assert basis.tolist() == want["basis"]
    assert -e.read_row(8)[16] == want["objective"]
    e.close()


# ---- working prime set (lps_exact_set_weights): same integers on a prefix of the primes ------------------------
@pytest.mark.parametrize("m,seed,hi", [(24, 0, 9), (60, 1, 127), (120, 2, 127)])
def test_working_prime_set_gives_the_same_integers_as_the_full_set(m, seed, hi):
    from linear_program_solver_b200.exact import ExactEngine
    c, A, b = problems.dense_lp(m, seed, 1, hi)
    grid = np.array(problems.to_grid(c, A, b), dtype=np.int64)
    runs = []
    for adaptive in (True, False):
        e = ExactEngine(m, 2 * m)
        e.adaptive = adaptive
        e.upload(grid)
        info0 = e.info()
        basis = np.arange(m, 2 * m, dtype=np.int32)
        st, k, trace = e.solve(basis, None, trace_cap=1 << 14)
        runs.append((st, k, trace.tolist(), basis.tolist(), e.integer_state(), info0, e.info()))
        e.close()
    (st_a, k_a, tr_a, bs_a, ints_a, i0_a, i1_a), (st_f, k_f, tr_f, bs_f, ints_f, i0_f, i1_f) = runs
    assert (st_a, k_a, tr_a, bs_a) == (st_f, k_f, tr_f, bs_f)
    assert ints_a == ints_f                                     # every N_ij and D, as Python ints
    assert i0_f["primes"] == i0_f["allocated_primes"] == i1_f["primes"] and i1_f["extensions"] == 0
    assert i0_a["primes"] < i0_a["allocated_primes"]            # starts on a short prefix ...
    assert i0_a["primes"] <= i1_a["primes"] <= i1_a["allocated_primes"]
    if m >= 60:                                                 # ... and grows once the basis holds enough dense columns
        assert i1_a["extensions"] >= 1 and i0_a["primes"] < i1_a["primes"]


def test_working_prime_set_covers_every_integer_of_a_solve():
    """The bound behind the working set: |entry| <= product of the norms of the basis columns and two more."""
    import math
    from linear_program_solver_b200.exact import column_weights
    from oracle import lpsol_oracle as O
    m = 12
    c, A, b = problems.dense_lp(m, 5, 1, 99)
    grid = [list(map(int, row)) for row in problems.to_grid(c, A, b)]
    w = column_weights(grid)
    assert w == column_weights(np.array(grid, dtype=np.int64)) or all(abs(x - y) <= 1 for x, y in zip(w, column_weights(np.array(grid, dtype=np.int64))))
    N, D = [row[:] for row in grid], 1
    xbasis = [-1] * m
    want = O.simplex(c, A, b, num=Fraction)
    for r, cc, _rule in want["pivots"]:
        D = O.bareiss_pivot(N, D, r, cc)
        xbasis[r] = cc
        cap = sum(w[j] for j in xbasis if j >= 0) + max(w)
        big = max(max(abs(v) for row in N for v in row), abs(D))
        assert 256 * math.log2(big) <= cap


# ---- full-size property checks: an exact optimality certificate -------------------------------------------------
def _certify_optimum(c, A, b, t, s):
    """Prove on the CPU, with Python integers and independently of the device arithmetic, that the solved exact
    tableau `t` / `s` is THE optimum of min c^T x, A x = b, x >= 0: the reported basic solution satisfies A x = b,
    x >= 0 exactly; the multipliers y read off the slack columns of the device's objective row reproduce EVERY reduced
    cost d_j = c_j - y^T A_j exactly, all d_j >= 0, d_j = 0 on the basis; and c^T x = y^T b = the reported objective
    (strong duality).  A is [structural | identity]."""
    m, n = len(A), len(c)
    basis = s.getBasicSequence()
    assert len(set(basis)) == m and t.isOptimal()
    bfs, z = s.getBFS(), s.getObjValue()
    d = t.getC()                                           # reduced costs, n Fractions (one Garner pass on the device)
    den = 1
    for v in bfs.values():
        den = den * v.denominator // math.gcd(den, v.denominator)
    X = np.zeros(n, dtype=object)
    for j, v in bfs.items():
        assert v >= 0
        X[j] = int(v * den)
    Aobj = np.array(A, dtype=object)
    cobj, bobj = np.array(c, dtype=object), np.array(b, dtype=object)
    assert ((Aobj @ X) == bobj * den).all()
    assert Fraction(int(np.dot(cobj, X)), den) == z
    dd = 1
    for v in d:
        dd = dd * v.denominator // math.gcd(dd, v.denominator)
    Dint = np.array([int(v * dd) for v in d], dtype=object)
    assert all(v >= 0 for v in Dint) and all(Dint[j] == 0 for j in basis)
    Y = -Dint[m:]                                          # slack columns are unit vectors with zero cost
    assert ((cobj * dd - (Y @ Aobj)) == Dint).all()
    assert Fraction(int(np.dot(Y, bobj)), dd) == z
    return z, bfs


def test_c3_full_solve_carries_an_exact_optimality_certificate():
    """C3 (exact 1000 x 2000, 8-bit coefficients) solved to optimality through the public API.  The reference needs
    hours for this size, so instead of a recorded answer the result is proved optimal on the CPU (`_certify_optimum`).
    The pivot sequence's first six pivots are pinned by the reference fixture (dense_exact.json)."""
    from linear_program_solver_b200 import Simplex, Tableau
    m = 1000
    c, A, b = problems.dense_lp(m, 0, 1, 127)
    grid = np.array([A[i] + [b[i]] for i in range(m)] + [c + [0]], dtype=np.int64)
    t = Tableau.from_grid(grid, mode="exact")
    s = Simplex(t)
    assert s.solve() is None                               # simplex.py:110 returns nothing; the count is an attribute
    npiv = s.last_pivots
    rec = next(r for r in load_golden("dense_exact") if r["m"] == m and r["hi"] == 127)
    assert [list(p) for p in s.pivot_log[:len(rec["pivots"])]] == rec["pivots"]
    assert npiv == len(s.pivot_log) > m // 2
    z, _ = _certify_optimum(c, A, b, t, s)
    assert z.denominator.bit_length() > 300                # a genuinely big rational, not a float-sized answer
    info = t._eng.info()
    assert info["primes"] < info["allocated_primes"]       # solved on the working prime set


def test_c2_fp64_is_within_1e9_of_the_certified_exact_optimum():
    """BASELINE.json config 2 at full size (2000 x 4000): the north star asks fp64 mode to be within 1e-9 relative of
    the exact objective and primal values.  The reference's Fraction simplex needs hours here; the exact mode solves
    the same LP in seconds, its answer is proved optimal on the CPU (`_certify_optimum`), and the fp64 result is
    compared with it."""
    from linear_program_solver_b200 import Simplex, Tableau
    m = 2000
    G = problems.hashed_dense_grid(m, 0)
    A = [[int(v) for v in row[:2 * m]] for row in G[:m]]
    b = [int(row[2 * m]) for row in G[:m]]
    c = [int(v) for v in G[m, :2 * m]]
    t = Tableau.from_grid(G.astype(np.int64), mode="exact")
    s = Simplex(t)
    s.solve()
    z, bfs = _certify_optimum(c, A, b, t, s)
    tf = Tableau.from_grid(G.astype(np.float64), mode="fp64")
    sf = Simplex(tf)
    sf.solve()
    zf = Fraction(sf.getObjValue())
    assert abs(zf - z) <= Fraction(1, 10 ** 9) * abs(z)
    # primal values: compared when the float path ends on the same vertex (with ties it may stop on another vertex of
    # the optimal face; this instance does not)
    assert set(sf.getBasicSequence()) == set(s.getBasicSequence())
    xf = sf.getBFS()
    worst = max(abs(Fraction(xf.get(j, 0.0)) - v) / max(abs(v), 1) for j, v in bfs.items())
    assert worst <= Fraction(1, 10 ** 9)


# ---- whole-tableau rules in exact mode (lps_exact_col_ratios) ---------------------------------------------------
def _walk_rules(s, states):
    for step, want in enumerate(states):
        got = s.findPivotMaxIncrease()
        assert (list(got) if isinstance(got, tuple) else got) == want["max_increase"], step
        assert [list(p) for p in s.findPivotAll()] == want["all"], step
        if isinstance(got, str):
            return
        nxt = s.findPivotMaxIncrease(True) if step % 2 else s.findPivotStandard(True)
        assert not isinstance(nxt, str)


@pytest.mark.parametrize("rec", load_golden("rules"), ids=lambda r: f"{r['family']}m{r['m']}s{r['seed']}")
def test_whole_tableau_rules_match_reference(rec):
    """findPivotMaxIncrease / findPivotAll (simplex.py:286-360) against the unmodified reference's answers at every
    state of the recorded walk; the ratio tests of all columns run on the device (many equal ratios in this data)."""
    from linear_program_solver_b200 import Simplex
    if rec["family"] == "dense":
        c, A, b = problems.dense_lp(rec["m"], rec["seed"], rec["lo"], rec["hi"])
    else:
        c, A, b = problems.equality_lp(rec["m"], rec["n"], rec["seed"])
    _walk_rules(Simplex(_tab(c, A, b)), rec["exact"])


def test_whole_tableau_rules_larger_against_the_oracle():
    """150 x 300 with 1..3 coefficients: heavy ties, several batches of columns; oracle on a downloaded snapshot."""
    from linear_program_solver_b200 import Simplex
    from oracle import lpsol_oracle as O
    c, A, b = problems.dense_lp(150, 2, 1, 3)
    s = Simplex(_tab(c, A, b))
    for step in range(6):
        snap = s._tab._eng.download()
        want = O.find_pivot_max_increase(snap)
        assert s.findPivotMaxIncrease() == want
        assert s.findPivotAll() == O.find_pivot_all(snap)
        if isinstance(want, str):
            break
        s.findPivotMaxIncrease(True) if step % 2 else s.findPivotStandard(True)
    unb = Simplex(_tab([-1, -2, 0], [[1, -1, 1]], [1]))
    assert unb.findPivotMaxIncrease() == "unbounded"
    opt = Simplex(_tab([1, 2, 0], [[1, 1, 1]], [1]))
    assert opt.findPivotMaxIncrease() == "optimal" and opt.findPivotAll() == [(0, 0), (0, 1), (0, 2)]


# ---- row operations outside a pivot on the device residues (lps_row_combine_exact) -----------------------------
def _view(t):
    """The valid copy of the tableau: the device's when it is the only valid one, else the host's."""
    return t._eng.download() if (t._dev_ok and not t._host_ok) else [list(row) for row in t._grid]


def test_row_operations_on_the_device_match_the_reference_arithmetic():
    """rowMult / rowDiv / rowAdd / rowSub / rowAddToObj / rowSubFromObj (tableau.py:254-293) on a tableau that lives on
    the device (after a pivot), integer and rational multipliers, against the oracle's restatement entry by entry;
    pivots and a whole solve afterwards still agree (the fraction-free divisions stay exact)."""
    from linear_program_solver_b200 import Simplex
    from oracle import lpsol_oracle as O
    c, A, b = problems.dense_lp(10, 4)
    t = _tab(c, A, b)
    grid = O.make_grid(c, A, b)
    t.pivot(2, 3)
    O.pivot(grid, 2, 3)
    ops = [("rowMult", (1, -1)), ("rowAdd", (0, 2, 3)), ("rowSub", (4, 2, Fraction(5, 2))), ("rowDiv", (5, 4)),
           ("rowAddToObj", (6, Fraction(-7, 3))), ("rowSubFromObj", (2, 2)), ("rowMult", (2, Fraction(3, 7))),
           ("rowAdd", (3, 3, 2)), ("rowAdd", (7, 5, Fraction(1, 4))), ("rowMult", (8, 1)), ("rowAdd", (8, 9, 0))]
    for name, args in ops:
        getattr(t, name)(*args)
        if name == "rowMult":
            O.row_scale(grid, args[0], Fraction(args[1]))
        elif name == "rowDiv":
            O.row_scale(grid, args[0], 1 / Fraction(args[1]))
        elif name in ("rowAdd", "rowSub"):
            O.row_axpy(grid, args[0], args[1], Fraction(args[2]) * (1 if name == "rowAdd" else -1))
        else:
            O.row_axpy(grid, 10, args[0], Fraction(args[1]) * (1 if name == "rowAddToObj" else -1))
        assert _view(t) == grid, (name, args)
    # the two no-ops (x1, +0) never reach the device; an operation whose allowance no longer fits the handle's primes
    # (this 10 x 20 tableau has two) moves the tableau to the host, where the rest is applied
    assert 5 <= t.row_ops_on_device <= 9
    t.pivot(1, 0)
    O.pivot(grid, 1, 0)
    t.pivot(2, 5)                                                  # a row that carries a scale again
    O.pivot(grid, 2, 5)
    assert _view(t) == grid
    t.rowAdd(4, 2, Fraction(-3, 2))                               # on the device again (new upload, new allowance)
    O.row_axpy(grid, 4, 2, Fraction(-3, 2))
    assert _view(t) == grid and not t._host_ok
    with pytest.raises(ZeroDivisionError):
        t.rowDiv(0, 0)
    # the objective row was edited: bring it back to canonical form, then solve from this basis
    basis = [-1] * 10
    O.is_canonical(grid, basis)
    got = [-1] * 10
    t.isCanonical(got)
    assert got == basis


def test_row_operations_then_solve_and_reencode_when_the_primes_run_out():
    """Row additions grow the magnitude allowance of the handle.  Accepted ones may leave too little room for the
    pivots that follow: the loop then reports LPS_NEED_PRIMES with every prime in use, the host re-encodes the
    tableau as a new matrix and the solve continues -- same pivots, basis and objective as the oracle either way.
    An operation the handle's primes cannot hold at all is refused by the device and applied to a host copy."""
    from linear_program_solver_b200 import Simplex
    from oracle import lpsol_oracle as O
    c, A, b = problems.dense_lp(8, 0)
    reenc, on_dev = [], []
    for k in range(1, 8):
        t = _tab(c, A, b)
        grid = O.make_grid(c, A, b)
        s = Simplex(t)
        basis = list(range(8, 16))
        r, cc = s.findPivotStandard(True)                         # the tableau now lives on the device
        O.pivot(grid, r, cc)
        basis[r] = cc
        for _ in range(k):                                        # row 5 doubled and halved: same values, larger allowance
            t.rowAdd(5, 5, 1)
            t.rowMult(5, Fraction(1, 2))
            O.row_axpy(grid, 5, 5, Fraction(1))
            O.row_scale(grid, 5, Fraction(1, 2))
        on_dev.append(t.row_ops_on_device)
        s.solve()
        trace = []
        O.solve(grid, basis, trace)
        assert [tuple(p) for p in s.pivot_log[1:]] == [tuple(p) for p in trace], k
        assert s.getBasicSequence() == basis and s.getObjValue() == O.objective(grid)
        assert _view(t) == grid
        reenc.append(t._eng.reencodes)
    assert on_dev[0] == 2 and reenc[0] == 0, (on_dev, reenc)
    # fill the handle's allowance to the last bit (each doubling of row 5 costs one bit; driven through the engine so
    # that the refused one is simply not applied), then solve: the next pivot that adds a column to the basis cannot
    # fit, every prime is in use, the tableau is re-encoded and the loop resumes
    t = _tab(c, A, b)
    grid = O.make_grid(c, A, b)
    s = Simplex(t)
    basis = list(range(8, 16))
    r, cc = s.findPivotStandard(True)
    O.pivot(grid, r, cc)
    basis[r] = cc
    applied = 0
    for k in range(400):
        mult = Fraction(2)
        if not t._eng.row_combine(5, None, mult):
            break
        O.row_scale(grid, 5, mult)
        applied += 1
    t._peek.clear()
    assert 4 <= applied < 400 and _view(t) == grid
    s.solve()
    trace = []
    O.solve(grid, basis, trace)
    assert len(trace) >= 2 and [tuple(p) for p in s.pivot_log[1:]] == [tuple(p) for p in trace]
    assert s.getBasicSequence() == basis and s.getObjValue() == O.objective(grid) and _view(t) == grid
    assert t._eng.reencodes >= 1
    big = Fraction((1 << 400) + 1, (1 << 399) + 7)
    col = next(j for j in range(16) if grid[0][j] != 0)
    t.pivot(0, col)                                               # device copy is the valid one again
    O.pivot(grid, 0, col)
    before = t.row_ops_on_device
    t.rowAdd(0, 1, big)                                           # cannot fit the handle's primes: host copy, re-encode
    O.row_axpy(grid, 0, 1, big)
    assert t.row_ops_on_device == before and t._host_ok
    assert t.getA() == [row[:-1] for row in grid[:-1]]
    col = next(j for j in range(16) if grid[1][j] != 0)
    t.pivot(1, col)
    O.pivot(grid, 1, col)
    assert _view(t) == grid


# ---- form predicates on the device (tableau.py:504-521) ---------------------------------------------------------
@pytest.mark.parametrize("mode", ["exact", "fp64"])
def test_form_predicates_match_the_reference_definitions(mode):
    """isUnbounded / isInfeasible / isDegenerate on random sparse-signed tableaux, before and after pivots, against the
    reference's definitions evaluated on a snapshot (tableau.py:504-521); the api golden cases pin the same calls to
    the unmodified reference's answers."""
    rng = random.Random(11)
    seen = set()
    for trial in range(24):
        m, n = rng.randint(2, 7), rng.randint(3, 9)
        A = [[rng.choice([-3, -1, 0, 0, 0, 1, 2]) if rng.random() < 0.7 else -rng.randint(0, 2) for _ in range(n)] for _ in range(m)]
        if trial % 3 == 0:
            A[rng.randrange(m)] = [-abs(v) for v in A[0]]           # a row without positive entries
        if trial % 4 == 0:
            jj = rng.randrange(n)
            for row in A:
                row[jj] = -abs(row[jj])                            # a column without positive entries
        b = [rng.randint(-1, 3) for _ in range(m)]
        c = [rng.randint(-3, 2) for _ in range(n)]
        t = _tab(c, A, b, mode=mode)
        for step in range(3):
            t._need_dev()
            snap = t._eng.download()
            g = [[float(x) for x in row] for row in snap] if mode == "fp64" else snap
            want_unb = any(g[m][j] < 0 and all(g[i][j] <= 0 for i in range(m)) for j in range(n))
            want_inf = any(g[i][n] > 0 and all(g[i][j] <= 0 for j in range(n)) for i in range(m))
            want_deg = any(g[i][n] == 0 for i in range(m))
            assert (t.isUnbounded(), t.isInfeasible(), t.isDegenerate()) == (want_unb, want_inf, want_deg), (trial, step)
            seen.add((want_unb, want_inf))
            piv = [(i, j) for i in range(m) for j in range(n) if g[i][j] != 0]
            if not piv:
                break
            t.pivot(*piv[rng.randrange(len(piv))])
    assert len(seen) == 4                                          # every combination came up
