"""The C restatement (oracle/pivot_oracle.c) is bit-identical to the pinned Python oracle."""
import numpy as np
import pytest

from conftest import load_golden
from linear_program_solver_b200 import problems
from oracle import c_oracle as C
from oracle import lpsol_oracle as O


@pytest.mark.parametrize("m,seed", [(8, 0), (8, 3), (50, 0), (50, 2), (100, 1)])
def test_c_solve_bit_identical_to_python_float_oracle(m, seed):
    c, A, b = problems.dense_lp(m, seed)
    res = O.simplex(c, A, b, num=float)
    T = np.array(problems.to_grid(c, A, b), dtype=np.float64)
    basis = np.arange(m, 2 * m, dtype=np.int32)
    st, npiv, trace = C.solve(T, basis)
    assert st == C.OPTIMAL
    assert trace.tolist() == [list(p) for p in res["pivots"]]
    assert basis.tolist() == res["basis"]
    want = np.array(res["grid"], dtype=np.float64)
    assert T.tobytes() == want.tobytes()


def test_c_solve_matches_golden_float_reference_200():
    rec = [r for r in load_golden("dense_float") if r["m"] == 200][0]
    c, A, b = problems.dense_lp(200, rec["seed"])
    T = np.array(problems.to_grid(c, A, b), dtype=np.float64)
    basis = np.arange(200, 400, dtype=np.int32)
    st, npiv, trace = C.solve(T, basis)
    assert st == C.OPTIMAL and trace.tolist() == rec["pivots"]
    assert basis.tolist() == rec["basis"]
    assert float(-T[200, 400]).hex() == rec["objective"]
    assert {str(int(basis[i])): float(T[i, 400]).hex() for i in range(200)} == rec["bfs"]


def test_c_threads_do_not_change_results():
    c, A, b = problems.dense_lp(64, 5)
    outs = []
    for k in (1, 3, 8):
        C.set_threads(k)
        T = np.array(problems.to_grid(c, A, b), dtype=np.float64)
        basis = np.arange(64, 128, dtype=np.int32)
        C.solve(T, basis)
        outs.append(T.tobytes())
    C.set_threads(0)
    assert outs[0] == outs[1] == outs[2]


def test_c_beale_and_statuses():
    c, A, b = problems.BEALE
    T = np.array(problems.to_grid([float(x) for x in c], [[float(x) for x in r] for r in A], [float(x) for x in b]))
    basis = np.array([0, 1, 2], dtype=np.int32)
    st, npiv, trace = C.solve(T, basis)
    assert st == C.OPTIMAL and trace.tolist() == load_golden("special")["beale_float"]["pivots"]
    T = np.array(problems.to_grid([-1.0, 0.0], [[-1.0, 1.0]], [1.0]))
    st, npiv, _ = C.solve(T, np.array([1], dtype=np.int32))
    assert st == C.UNBOUNDED and npiv == 0
    T = np.array(problems.to_grid(*problems.TEXTBOOK), dtype=np.float64)
    assert C.pivot(T, 0, 3) == C.ZERO_PIVOT


def test_c_modp_pivot_matches_bareiss_integers():
    q = 2147483629
    c, A, b = problems.dense_lp(10, 4)
    grid = problems.to_grid(c, A, b)
    N = [[int(x) for x in row] for row in grid]
    R = np.array([[x % q for x in row] for row in N], dtype=np.uint32)
    D = 1
    res = O.simplex(c, A, b)
    for (r, cc, _) in res["pivots"]:
        C.pivot_modp(R, r, cc, q, pow(D % q, -1, q))
        D = O.bareiss_pivot(N, D, r, cc)
        assert R.tolist() == [[x % q for x in row] for row in N]


def _irregular_canonical(m, seed):
    """Canonical float tableau with mixed signs, zeros, degenerate (b = 0) rows and non-integer data: slack basis, a
    bounding row so that most draws are bounded, ties in the ratio test from repeated small integers."""
    import random
    rng = random.Random(seed)
    T = np.zeros((m + 1, 2 * m + 1))
    for i in range(m - 1):
        for j in range(m):
            if rng.random() < 0.6:
                T[i, j] = rng.choice([-2, -1, 1, 1, 2, 3, 0.5, 1.25, 1 / 3])
        T[i, 2 * m] = rng.choice([0, 0, 1, 2, 2, 4, 4.5])
    T[m - 1, :m] = [rng.choice([1, 2, 3]) for _ in range(m)]
    T[m - 1, 2 * m] = rng.choice([6, 10, 12.5])
    T[:m, m:2 * m] = np.eye(m)
    T[m, :m] = [rng.choice([-3, -2, -1, -1, 0, 1, -0.75]) for _ in range(m)]
    return T


@pytest.mark.parametrize("m,seed", [(6, 0), (9, 1), (14, 2), (20, 3), (33, 4), (48, 5), (64, 6), (20, 7), (33, 8)])
def test_c_solve_bit_identical_on_irregular_and_tie_heavy_tableaux(m, seed):
    """Degenerate pivots, ties, negative and fractional entries, the stuck counter: trace, status and every byte of
    the final tableau equal the Python float oracle's (which tests/test_wild.py pins to the float-rebound reference)."""
    for T0 in (_irregular_canonical(m, seed), problems.tie_heavy_grid(m - m % 2, 2, seed)):
        mm = T0.shape[0] - 1
        grid = [[float(x) for x in row] for row in T0]
        pb = list(range(mm, 2 * mm))
        ptrace = []
        try:
            O.solve(grid, pb, ptrace)
            want_status = C.OPTIMAL
        except O.UnboundedError:
            want_status = C.UNBOUNDED
        T = T0.copy()
        basis = np.arange(mm, 2 * mm, dtype=np.int32)
        st, npiv, trace = C.solve(T, basis)
        assert st == want_status
        assert trace.tolist() == [list(p) for p in ptrace] and basis.tolist() == pb
        assert T.tobytes() == np.array(grid, dtype=np.float64).tobytes()


@pytest.mark.parametrize("m,row0,ml,seed,lo,hi", [(16, 0, 16, 0, 1, 9), (31, 10, 11, 3, 1, 9), (64, 48, 16, 1, 1, 127),
                                                  (40, 0, 13, 7, 2, 5)])
def test_c_fill_equals_the_python_statement_of_the_hashed_family(m, row0, ml, seed, lo, hi):
    """The synthetic C4 input exists three times -- `k_fill_dense_f64` on the device, `orc_fill_dense_f64` for the CPU
    arm of bench.py, `problems.hashed_dense_grid` in Python -- and every rank generates only its own rows: the C and
    the Python statement agree cell for cell on whole tableaux and on row slabs (the device one is compared with the
    Python one in tests/test_api_fp64_gpu.py)."""
    T = np.full((ml + 1, 2 * m + 1 + 3), -7.0)                  # padded rows: the fill must respect the leading dimension
    C.fill_dense(T, m, row0, seed, lo, hi)
    want = problems.hashed_dense_grid(m, seed, lo, hi, row0, ml).astype(np.float64)
    assert T[:, :2 * m + 1].tobytes() == np.ascontiguousarray(want).tobytes()
    assert (T[:, 2 * m + 1:] == -7.0).all() or (T[:, 2 * m + 1:] == 0.0).all()
