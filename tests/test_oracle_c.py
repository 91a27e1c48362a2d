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
