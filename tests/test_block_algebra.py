"""The algebra of the blocked fp64 loop (csrc/lps_f64_blk.cuh), restated with numpy on the CPU and checked against the
C oracle: deciding K pivots from a STALE tableau -- objective row and right-hand side kept current, entering column
and leaving row gathered from the stale tableau and sent through the block's earlier pivots one after the other --
and then applying all K to every cell gives, bit for bit, the tableau and the pivots of a run that rewrites the whole
tableau after every pivot (oracle/pivot_oracle.c, the reference's float path: tableau.py:295-308, simplex.py:251-284).

numpy rounds a product and a sum separately (no fused multiply-add), like the reference's Python floats; a zero
multiplier leaves the cell alone (tableau.py:272); a pivot row is replaced by its normalised copy (tableau.py:262-267).
This is the design check that runs without a GPU; tests/test_f64_blocks_gpu.py holds the CUDA kernels to the same oracle."""
import random

import numpy as np
import pytest

from linear_program_solver_b200 import problems


def _through(x, u, v, is_pivot_row):
    """cells x (array) through one pivot: multipliers u (array or scalar), pivot-row cells v (array or scalar)."""
    moved = x + (-u) * v                       # product rounded, then sum rounded
    out = np.where(u == 0.0, x, moved)
    return np.where(is_pivot_row, v, out)


def blocked_solve(T, basis, K, max_pivots=-1):
    """Simplex.solve on T (modified in place) in blocks of K pivots; returns (status, trace)."""
    m, n = T.shape[0] - 1, T.shape[1] - 1
    O = T[m].copy()                            # objective row, kept current
    B = T[:m, n].copy()                        # right-hand side, kept current
    z_entry = O[n]
    rule, stuck, pending, trace, status = 0, 0, False, [], None
    while status is None:
        U, V, R = [], [], []                   # the block: column snapshots (m+1), normalised rows (n+1), leaving rows
        for _ in range(K):
            if 0 <= max_pivots <= len(trace):
                status = 4
                break
            stuck_now = stuck
            if pending:
                if not (O[n] >= z_entry):
                    status = 6
                    break
                stuck_now = stuck + 1 if O[n] == z_entry else 0
            rule_now = 1 if (rule == 1 or not stuck_now < m + n) else 0
            neg = np.nonzero(O[:n] < 0.0)[0]
            if len(neg) == 0:
                status = 1
                break
            c = int(neg[0]) if rule_now == 1 else int(neg[np.argmin(O[neg])])   # argmin: first among equal minima
            # column c as it stands after the block's earlier pivots
            col = T[:m, c].copy()
            rows = np.arange(m)
            for lp in range(len(U)):
                col = _through(col, U[lp][:m], V[lp][c], rows == R[lp])
            pos = np.nonzero(col > 0.0)[0]
            if len(pos) == 0:
                status = 2
                break
            ratios = B[pos] / col[pos]
            r = int(pos[np.argmin(ratios)])                                      # lowest row among equal ratios
            # row r through the earlier pivots, then normalised
            row = T[r].copy()
            for lp in range(len(U)):
                row = V[lp].copy() if R[lp] == r else _through(row, U[lp][r], V[lp], False)
            inv = 1.0 / col[r]
            v = row * inv if inv != 1.0 else row
            u = np.append(col, O[c])
            # objective row and right-hand side move on
            O = _through(O, u[m], v, False)
            B = _through(B, col, v[n], rows == r)
            U.append(u); V.append(v); R.append(r)
            basis[r] = c
            trace.append((r, c, rule_now))
            rule, stuck, pending = rule_now, stuck_now, rule_now == 0
        # the fused update: every cell through the block's pivots, in order
        allrows = np.arange(m + 1)[:, None]
        for lp in range(len(U)):
            T[:] = _through(T, U[lp][:, None], V[lp][None, :], allrows == R[lp])
        if status is None and 0 <= max_pivots <= len(trace):
            status = 4
    return status, trace


def _dense(m, seed):
    return np.array(problems.to_grid(*problems.dense_lp(m, seed)), dtype=np.float64)


def _sparse(m, ns, seed):
    rng = random.Random(seed)
    T = np.zeros((m + 1, ns + m + 1))
    for i in range(m - 1):
        for j in range(ns):
            if rng.random() < 0.35:
                T[i, j] = rng.choice([-3, -2, -1, 1, 2, 3, 4])
        T[i, -1] = rng.randint(0, 12)
    T[m - 1, :ns] = 1.0
    T[m - 1, -1] = 40 + rng.randint(0, 20)
    for i in range(m):
        T[i, ns + i] = 1.0
    T[m, :ns] = [-rng.randint(0, 5) for _ in range(ns)]
    return T


CASES = [("dense", 12, 0), ("dense", 40, 3), ("dense", 70, 1), ("sparse", 20, 2), ("sparse", 33, 5), ("ties", 48, 0),
         ("ties", 64, 6)]


@pytest.mark.parametrize("K", [1, 2, 5, 16])
@pytest.mark.parametrize("family,m,seed", CASES)
def test_blocked_decisions_and_fused_update_equal_pivot_by_pivot(family, m, seed, K):
    from oracle import c_oracle as C
    if family == "dense":
        T0, b0 = _dense(m, seed), np.arange(m, 2 * m, dtype=np.int32)
    elif family == "sparse":
        ns = m + 10
        T0, b0 = _sparse(m, ns, seed), np.arange(ns, ns + m, dtype=np.int32)
    else:
        T0, b0 = problems.tie_heavy_grid(m, 4, seed), np.arange(m, 2 * m, dtype=np.int32)
    want = T0.copy()
    wb = b0.copy()
    wst, wk, wtrace = C.solve(want, wb)
    got = T0.copy()
    gb = b0.copy()
    st, trace = blocked_solve(got, gb, K)
    assert st == wst and [list(t) for t in trace] == wtrace.tolist()
    assert gb.tolist() == wb.tolist()
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("max_pivots", [0, 1, 7, 16, 17])
def test_blocked_prefix_stops(max_pivots):
    from oracle import c_oracle as C
    T0, b0 = _dense(30, 4), np.arange(30, 60, dtype=np.int32)
    want, wb = T0.copy(), b0.copy()
    wst, wk, wtrace = C.solve(want, wb, max_pivots=max_pivots)
    got, gb = T0.copy(), b0.copy()
    st, trace = blocked_solve(got, gb, 16, max_pivots=max_pivots)
    assert st == wst and [list(t) for t in trace] == wtrace.tolist() and got.tobytes() == want.tobytes()


def test_blocked_algebra_reproduces_the_reference_float_runs():
    """tests/golden/dense_float.json was recorded from the UNMODIFIED reference with its arithmetic type rebound to float
    (tools/make_golden.py): pivots, basis, objective and basic solution as hex floats.  The blocked loop's algebra has
    to reproduce them exactly -- not only the C oracle."""
    from conftest import load_golden
    for rec in load_golden("dense_float"):
        m = rec["m"]
        T = np.array(problems.to_grid(*problems.dense_lp(m, rec["seed"], rec["lo"], rec["hi"])), dtype=np.float64)
        basis = np.arange(m, 2 * m, dtype=np.int32)
        for K in (16, 3):
            got, gb = T.copy(), basis.copy()
            st, trace = blocked_solve(got, gb, K)
            assert st == 1 and [list(t) for t in trace] == rec["pivots"], (m, rec["seed"], K)
            assert gb.tolist() == rec["basis"]
            assert float(-got[m, 2 * m]).hex() == rec["objective"]
            assert {str(int(gb[i])): float(got[i, 2 * m]).hex() for i in range(m)} == rec["bfs"]
