"""GPU parity of the fp64 engine (C-ABI level) against the CPU oracle.

Bit-exact: the kernels round exactly like the reference's float-rebound run
(product then sum, IEEE division), so whole tableaux must be byte-identical.
"""
import numpy as np
import pytest

from conftest import load_golden
from linear_program_solver_b200 import problems

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng_mod():
    from linear_program_solver_b200 import engine
    return engine


def _grid(m, seed, lo=1, hi=9):
    c, A, b = problems.dense_lp(m, seed, lo, hi)
    return np.array(problems.to_grid(c, A, b), dtype=np.float64)


def test_textbook_pivots_match_reference_fixture(eng_mod):
    g = load_golden("textbook")
    T = np.array(problems.to_grid(g["c"], g["A"], g["b"]), dtype=np.float64)
    e = eng_mod.F64Engine(2, 4)
    e.upload(T)
    for step in g["steps"]:
        assert e.pivot(*step["pivot"]) == 0
        out = e.download()
        from fractions import Fraction
        want_c = [float(Fraction(x)) for x in step["c"]]
        assert out[2, :4].tolist() == want_c
        assert out[2, 4] == float(Fraction(step["z_cell"]))
        assert out[:2, 4].tolist() == [float(Fraction(x)) for x in step["b"]]
        assert out[:2, :4].tolist() == [[float(Fraction(x)) for x in row] for row in step["a"]]
    e.close()


def test_zero_pivot_status(eng_mod):
    T = np.array(problems.to_grid(*problems.TEXTBOOK), dtype=np.float64)
    e = eng_mod.F64Engine(2, 4)
    e.upload(T)
    assert e.pivot(0, 3) == 3           # LPS_ZERO_PIVOT, tableau untouched
    assert e.download().tobytes() == T.tobytes()
    e.close()


@pytest.mark.parametrize("m,seed", [(8, 0), (50, 1), (100, 0), (200, 0), (333, 2)])
def test_single_pivots_bit_identical_to_oracle(eng_mod, m, seed):
    from oracle import c_oracle as C
    T = _grid(m, seed)
    e = eng_mod.F64Engine(m, 2 * m)
    e.upload(T)
    rng = np.random.default_rng(seed)
    for _ in range(6):
        r, c = int(rng.integers(0, m)), int(rng.integers(0, 2 * m))
        while T[r, c] == 0.0:
            c = int(rng.integers(0, 2 * m))
        assert C.pivot(T, r, c) == 0
        assert e.pivot(r, c) == 0
        assert e.download().tobytes() == T.tobytes()
    e.close()


@pytest.mark.parametrize("m,seed", [(8, 0), (50, 0), (50, 3), (100, 1), (200, 0), (400, 0)])
def test_solve_bit_identical_to_oracle(eng_mod, m, seed):
    from oracle import c_oracle as C
    T = _grid(m, seed)
    e = eng_mod.F64Engine(m, 2 * m)
    e.upload(T)
    basis = np.arange(m, 2 * m, dtype=np.int32)
    marks = np.zeros(2 * m, dtype=np.int32)
    marks[m:] = 1
    st, npiv, trace = e.solve(basis, marks, trace_cap=100 * m)
    obasis = np.arange(m, 2 * m, dtype=np.int32)
    ost, onpiv, otrace = C.solve(T, obasis)
    assert st == 1 and ost == C.OPTIMAL
    assert npiv == onpiv and trace.tolist() == otrace.tolist()
    assert basis.tolist() == obasis.tolist()
    assert sorted(np.nonzero(marks)[0].tolist()) == sorted(obasis.tolist())
    assert e.download().tobytes() == T.tobytes()
    e.close()


def test_solve_matches_float_rebound_reference_golden(eng_mod):
    for rec in load_golden("dense_float"):
        m = rec["m"]
        T = _grid(m, rec["seed"])
        e = eng_mod.F64Engine(m, 2 * m)
        e.upload(T)
        basis = np.arange(m, 2 * m, dtype=np.int32)
        st, npiv, trace = e.solve(basis, None, trace_cap=100 * m)
        assert st == 1 and trace.tolist() == rec["pivots"]
        assert basis.tolist() == rec["basis"]
        out = e.download()
        assert float(-out[m, 2 * m]).hex() == rec["objective"]
        assert {str(int(basis[i])): float(out[i, 2 * m]).hex() for i in range(m)} == rec["bfs"]
        e.close()


def test_beale_cycling_rule_switch(eng_mod):
    c, A, b = problems.BEALE
    T = np.array(problems.to_grid([float(x) for x in c], [[float(x) for x in r] for r in A], [float(x) for x in b]))
    e = eng_mod.F64Engine(3, 7)
    e.upload(T)
    basis = np.array([0, 1, 2], dtype=np.int32)
    st, npiv, trace = e.solve(basis, None, trace_cap=64)
    g = load_golden("special")["beale_float"]
    assert st == 1 and trace.tolist() == g["pivots"] and basis.tolist() == g["basis"]
    assert float(-e.download()[3, 7]).hex() == g["objective"]
    e.close()


def test_find_pivot_and_statuses(eng_mod):
    from oracle import c_oracle as C
    T = _grid(40, 7)
    e = eng_mod.F64Engine(40, 80)
    e.upload(T)
    for rule in (0, 1):
        c = C.price(T, rule)
        assert e.find_pivot(rule) == (C.ratio(T, c), c)
    assert not e.is_optimal()
    basis = np.arange(40, 80, dtype=np.int32)
    e.solve(basis)
    assert e.find_pivot(0) == "optimal" and e.find_pivot(1) == "optimal" and e.is_optimal()
    e.close()
    U = np.array(problems.to_grid([-1.0, 0.0], [[-1.0, 1.0]], [1.0]))
    e = eng_mod.F64Engine(1, 2)
    e.upload(U)
    assert e.find_pivot(0) == "unbounded" and e.find_pivot(1) == "unbounded"
    st, npiv, _ = e.solve(np.array([1], dtype=np.int32))
    assert st == 2 and npiv == 0
    e.close()


def test_max_pivots_prefix_then_continue(eng_mod):
    from oracle import c_oracle as C
    m = 120
    T = _grid(m, 4)
    e = eng_mod.F64Engine(m, 2 * m)
    e.upload(T)
    basis = np.arange(m, 2 * m, dtype=np.int32)
    st, npiv, trace = e.solve(basis, None, max_pivots=7, trace_cap=16)
    assert st == 4 and npiv == 7
    ob = np.arange(m, 2 * m, dtype=np.int32)
    ost, onpiv, otrace = C.solve(T, ob, max_pivots=7)
    assert trace.tolist() == otrace.tolist() and e.download().tobytes() == T.tobytes()
    e.close()


def test_is_canonical(eng_mod):
    from oracle import lpsol_oracle as O
    m = 30
    T = _grid(m, 1)
    e = eng_mod.F64Engine(m, 2 * m)
    e.upload(T)
    ok, cols = e.is_canonical()
    want = [-1] * m
    assert O.is_canonical(T.tolist(), want) and ok and cols.tolist() == want == list(range(m, 2 * m))
    T2 = T.copy()
    T2[3, m + 3] = 2.0          # row 3 loses its unit column
    T2[5, 2 * m] = -1.0         # negative rhs: not canonical, bcols untouched
    e.upload(T2)
    ok, cols = e.is_canonical()
    assert not ok and cols.tolist() == [-1] * m
    T2[5, 2 * m] = 1.0
    e.upload(T2)
    ok, cols = e.is_canonical()
    want = [-1] * m
    assert not O.is_canonical(T2.tolist(), want) and not ok and cols.tolist() == want
    e.close()


def test_large_prefix_matches_oracle(eng_mod):
    """2000 x 4000 (config 2): 12-pivot prefix, whole tableau byte-identical to the C oracle."""
    from oracle import c_oracle as C
    m = 2000
    rng = np.random.default_rng(0)
    T = np.zeros((m + 1, 2 * m + 1))
    T[:m, :m] = rng.integers(1, 10, size=(m, m))
    T[:m, m:2 * m] = np.eye(m)
    T[:m, 2 * m] = rng.integers(9 * m // 4, 9 * m // 2, size=m)
    T[m, :m] = -rng.integers(1, 10, size=m)
    e = eng_mod.F64Engine(m, 2 * m)
    e.upload(T)
    basis = np.arange(m, 2 * m, dtype=np.int32)
    st, npiv, trace = e.solve(basis, None, max_pivots=12, trace_cap=12)
    ob = np.arange(m, 2 * m, dtype=np.int32)
    C.solve(T, ob, max_pivots=12)
    _, _, otrace = (None, None, None)
    assert st == 4 and npiv == 12 and basis.tolist() == ob.tolist()
    assert e.download().tobytes() == T.tobytes()
    e.close()


def test_c2_full_solve_bit_identical_to_oracle(eng_mod):
    """Config 2 at full size, to optimality (about 3 600 pivots through the captured-graph loop): pivot count, trace,
    basis and the WHOLE final tableau byte-identical to the C oracle.  Size-independent properties of the result on
    top: every reduced cost and right-hand side is >= 0 and the basis columns are unit vectors with zero cost up to
    rounding (rowDiv multiplies by 1/a_rc, tableau.py:262-267, so they are not exact in floating point)."""
    from linear_program_solver_b200 import problems
    from oracle import c_oracle as C
    m = 2000
    T = problems.hashed_dense_grid(m, 0).astype(np.float64)
    e = eng_mod.F64Engine(m, 2 * m)
    e.upload(T)
    basis = np.arange(m, 2 * m, dtype=np.int32)
    st, npiv, trace = e.solve(basis, None, trace_cap=1 << 14)
    ob = np.arange(m, 2 * m, dtype=np.int32)
    ost, ok, otrace = C.solve(T, ob)
    assert (st, npiv) == (1, ok) and npiv > m and trace.tolist() == otrace.tolist() and basis.tolist() == ob.tolist()
    got = e.download()
    assert got.tobytes() == T.tobytes()
    assert (got[m, :2 * m] >= 0).all() and (got[:m, 2 * m] >= 0).all()
    cols = got[:m, basis]
    assert np.abs(cols - np.eye(m)).max() < 1e-9 and np.abs(got[m, basis]).max() < 1e-9
    e.close()


def _mem_available() -> int:
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 0


def _same_bits(a: np.ndarray, b: np.ndarray, rows: int = 1024) -> bool:
    """byte equality of two large float64 arrays without copying them (`tobytes` would)"""
    ua, ub = a.view(np.uint64), b.view(np.uint64)
    return a.shape == b.shape and all(np.array_equal(ua[i:i + rows], ub[i:i + rows]) for i in range(0, a.shape[0], rows))


def test_c4_full_size_prefix_bit_identical_to_oracle(eng_mod):
    """BASELINE.json config 4 at its full size, 32768 x 65536 (17.2 GB): the seeded family generated on the device
    equals the oracle's, and after a 19-pivot prefix of the solve loop -- one full block of 16 through the grid-wide
    panel and the fused update's fast path, then a short block of 3 -- the whole tableau, the pivot trace and the basis
    are byte-identical to the C oracle's (about 25 s: two downloads and 19 CPU pivots)."""
    import torch
    from oracle import c_oracle as C
    m = 32768
    need = (m + 1) * (2 * m + 1) * 8
    if _mem_available() < 2.4 * need or torch.cuda.mem_get_info()[0] < 1.1 * need:
        pytest.skip("needs about 40 GB of host memory and 19 GB of device memory")
    e = eng_mod.F64Engine(m, 2 * m)
    assert e.block_size() == 16                                   # the blocked loop is what runs at this size
    e.fill_dense(0, 1, 9)
    T = np.empty((m + 1, 2 * m + 1), dtype=np.float64)
    C.fill_dense(T, m, 0, 0, 1, 9)
    got = e.download()
    assert _same_bits(got, T)                                     # same input on both sides
    del got
    basis = np.arange(m, 2 * m, dtype=np.int32)
    st, npiv, trace = e.solve(basis, None, max_pivots=19, trace_cap=19)
    ob = np.arange(m, 2 * m, dtype=np.int32)
    ost, ok, otrace = C.solve(T, ob, max_pivots=19)
    assert (st, npiv) == (4, 19) and ok == 19
    assert trace.tolist() == otrace.tolist() and basis.tolist() == ob.tolist()
    got = e.download()
    assert _same_bits(got, T)
    e.close()
