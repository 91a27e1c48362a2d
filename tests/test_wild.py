"""The irregular corpus (tests/golden/wild.json: 400 small LPs with mixed signs, zeros, rational data, negative right-hand
sides, partial slack bases, dependent rows, infeasible and unbounded problems, recorded from the UNMODIFIED reference by
tools/make_golden_wild.py) against the CPU oracle (`not gpu`) and against the CUDA path through the public classes
(`gpu`): constructor errors (type and message), phase-1 and phase-2 pivot traces, bases, objective and basic solution
as exact strings, marks and the size the constructor leaves."""
from fractions import Fraction

import pytest

from conftest import load_golden

WILD = load_golden("wild")
ERR = {"ValueError": ValueError, "IndexError": IndexError, "AssertionError": AssertionError}


def _data(rec, num=Fraction):
    i = rec["input"]
    return ([num(Fraction(v)) for v in i["c"]], [[num(Fraction(v)) for v in r] for r in i["A"]],
            [num(Fraction(v)) for v in i["b"]])


def _fmt(x):
    """Exact results as 'p/q' strings; floats as C99 hex with -0.0 written as 0.0 (bit-identical except for the sign of
    a zero: whether `_z` starts as +0.0 or as -(0.0) depends on whether the caller ever called setZ)."""
    return (x + 0.0).hex() if isinstance(x, float) else str(Fraction(x))


def _w(s):
    return s[1:] if s == "-0x0.0p+0" else s


def test_corpus_covers_every_outcome():
    kinds = set()
    for r in WILD:
        kinds.add(r["ctor_error"][0] if "ctor_error" in r else ("unbounded" if "solve_error" in r else
                  ("phase1" if r["phase1_pivots"] else "canonical")))
    assert kinds == {"ValueError", "IndexError", "AssertionError", "unbounded", "phase1", "canonical"}
    assert not any(r.get("truncated") for r in WILD)


@pytest.mark.parametrize("arith", ["exact", "float"])
def test_oracle_matches_reference_on_the_wild_corpus(arith):
    from oracle import lpsol_oracle as O
    n_checked = 0
    for rec in WILD:
        want = rec if arith == "exact" else rec.get("float")
        if want is None:
            continue
        c, A, b = _data(rec, Fraction if arith == "exact" else float)
        n_checked += 1
        if "ctor_error" in want:
            with pytest.raises(ERR[want["ctor_error"][0]]) as ei:
                O.find_bfs(O.make_grid(c, A, b, 0, type(c[0]) if c else Fraction), [])
            if want["ctor_error"][0] != "IndexError":
                assert str(ei.value) == want["ctor_error"][1], rec["id"]
            continue
        if "solve_error" in want:
            with pytest.raises(AssertionError) as ei:
                O.simplex(c, A, b, num=type(c[0]))
            assert str(ei.value) == want["solve_error"][1], rec["id"]
            continue
        got = O.simplex(c, A, b, num=type(c[0]))
        assert [list(p) for p in got["phase1_pivots"]] == want["phase1_pivots"], rec["id"]
        assert [list(p) for p in got["pivots"]] == want["pivots"], rec["id"]
        assert got["initial_basis"] == want["initial_basis"] and got["basis"] == want["basis"], rec["id"]
        assert _fmt(got["objective"]) == _w(want["objective"]), rec["id"]
        assert {str(j): _fmt(v) for j, v in got["bfs"].items()} == {j: _w(v) for j, v in want["bfs"].items()}, rec["id"]
    assert n_checked >= 200


def _replay(rec, want, mode):
    from linear_program_solver_b200 import Simplex, Tableau
    c, A, b = _data(rec, Fraction if mode == "exact" else float)
    t = Tableau(len(A), len(c), mode=mode)
    t.setC(c)
    t.setB(b)
    t.setA(A)
    if "ctor_error" in want:
        with pytest.raises(ERR[want["ctor_error"][0]]) as ei:
            Simplex(t)
        if want["ctor_error"][0] != "IndexError":        # CPython's own message, not part of the contract
            assert str(ei.value) == want["ctor_error"][1]
        return
    s = Simplex(t)
    n1 = len(want["phase1_pivots"])
    assert [list(p) for p in s.pivot_log] == want["phase1_pivots"]
    assert s.getBasicSequence() == want["initial_basis"]
    if "solve_error" in want:
        with pytest.raises(AssertionError) as ei:
            s.solve()
        assert str(ei.value) == want["solve_error"][1]
    else:
        s.solve()
    assert [list(p) for p in s.pivot_log[n1:]] == want["pivots"]
    assert s.getBasicSequence() == want["basis"]
    assert _fmt(s.getObjValue()) == _w(want["objective"])
    assert {str(j): _fmt(v) for j, v in s.getBFS().items()} == {j: _w(v) for j, v in want["bfs"].items()}
    assert [bool(x) for x in t.getVarMarks()] == want["marks"]
    assert list(t.getTableauSize()) == want["m_n"]


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", range(8))
def test_exact_mode_matches_reference_on_the_wild_corpus(chunk):
    bad = []
    for rec in WILD[chunk::8]:
        try:
            _replay(rec, rec, "exact")
        except BaseException as e:  # collect: one report names every divergent record of the chunk
            bad.append((rec["id"], type(e).__name__, str(e)[:300]))
    assert not bad, bad


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", range(8))
def test_fp64_mode_matches_float_rebound_reference_on_the_wild_corpus(chunk):
    bad = []
    for rec in [r for r in WILD if "float" in r][chunk::8]:
        try:
            _replay(rec, rec["float"], "fp64")
        except BaseException as e:
            bad.append((rec["id"], type(e).__name__, str(e)[:300]))
    assert not bad, bad


@pytest.mark.gpu
def test_batched_node_lp_kernel_matches_reference_on_the_wild_corpus():
    """The batched kernel (`k_bnb_nodes`: its own constructor + solve, one LP per thread-block cluster) on the
    integer-data records, grouped by shape into batches: outcome (the exception the reference raises, through
    `raise_for_status`), pivot trace with rules, basis, objective and basic solution."""
    import numpy as np

    from linear_program_solver_b200 import _native as N
    from linear_program_solver_b200 import problems
    from linear_program_solver_b200.batch import solve_batch
    by_shape = {}
    for rec in WILD:
        if "float" in rec:
            c, A, b = _data(rec, int)
            by_shape.setdefault((len(A), len(c)), []).append((rec, np.asarray(problems.to_grid(c, A, b), dtype=np.int64)))
    assert len(by_shape) >= 40
    bad = []
    for (m, n), group in sorted(by_shape.items()):
        res = solve_batch(np.stack([g for _, g in group]), trace_cap=128)
        for k, (rec, _) in enumerate(group):
            try:
                if "ctor_error" in rec or "solve_error" in rec:
                    kind, msg = rec.get("ctor_error") or rec["solve_error"]
                    with pytest.raises(ERR[kind]) as ei:
                        res.raise_for_status(k)
                    assert not isinstance(ei.value, N.NativeError)
                    if kind != "IndexError":
                        assert str(ei.value) == msg
                    if "solve_error" in rec:
                        assert [list(p) for p in res.pivots(k)] == rec["phase1_pivots"] + rec["pivots"]
                    continue
                assert int(res.status[k]) == N.OPTIMAL, res.status_name(k)
                assert [list(p) for p in res.pivots(k)] == rec["phase1_pivots"] + rec["pivots"]
                assert [int(j) for j in res.basis[k][:len(rec["basis"])]] == rec["basis"]
                assert str(res.objective(k)) == rec["objective"]
                assert {str(j): str(v) for j, v in res.bfs(k).items()} == rec["bfs"]
            except BaseException as e:
                bad.append((rec["id"], (m, n), type(e).__name__, str(e)[:300]))
    assert not bad, "\n".join(map(str, bad))


MEDIUM = [(12, 20, 0), (12, 20, 7), (12, 20, 9), (20, 30, 0), (20, 30, 3), (20, 30, 4), (20, 30, 5), (20, 30, 8),
          (30, 50, 0), (30, 50, 3), (30, 50, 4), (30, 50, 7), (45, 70, 4), (45, 70, 7), (45, 70, 8)]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["exact", "fp64"])
@pytest.mark.parametrize("m,n,seed", MEDIUM, ids=lambda v: str(v))
def test_medium_irregular_lps_match_the_oracle(m, n, seed, mode):
    """The irregular family at 12..46 rows (optimal after 14..60 phase-1 pivots, infeasible, unbounded in phase 2,
    dependent rows; some with rational data) against the oracle, which the small corpus above pins to the reference:
    same exception, or same traces / basis / objective / solution -- exact strings, or floats bit for bit."""
    import random

    from linear_program_solver_b200 import Simplex, Tableau, problems
    from oracle import lpsol_oracle as O
    num = Fraction if mode == "exact" else float
    c, A, b = problems.irregular_lp(random.Random(7000 + seed * 31 + m), m, n)
    c, A, b = [num(v) for v in c], [[num(v) for v in r] for r in A], [num(v) for v in b]
    try:
        want = O.simplex(c, A, b, num=num)
    except (ValueError, IndexError, AssertionError) as e:
        want = e
    t = Tableau(len(A), len(c), mode=mode)
    t.setC(c)
    t.setB(b)
    t.setA(A)
    if isinstance(want, Exception):
        kind = next(k for k in (ValueError, IndexError, AssertionError) if isinstance(want, k))
        with pytest.raises(kind) as ei:
            Simplex(t).solve()
        if kind is not IndexError:
            assert str(ei.value) == str(want)
        return
    s = Simplex(t)
    assert [list(p) for p in s.pivot_log] == [list(p) for p in want["phase1_pivots"]]
    s.solve()
    assert [list(p) for p in s.pivot_log[len(want["phase1_pivots"]):]] == [list(p) for p in want["pivots"]]
    assert s.getBasicSequence() == want["basis"]
    assert _fmt(s.getObjValue()) == _fmt(want["objective"])
    assert {j: _fmt(v) for j, v in s.getBFS().items()} == {j: _fmt(v) for j, v in want["bfs"].items()}
