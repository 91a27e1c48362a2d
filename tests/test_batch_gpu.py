"""GPU parity of the batched node-LP kernel (`lps_batch_*`, one LP per thread-block cluster) against the
reference fixtures and the oracle: the whole `Simplex(tab); solve()` -- phase 1, drive-out pivots, rule
switch -- must give the reference's pivot sequence, basis, objective and basic solution bit for bit."""
import random
from fractions import Fraction

import numpy as np
import pytest

from conftest import load_golden
from linear_program_solver_b200 import problems

pytestmark = pytest.mark.gpu


def _grid(c, A, b, z=0):
    return np.asarray(problems.to_grid(c, A, b, z), dtype=np.int64)


def _oracle(c, A, b):
    """What the reference does with this LP, through the pinned oracle: a dict or an exception."""
    from oracle import lpsol_oracle as O
    try:
        return O.simplex(c, A, b)
    except O.UnboundedError as e:
        return e
    except (ValueError, IndexError, AssertionError) as e:
        return e


def _check_against(res, k, want, m):
    """LP k of a BatchResult against an oracle result (dict) or the exception the reference raises."""
    from linear_program_solver_b200 import _native as N
    if isinstance(want, Exception):
        kind = next(t for t in (ValueError, IndexError, AssertionError) if isinstance(want, t))
        with pytest.raises(kind) as info:
            res.raise_for_status(k)
        if isinstance(want, (ValueError, IndexError)):
            assert str(info.value) == str(want)
        return
    assert int(res.status[k]) == N.OPTIMAL, res.status_name(k)
    assert res.pivots(k) == [tuple(p) for p in want["phase1_pivots"]] + [tuple(p) for p in want["pivots"]]
    assert [int(j) for j in res.basis[k]] == list(want["basis"])
    assert res.objective(k) == want["objective"]
    assert res.bfs(k) == want["bfs"]


def _check_fixture(res, k, rec):
    from linear_program_solver_b200 import _native as N
    assert int(res.status[k]) == N.OPTIMAL, res.status_name(k)
    assert res.pivots(k) == [tuple(p) for p in rec.get("phase1_pivots", [])] + [tuple(p) for p in rec["pivots"]]
    assert [int(j) for j in res.basis[k]] == rec["basis"]
    assert str(res.objective(k)) == rec["objective"]
    assert {str(j): str(v) for j, v in res.bfs(k).items()} == rec["bfs"]


def test_dense_family_matches_the_reference_fixtures():
    """tests/golden/dense_exact.json (unmodified reference), grouped by shape into batches."""
    from linear_program_solver_b200.batch import solve_batch
    recs = [r for r in load_golden("dense_exact") if r["hi"] == 9 and r["m"] <= 64 and not r.get("truncated")]
    by_m = {}
    for r in recs:
        by_m.setdefault(r["m"], []).append(r)
    assert 64 in by_m and 50 in by_m
    for m, group in sorted(by_m.items()):
        grids = np.stack([_grid(*problems.dense_lp(m, r["seed"], r["lo"], r["hi"])) for r in group])
        res = solve_batch(grids, trace_cap=4 * m)
        for k, r in enumerate(group):
            _check_fixture(res, k, r)
            assert res.npiv[k, 0] == 0 and res.npiv[k, 3] == 0      # canonical input: no phase 1


def test_phase1_family_matches_the_reference_fixtures():
    """Equality-form LPs: artificial columns, phase-1 solve, drive-out pivots, objective restore
    (simplex.py:36-108) against tests/golden/phase1.json."""
    from linear_program_solver_b200 import _native as N
    from linear_program_solver_b200.batch import solve_batch
    recs = load_golden("phase1")
    by_shape = {}
    for r in recs:
        by_shape.setdefault((r["m"], r["n"]), []).append(r)
    seen_phase1 = 0
    for (m, n), group in sorted(by_shape.items()):
        grids = np.stack([_grid(*problems.equality_lp(m, n, r["seed"])) for r in group])
        res = solve_batch(grids, trace_cap=8 * (m + n))
        for k, r in enumerate(group):
            if "ctor_error" in r:
                kind, msg = r["ctor_error"]
                with pytest.raises({"ValueError": ValueError, "IndexError": IndexError}[kind]) as info:
                    res.raise_for_status(k)
                assert str(info.value) == msg
                continue
            _check_fixture(res, k, r)
            seen_phase1 += len(r["phase1_pivots"]) > 0
            assert int(res.npiv[k, 3]) == m                         # no slack columns: every row gets an artificial
    assert seen_phase1 >= 10


def test_special_cases():
    """infeasible / redundant row / unbounded / textbook LP of the reference's own test."""
    from linear_program_solver_b200 import _native as N
    from linear_program_solver_b200.batch import solve_batch
    sp = load_golden("special")
    res = solve_batch(_grid([0, 0], [[1, 1], [1, 1]], [1, 2])[None], trace_cap=16)
    assert int(res.status[0]) == N.INFEASIBLE
    with pytest.raises(ValueError) as info:
        res.raise_for_status(0)
    assert [type(info.value).__name__, str(info.value)] == sp["infeasible"]["ctor_error"]
    res = solve_batch(_grid([1, 1, 1], [[1, 1, 0], [2, 2, 0], [0, 1, 1]], [2, 4, 3])[None], trace_cap=16)
    assert int(res.status[0]) == N.REDUNDANT_ROW
    with pytest.raises(IndexError):
        res.raise_for_status(0)
    assert res.pivots(0) == [tuple(p) for p in sp["redundant_row"]["pivots"]]
    res = solve_batch(_grid([-1, 0], [[-1, 1]], [1])[None], trace_cap=16)
    assert int(res.status[0]) == N.UNBOUNDED
    with pytest.raises(AssertionError):
        res.raise_for_status(0)
    tb = load_golden("textbook")
    res = solve_batch(_grid(tb["c"], tb["A"], tb["b"])[None], trace_cap=16)
    _check_fixture(res, 0, tb["simplex"])


def test_rule_switch_to_bland():
    """Loop B of Simplex.solve (simplex.py:138-146): degenerate LPs with the stuck limit lowered from the
    reference's m+n to 1 and 2 (no integer LP that cycles under the reference's Dantzig rule is known to
    us; Beale's has fractional rows and is covered by the single-LP exact engine).  The stale-objective
    counter (simplex.py:118/:134), the one-way switch and Bland's lowest-index pricing must follow the oracle
    run with the same limit."""
    from linear_program_solver_b200.batch import solve_batch
    for limit in (1, 2):
        for (m, k) in [(3, 4), (5, 6)]:
            rng = random.Random(77 * m + k + limit)
            lps = []
            while len(lps) < 64:
                A = [[rng.randint(-3, 5) for _ in range(k)] + [1 if i == j else 0 for j in range(m)] for i in range(m)]
                b = [0 if rng.random() < 0.7 else rng.randint(1, 4) for _ in range(m)]
                c = [rng.randint(-5, 1) for _ in range(k)] + [0] * m
                lps.append((c, A, b))
            res = solve_batch(np.stack([_grid(*lp) for lp in lps]), trace_cap=256, stuck_limit=limit)
            bland = 0
            for idx, (c, A, b) in enumerate(lps):
                from oracle import lpsol_oracle as O
                try:
                    want = O.simplex(c, A, b, stuck_limit=limit)
                except (ValueError, IndexError, AssertionError) as e:
                    want = e
                _check_against(res, idx, want, m)
                if isinstance(want, dict):
                    bland += sum(1 for p in want["pivots"] if p[2] == 1)
            assert bland >= 10


def _random_lp(rng, m, k, with_slack):
    """Small integer LP in the shapes a B&B node produces: <= rows with slacks (possibly negative rhs, which
    forces artificials), optionally equality rows without a slack."""
    A, b = [], []
    for i in range(m):
        row = [rng.randint(-4, 6) for _ in range(k)]
        slack = [0] * m
        if with_slack or rng.random() < 0.6:
            slack[i] = 1
        A.append(row + slack)
        b.append(rng.randint(-6, 20) if rng.random() < 0.8 else 0)
    c = [rng.randint(-6, 3) for _ in range(k)] + [0] * m
    return c, A, b


@pytest.mark.parametrize("m,k,with_slack", [(3, 3, True), (4, 5, False), (6, 6, True), (8, 10, False), (12, 9, False)])
def test_random_node_lps_against_the_oracle(m, k, with_slack):
    """Mixed outcomes in one batch (optimal, infeasible, unbounded, redundant rows) against the oracle."""
    from linear_program_solver_b200.batch import solve_batch
    rng = random.Random(1000 * m + k)
    lps = [_random_lp(rng, m, k, with_slack) for _ in range(96)]
    res = solve_batch(np.stack([_grid(*lp) for lp in lps]), trace_cap=16 * (m + k))
    kinds = set()
    for idx, lp in enumerate(lps):
        want = _oracle(*lp)
        kinds.add(type(want).__name__)
        _check_against(res, idx, want, m)
    assert "dict" in kinds and len(kinds) >= 2


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_every_cluster_width_gives_the_same_answer(cluster):
    """The distributed-shared-memory path (residue planes spread over 2/4/8 CTAs) against the fixtures."""
    from linear_program_solver_b200.batch import solve_batch
    recs = [r for r in load_golden("dense_exact") if r["hi"] == 9 and r["m"] == 16]
    grids = np.stack([_grid(*problems.dense_lp(16, r["seed"], 1, 9)) for r in recs] * 4)
    res = solve_batch(grids, trace_cap=64, cluster=cluster)
    for k in range(len(res)):
        _check_fixture(res, k, recs[k % len(recs)])
    recs = load_golden("phase1")
    group = [r for r in recs if (r["m"], r["n"]) == (8, 14) and "ctor_error" not in r]
    grids = np.stack([_grid(*problems.equality_lp(8, 14, r["seed"])) for r in group])
    res = solve_batch(grids, trace_cap=256, cluster=cluster)
    for k, r in enumerate(group):
        _check_fixture(res, k, r)


def test_c5_batch_of_64x128_nodes():
    """BASELINE.json config 5 shape: 64 x 128 node LPs, seeds 0..N-1 of the dense family.  Seeds with a
    reference fixture are compared bit for bit; every LP is checked through size-independent properties:
    the returned basic solution satisfies A x = b exactly, x >= 0, and its cost equals the objective."""
    from linear_program_solver_b200 import _native as N
    from linear_program_solver_b200.batch import plan_batch, solve_batch
    count, m = 256, 64
    lps = [problems.dense_lp(m, seed) for seed in range(count)]
    grids = np.stack([_grid(*lp) for lp in lps])
    art_cap, primes, bits = plan_batch(grids.astype(np.int32))
    assert art_cap == 0 and len(primes) <= 16
    res = solve_batch(grids, trace_cap=4 * m)
    fixtures = {r["seed"]: r for r in load_golden("dense_exact") if r["m"] == 64 and r["hi"] == 9}
    assert fixtures
    for seed, r in fixtures.items():
        _check_fixture(res, seed, r)
    allobj = res.objectives()                                    # vectorised decode == per-LP decode
    assert all(allobj[k] == res.objective(k) for k in range(0, count, 5))
    for k in range(0, count, 7):
        c, A, b = lps[k]
        assert int(res.status[k]) == N.OPTIMAL
        x = [Fraction(0)] * (2 * m)
        for j, v in res.bfs(k).items():
            x[j] = v
        assert all(v >= 0 for v in x)
        for i in range(m):
            assert sum(A[i][j] * x[j] for j in range(2 * m)) == b[i]
        assert sum(c[j] * x[j] for j in range(2 * m)) == res.objective(k)
        assert len(res.pivots(k)) == int(res.npiv[k, 2])


def test_device_planner_agrees_with_the_numpy_statement():
    """`lps_batch_plan` (artificial columns to reserve, Hadamard bound) against batch.missing_basic_rows / bound_bits."""
    from linear_program_solver_b200.batch import plan_batch
    rng = random.Random(5)
    dense = np.stack([_grid(*problems.dense_lp(20, seed)) for seed in range(40)]).astype(np.int32)
    eq = np.stack([_grid(*problems.equality_lp(8, 14, seed)) for seed in range(40)]).astype(np.int32)
    mixed = np.stack([_grid(*_random_lp(rng, 6, 6, False)) for _ in range(64)]).astype(np.int32)
    for grids in (dense, eq, mixed):
        k1, p1, b1 = plan_batch(grids, on_device=True)
        k2, p2, b2 = plan_batch(grids, on_device=False)
        assert k1 == k2
        assert abs(b1 - b2) <= 1 and abs(len(p1) - len(p2)) <= 2


@pytest.mark.parametrize("m,hi,count", [(24, 9, 48), (40, 127, 32), (64, 9, 96)])
def test_two_tier_batches_give_the_single_tier_answers(m, hi, count):
    """solve_batch first runs every LP on half the primes with the per-LP capacity check and re-runs the LPs that
    asked for more; pivots, bases, statuses and rational results must equal the run on the full prime set."""
    from linear_program_solver_b200.batch import solve_batch
    grids = np.stack([_grid(*problems.dense_lp(m, seed, 1, hi)) for seed in range(count)])
    one = solve_batch(grids, trace_cap=8 * m, two_tier=False)
    two = solve_batch(grids, trace_cap=8 * m, two_tier=True)
    assert not hasattr(one, "tiers")
    if hasattr(two, "tiers"):                                     # (a 2-prime plan has no shorter tier)
        assert two.tiers[0] == count and 0 <= two.tiers[1] <= count
    assert one.status.tolist() == two.status.tolist()
    assert one.npiv.tolist() == two.npiv.tolist() and one.basis.tolist() == two.basis.tolist()
    for k in range(count):
        assert one.pivots(k) == two.pivots(k)
        assert one.integers(k) == two.integers(k)
    assert one.objectives() == two.objectives()


def test_capacity_check_stops_an_lp_before_a_short_prime_set_overflows():
    """Two primes (about 60 bits) cannot hold a 24 x 48 solve: with the check on, every LP either finishes with the
    answers of the full prime set or ends with NEED_PRIMES -- never with a wrong number; merging the re-run LPs
    reproduces the single-tier batch."""
    from linear_program_solver_b200 import _native as N
    from linear_program_solver_b200.batch import BatchExactSolver, MixedRadix, _merge_tiers, plan_batch, solve_batch
    m, count = 24, 40
    grids = np.stack([_grid(*problems.dense_lp(m, seed)) for seed in range(count)]).astype(np.int32)
    art_cap, primes, _ = plan_batch(grids)
    primes = sorted(primes)
    full = solve_batch(grids, trace_cap=8 * m, two_tier=False)
    s = BatchExactSolver(count, m, 2 * m, primes[:2], art_cap=art_cap, trace_cap=8 * m)
    s.set_capacity_check(True)
    first = s.solve(grids, 64 * 3 * m)
    s.close()
    redo = np.nonzero(first.status == N.NEED_PRIMES)[0]
    assert 0 < redo.size                                          # the short set is too small for some LPs at least
    assert set(first.status.tolist()) <= {N.OPTIMAL, N.NEED_PRIMES}
    for k in np.nonzero(first.status == N.OPTIMAL)[0]:            # finished LPs carry the exact answers
        assert first.pivots(int(k)) == full.pivots(int(k)) and first.basis[k].tolist() == full.basis[k].tolist()
    s2 = BatchExactSolver(int(redo.size), m, 2 * m, primes, art_cap=art_cap, trace_cap=8 * m)
    second = s2.solve(np.ascontiguousarray(grids[redo]), 64 * 3 * m)
    s2.close()
    merged = _merge_tiers(first, second, redo, MixedRadix(primes))
    assert merged.status.tolist() == full.status.tolist() and merged.basis.tolist() == full.basis.tolist()
    assert merged.objectives() == full.objectives()
    assert all(merged.integers(k) == full.integers(k) for k in range(count))
