"""Pivots in blocks (csrc/lps_f64_blk.cuh): the look-ahead panel + fused update must be bit-identical to the
pivot-by-pivot reference path -- pivot trace, basis, marks and every byte of the tableau against the C oracle
(oracle/pivot_oracle.c, itself pinned to the float-rebound reference run), for every block size, for both panel
launch shapes, at prefix stops inside a block, on tie-heavy and irregular tableaux."""
import random

import numpy as np
import pytest

from linear_program_solver_b200 import problems

pytestmark = pytest.mark.gpu


def _grid(m, seed, lo=1, hi=9):
    c, A, b = problems.dense_lp(m, seed, lo, hi)
    return np.array(problems.to_grid(c, A, b), dtype=np.float64)


def _engine(T, block, ctas=None):
    from linear_program_solver_b200.engine import F64Engine
    m, n = T.shape[0] - 1, T.shape[1] - 1
    e = F64Engine(m, n)
    e.set_block_size(block)
    if ctas is not None:
        e.set_panel_ctas(ctas)
    e.upload(T)
    return e


def _check_solve(T0, basis0, block, ctas=None, max_pivots=-1):
    from oracle import c_oracle as C
    m, n = T0.shape[0] - 1, T0.shape[1] - 1
    e = _engine(T0, block, ctas)
    basis = basis0.copy()
    marks = np.zeros(n, dtype=np.int32)
    marks[basis] = 1
    st, npiv, trace = e.solve(basis, marks, max_pivots=max_pivots, trace_cap=64 * (m + n))
    T = T0.copy()
    ob = basis0.copy()
    ost, ok, otrace = C.solve(T, ob, max_pivots=max_pivots)
    assert (st, npiv) == (ost, ok)
    assert trace.tolist() == otrace.tolist()
    assert basis.tolist() == ob.tolist()
    assert sorted(np.nonzero(marks)[0].tolist()) == sorted(set(ob.tolist()))
    assert e.download().tobytes() == T.tobytes()
    e.close()
    return ok


@pytest.mark.parametrize("block", [1, 2, 3, 5, 8, 13, 16])
@pytest.mark.parametrize("m,seed", [(8, 0), (50, 3), (200, 0)])
def test_block_sizes_bit_identical_to_oracle(block, m, seed):
    _check_solve(_grid(m, seed), np.arange(m, 2 * m, dtype=np.int32), block)


@pytest.mark.parametrize("ctas", [1, 2, 7, 64, 148, -8, -16])
@pytest.mark.parametrize("m,seed", [(50, 1), (333, 2)])
def test_panel_launch_shapes_agree(ctas, m, seed):
    """One CTA of 1024 threads, grids that synchronise through memory and one thread-block cluster with the panel in
    distributed shared memory (-8, -16) make the same decisions."""
    _check_solve(_grid(m, seed), np.arange(m, 2 * m, dtype=np.int32), 16, ctas=ctas)


@pytest.mark.parametrize("max_pivots", [0, 1, 7, 15, 16, 17, 33, 64])
def test_prefix_stops_inside_a_block(max_pivots):
    m = 120
    T = _grid(m, 5)
    _check_solve(T, np.arange(m, 2 * m, dtype=np.int32), 16, max_pivots=max_pivots)
    _check_solve(T, np.arange(m, 2 * m, dtype=np.int32), 16, ctas=5, max_pivots=max_pivots)
    _check_solve(T, np.arange(m, 2 * m, dtype=np.int32), 16, ctas=-8, max_pivots=max_pivots)


@pytest.mark.parametrize("m,parts,seed", [(48, 3, 0), (64, 4, 6), (80, 4, 2)])
@pytest.mark.parametrize("ctas", [1, 6, -8, -16])
def test_tie_heavy_tableaux(m, parts, seed, ctas):
    """Ties in the ratio test, degenerate pivots (zero right-hand sides), rows that leave again within a block."""
    T = problems.tie_heavy_grid(m, parts, seed)
    _check_solve(T, np.arange(m, 2 * m, dtype=np.int32), 16, ctas=ctas)


def _sparse_canonical(m, n_struct, seed):
    """Canonical tableau with many exact zeros (zero multipliers: rows the reference leaves alone, tableau.py:272),
    mixed signs and a bounding row so that it stays bounded."""
    rng = random.Random(seed)
    n = n_struct + m
    T = np.zeros((m + 1, n + 1))
    for i in range(m - 1):
        for j in range(n_struct):
            if rng.random() < 0.35:
                T[i, j] = rng.choice([-3, -2, -1, 1, 2, 3, 4])
        T[i, n] = rng.randint(0, 12)
    T[m - 1, :n_struct] = 1.0
    T[m - 1, n] = 40 + rng.randint(0, 20)
    for i in range(m):
        T[i, n_struct + i] = 1.0
    T[m, :n_struct] = [-rng.randint(0, 5) for _ in range(n_struct)]
    return T


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("ctas", [1, 3, -8])
def test_sparse_mixed_sign_tableaux(seed, ctas):
    m, ns = 30 + 3 * seed, 40 + 5 * seed
    T = _sparse_canonical(m, ns, seed)
    k = _check_solve(T, np.arange(ns, ns + m, dtype=np.int32), 16, ctas=ctas)
    assert k > 0


def test_beale_cycling_rule_switch_inside_blocks():
    """10 Dantzig pivots at objective 0 (the stuck counter trips at m+n = 10), then Bland: the rule switch falls
    inside a block."""
    from conftest import load_golden
    c, A, b = problems.BEALE
    T = np.array(problems.to_grid([float(x) for x in c], [[float(x) for x in r] for r in A], [float(x) for x in b]))
    g = load_golden("special")["beale_float"]
    for block, ctas in [(16, 1), (16, 4), (3, 1), (7, 2), (16, -8), (5, -16)]:
        e = _engine(T, block, ctas)
        basis = np.array([0, 1, 2], dtype=np.int32)
        st, npiv, trace = e.solve(basis, None, trace_cap=64)
        assert st == 1 and trace.tolist() == g["pivots"] and basis.tolist() == g["basis"]
        assert float(-e.download()[3, 7]).hex() == g["objective"]
        e.close()


def test_unbounded_is_reported_after_the_decided_pivots_are_applied():
    """min -x1 - x2, x1 - x2 + s1 = 1, -x1 + ... : the first pivots are fine, then a column has no positive entry."""
    from oracle import c_oracle as C
    T0 = np.array([[1.0, -1.0, 1.0, 0.0, 1.0],
                   [-1.0, 0.0, 0.0, 1.0, 2.0],
                   [-1.0, -1.0, 0.0, 0.0, 0.0]])
    for block, ctas in [(16, 1), (16, 2), (1, 1), (16, -8)]:
        e = _engine(T0, block, ctas)
        basis = np.array([2, 3], dtype=np.int32)
        st, npiv, trace = e.solve(basis, None, trace_cap=16)
        T = T0.copy()
        ob = np.array([2, 3], dtype=np.int32)
        ost, ok, otrace = C.solve(T, ob)
        assert ost == C.UNBOUNDED and (st, npiv) == (ost, ok) and trace.tolist() == otrace.tolist()
        assert e.download().tobytes() == T.tobytes()
        e.close()


def test_stepping_blocks_and_rank1_steps_can_be_mixed():
    """lps_step_block and lps_step_select / lps_step_update on the same solve: each loop re-reads the objective row
    after the other one touched the grid."""
    from oracle import c_oracle as C
    from linear_program_solver_b200.engine import F64Engine
    m = 90
    T0 = _grid(m, 11)
    e = F64Engine(m, 2 * m)
    e.set_block_size(4)
    e.upload(T0)
    basis = np.arange(m, 2 * m, dtype=np.int32)
    e.solve_begin(basis, None, -1, 4096)
    for rnd in range(400):
        if rnd % 3 == 0:
            e.step_block()
        else:
            e.step_select()
            e.step_update()
        if rnd % 8 == 7 and e.poll()[0] != 0:
            break
    st, k, trace = e.solve_end(basis, None, 4096)
    T = T0.copy()
    ob = np.arange(m, 2 * m, dtype=np.int32)
    ost, ok, otrace = C.solve(T, ob)
    assert (st, k) == (ost, ok) and trace.tolist() == otrace.tolist() and basis.tolist() == ob.tolist()
    assert e.download().tobytes() == T.tobytes()
    e.close()


@pytest.mark.parametrize("m,seed,block", [(1200, 0, 16), (2000, 0, 16), (700, 3, 7)])
def test_medium_tableaux_take_the_cluster_panel_by_default(m, seed, block):
    """The default plan of a medium tableau: the panel as one cluster (8 CTAs here), whole solve against the oracle."""
    T = problems.hashed_dense_grid(m, seed).astype(np.float64)
    from linear_program_solver_b200.engine import F64Engine
    from oracle import c_oracle as C
    e = F64Engine(m, 2 * m)
    e.set_block_size(block)
    e.upload(T)
    basis = np.arange(m, 2 * m, dtype=np.int32)
    st, npiv, trace = e.solve(basis, None, trace_cap=1 << 16)
    ob = np.arange(m, 2 * m, dtype=np.int32)
    ost, ok, otrace = C.solve(T, ob)
    assert (st, npiv) == (ost, ok) and trace.tolist() == otrace.tolist() and basis.tolist() == ob.tolist()
    assert e.download().tobytes() == T.tobytes()
    e.close()


def test_upper_medium_tableau_takes_the_16_cta_cluster_panel_by_default():
    """5000 x 10000 (400 MB): the panel does not fit 8 CTAs' shared memory any more, the default plan is the cluster of
    16 (non-portable cluster size); a 40-pivot prefix is byte-identical to the oracle."""
    from oracle import c_oracle as C
    from linear_program_solver_b200.engine import F64Engine
    m = 5000
    e = F64Engine(m, 2 * m)
    assert e.block_size() == 16
    e.fill_dense(1, 1, 9)
    basis = np.arange(m, 2 * m, dtype=np.int32)
    st, npiv, trace = e.solve(basis, None, max_pivots=40, trace_cap=64)
    T = np.empty((m + 1, 2 * m + 1))
    C.fill_dense(T, m, 0, 1, 1, 9)
    ob = np.arange(m, 2 * m, dtype=np.int32)
    ost, ok, otrace = C.solve(T, ob, max_pivots=40)
    assert (st, npiv) == (ost, ok) and trace.tolist() == otrace.tolist() and basis.tolist() == ob.tolist()
    assert e.download().tobytes() == T.tobytes()
    e.close()


def test_large_tableau_takes_the_multi_cta_panel_by_default():
    """9000 x 18000 (1.3 GB): the default plan is a multi-CTA panel and the streaming update; a 40-pivot prefix is
    byte-identical to the oracle."""
    from oracle import c_oracle as C
    from linear_program_solver_b200.engine import F64Engine
    m = 9000
    e = F64Engine(m, 2 * m)
    e.fill_dense(0, 1, 9)
    basis = np.arange(m, 2 * m, dtype=np.int32)
    st, npiv, trace = e.solve(basis, None, max_pivots=40, trace_cap=64)
    T = np.empty((m + 1, 2 * m + 1))
    C.fill_dense(T, m, 0, 0, 1, 9)
    ob = np.arange(m, 2 * m, dtype=np.int32)
    ost, ok, otrace = C.solve(T, ob, max_pivots=40)
    assert (st, npiv) == (ost, ok) and trace.tolist() == otrace.tolist() and basis.tolist() == ob.tolist()
    assert e.download().tobytes() == T.tobytes()
    e.close()


@pytest.mark.parametrize("m,parts,seed,block", [(64, 2, 0, 16), (250, 4, 2, 16), (333, 3, 3, 5), (80, 4, 2, 16)])
def test_blocks_with_the_peer_store_exchange_between_engines_of_one_process(m, parts, seed, block):
    """Row-sharded blocks: every engine runs its panel concurrently on its own stream; per pivot the panels exchange
    candidate headers and the owner stores the normalised row into every engine's panel (peer stores), exactly like
    ranks on different GPUs."""
    import torch
    from linear_program_solver_b200.engine import F64Engine
    from linear_program_solver_b200.sharded import shard_rows
    from oracle import c_oracle as C
    n = 2 * m
    tie = (m == 80)
    T0 = problems.tie_heavy_grid(m, parts, seed) if tie else problems.hashed_dense_grid(m, seed).astype(np.float64)
    ndev = torch.cuda.device_count()
    engs = []
    for g in range(parts):
        row0, ml = shard_rows(m, g, parts)
        engs.append(F64Engine(ml, n, device=g % ndev))
    blocks = [e.p2p_block(parts) for e in engs]
    for g, e in enumerate(engs):
        row0, ml = shard_rows(m, g, parts)
        e.set_block_size(block)
        e.p2p_attach_local(row0, m, g, parts, blocks)
        e.upload(np.ascontiguousarray(np.vstack([T0[row0:row0 + ml], T0[m:m + 1]])))
    basis = np.arange(m, n, dtype=np.int32)
    for e in engs:
        e.solve_begin(basis, None, -1, 8192)
    done = False
    for _ in range(400):
        for _ in range(4):
            for e in engs:
                e.step_block()
        st = {e.poll() for e in engs}
        assert len(st) == 1, st
        if next(iter(st))[0] != 0:
            done = True
            break
    assert done
    T = T0.copy()
    ob = np.arange(m, n, dtype=np.int32)
    ost, ok, otrace = C.solve(T, ob)
    for g, e in enumerate(engs):
        b = basis.copy()
        st, k, trace = e.solve_end(b, None, 8192)
        assert (st, k) == (ost, ok) and trace.tolist() == otrace.tolist() and b.tolist() == ob.tolist()
        row0, ml = shard_rows(m, g, parts)
        got = e.download()
        assert got[:ml].tobytes() == T[row0:row0 + ml].tobytes() and got[ml].tobytes() == T[m].tobytes()
        e.close()
