"""Row-sharded fp64 path on real GPUs.  With one visible GPU the shards are emulated sequentially by
exchanging the candidate slots by hand (no concurrent kernels wait on each other); with >= 2 GPUs the
NCCL path is covered by tools/check_sharded_nccl.py under torchrun (see tools/gpu_round.sh)."""
import numpy as np
import pytest

from linear_program_solver_b200 import problems

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,parts,seed", [(64, 2, 0), (96, 3, 1), (250, 4, 2)])
def test_shards_emulated_on_one_gpu_match_oracle(m, parts, seed):
    import torch
    from linear_program_solver_b200.engine import F64Engine
    from linear_program_solver_b200.sharded import shard_rows
    from oracle import c_oracle as C
    n = 2 * m
    engs = []
    for g in range(parts):
        row0, ml = shard_rows(m, g, parts)
        e = F64Engine(ml, n)
        e.set_shard(row0, m, g, parts)
        e.fill_dense(seed, 1, 9)
        engs.append(e)
    basis = np.arange(m, n, dtype=np.int32)
    for e in engs:
        e.solve_begin(basis, None, -1, 4096)
    slot = engs[0].cand_send.numel()
    for _ in range(20 * m):
        for e in engs:
            e.step_select()
        torch.cuda.synchronize()
        gathered = torch.cat([e.cand_send for e in engs])          # what the NCCL all-gather delivers
        for e in engs:
            e.cand_all.copy_(gathered)
        torch.cuda.synchronize()
        for e in engs:
            e.step_update()
        st = {e.poll() for e in engs}
        assert len(st) == 1
        if next(iter(st))[0] != 0:
            break
    T = problems.hashed_dense_grid(m, seed).astype(np.float64)
    ob = np.arange(m, n, dtype=np.int32)
    ost, ok, otrace = C.solve(T, ob)
    for g, e in enumerate(engs):
        b = basis.copy()
        st, k, trace = e.solve_end(b, None, 4096)
        assert (st, k) == (ost, ok) and trace.tolist() == otrace.tolist() and b.tolist() == ob.tolist()
        row0, ml = shard_rows(m, g, parts)
        got = e.download()
        assert got[:ml].tobytes() == T[row0:row0 + ml].tobytes() and got[ml].tobytes() == T[m].tobytes()
        e.close()


@pytest.mark.parametrize("m,parts,seed", [(64, 2, 0), (250, 4, 2), (333, 3, 3)])
def test_peer_store_exchange_between_engines_of_one_process(m, parts, seed):
    """The fused exchange of k_select_f64 (headers and the pivot row written straight into every rank's exchange
    block, sequence flags) with all ranks on the visible GPU(s): each engine has its own stream, so the select
    kernels of a pivot run concurrently and wait for each other exactly like ranks on different GPUs do."""
    import torch
    from linear_program_solver_b200.engine import F64Engine
    from linear_program_solver_b200.sharded import shard_rows
    from oracle import c_oracle as C
    n = 2 * m
    ndev = torch.cuda.device_count()
    engs = []
    for g in range(parts):
        row0, ml = shard_rows(m, g, parts)
        engs.append(F64Engine(ml, n, device=g % ndev))
    blocks = [e.p2p_block(parts) for e in engs]
    for g, e in enumerate(engs):
        row0, ml = shard_rows(m, g, parts)
        e.p2p_attach_local(row0, m, g, parts, blocks)
        e.fill_dense(seed, 1, 9)
    basis = np.arange(m, n, dtype=np.int32)
    for e in engs:
        e.solve_begin(basis, None, -1, 4096)
    done = False
    for _ in range(40 * m // 16 + 2):
        for _ in range(16):                                         # a burst without any host synchronisation
            for e in engs:
                e.step_select()
            for e in engs:
                e.step_update()
        st = {e.poll() for e in engs}
        assert len(st) == 1, st
        if next(iter(st))[0] != 0:
            done = True
            break
    assert done
    T = problems.hashed_dense_grid(m, seed).astype(np.float64)
    ob = np.arange(m, n, dtype=np.int32)
    ost, ok, otrace = C.solve(T, ob)
    for g, e in enumerate(engs):
        b = basis.copy()
        st, k, trace = e.solve_end(b, None, 4096)
        assert (st, k) == (ost, ok) and trace.tolist() == otrace.tolist() and b.tolist() == ob.tolist()
        row0, ml = shard_rows(m, g, parts)
        got = e.download()
        assert got[:ml].tobytes() == T[row0:row0 + ml].tobytes() and got[ml].tobytes() == T[m].tobytes()
        e.close()


@pytest.mark.parametrize("m,parts,seed", [(48, 3, 0), (64, 4, 6), (60, 2, 2), (80, 4, 2)])
def test_ties_across_shards_pick_the_lowest_global_row(m, parts, seed):
    """Peer-store exchange on a tableau built so that the minimum ratio is attained on several shards at once."""
    import torch
    from linear_program_solver_b200.engine import F64Engine
    from linear_program_solver_b200.sharded import shard_rows
    from oracle import c_oracle as C
    n = 2 * m
    T = problems.tie_heavy_grid(m, parts, seed)
    ndev = torch.cuda.device_count()
    engs = []
    for g in range(parts):
        row0, ml = shard_rows(m, g, parts)
        engs.append(F64Engine(ml, n, device=g % ndev))
    blocks = [e.p2p_block(parts) for e in engs]
    for g, e in enumerate(engs):
        row0, ml = shard_rows(m, g, parts)
        e.p2p_attach_local(row0, m, g, parts, blocks)
        e.upload(np.ascontiguousarray(np.vstack([T[row0:row0 + ml], T[m:m + 1]])))
    basis = np.arange(m, n, dtype=np.int32)
    for e in engs:
        e.solve_begin(basis, None, -1, 8192)
    done = False
    for _ in range(200):
        for _ in range(16):
            for e in engs:
                e.step_select()
            for e in engs:
                e.step_update()
        st = {e.poll() for e in engs}
        assert len(st) == 1, st
        if next(iter(st))[0] != 0:
            done = True
            break
    assert done
    want = T.copy()
    ob = np.arange(m, n, dtype=np.int32)
    ost, ok, otrace = C.solve(want, ob)
    ties = _count_cross_shard_ties(T, otrace, m, parts)
    assert ties >= 3, ties                                   # the construction really produced cross-shard ties
    for g, e in enumerate(engs):
        b = basis.copy()
        st, k, trace = e.solve_end(b, None, 8192)
        assert (st, k) == (ost, ok) and trace.tolist() == otrace.tolist() and b.tolist() == ob.tolist()
        row0, ml = shard_rows(m, g, parts)
        got = e.download()
        assert got[:ml].tobytes() == want[row0:row0 + ml].tobytes() and got[ml].tobytes() == want[m].tobytes()
        e.close()


def _count_cross_shard_ties(T0, trace, m, parts):
    """Replays the oracle's trace and counts the pivots whose minimum ratio is attained on more than one shard."""
    from linear_program_solver_b200.sharded import shard_rows
    from oracle import c_oracle as C
    T = T0.copy()
    owner = np.empty(m, dtype=int)
    for g in range(parts):
        row0, ml = shard_rows(m, g, parts)
        owner[row0:row0 + ml] = g
    ties = 0
    for r, c, _ in trace.tolist():
        col, rhs = T[:m, c], T[:m, -1]
        pos = col > 0
        ratios = np.where(pos, rhs / np.where(pos, col, 1.0), np.inf)
        at_min = np.nonzero(ratios == ratios.min())[0]
        assert at_min[0] == r
        ties += len(set(owner[at_min])) > 1
        C.pivot(T, r, c)
    return ties
