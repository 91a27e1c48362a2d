"""Host-side protocol of the row-sharded solver on CPU: world_size 2 and 3 over gloo.

The product engine is CUDA-only, so the ranks here plug a stand-in engine built on the CPU oracle
(test infrastructure) into `ShardedF64Solver(engine_factory=...)`.  What is under test is the
driver: row partition, candidate all-gather, global leaving-row choice (lowest ratio, then lowest
global row), identical decisions on every rank, burst / poll / max_pivots logic.  The result must
equal the single-process oracle solve of the whole tableau bit-for-bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HEADER = 64      # bytes of the candidate header (ratio f64, row i64, column i32, rule i32, stuck i64, pad)


class OracleShardEngine:
    """CPU stand-in with the stepping surface of F64Engine (candidate slot layout included)."""

    def __init__(self, m, n, device=None):
        self.m, self.n = m, n
        self.T = np.zeros((m + 1, n + 1))
        self.row0, self.m_total, self.rank, self.nranks = 0, m, 0, 1
        self._alloc(1)

    def _alloc(self, nranks):
        self.slot = HEADER + 8 * (self.n + 1)
        self.cand_send = torch.zeros(self.slot, dtype=torch.uint8)
        self.cand_all = torch.zeros(self.slot * nranks, dtype=torch.uint8) if nranks > 1 else self.cand_send

    def set_shard(self, row0, m_total, rank, nranks):
        self.row0, self.m_total, self.rank, self.nranks = row0, m_total, rank, nranks
        self._alloc(nranks)

    def close(self):
        pass

    def solve_begin(self, basis, marks, max_pivots, trace_cap):
        self.basis = basis.copy()
        self.z_entry = self.T[self.m, self.n]
        self.status, self.rule, self.stuck, self.npiv, self.max_pivots = 0, 0, 0, 0, max_pivots
        self.pending = False
        self.trace = []

    def step_select(self):
        from oracle import c_oracle as C
        if self.status != 0:
            return
        if self.max_pivots >= 0 and self.npiv >= self.max_pivots:
            self.status = 4
            return
        stuck, rule = self.stuck, self.rule
        if self.pending:
            if not (self.T[self.m, self.n] >= self.z_entry):      # simplex.py:133 assert z <= obj_val
                self.status = 6
                return
            stuck = stuck + 1 if self.T[self.m, self.n] == self.z_entry else 0
        if rule == 0 and not stuck < self.m_total + self.n:
            rule = 1
        c = C.price(self.T, rule)
        if c < 0:
            self.status = 1
            return
        r = C.ratio(self.T, c)
        hdr = np.zeros(8, dtype=np.int64)
        buf = self.cand_send.numpy()
        ratio = self.T[r, self.n] / self.T[r, c] if r >= 0 else 0.0
        buf[0:8] = np.frombuffer(np.float64(ratio).tobytes(), dtype=np.uint8)
        buf[8:16] = np.frombuffer(np.int64(self.row0 + r if r >= 0 else -1).tobytes(), dtype=np.uint8)
        buf[16:24] = np.frombuffer(np.array([c, rule], dtype=np.int32).tobytes(), dtype=np.uint8)
        buf[24:32] = np.frombuffer(np.int64(stuck).tobytes(), dtype=np.uint8)
        if r >= 0:
            p = self.T[r, c]
            inv = 1.0 / p
            row = self.T[r] if inv == 1.0 else self.T[r] * inv
            buf[HEADER:] = np.frombuffer(row.tobytes(), dtype=np.uint8)

    def step_update(self):
        if self.status != 0:
            return
        buf = self.cand_all.numpy()
        best = None
        for k in range(self.nranks):
            s = buf[k * self.slot:(k + 1) * self.slot]
            ratio = float(np.frombuffer(s[0:8].tobytes(), dtype=np.float64)[0])
            row = int(np.frombuffer(s[8:16].tobytes(), dtype=np.int64)[0])
            if row < 0:
                continue
            if best is None or ratio < best[0] or (ratio == best[0] and row < best[1]):
                best = (ratio, row, k)
        if best is None:
            self.status = 2
            return
        _, grow, k = best
        s = buf[k * self.slot:(k + 1) * self.slot]
        c, rule = np.frombuffer(s[16:24].tobytes(), dtype=np.int32)
        stuck = int(np.frombuffer(s[24:32].tobytes(), dtype=np.int64)[0])
        v = np.frombuffer(s[HEADER:].tobytes(), dtype=np.float64)
        lr = grow - self.row0
        mine = 0 <= lr < self.m                    # (lr == self.m is the next rank's first row, not the objective row)
        for i in range(self.m + 1):
            if mine and i == lr:
                continue
            mult = -self.T[i, c]
            if mult != 0.0:
                self.T[i] = self.T[i] + mult * v       # product then sum, like the reference
        if mine:
            self.T[lr] = v
        self.basis[grow] = c
        self.trace.append((grow, int(c), int(rule)))
        self.npiv += 1
        self.rule, self.stuck, self.pending = int(rule), stuck, rule == 0

    def poll(self):
        return self.status, self.npiv

    def solve_end(self, basis, marks, trace_cap):
        basis[:] = self.basis
        return self.status, self.npiv, np.array(self.trace[:trace_cap], dtype=np.int32).reshape(-1, 3)


def _grid(m, seed, world, family):
    from linear_program_solver_b200 import problems
    if family == "ties":            # minimum ratios attained on several shards at once (lowest GLOBAL row must win)
        return problems.tie_heavy_grid(m, world, seed)
    return problems.hashed_dense_grid(m, seed).astype(np.float64)


def _worker(rank, world, port, m, seed, max_pivots, family, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from linear_program_solver_b200 import problems
    from linear_program_solver_b200.sharded import ShardedF64Solver, shard_rows
    sv = ShardedF64Solver(m, 2 * m, rank, world, engine_factory=OracleShardEngine, burst=3)
    row0, ml = shard_rows(m, rank, world)
    if family == "dense":         # every rank generates its own rows, like the device fill
        sv.eng.T[:] = problems.hashed_dense_grid(m, seed, 1, 9, row0, ml).astype(np.float64)
    else:
        T = _grid(m, seed, world, family)
        sv.eng.T[:] = np.vstack([T[row0:row0 + ml], T[m:m + 1]])
    basis = np.arange(m, 2 * m, dtype=np.int32)
    st, k, trace = sv.solve(basis, None, max_pivots=max_pivots, trace_cap=10 ** 6)
    out[rank] = (st, k, trace.tolist(), basis.tolist(), row0, sv.eng.T.copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,m,seed,max_pivots,family", [(2, 24, 0, -1, "dense"), (3, 31, 1, -1, "dense"),
                                                             (2, 40, 2, 7, "dense"), (2, 60, 2, -1, "ties"),
                                                             (3, 48, 0, -1, "ties")])
def test_sharded_protocol_equals_single_process_oracle(world, m, seed, max_pivots, family):
    from linear_program_solver_b200 import problems
    from oracle import c_oracle as C
    mgr = mp.get_context("spawn").Manager()     # not fork: the pytest process is multi-threaded by now
    out = mgr.dict()
    port = 29500 + (os.getpid() + 7 * world + m) % 2000
    mp.spawn(_worker, args=(world, port, m, seed, max_pivots, family, out), nprocs=world, join=True)
    T = _grid(m, seed, world, family)
    basis = np.arange(m, 2 * m, dtype=np.int32)
    st, k, trace = C.solve(T, basis, max_pivots=max_pivots)
    assert len(out) == world and (family != "ties" or k >= 7)
    for rank in range(world):
        rst, rk, rtrace, rbasis, row0, shard = out[rank]
        assert (rst, rk) == (st, k) and rtrace == trace.tolist() and rbasis == basis.tolist()
        ml = shard.shape[0] - 1
        assert shard[:ml].tobytes() == T[row0:row0 + ml].tobytes()       # the rank's rows
        assert shard[ml].tobytes() == T[m].tobytes()                      # its replica of the objective row


def test_default_burst_is_the_same_on_every_rank_with_uneven_shards():
    """Every enqueued pivot is a collective with exchange="nccl": ranks whose shards differ by a row must not round
    their default burst differently (round-1 advisor finding)."""
    from linear_program_solver_b200.sharded import ShardedF64Solver, shard_rows
    differing = 0
    for m_total, n, world in [(142, 49999, 3), (1001, 3000, 3), (32769, 65538, 8), (233, 49999, 3), (12345, 24690, 7)]:
        bursts, old = set(), set()
        for rank in range(world):
            sv = ShardedF64Solver(m_total, n, rank, world, engine_factory=OracleShardEngine, exchange="nccl")
            bursts.add(sv.burst)
            ml = shard_rows(m_total, rank, world)[1]
            old.add(int(max(4, min(256, 2000.0 / max(4.0, (ml + 1) * (n + 1) * 16.0 / 5.0e6)))))
        assert len(bursts) == 1, (m_total, n, world, bursts)
        differing += len(old) > 1
    assert differing >= 1          # the per-rank formula really did disagree on one of these shapes


def test_every_rank_picks_the_same_loop_with_uneven_shards():
    """The library picks the blocked loop by the handle's OWN rows; the sharded driver decides from the largest shard
    (`default_block`) so that ranks whose shards straddle the 1.5 M-cell threshold still run the same loop."""
    from linear_program_solver_b200.sharded import default_block, shard_rows
    ld = (15998 + 1 + 15) // 16 * 16
    straddling = 0
    for world in (2, 3, 8):
        for m_total in range(100, 1200):
            own = {16 if (shard_rows(m_total, r, world)[1] + 1) * ld >= 1_500_000 else 0 for r in range(world)}
            assert default_block(m_total, world, ld) in (0, 16)
            if len(own) == 2:                              # the per-handle rule disagrees between these ranks ...
                straddling += 1
                assert default_block(m_total, world, ld) == 16   # ... the driver's rule gives one answer (the largest shard's)
    assert straddling >= 3


def test_node_lps_shard_by_lp_without_overlap():
    """C5 over several GPUs: contiguous, balanced, exhaustive shares."""
    from linear_program_solver_b200.batch import shard_lps
    for count in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            parts = [shard_lps(count, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == count
            assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
