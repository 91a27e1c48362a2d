"""Row-sharded fp64 tableau: one process per GPU, one exchange per pivot.

Rank g holds constraint rows [g*m/G, (g+1)*m/G) plus its own replica of the objective row.  Per
pivot: every rank prices the (identical) objective row and ratio-tests its own rows
(`lps_step_select`), the ranks exchange their candidates, every rank picks the same winner (lowest
ratio, then lowest global row: simplex.py:277) and applies the rank-1 update to its rows
(`lps_step_update`).  No host round trip inside a burst of pivots.  Two exchanges:

  exchange="p2p"   (default on GPUs) the kernels do it themselves with peer stores over NVLink: each rank
                   writes a 64-byte header into every peer's exchange block and waits for theirs, the
                   owner of the winning row writes the normalised row into every peer.  No collective
                   call; torch.distributed is used once, to pass the CUDA IPC handles around.  Pivots
                   run in BLOCKS (`lps_step_block`, csrc/lps_f64_blk.cuh): a look-ahead panel decides
                   up to 16 pivots -- one exchange per pivot, on O(m+n) numbers -- and one pass over
                   the local rows applies them all, so the tableau is read and written once per block
                   instead of once per pivot.  `block=0` keeps the rank-1 loop (`lps_step_select` +
                   `lps_step_update` per pivot), the previous round's path.
  exchange="nccl"  the candidates -- header followed by the normalised candidate row -- are all-gathered
                   with one NCCL collective per pivot (the baseline the p2p path is measured against, and
                   the protocol the CPU/gloo test exercises).

`engine_factory` exists so that the host-side protocol can be exercised on CPU (gloo) with a
stand-in engine supplied by the tests; the product default is the CUDA `F64Engine`.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

from . import _native as N


def shard_rows(m_total: int, rank: int, nranks: int) -> Tuple[int, int]:
    """(row0, m_local) of the contiguous row block owned by `rank`."""
    lo = m_total * rank // nranks
    hi = m_total * (rank + 1) // nranks
    return lo, hi - lo


def default_block(m_total: int, nranks: int, ld: int) -> int:
    """Pivots per block of the blocked loop for a row-sharded tableau: 16 from 1.5 M cells per shard on, else 0 (the
    rank-1 loop).  The library makes this choice from a handle's OWN rows; shards differ by a row and every rank must
    run the same loop, so the driver decides from the LARGEST shard and tells every handle."""
    return 16 if (-(-m_total // nranks) + 1) * ld >= 1_500_000 else 0


class ShardedF64Solver:
    def __init__(self, m_total: int, n: int, rank: int = 0, nranks: int = 1, device: Optional[int] = None,
                 group=None, engine_factory: Optional[Callable] = None, burst: Optional[int] = None,
                 exchange: Optional[str] = None, block: Optional[int] = None):
        self.m_total, self.n, self.rank, self.nranks, self.group = int(m_total), int(n), rank, nranks, group
        self.row0, self.m_local = shard_rows(m_total, rank, nranks)
        if self.m_local <= 0:
            raise ValueError(f"rank {rank} of {nranks} would own no rows of a {m_total}-row tableau")
        if engine_factory is None:
            from .engine import F64Engine
            engine_factory = F64Engine
        self.eng = engine_factory(self.m_local, self.n, device)
        if exchange is None:
            exchange = "p2p" if hasattr(self.eng, "p2p_export") else "nccl"
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        self.exchange = exchange
        if nranks > 1:
            if exchange == "p2p" and not self._attach_peers():
                # some rank could not map a peer's exchange block (no peer access / IPC between these GPUs): every
                # rank falls back to the collective exchange together
                self.exchange = "nccl"
            if self.exchange == "nccl":
                self.eng.set_shard(self.row0, self.m_total, rank, nranks)
        # pivots per block of the blocked loop (peer-store exchange only; the collective exchange is per pivot)
        self.block = 0
        if self.exchange == "p2p" and hasattr(self.eng, "block_size"):
            if block is None:
                block = default_block(self.m_total, nranks, int(self.eng.L.lps_grid_ld(self.n)))
            self.eng.set_block_size(block)
            self.block = self.eng.block_size()
        # pivots enqueued between two host checks: ~2 ms of device work, between 4 and 256.  The same number on every
        # rank -- every enqueued pivot is a collective with exchange="nccl" -- so it is derived from the LARGEST
        # shard, not from this rank's own (uneven shards differ by a row).
        rows_max = -(-self.m_total // nranks)
        est_us = max(4.0, (rows_max + 1) * (self.n + 1) * 16.0 / 5.0e6)
        self.burst = burst or int(max(4, min(256, 2000.0 / est_us)))
        if self.block:
            # blocks between two host checks: about 8 ms of device work (a check costs a stream synchronisation and a
            # small copy: tens of microseconds against a block of about 1 ms on an 8-GPU shard of C4)
            blocks = max(1, min(64, int(8000.0 / (est_us + 20.0 * self.block))))
            self.burst = burst or blocks * self.block
        self._active = False

    def close(self):
        self.eng.close()

    def _attach_peers(self):
        """All-gather the 64-byte CUDA IPC handles of the exchange blocks, map the peers' blocks.  True when every
        rank succeeded."""
        import torch
        import torch.distributed as dist
        mine = torch.from_numpy(self.eng.p2p_export(self.nranks)).to(f"cuda:{self.eng.device}")
        everyone = torch.empty(64 * self.nranks, dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        ok = 1
        try:
            self.eng.p2p_attach(self.row0, self.m_total, self.rank, self.nranks, everyone.cpu().numpy())
        except N.NativeError:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=mine.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)   # also: every rank has mapped every block
        return bool(flag.item())

    # -- one pivot: select, exchange, update -------------------------------------------------------
    def _exchange(self):
        if self.exchange == "p2p":
            return                                    # fused into lps_step_select
        if self.nranks > 1:
            import torch.distributed as dist
            if hasattr(self.eng, "on_stream"):        # order the collective on the engine's own stream
                with self.eng.on_stream():
                    dist.all_gather_into_tensor(self.eng.cand_all, self.eng.cand_send, group=self.group)
            else:
                dist.all_gather_into_tensor(self.eng.cand_all, self.eng.cand_send, group=self.group)

    def enqueue(self, pivots: int):
        """Enqueue `pivots` pivots without waiting for the device: blocks of `self.block` pivots (the last one of a
        pivot prefix stops early on the device), or (select, exchange, update) triples."""
        eng = self.eng
        if self.block:
            for _ in range(-(-pivots // self.block)):
                eng.step_block()
            return
        for _ in range(pivots):
            eng.step_select()
            self._exchange()
            eng.step_update()

    # -- Simplex.solve (simplex.py:110-148) over all shards ----------------------------------------
    def begin(self, basis: np.ndarray, marks: Optional[np.ndarray] = None, max_pivots: int = -1, trace_cap: int = 0):
        assert basis.dtype == np.int32 and basis.size == self.m_total
        self.eng.solve_begin(basis, marks, max_pivots, trace_cap)
        self._active = True

    def poll(self) -> Tuple[int, int]:
        return self.eng.poll()

    def end(self, basis: np.ndarray, marks: Optional[np.ndarray] = None, trace_cap: int = 0):
        self._active = False
        return self.eng.solve_end(basis, marks, trace_cap)

    def solve(self, basis: np.ndarray, marks: Optional[np.ndarray] = None, max_pivots: int = -1, trace_cap: int = 0):
        """Returns (status, pivots, trace) -- identical on every rank."""
        self.begin(basis, marks, max_pivots, trace_cap)
        done = 0
        while True:
            now = self.burst if max_pivots < 0 else min(self.burst, max_pivots - done)
            self.enqueue(now)
            done += now
            if now == 0:
                # lets the device report LPS_MAX_PIVOTS (or optimality)
                self.eng.step_block() if self.block else self.eng.step_select()
            status, _ = self.poll()
            if status != N.OK:
                break
        return self.end(basis, marks, trace_cap)
