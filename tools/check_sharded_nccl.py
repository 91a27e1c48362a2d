#!/usr/bin/env python3
"""Multi-GPU parity check of the row-sharded fp64 solver over NCCL (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_sharded_nccl.py --rows 512
Every rank solves its shard; rank 0 checks pivot trace, basis and its rows against the CPU oracle."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=512)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--exchange", choices=["p2p", "nccl"], default="p2p")
    ap.add_argument("--block", type=int, default=None, help="pivots per block with the peer-store exchange (0: rank-1 loop)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from linear_program_solver_b200 import problems
    from linear_program_solver_b200.sharded import ShardedF64Solver
    from oracle import c_oracle as C
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m, n = a.rows, 2 * a.rows
    sv = ShardedF64Solver(m, n, rank, world, device=local, exchange=a.exchange, block=a.block)
    sv.eng.fill_dense(a.seed, 1, 9)
    basis = np.arange(m, n, dtype=np.int32)
    st, k, trace = sv.solve(basis, None, trace_cap=1 << 20)
    T = problems.hashed_dense_grid(m, a.seed).astype(np.float64)
    ob = np.arange(m, n, dtype=np.int32)
    ost, ok, otrace = C.solve(T, ob)
    got = sv.eng.download()
    good = (st, k) == (ost, ok) and trace.tolist() == otrace.tolist() and basis.tolist() == ob.tolist() \
        and got[:sv.m_local].tobytes() == T[sv.row0:sv.row0 + sv.m_local].tobytes() and got[sv.m_local].tobytes() == T[m].tobytes()
    flag = torch.tensor([1 if good else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"sharded check ({a.exchange} exchange, block {sv.block}): world={world} m={m} pivots={k} status={st} -> {'OK' if flag.item() else 'MISMATCH'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
