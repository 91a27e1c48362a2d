#!/usr/bin/env python3
"""C4 to optimality, row-sharded over the GPUs of one box (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        tools/bench_c4_tto_sharded.py
The fp64 32768 x 65536 tableau of the seeded dense family is generated on the devices and solved with no pivot limit
(blocks of 16, one peer-store exchange per pivot).  Rank 0 prints pivots, seconds, objective; the one-GPU solve
(profiles/r02d_c4_time_to_optimal.json) made 193 374 pivots and ended at -136829.96482983703 -- the sharded solve must
print the same two numbers, bit for bit."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=32768)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from linear_program_solver_b200.sharded import ShardedF64Solver
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m, n = a.m, 2 * a.m
    sv = ShardedF64Solver(m, n, rank, world, device=local)
    sv.eng.fill_dense(a.seed, 1, 9)
    basis = np.arange(m, n, dtype=np.int32)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st, k, _ = sv.solve(basis, None)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    obj = sv.eng.read_row(sv.m_local)
    if rank == 0:
        print(json.dumps({"case": f"C4 fp64 {m}x{n}, seed {a.seed}, to optimality on {world} GPUs", "status": int(st), "pivots": int(k),
                          "seconds": dt, "pivots_per_s": k / dt, "block": sv.block, "objective": float(-obj[n]),
                          "objective_hex": float(-obj[n]).hex(), "min_reduced_cost": float(obj[:n].min()),
                          "structural_in_basis": int((basis < m).sum())}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
