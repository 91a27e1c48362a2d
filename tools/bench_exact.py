#!/usr/bin/env python3
"""Micro-benchmark of the exact (multi-modular Bareiss) pivot loop at a given size.

    python tools/bench_exact.py --m 1000 --hi 127 --pivots 20        # config C3
Prints primes, residue bytes, per-update-kernel time (CUDA events on the launching stream), achieved
GB/s on the algorithmic bytes 8 P (m+1)(n+1), and pivots/s of the whole loop (CRT + decisions + update).
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
    ap.add_argument("--m", type=int, default=1000)
    ap.add_argument("--hi", type=int, default=127)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--pivots", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-prof", action="store_true", help="no per-launch events: the solve replays a CUDA graph on small sizes")
    a = ap.parse_args()
    import torch
    from linear_program_solver_b200 import problems
    from linear_program_solver_b200.exact import ExactEngine
    m, n = a.m, 2 * a.m
    c, A, b = problems.dense_lp(m, a.seed, 1, a.hi)
    grid = np.array(problems.to_grid(c, A, b), dtype=np.int64)
    eng = ExactEngine(m, n)
    t0 = time.perf_counter()
    eng.upload(grid)
    eng.sync()
    up = time.perf_counter() - t0
    info = eng.info()
    basis = np.arange(m, 2 * m, dtype=np.int32)
    st, k0, _ = eng.solve(basis, None, max_pivots=a.warmup)
    if not a.no_prof:
        eng.prof_reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st, k, trace = eng.solve(basis, None, max_pivots=a.pivots, trace_cap=a.pivots)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms, nu, na = eng.prof_read() if not a.no_prof else (0.0, 0, 0)
    info_end = eng.info()
    P = info_end["primes"]
    bytes_per = 8 * P * (m + 1) * (n + 1)
    out = {"m": m, "n": n, "hi": a.hi, "primes": P, "primes_at_upload": info["primes"], "allocated_primes": info_end["allocated_primes"],
           "extensions": info_end["extensions"], "bound_bits": info["bound_bits"], "residue_GB": info["residue_bytes"] / 1e9,
           "upload_s": up, "status": st, "pivots": k, "wall_s": wall, "pivots_per_s": k / wall,
           "update_ms_avg": ms / max(nu, 1), "update_launches_timed": nu, "launches": na,
           "update_GBps": bytes_per / (ms / max(nu, 1) * 1e-3) / 1e9 if nu else None,
           "loop_GBps": bytes_per * k / wall / 1e9, "first_pivots": trace[:6].tolist() if trace is not None else None}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
