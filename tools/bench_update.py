#!/usr/bin/env python3
"""Micro-benchmark of the fp64 rank-1 update kernel at a given size (device-resident input).

    python tools/bench_update.py --m 32768 --pivots 20
Prints per-update-kernel time (CUDA events on the launching stream), achieved GB/s on the
algorithmic bytes 16(m+1)(n+1) + 8(m+1) + 8(n+1), and pivots/s of the whole loop.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def fill_dense(eng, seed=0):
    """Seeded dense family of SURVEY 8(d), generated on the device (same distribution)."""
    import torch
    m, n = eng.m, eng.n
    g = torch.Generator(device=eng.grid.device)
    g.manual_seed(seed)
    T = eng.grid
    T.zero_()
    T[:m, :n - m] = torch.randint(1, 10, (m, n - m), generator=g, device=T.device).to(torch.float64)
    idx = torch.arange(m, device=T.device)
    T[idx, n - m + idx] = 1.0
    T[:m, n] = torch.randint(9 * m // 4, 9 * m // 2 + 1, (m,), generator=g, device=T.device).to(torch.float64)
    T[m, :n - m] = -torch.randint(1, 10, (n - m,), generator=g, device=T.device).to(torch.float64)
    torch.cuda.synchronize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=32768)
    ap.add_argument("--n", type=int, default=0, help="variables (default 2m); wider than 2m = the shape of one shard")
    ap.add_argument("--pivots", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-prof", action="store_true", help="no per-launch events: lps_solve replays a CUDA graph on small sizes")
    a = ap.parse_args()
    import torch
    from linear_program_solver_b200.engine import F64Engine
    m, n = a.m, (a.n or 2 * a.m)
    eng = F64Engine(m, n)
    fill_dense(eng)
    basis = np.arange(n - m, n, dtype=np.int32)
    st, k, _ = eng.solve(basis, None, max_pivots=a.warmup)
    if not a.no_prof:
        eng.prof_reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st, k, _ = eng.solve(basis, None, max_pivots=a.pivots)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms, nu, na = eng.prof_read() if not a.no_prof else (0.0, 0, 0)
    bytes_per = 16 * (m + 1) * (n + 1) + 8 * (m + 1) + 8 * (n + 1)
    kb = eng.block_size()
    # one update launch applies up to kb pivots (kb = 0: the rank-1 loop): bytes the launch has to move
    bytes_launch = 16 * (m + 1) * (n + 1) + max(kb, 1) * (8 * (m + 1) + 8 * (n + 1))
    out = {"m": m, "n": n, "block": kb, "status": st, "pivots": k, "wall_s": wall, "pivots_per_s": k / wall,
           "update_ms_avg": ms / max(nu, 1), "update_launches_timed": nu, "launches": na,
           "update_GBps": bytes_launch / (ms / max(nu, 1) * 1e-3) / 1e9 if nu else None,
           "rank1_equivalent_GBps": bytes_per * k / wall / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
