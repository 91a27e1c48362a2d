#!/usr/bin/env python3
"""cProfile of the end-to-end batched call (`solve_batch` + `objectives()`) on the C5 workload."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from linear_program_solver_b200 import problems
    from linear_program_solver_b200.batch import solve_batch
    count, m = 4096, 64
    grids = np.stack([problems.hashed_dense_grid(m, seed) for seed in range(count)]).astype(np.int32)
    solve_batch(grids[:8]).objectives()
    for _ in range(3):
        t0 = time.perf_counter()
        r = solve_batch(grids)
        t1 = time.perf_counter()
        r.objectives()
        t2 = time.perf_counter()
        print(f"solve_batch {t1 - t0:.4f} s, objectives {t2 - t1:.4f} s -> {count / (t2 - t0):.0f} node LPs/s")
    pr = cProfile.Profile()
    pr.enable()
    solve_batch(grids).objectives()
    pr.disable()
    buf = io.StringIO()
    pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(14)
    print(buf.getvalue()[:4000])


if __name__ == "__main__":
    main()
