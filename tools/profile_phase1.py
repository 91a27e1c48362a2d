#!/usr/bin/env python3
"""cProfile of one equality-form exact solve (phase 1 + phase 2) through the public API, after a warm-up solve."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def solve(grid):
    from linear_program_solver_b200 import Simplex, Tableau
    t = Tableau.from_grid(grid, mode="exact")
    s = Simplex(t)
    s.solve()
    z = s.getObjValue()
    t._drop_engine()
    return z, len(s.pivot_log)


def main():
    from linear_program_solver_b200 import problems
    m, n = int(sys.argv[1]), int(sys.argv[2])
    grid = np.array(problems.to_grid(*problems.equality_lp(m, n, 0)), dtype=np.int64)
    solve(np.array(problems.to_grid(*problems.equality_lp(8, 14, 0)), dtype=np.int64))
    for _ in range(2):
        t0 = time.perf_counter()
        z, k = solve(grid)
        print(f"{m}x{n}: {k} pivots, {time.perf_counter() - t0:.3f} s")
    pr = cProfile.Profile()
    pr.enable()
    solve(grid)
    pr.disable()
    for key in ("cumulative", "tottime"):
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats(key).print_stats(18)
        print(buf.getvalue()[:5000])


if __name__ == "__main__":
    main()
