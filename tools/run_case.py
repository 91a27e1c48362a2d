#!/usr/bin/env python3
"""One solve through the public API, for profiling under ncu:
    python tools/run_case.py phase1 200 340        exact equality-form LP: phase 1, drive-out, device phase switch, phase 2
    python tools/run_case.py rules 1000            exact dense LP, 40 pivots, then findPivotAll + findPivotMaxIncrease
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from linear_program_solver_b200 import Simplex, Tableau, problems
    kind = sys.argv[1]
    if kind == "phase1":
        m, n = int(sys.argv[2]), int(sys.argv[3])
        grid = np.array(problems.to_grid(*problems.equality_lp(m, n, 0)), dtype=np.int64)
        t0 = time.perf_counter()
        t = Tableau.from_grid(grid, mode="exact")
        s = Simplex(t)
        n1 = len(s.pivot_log)
        s.solve()
        print(f"phase1 {m}x{n}: {n1} + {len(s.pivot_log) - n1} pivots, {time.perf_counter() - t0:.3f} s, "
              f"on device: {t.phase2_on_device}, primes {t._eng.info()}")
    else:
        m = int(sys.argv[2])
        hi = 127 if m >= 1000 else 9
        grid = np.array(problems.to_grid(*problems.dense_lp(m, 0, 1, hi)), dtype=np.int64)
        t = Tableau.from_grid(grid, mode="exact")
        s = Simplex(t)
        s.solve(max_pivots=40)
        t0 = time.perf_counter()
        allp = s.findPivotAll()
        t1 = time.perf_counter()
        mx = s.findPivotMaxIncrease()
        t2 = time.perf_counter()
        print(f"rules {m}x{2 * m}: findPivotAll {len(allp)} pivots in {t1 - t0:.4f} s, findPivotMaxIncrease {mx} in {t2 - t1:.4f} s")


if __name__ == "__main__":
    main()
