#!/usr/bin/env python3
"""Where the host time of a small solve goes: cProfile of `Tableau.from_grid -> Simplex -> solve -> getObjValue`
over the ten C1 problems (exact 50 x 100), after a warm-up.  Prints the top entries by cumulative and own time."""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(grids, mode):
    from linear_program_solver_b200 import Simplex, Tableau
    out = []
    for g in grids:
        t = Tableau.from_grid(g, mode=mode)
        s = Simplex(t)
        s.solve()
        out.append(s.getObjValue())
        t._drop_engine()
    return out


def main():
    import numpy as np
    from linear_program_solver_b200 import problems
    mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    as_array = len(sys.argv) > 3 and sys.argv[3] == "array"
    grids = [problems.to_grid(*problems.dense_lp(m, seed)) for seed in range(10)]
    if mode == "fp64":
        grids = [np.array(g, dtype=np.float64) for g in grids]
    elif as_array:
        grids = [np.array(g, dtype=np.int64) for g in grids]
    run(grids[:2], mode)
    for _ in range(4):
        t0 = time.perf_counter()
        run(grids, mode)
        print(f"{mode} m={m} array={as_array}: {1e3 * (time.perf_counter() - t0) / len(grids):.2f} ms per solve (unprofiled)")
    pr = cProfile.Profile()
    pr.enable()
    run(grids, mode)
    pr.disable()
    for key in ("cumulative", "tottime"):
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats(key).print_stats(22)
        print(buf.getvalue()[:6000])


if __name__ == "__main__":
    main()
