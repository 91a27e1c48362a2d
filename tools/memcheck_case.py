#!/usr/bin/env python3
"""A few tiny solves through every kernel family, meant to run under `compute-sanitizer --tool memcheck` where that
tool is allowed (it is closed on this round's GPU pool, so no sanitizer record exists yet):
fp64 and exact single tableaux (phase 1, drive-out, phase switch, phase 2, whole-tableau rules, row operations) and a
batch of node LPs.  Results are checked against the oracle so that a silent corruption cannot pass as a clean run."""
import os
import sys
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from linear_program_solver_b200 import Simplex, Tableau, problems
    from linear_program_solver_b200.batch import solve_batch
    from oracle import lpsol_oracle as O
    c, A, b = problems.equality_lp(6, 11, seed=1)
    for mode, num in (("fp64", float), ("exact", Fraction)):
        want = O.simplex([num(v) for v in c], [[num(v) for v in r] for r in A], [num(v) for v in b], num=num)
        t = Tableau(len(A), len(c), mode=mode)
        t.setC(c), t.setB(b), t.setA(A)
        s = Simplex(t)
        s.findPivotAll(), s.findPivotMaxIncrease()
        s.solve()
        assert s.getBasicSequence() == want["basis"] and s.getObjValue() == want["objective"], mode
        t.rowMult(0, 2), t.rowAddToObj(1, 3)
        t.getA()
        print(mode, "ok", len(s.pivot_log), "pivots", flush=True)
    c, A, b = problems.dense_lp(10, seed=2)
    t = Tableau.from_grid(np.array(problems.to_grid(c, A, b), dtype=np.int64), mode="exact")
    s = Simplex(t)
    s.solve()
    assert s.getObjValue() == O.simplex(c, A, b)["objective"]
    print("exact dense ok", flush=True)
    lps = [problems.dense_lp(8, seed=k) for k in range(3)]
    res = solve_batch(np.stack([np.asarray(problems.to_grid(*lp), dtype=np.int64) for lp in lps]), trace_cap=64)
    for k, lp in enumerate(lps):
        assert res.objective(k) == O.simplex(*lp)["objective"]
    print("batch ok", flush=True)


if __name__ == "__main__":
    main()
