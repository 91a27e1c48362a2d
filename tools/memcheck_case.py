#!/usr/bin/env python3
"""A few tiny solves through every kernel family, meant to run under `compute-sanitizer --tool memcheck` where that
tool is allowed (it is closed on this pool in both rounds; the script is then a plain end-to-end check):
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
        t = Tableau(len(A), len(c), mode=mode)
        t.setC(c), t.setB(b), t.setA(A)
        try:
            want = O.simplex([num(v) for v in c], [[num(v) for v in r] for r in A], [num(v) for v in b], num=num)
        except O.InfeasibleError as e:
            # the reference's float control flow calls this LP infeasible (a phase-1 optimum of -1.7e-13 is not 0):
            # the device path must say the same, with the same number
            try:
                Simplex(t)
                raise AssertionError("the device path found a basis where the reference's float run does not")
            except ValueError as e2:
                assert str(e2) == str(e), (str(e2), str(e))
            print(mode, "ok: infeasible in float arithmetic, like the reference:", e, flush=True)
            continue
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
    # the blocked fp64 loop: cluster panel, one-CTA panel, grid-wide panel, two emulated row shards with the peer-store
    # exchange; the persistent TMA-ring update behind each of them
    from linear_program_solver_b200.engine import F64Engine
    from linear_program_solver_b200.sharded import shard_rows
    from oracle import c_oracle as C
    m = 40
    T0 = np.array(problems.to_grid(*problems.dense_lp(m, seed=4)), dtype=np.float64)
    for ctas in (-8, 1, 3):
        e = F64Engine(m, 2 * m)
        e.set_block_size(16)
        e.set_panel_ctas(ctas)
        e.upload(T0)
        basis = np.arange(m, 2 * m, dtype=np.int32)
        st, k, _ = e.solve(basis, None)
        T, ob = T0.copy(), np.arange(m, 2 * m, dtype=np.int32)
        C.solve(T, ob)
        assert basis.tolist() == ob.tolist() and e.download().tobytes() == T.tobytes(), ctas
        e.close()
        print("blocked loop, panel shape", ctas, "ok", k, "pivots", flush=True)
    engs = []
    for g in range(2):
        row0, ml = shard_rows(m, g, 2)
        engs.append(F64Engine(ml, 2 * m))
    blocks = [e.p2p_block(2) for e in engs]
    for g, e in enumerate(engs):
        row0, ml = shard_rows(m, g, 2)
        e.set_block_size(16)
        e.p2p_attach_local(row0, m, g, 2, blocks)
        e.upload(np.ascontiguousarray(np.vstack([T0[row0:row0 + ml], T0[m:m + 1]])))
    basis = np.arange(m, 2 * m, dtype=np.int32)
    for e in engs:
        e.solve_begin(basis, None, -1, 0)
    for _ in range(40):
        for e in engs:
            e.step_block()
        if engs[0].poll()[0] != 0:
            break
    T, ob = T0.copy(), np.arange(m, 2 * m, dtype=np.int32)
    C.solve(T, ob)
    for g, e in enumerate(engs):
        bb = basis.copy()
        e.solve_end(bb, None, 0)
        row0, ml = shard_rows(m, g, 2)
        assert bb.tolist() == ob.tolist() and e.download()[:ml].tobytes() == T[row0:row0 + ml].tobytes()
        e.close()
    print("blocked loop, two shards with the peer-store exchange ok", flush=True)
    lps = [problems.dense_lp(8, seed=k) for k in range(3)]
    res = solve_batch(np.stack([np.asarray(problems.to_grid(*lp), dtype=np.int64) for lp in lps]), trace_cap=64)
    for k, lp in enumerate(lps):
        assert res.objective(k) == O.simplex(*lp)["objective"]
    print("batch ok", flush=True)


if __name__ == "__main__":
    main()
