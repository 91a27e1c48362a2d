#!/usr/bin/env python3
"""Time-to-optimal through the public API (`Tableau.from_grid` -> `Simplex` -> `solve()` -> `getObjValue()`),
next to the CPU oracle where the oracle finishes in bounded time (BASELINE.json metric: "time-to-optimal vs ref").

    python tools/bench_tto.py [--cpu-seconds 60]

Cases (seeded dense family of SURVEY.md section 8d):
  C1  exact 50 x 100, seeds 0..9       GPU vs oracle/lpsol_oracle.py (the reference's Fraction simplex, 1 core)
  X2  exact 200 x 400, seed 0          GPU vs the oracle when it fits the CPU budget
  C3  exact 1000 x 2000, 8-bit         GPU only (the oracle needs hours); objective digits reported
  C2  fp64 2000 x 4000                 GPU vs oracle/pivot_oracle.c (all host cores), bit-identical objective checked
  P1  exact equality-form LPs          phase 1 + drive-out + phase switch on the device + phase 2; 60 x 100 vs the oracle
One JSON line per case.  Upload, set-up, solve and read-back are all inside the timed region.
"""
import argparse
import json
import os
import sys
import time
from fractions import Fraction

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def gpu_solve(grid, mode):
    import torch
    from linear_program_solver_b200 import Simplex, Tableau
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t = Tableau.from_grid(grid, mode=mode)
    s = Simplex(t)
    s.solve()
    z = s.getObjValue()
    basis = list(s.getBasicSequence())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    npiv = len(s.pivot_log)
    t._drop_engine()
    return dt, z, basis, npiv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-seconds", type=float, default=60.0, help="skip an oracle run expected to take longer")
    ap.add_argument("--skip-c3", action="store_true")
    ap.add_argument("--only-phase1", action="store_true", help="only the equality-form (phase 1 + phase 2) cases")
    a = ap.parse_args()
    from linear_program_solver_b200 import problems
    from oracle import c_oracle as C
    from oracle import lpsol_oracle as O

    gpu_solve(problems.to_grid(*problems.dense_lp(8, 1)), "exact")          # warm-up: library load, contexts
    gpu_solve(np.array(problems.to_grid(*problems.dense_lp(8, 1)), dtype=np.float64), "fp64")

    # ---- P1: equality-form LPs (no slack basis): artificial problem, drive-out pivots, phase switch on the device ----
    for (m, n, with_oracle) in ((60, 100, True), (200, 340, False), (400, 680, False)):
        c, A, b = problems.equality_lp(m, n, 0)
        dt, z, basis, k = gpu_solve(np.array(problems.to_grid(c, A, b), dtype=np.int64), "exact")
        z = Fraction(z)
        line = {"case": f"P1 exact equality-form {m}x{n}, seed 0 (phase 1 + phase 2)", "gpu_seconds": dt, "pivots": k,
                "objective_numerator_bits": z.numerator.bit_length(), "objective_denominator_bits": z.denominator.bit_length(),
                "cpu_oracle_seconds": None}
        if with_oracle:
            t0 = time.perf_counter()
            want = O.simplex(c, A, b, num=Fraction)
            line.update(cpu_oracle_seconds=time.perf_counter() - t0, identical=bool(z == want["objective"] and basis == want["basis"]))
            line["speedup"] = line["cpu_oracle_seconds"] / dt
        print(json.dumps(line), flush=True)
    if a.only_phase1:
        return

    # ---- C1: exact 50 x 100, ten seeds -------------------------------------------------------------
    tg = tc = 0.0
    piv = 0
    for seed in range(10):
        c, A, b = problems.dense_lp(50, seed)
        dt, z, basis, k = gpu_solve(problems.to_grid(c, A, b), "exact")
        t0 = time.perf_counter()
        want = O.simplex(c, A, b, num=Fraction)
        tc += time.perf_counter() - t0
        assert z == want["objective"] and basis == want["basis"], f"C1 seed {seed}: results differ"
        tg += dt
        piv += k
    print(json.dumps({"case": "C1 exact 50x100, seeds 0..9", "gpu_seconds": tg, "cpu_oracle_seconds": tc, "pivots": piv,
                      "speedup": tc / tg, "cpu": "oracle/lpsol_oracle.py, Fraction, 1 core", "identical": True}), flush=True)

    # ---- X2: exact 200 x 400 -------------------------------------------------------------------------
    c, A, b = problems.dense_lp(200, 0)
    dt, z, basis, k = gpu_solve(np.array(problems.to_grid(c, A, b), dtype=np.int64), "exact")
    line = {"case": "exact 200x400, seed 0", "gpu_seconds": dt, "pivots": k, "objective": str(z)}
    t0 = time.perf_counter()
    part = O.simplex(c, A, b, num=Fraction, max_pivots=3)
    per = (time.perf_counter() - t0) / 3
    if per * k * 2.0 <= a.cpu_seconds:                      # Fractions grow: later pivots cost more than the first ones
        t0 = time.perf_counter()
        want = O.simplex(c, A, b, num=Fraction)
        line.update(cpu_oracle_seconds=time.perf_counter() - t0, identical=bool(z == want["objective"] and basis == want["basis"]))
        line["speedup"] = line["cpu_oracle_seconds"] / dt
    else:
        line.update(cpu_oracle_seconds=None, cpu_first_pivots_seconds_each=per)
    print(json.dumps(line), flush=True)

    # ---- C3: exact 1000 x 2000, 8-bit -----------------------------------------------------------------
    if not a.skip_c3:
        c, A, b = problems.dense_lp(1000, 0, 1, 127)
        dt, z, basis, k = gpu_solve(np.array(problems.to_grid(c, A, b), dtype=np.int64), "exact")
        z = Fraction(z)
        print(json.dumps({"case": "C3 exact 1000x2000, 8-bit, seed 0", "gpu_seconds": dt, "pivots": k,
                          "objective_numerator_bits": z.numerator.bit_length(),
                          "objective_denominator_bits": z.denominator.bit_length(), "objective_float": float(z),
                          "cpu_oracle_seconds": None}), flush=True)

    # ---- C2: fp64 2000 x 4000, and the same LP in exact mode -------------------------------------------------
    m = 2000
    T = problems.hashed_dense_grid(m, 0).astype(np.float64)
    dtx, zx, basisx, kx = gpu_solve(problems.hashed_dense_grid(m, 0).astype(np.int64), "exact")
    dt, z, basis, k = gpu_solve(T.copy(), "fp64")
    zx = Fraction(zx)
    print(json.dumps({"case": "C2's LP in exact mode, 2000x4000, seed 0", "gpu_seconds": dtx, "pivots": kx,
                      "objective_numerator_bits": zx.numerator.bit_length(), "objective_denominator_bits": zx.denominator.bit_length(),
                      "fp64_relative_error_of_objective": float(abs(Fraction(z) - zx) / abs(zx)), "cpu_oracle_seconds": None}), flush=True)
    Tc = T.copy()
    ob = np.arange(m, 2 * m, dtype=np.int32)
    t0 = time.perf_counter()
    st, ok, _ = C.solve(Tc, ob, want_trace=False)
    tc = time.perf_counter() - t0
    same = (float(z).hex() == float(-Tc[m, 2 * m]).hex()) and basis == ob.tolist() and ok == k
    print(json.dumps({"case": "C2 fp64 2000x4000, seed 0", "gpu_seconds": dt, "cpu_oracle_seconds": tc, "pivots": k,
                      "speedup": tc / dt, "cpu": f"oracle/pivot_oracle.c, {C.num_threads()} threads",
                      "identical": bool(same), "objective": float(z)}), flush=True)


if __name__ == "__main__":
    main()
