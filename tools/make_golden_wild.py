#!/usr/bin/env python3
"""Generate tests/golden/wild.json: small irregular LPs solved by the UNMODIFIED reference (`lpsol`).

    python tools/make_golden_wild.py            # build container only (needs /root/reference)

The seeded families of tools/make_golden.py are regular (dense positive data, or equality rows with a known feasible
point).  This corpus is the opposite: mixed signs, zeros, rational data, negative and zero right-hand sides (the sign
normalisation of simplex.py:43-45), partial slack bases, duplicated and scaled rows (the row deletion of
simplex.py:86-94 and its IndexError), infeasible and unbounded problems, degenerate ties.  The inputs are stored with the
answers (they come from `random.Random`, but a fixture that carries its data does not depend on the generator).
Every record holds what `tools/make_golden.run` records: constructor error or initial basis, phase-1 and phase-2 pivot
traces with the rule of each pivot, final basis / objective / basic solution / marks / size, or the solve error.
Integer-data records also carry the float-rebound reference's answers (`float`), for the fp64 mode.
"""
from __future__ import annotations

import json
import os
import random
import sys
from fractions import Fraction

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import make_golden as MG  # noqa: E402
from linear_program_solver_b200 import problems  # noqa: E402

PIVOT_CAP = 400     # the reference's Bland variant (lowest ROW among ties) has no termination proof


def enc(x):
    return str(Fraction(x))


def main():
    ref, ref_f = MG.ref, MG.ref_f         # the exact reference and its float-rebound second copy
    out = []
    stats = {}
    for k in range(int(os.environ.get("WILD_COUNT", "400"))):
        rng = random.Random(0xB200 + k)
        c, A, b = problems.irregular_lp(rng)
        rec = MG.run(ref, c, A, b, max_pivots=PIVOT_CAP)
        rec["id"] = k
        rec["input"] = {"c": [enc(v) for v in c], "A": [[enc(v) for v in r] for r in A], "b": [enc(v) for v in b]}
        if all(v.denominator == 1 for row in A + [b, c] for v in row):
            cf, Af, bf = [float(v) for v in c], [[float(v) for v in r] for r in A], [float(v) for v in b]
            rec["float"] = MG.run(ref_f, cf, Af, bf, max_pivots=PIVOT_CAP)
        kind = ("ctor:" + rec["ctor_error"][0]) if "ctor_error" in rec else \
            ("solve_error" if "solve_error" in rec else ("truncated" if rec.get("truncated") else "optimal"))
        if kind == "optimal" and rec["phase1_pivots"]:
            kind = "optimal after phase 1"
        stats[kind] = stats.get(kind, 0) + 1
        out.append(rec)
    path = os.path.join(ROOT, "tests", "golden", "wild.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
        f.write("\n")
    print("wrote", path, len(out), "records", stats, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
