#!/usr/bin/env python3
"""tests/golden/rules.json: the whole-tableau pivot rules of the UNMODIFIED reference (`Simplex.findPivotMaxIncrease`,
`Simplex.findPivotAll`, simplex.py:286-360) at every state along solves that alternate max-increase and Dantzig
pivots.  Build container only (needs /root/reference):

    python tools/make_golden_rules.py

Exact runs, and float runs with the module globals `Frac/ZERO/ONE` rebound to float (as tools/make_golden.py).
Small integer data on purpose: many equal ratios, so the tie rules of :308, :316-318 and :349-351 are exercised.
"""
from __future__ import annotations

import json
import os
import sys

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from make_golden import build, ref, ref_f  # noqa: E402  (loads the reference twice: exact and float-rebound)
from linear_program_solver_b200 import problems  # noqa: E402


def walk(pkg, c, A, b, limit):
    s = pkg.Simplex(build(pkg, c, A, b))
    states = []
    for step in range(limit):
        mx = s.findPivotMaxIncrease()
        states.append({"max_increase": list(mx) if isinstance(mx, tuple) else mx,
                       "all": [list(p) for p in s.findPivotAll()]})
        if isinstance(mx, str):
            break
        nxt = s.findPivotMaxIncrease(True) if step % 2 else s.findPivotStandard(True)
        assert not isinstance(nxt, str)
    return states


def main():
    out = []
    for (m, seed, hi) in [(6, 0, 3), (8, 1, 9), (12, 5, 9), (16, 2, 3), (24, 3, 2), (40, 1, 3)]:
        c, A, b = problems.dense_lp(m, seed, 1, hi)
        rec = {"family": "dense", "m": m, "seed": seed, "lo": 1, "hi": hi,
               "exact": walk(ref, c, A, b, 4 * m),
               "float": walk(ref_f, [float(x) for x in c], [[float(x) for x in r] for r in A], [float(x) for x in b], 4 * m)}
        out.append(rec)
        print("dense", m, seed, hi, len(rec["exact"]), len(rec["float"]), rec["exact"][-1]["max_increase"], flush=True)
    for (m, n, seed) in [(5, 9, 0), (8, 14, 1), (12, 20, 2)]:
        c, A, b = problems.equality_lp(m, n, seed)
        rec = {"family": "equality", "m": m, "n": n, "seed": seed, "exact": walk(ref, c, A, b, 4 * n)}
        out.append(rec)
        print("equality", m, n, seed, len(rec["exact"]), rec["exact"][-1]["max_increase"], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "rules.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
        f.write("\n")


if __name__ == "__main__":
    main()
