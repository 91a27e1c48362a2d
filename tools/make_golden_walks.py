#!/usr/bin/env python3
"""Generate tests/golden/walks.json: random walks over the Simplex / Tableau device surface, answered by the
UNMODIFIED reference (`lpsol`).

    python tools/make_golden_walks.py           # build container only (needs /root/reference)

Start: a feasible LP of the irregular corpus (tools/make_golden_wild.py), `s = Simplex(t)`.  Then 8-16 random steps from
  * the four pivot rules, asked (`findPivotStandard()`, ...) or applied (`findPivotMaxIncrease(True)`, ...)
    -- simplex.py:218-360;
  * the checked pivot `s.pivot(r, c)` on a random cell (mostly refused: ValueError / ZeroDivisionError) or on an entry
    of `findPivotAll()` -- simplex.py:199-216;
  * the form predicates `isOptimal / isUnbounded / isInfeasible / isDegenerate / isCanonical(bcols)` --
    tableau.py:466-521;
  * row operations that keep b >= 0: `rowMult(r, k)` with k > 0 and `rowAddToObj / rowSubFromObj(r, k)` --
    tableau.py:254-293 (a basis column then stops being a unit column or gets a cost; the rules do not care);
  * a snapshot of basis, basic solution, objective, marks, and now and then the whole tableau;
and finally `s.solve()`.  Programs are source text in the format of tests/api_cases.py (`run_case`); every expression's
repr -- or exception type and message -- is recorded.  Integer-data walks are also answered by the float-rebound second
copy of the reference (`float_results`), for the fp64 mode.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

import api_cases  # noqa: E402
import make_golden as MG  # noqa: E402

STATE = "(list(s.getBasicSequence()), dict(s.getBFS()), s.getObjValue(), list(t.getVarMarks()))"
WHOLE = "(t.getTableauSize(), t.getZ(), snap(t.getC()), snap(t.getB()), snap(t.getA()))"
RULES = ["findPivotStandard", "findPivotMinIndex", "findPivotMaxIncrease"]


def lit(v: Fraction, as_float: bool) -> str:
    if as_float:
        return repr(float(v))
    return str(v.numerator) if v.denominator == 1 else repr(str(v))


def setup_source(rec, as_float: bool) -> str:
    i = rec["input"]
    row = lambda r: "[" + ", ".join(lit(Fraction(v), as_float) for v in r) + "]"  # noqa: E731
    return (f"t = mk({row(i['c'])}, [{', '.join(row(r) for r in i['A'])}], {row(i['b'])})\n"
            f"s = Simplex(t)\n")


def walk(rng: random.Random, m: int, n: int, integer_only: bool):
    exprs = [STATE]
    for _ in range(rng.randint(8, 16)):
        k = rng.randrange(12)
        if k <= 1:
            exprs.append(f"s.{rng.choice(RULES)}()")
        elif k == 2:
            exprs.append("s.findPivotAll()")
        elif k <= 4:
            exprs.append(f"s.{rng.choice(RULES)}(True)")
            exprs.append(STATE)
        elif k == 5:
            exprs.append(f"s.pivot({rng.randrange(m)}, {rng.randrange(n)})")
            exprs.append(STATE)
        elif k == 6:
            exprs.append(f"[s.pivot(*p) for p in s.findPivotAll()[{rng.randrange(3)}:][:1]]")
            exprs.append(STATE)
        elif k == 7:
            exprs.append("(t.isOptimal(), t.isUnbounded(), t.isInfeasible(), t.isDegenerate())")
        elif k == 8:
            exprs.append(f"(lambda bc: (t.isCanonical(bc), bc))([-7] * {m})")
        elif k == 9:
            mult = Fraction(rng.randint(1, 4), 1 if integer_only else rng.choice([1, 1, 2, 3]))
            exprs.append(f"t.rowMult({rng.randrange(m)}, {lit(mult, False)})")
            exprs.append(STATE)
        elif k == 10:
            mult = Fraction(rng.randint(-3, 3), 1 if integer_only else rng.choice([1, 1, 2]))
            exprs.append(f"t.{rng.choice(['rowAddToObj', 'rowSubFromObj'])}({rng.randrange(m)}, {lit(mult, False)})")
            exprs.append("(t.getZ(), snap(t.getC()))")
        elif rng.random() < 0.3:
            exprs.append(WHOLE)
        else:
            i, j = rng.randrange(m), rng.randrange(n)
            exprs.append(f"(t.getZ(), snap(t.getB()), t.getCj({j}), t.getAij({i}, {j}), t.getBi({i}))")
    exprs += ["s.solve()", STATE, "t.isOptimal()", WHOLE]
    return exprs


def main():
    with open(os.path.join(ROOT, "tests", "golden", "wild.json")) as f:
        wild = json.load(f)
    out = []
    errors = {}
    for rec in wild:
        if "ctor_error" in rec or len(out) >= int(os.environ.get("WALK_COUNT", "120")):
            continue
        m, n = rec["m_n"]
        rng = random.Random(0xA11 + rec["id"])
        integer_only = "float" in rec and "ctor_error" not in rec["float"]
        exprs = walk(rng, m, n, integer_only)
        mk = lambda pkg: (lambda mm, nn: pkg.Tableau(mm, nn))  # noqa: E731
        item = {"id": rec["id"], "setup": setup_source(rec, False), "exprs": exprs,
                "results": api_cases.run_case(MG.ref, setup_source(rec, False), exprs, make_tableau=mk(MG.ref))}
        if integer_only:
            item["float_setup"] = setup_source(rec, True)
            item["float_results"] = api_cases.run_case(MG.ref_f, item["float_setup"], exprs, make_tableau=mk(MG.ref_f))
        for r in item["results"]:
            if r.startswith("!"):
                errors[r.split(":")[0]] = errors.get(r.split(":")[0], 0) + 1
        out.append(item)
    path = os.path.join(ROOT, "tests", "golden", "walks.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
        f.write("\n")
    print("wrote", path, len(out), "walks,", sum(len(w["exprs"]) for w in out), "steps,",
          sum("float_results" in w for w in out), "with float answers; errors recorded:", errors,
          os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
