#!/usr/bin/env python3
"""Generate tests/golden/*.json by running the UNMODIFIED reference (`lpsol`).

Run in the build container only (it needs /root/reference, which does not exist
on the GPU box):

    python tools/make_golden.py

The reference is imported read-only under the alias `lpsol_ref` (no bytecode is
written into its tree).  For every case we record the inputs' recipe (family,
size, seed) and the reference's outputs: initial basis, pivot trace with the
rule that chose each pivot, final basis, objective and basic solution as exact
"p/q" strings, or as C99 hex floats for the float-rebound runs.

Float-rebound runs: the reference has no float path,
but its control flow runs unmodified once the module globals `Frac/ZERO/ONE`
of `lpsol.tableau` and `lpsol.simplex` are rebound to float.  We do that on a
second, separately loaded copy of the package so the exact one stays pristine.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
from fractions import Fraction

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("LPSOL_REFERENCE", "/root/reference")

from linear_program_solver_b200 import problems  # noqa: E402


def load_reference(alias: str):
    pkg_dir = os.path.join(REF, "lpsol")
    spec = importlib.util.spec_from_file_location(
        alias, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


ref = load_reference("lpsol_ref")
ref_f = load_reference("lpsol_ref_float")
for name in ("tableau", "simplex"):
    sub = sys.modules[f"lpsol_ref_float.{name}"]
    sub.Frac, sub.ZERO, sub.ONE = float, 0.0, 1.0


def build(pkg, c, A, b, names=None):
    m, n = len(A), len(c)
    t = pkg.Tableau(m, n)
    t.setC(list(c))
    t.setB(list(b))
    t.setA([list(r) for r in A])
    if names:
        t.setVarNames(names)
    return t


def fmt(x):
    return x.hex() if isinstance(x, float) else str(Fraction(x))


def run(pkg, c, A, b, max_pivots=None):
    """Reference `Simplex(tab); solve()` with a tracing wrapper on `_pivot`."""
    tab = build(pkg, c, A, b)
    trace = []
    sx_cls = pkg.Simplex
    rule_now = {"rule": 2}
    orig_pivot = sx_cls._pivot
    orig_std, orig_min = sx_cls.findPivotStandard, sx_cls.findPivotMinIndex

    class Stop(Exception):
        pass

    def traced_pivot(self, r, c_):
        if max_pivots is not None and len(trace) >= max_pivots:
            raise Stop()
        orig_pivot(self, r, c_)
        trace.append([r, c_, rule_now["rule"]])

    def std(self, do_pivot=False):
        rule_now["rule"] = 0
        try:
            return orig_std(self, do_pivot)
        finally:
            rule_now["rule"] = 2

    def mn(self, do_pivot=False):
        rule_now["rule"] = 1
        try:
            return orig_min(self, do_pivot)
        finally:
            rule_now["rule"] = 2

    sx_cls._pivot, sx_cls.findPivotStandard, sx_cls.findPivotMinIndex = traced_pivot, std, mn
    out = {}
    try:
        try:
            sx = sx_cls(tab)
        except Exception as e:  # infeasible / redundant rows (F5b)
            out["ctor_error"] = [type(e).__name__, str(e)]
            out["pivots"] = trace
            return out
        out["initial_basis"] = list(sx.getBasicSequence())
        n_phase1 = len(trace)
        try:
            sx.solve()
        except Stop:
            out["truncated"] = True
        except AssertionError as e:
            out["solve_error"] = ["AssertionError", str(e)]
        out["phase1_pivots"] = trace[:n_phase1]
        out["pivots"] = trace[n_phase1:]
        out["basis"] = list(sx.getBasicSequence())
        out["objective"] = fmt(sx.getObjValue())
        out["bfs"] = {str(j): fmt(v) for j, v in sx.getBFS().items()}
        out["marks"] = [bool(x) for x in tab.getVarMarks()]
        out["m_n"] = list(tab.getTableauSize())
    finally:
        sx_cls._pivot, sx_cls.findPivotStandard, sx_cls.findPivotMinIndex = orig_pivot, orig_std, orig_min
    return out


def lp_via_reference(c, A, b):
    """Node-LP solver for the B&B oracle: reference `Simplex(tab); solve()`, exceptions mapped to outcomes."""
    tab = build(ref, c, A, b)
    try:
        sx = ref.Simplex(tab)
    except ValueError as e:
        if str(e).startswith("infeasible problem"):
            return ("infeasible",)
        raise
    try:
        sx.solve()
    except AssertionError as e:
        if "unbounded" in str(e):
            return ("unbounded",)
        raise
    return ("optimal", sx.getObjValue(), dict(sx.getBFS()))


def make_milp():
    import random
    from oracle import bnb_oracle as B
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_bnb_oracle import random_ip
    out = []
    cases = []
    for seed in range(14):
        rng = random.Random(500 + seed)
        cases.append(random_ip(rng, rng.randint(3, 6), rng.randint(2, 5), pure=(seed % 2 == 0)))
    # a knapsack with a known optimum and an infeasible program
    cases.append(([("a", True, 0, 1), ("b", True, 0, 1), ("c", True, 0, 1), ("d", True, 0, 1)],
                  [({"a": 5, "b": 7, "c": 4, "d": 3}, "<=", 14)], ({"a": 8, "b": 11, "c": 6, "d": 4}, 0, "max")))
    cases.append(([("x", True, 0, 5), ("y", True, 0, 5)],
                  [({"x": 2, "y": 2}, "==", 3)], ({"x": 1, "y": 1}, 0, "min")))
    for variables, cons, obj in cases:
        got = B.branch_and_bound(variables, cons, obj, solve_lp=lp_via_reference)
        out.append({
            "variables": [[n, i, None if lb is None else str(Fraction(lb)), None if ub is None else str(Fraction(ub))]
                          for n, i, lb, ub in variables],
            "constraints": [[{k: str(Fraction(v)) for k, v in c.items()}, comp, str(Fraction(rhs))] for c, comp, rhs in cons],
            "objective": [{k: str(Fraction(v)) for k, v in obj[0].items()}, str(Fraction(obj[1])), obj[2]],
            "status": got["status"],
            "objective_value": None if got["objective"] is None else str(got["objective"]),
            "values": {k: str(v) for k, v in got["values"].items()},
            "nodes": got["nodes"], "lps": got["lps"],
            "trace": [[d, s_, None if o is None else str(o)] for d, s_, o in got["trace"]],
        })
        print("milp", got["status"], got["objective"], got["nodes"], flush=True)
    return out


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    # 1. the reference's own known-answer test on the path (test_tableau.py:220-227)
    c, A, b = problems.TEXTBOOK
    t = build(ref, c, A, b)
    steps = []
    for (r, cc) in [(1, 0), (0, 1)]:
        t.pivot(r, cc)
        steps.append({"pivot": [r, cc], "z_cell": str(t._z), "c": [str(x) for x in t.getC()],
                      "b": [str(x) for x in t.getB()], "a": [[str(x) for x in row] for row in t.getA()]})
    textbook = {"source": "lpsol/test_tableau.py:9-29,220-227", "c": c, "A": A, "b": b, "steps": steps,
                "simplex": run(ref, c, A, b)}

    # 2. dense seeded family, exact
    dense = []
    for m, seeds in [(4, range(6)), (8, range(6)), (16, range(4)), (32, range(2)), (50, range(10)), (64, range(4)), (100, [0])]:
        for seed in seeds:
            c, A, b = problems.dense_lp(m, seed)
            rec = run(ref, c, A, b)
            rec.update({"family": "dense", "m": m, "seed": seed, "lo": 1, "hi": 9})
            dense.append(rec)
            print("dense exact", m, seed, len(rec["pivots"]), rec["objective"], flush=True)
    # 8-bit coefficients (config 3 family), small sizes + a pivot prefix at a larger one
    for m, seeds, mp in [(24, range(3), None), (60, [0], None), (200, [0], 12), (1000, [0], 6)]:
        for seed in seeds:
            c, A, b = problems.dense_lp(m, seed, 1, 127)
            rec = run(ref, c, A, b, max_pivots=mp)
            rec.update({"family": "dense", "m": m, "seed": seed, "lo": 1, "hi": 127, "max_pivots": mp})
            dense.append(rec)
            print("dense8 exact", m, seed, len(rec["pivots"]), rec["objective"], flush=True)

    # 3. same family through the float-rebound reference
    dense_f = []
    for m, seeds in [(8, range(4)), (50, range(4)), (100, range(2)), (200, [0])]:
        for seed in seeds:
            c, A, b = problems.dense_lp(m, seed)
            rec = run(ref_f, c, A, b)
            rec.update({"family": "dense", "m": m, "seed": seed, "lo": 1, "hi": 9})
            dense_f.append(rec)
            print("dense float", m, seed, len(rec["pivots"]), rec["objective"], flush=True)

    # 4. special cases
    special = {}
    c, A, b = problems.BEALE
    special["beale"] = run(ref, c, A, b)
    special["beale"]["input"] = {"c": [str(x) for x in c], "A": [[str(x) for x in r] for r in A], "b": [str(x) for x in b]}
    special["beale_float"] = run(ref_f, [float(x) for x in c], [[float(x) for x in r] for r in A], [float(x) for x in b])
    special["redundant_row"] = run(ref, [1, 1, 1], [[1, 1, 0], [2, 2, 0], [0, 1, 1]], [2, 4, 3])
    special["infeasible"] = run(ref, [0, 0], [[1, 1], [1, 1]], [1, 2])
    special["unbounded"] = run(ref, [-1, 0], [[-1, 1]], [1])
    tu = build(ref, [-1, 0], [[-1, 1]], [1])
    su = ref.Simplex(tu)
    special["unbounded"]["findPivotStandard"] = su.findPivotStandard()
    special["unbounded"]["findPivotMinIndex"] = su.findPivotMinIndex()

    # 5. phase-1 family (equality rows, no slack basis)
    phase1 = []
    for (m, n) in [(3, 6), (5, 9), (8, 14), (12, 20), (20, 36)]:
        for seed in range(4):
            c, A, b = problems.equality_lp(m, n, seed)
            rec = run(ref, c, A, b)
            rec.update({"family": "equality", "m": m, "n": n, "seed": seed})
            phase1.append(rec)
            print("phase1", m, n, seed, rec.get("ctor_error"), len(rec.get("phase1_pivots", [])), rec.get("objective"), flush=True)

    # 6. MILP layer: the branch-and-bound driver of oracle/bnb_oracle.py with every node LP solved by the
    #    UNMODIFIED reference Simplex (the reference itself has no MILP solver: its LinProg is an empty stub)
    milp = make_milp()

    only = os.environ.get("GOLDEN_ONLY")
    for name, obj in [("textbook", textbook), ("dense_exact", dense), ("dense_float", dense_f),
                      ("special", special), ("phase1", phase1), ("milp", milp)]:
        if only and name not in only.split(","):
            continue
        with open(os.path.join(out_dir, name + ".json"), "w") as f:
            json.dump(obj, f, separators=(",", ":"))
            f.write("\n")
    print("wrote", out_dir)


if __name__ == "__main__":
    main()
