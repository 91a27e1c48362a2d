#!/usr/bin/env python3
"""Generate tests/golden/api.json: API-level known answers recorded from the UNMODIFIED reference (`lpsol`).

Run in the build container only (needs /root/reference):

    python tools/make_golden_api.py

tests/api_cases.py holds small programs written against the lpsol class surface (`LinExpr`, `LinCon`,
`LinVar`, `Tableau`, `Simplex`): a setup block and a list of expressions.  Each program is executed against
the reference; for every expression we record `repr(value)` -- or `!ExcType: message` when it raises.
tests/test_api_golden.py replays the same programs against linear_program_solver_b200 and compares the strings:
`host` programs on the CPU (model layer, tableau bookkeeping, JSON and print formats), `device` programs on the
GPU (row operations, pivots, form checks, every Simplex pivot rule and accessor).
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("LPSOL_REFERENCE", "/root/reference")

import api_cases  # noqa: E402


def load_reference(alias: str):
    pkg_dir = os.path.join(REF, "lpsol")
    spec = importlib.util.spec_from_file_location(
        alias, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference("lpsol_ref")
    out = {}
    for kind, cases in (("host", api_cases.HOST), ("device", api_cases.DEVICE)):
        for name, setup, exprs in cases:
            got = api_cases.run_case(ref, setup, exprs, make_tableau=lambda m, n: ref.Tableau(m, n))
            assert name not in out, name
            out[name] = {"kind": kind, "results": got}
    path = os.path.join(ROOT, "tests", "golden", "api.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    n = sum(len(v["results"]) for v in out.values())
    print(f"wrote {path}: {len(out)} programs, {n} recorded answers")


if __name__ == "__main__":
    main()
