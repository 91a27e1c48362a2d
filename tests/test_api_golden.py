"""API-level parity: programs of tests/api_cases.py replayed against this package, compared with the answers the
UNMODIFIED reference gave (tests/golden/api.json, written by tools/make_golden_api.py).

`host` programs (model layer of linprog.py:22-381, tableau bookkeeping / JSON / print formats of
tableau.py:70-250,312-462) run on the CPU; `device` programs (row operations and pivots tableau.py:254-308, form
checks :466-521, every Simplex rule and accessor simplex.py:24-379) run on the GPU, in exact mode.
Error answers compare the exception type and message, except for IndexError / TypeError, where the reference's
message is CPython's own ("list index out of range", ...) and only the type is part of the contract.
"""
import json
import os

import pytest

import api_cases

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "api.json")) as f:
    GOLDEN = json.load(f)


def _same(got: str, want: str) -> bool:
    if got == want:
        return True
    for exc in ("!IndexError", "!TypeError"):
        if got.startswith(exc) and want.startswith(exc):
            return True
    return False


def _replay(name, setup, exprs):
    import linear_program_solver_b200 as pkg
    got = api_cases.run_case(pkg, setup, exprs, make_tableau=lambda m, n: pkg.Tableau(m, n))
    want = GOLDEN[name]["results"]
    assert len(got) == len(want), (got, want)
    bad = [(src, g, w) for src, g, w in zip(exprs if len(got) == len(exprs) else ["setup"], got, want) if not _same(g, w)]
    assert not bad, "\n".join(f"{src}\n   got  {g}\n   want {w}" for src, g, w in bad)


def test_every_program_has_a_recorded_answer():
    names = [c[0] for c in api_cases.HOST + api_cases.DEVICE]
    assert sorted(names) == sorted(GOLDEN.keys())
    assert all(not r.startswith("!setup") for v in GOLDEN.values() for r in v["results"])


@pytest.mark.parametrize("case", api_cases.HOST, ids=[c[0] for c in api_cases.HOST])
def test_host_programs(case):
    _replay(*case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", api_cases.DEVICE, ids=[c[0] for c in api_cases.DEVICE])
def test_device_programs(case):
    _replay(*case)
