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


def test_integer_host_copy_takes_integer_edits_in_place():
    """`Tableau.from_grid` with integers keeps an int64 array as the host copy; integer setters, `addVars` and integer
    row operations edit it in place (what phase 1 does to a large integer tableau), a rational edit falls back to
    Fractions.  Same tableau as the entry-by-entry path either way."""
    import numpy as np
    from fractions import Fraction
    from linear_program_solver_b200 import Tableau
    rng = np.random.default_rng(3)
    grid = rng.integers(-9, 10, size=(5, 8))
    a = Tableau.from_grid(grid, mode="exact")
    b = Tableau.from_grid([[Fraction(int(x)) for x in row] for row in grid], mode="exact")
    assert a._int_grid is not None and b._int_grid is None

    def both(f):
        f(a)
        f(b)

    both(lambda t: t.rowMult(1, -1))
    both(lambda t: t.setC([0] * 7))
    both(lambda t: t.setZ(0))
    both(lambda t: t.addVars(["$a0", "$a2"]))
    both(lambda t: (t.setCj(7, 1), t.setAij(0, 7, 1), t.rowSubFromObj(0, 1)))
    both(lambda t: (t.setCj(8, 1), t.setAij(2, 8, 1), t.rowSubFromObj(2, 1)))
    both(lambda t: t.rowAdd(3, 1, 4))
    both(lambda t: t.rowSub(3, 2, 2))
    assert a._int_grid is not None and a._grid is None or not a._host_ok      # still the integer array
    assert a.getTableauSize() == (4, 9) and a.getVarNames()[-2:] == ["$a0", "$a2"]
    assert a.getC() == b.getC() and a.getB() == b.getB() and a.getZ() == b.getZ()
    assert a._int_grid is not None
    both(lambda t: t.rowMult(2, Fraction(1, 3)))                   # rational: the Fraction path takes over
    assert a._int_grid is None
    both(lambda t: t.rowAdd(0, 2, 10 ** 30))                       # and so does an integer too large for the array
    assert a == b
