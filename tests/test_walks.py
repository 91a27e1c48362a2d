"""Random walks over the Simplex / Tableau device surface (tests/golden/walks.json, answered by the UNMODIFIED reference
through tools/make_golden_walks.py) replayed on the CUDA path: every pivot rule asked and applied, checked pivots
(accepted and refused), form predicates, row operations between pivots, snapshots of basis / solution / objective /
marks / whole tableau after every change, and the closing solve().  Exact mode compares the recorded strings; fp64
mode compares with the float-rebound reference's strings, a zero's sign aside.  IndexError / TypeError /
ZeroDivisionError messages are CPython's own (they differ between Fraction and float division), so only the type counts."""
import re

import pytest

import api_cases
from conftest import load_golden

WALKS = load_golden("walks")
TYPE_ONLY = ("!IndexError", "!TypeError", "!ZeroDivisionError")


def _same(got: str, want: str) -> bool:
    if got == want:
        return True
    return any(got.startswith(x) and want.startswith(x) for x in TYPE_ONLY)


def _unsigned_zero(s: str) -> str:
    return re.sub(r"-0\.0(?![0-9e])", "0.0", s)


def test_corpus_shape():
    assert len(WALKS) >= 100 and sum(len(w["exprs"]) for w in WALKS) >= 2000
    assert sum("float_results" in w for w in WALKS) >= 30
    text = " ".join(r for w in WALKS for r in w["results"])
    for needle in ("'optimal'", "'unbounded'", "!ValueError: bad pivot by min ratio test", "!ZeroDivisionError",
                   "!AssertionError: unbounded artificial problem"):
        assert needle in text, needle


def _replay(chunk, mode):
    import linear_program_solver_b200 as pkg
    bad = []
    walks = [w for w in WALKS if mode == "exact" or "float_results" in w][chunk::4]
    for w in walks:
        setup = w["setup"] if mode == "exact" else w["float_setup"]
        want = w["results"] if mode == "exact" else w["float_results"]
        got = api_cases.run_case(pkg, setup, w["exprs"], make_tableau=lambda m, n: pkg.Tableau(m, n, mode=mode))
        if len(got) != len(want):
            bad.append((w["id"], "setup", got[:1], want[:1]))
            continue
        for src, g, x in zip(w["exprs"], got, want):
            if mode == "fp64":
                g, x = _unsigned_zero(g), _unsigned_zero(x)
            if not _same(g, x):
                bad.append((w["id"], src, g[:400], x[:400]))
                break
    assert not bad, "\n".join(map(str, bad))
    assert walks


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", range(4))
def test_exact_mode_walks_match_reference(chunk):
    _replay(chunk, "exact")


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", range(4))
def test_fp64_mode_walks_match_float_rebound_reference(chunk):
    _replay(chunk, "fp64")
