"""Differential fuzz of the host-side class surface against the UNMODIFIED reference, live (build container only: it
imports /root/reference, which the GPU box does not have -- skipped there; the recorded fixtures of test_api_golden.py,
test_wild.py and test_walks.py are what travels)."""
import os
import random
import sys

import pytest

import api_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference/lpsol"

pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree is not present on this machine")


@pytest.mark.parametrize("kind", ["model", "tableau"])
def test_random_host_programs_agree_with_the_reference(kind):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.dont_write_bytecode = True
    import fuzz_host as F
    import linear_program_solver_b200 as pkg
    ref = F.load_reference()
    gen = F.model_program if kind == "model" else F.tableau_program
    bad = []
    for k in range(200):
        setup, exprs = gen(random.Random(424242 + k))
        want = api_cases.run_case(ref, setup, exprs, make_tableau=lambda m, n: ref.Tableau(m, n))
        got = api_cases.run_case(pkg, setup, exprs, make_tableau=lambda m, n: pkg.Tableau(m, n))
        assert len(got) == len(want), (setup, got[:1], want[:1])
        for src, g, w in zip(exprs, got, want):
            if g != w and not any(g.startswith(x) and w.startswith(x) for x in ("!IndexError", "!TypeError")):
                bad.append((k, setup, src, g[:200], w[:200]))
                break
    assert not bad, "\n".join(map(str, bad[:5]))
