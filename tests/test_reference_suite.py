"""The reference's OWN unit tests (lpsol/test_*.py) run against this package, in place, without copying them: the
test modules are loaded from /root/reference under a shim package whose `Tableau`, `Simplex`, `LinExpr`, `LinCon`,
`LinVar`, `LinProg` are ours, so their `from . import Tableau` binds to the drop-in.  Build container only (the tree
does not exist on the GPU box; skipped there).  Tests that reach the device (`test_pivot`, tableau.py:295-308) need
a GPU: without one they must fail loudly with `NativeUnavailable` -- there is no CPU fallback -- and the same
known-answer pivot sequence is checked on the GPU through tests/golden/textbook.json."""
import importlib.util
import io
import os
import sys
import types
import unittest

import pytest

REFERENCE = "/root/reference/lpsol"
MODULES = ["test_tableau", "test_linexpr", "test_lincon", "test_linvar", "test_simplex", "test_linprog"]

pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree is not present on this machine")


@pytest.mark.parametrize("name", MODULES)
def test_reference_unit_tests_pass_against_the_drop_in(name):
    import torch

    import linear_program_solver_b200 as pkg
    sys.dont_write_bytecode = True
    shim = types.ModuleType("lpsol_dropin")
    shim.__path__ = []
    for k in ("Tableau", "Simplex", "LinExpr", "LinCon", "LinVar", "LinProg"):
        setattr(shim, k, getattr(pkg, k))
    sys.modules["lpsol_dropin"] = shim
    spec = importlib.util.spec_from_file_location(f"lpsol_dropin.{name}", os.path.join(REFERENCE, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    try:
        spec.loader.exec_module(mod)
        result = unittest.TextTestRunner(stream=io.StringIO(), verbosity=0).run(
            unittest.defaultTestLoader.loadTestsFromModule(mod))
    finally:
        sys.modules.pop(spec.name, None)
        sys.modules.pop("lpsol_dropin", None)
    assert result.testsRun > 0
    problems = result.failures + result.errors
    if not torch.cuda.is_available():
        needs_device = [(t, tb) for t, tb in problems if "NativeUnavailable" in tb]
        problems = [p for p in problems if p not in needs_device]
        assert [t.id().rsplit(".", 1)[-1] for t, _ in needs_device] == (["test_pivot"] if name == "test_tableau" else [])
    assert not problems, "\n".join(f"{t.id()}: {tb.strip().splitlines()[-1]}" for t, tb in problems)
