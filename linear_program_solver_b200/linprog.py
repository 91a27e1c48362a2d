"""`LinProg`: a linear (integer) program in general form, lowered to a tableau for the GPU engine.

The reference's `LinProg` is an empty stub (/root/reference/lpsol/linprog.py:383-393; SURVEY.md
F1), so everything below the constructor is new design built from the reference's own
primitives (`LinExpr.substitute`, `LinCon.simplify`, `LinVar.boundAbove/boundBelow`).
"""
from __future__ import annotations

from .model import LinCon, LinExpr, LinVar  # noqa: F401  (re-exported like the reference module)


class LinProg:
    """representation of a linear program in possibly non standard form (linprog.py:383-393)"""

    def __init__(self):
        pass
