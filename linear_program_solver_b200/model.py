"""Model layer of the `lpsol` API: LinExpr, LinCon, LinVar (host-side, exact rationals).

Behaviour-compatible with /root/reference/lpsol/linprog.py:22-381 (same public
names -- including the `constrantEq` spelling --, same `str`/`repr` formats,
same normalisation rules); written from scratch.  These classes carry no
data-parallel work: they describe a program that `LinProg` (linprog.py in this
package) lowers to a tableau for the GPU pivot engine.
"""
from __future__ import annotations

import math
import re
from fractions import Fraction
from typing import Any, Dict, Optional, Tuple

_NAME = re.compile(r"[_A-Za-z][_A-Za-z0-9]*")
_FLIP = {"==": "==", "<=": ">=", ">=": "<="}


def _check_name(x) -> None:
    if not _NAME.fullmatch(x):
        raise ValueError(f"invalid var name: {x!r}")


def _lit(q: Fraction):
    """How a rational is written inside repr(): an int when integral, else the 'p/q' string."""
    return q.numerator if q.denominator == 1 else str(q)


class LinExpr:
    """Sum of coefficient*variable terms plus a constant; zero coefficients are never stored.

    reference: linprog.py:22-222.  `m` maps variable name -> Fraction, `c` is the constant.
    """

    def __init__(self, *args):
        # (c1, x1, c2, x2, ..., [constant])  -- linprog.py:30-52
        self.m: Dict[str, Fraction] = {}
        self.c: Fraction = Fraction(0)
        npairs, tail = divmod(len(args), 2)
        for k in range(npairs):
            coef, name = Fraction(args[2 * k]), args[2 * k + 1]
            if not isinstance(name, str):
                raise TypeError("var name must be str")
            _check_name(name)
            self._accumulate(name, coef)
        if tail:
            self.c = Fraction(args[-1])

    def _accumulate(self, name: str, coef: Fraction) -> None:
        total = self.m.get(name, Fraction(0)) + coef
        if total == 0:
            self.m.pop(name, None)
        else:
            self.m[name] = total

    # -- accessors ---------------------------------------------------------------------------
    def getConstant(self) -> Fraction:
        return self.c

    def getCoefficient(self, x: str) -> Fraction:
        return self.m.get(x, Fraction(0))

    def copy(self) -> "LinExpr":
        dup = LinExpr()
        dup.m = dict(self.m)
        dup.c = self.c
        return dup

    def __eq__(self, other) -> bool:
        # a scalar compares equal to a constant expression (linprog.py:64-68)
        if isinstance(other, LinExpr):
            return self.m == other.m and self.c == other.c
        return self.c == Fraction(other) and not self.m      # the scalar is coerced FIRST: a bad one raises even then

    __hash__ = None

    # -- text ------------------------------------------------------------------------------
    def __str__(self) -> str:
        if not self.m:
            return str(self.c)
        out = []
        for k, name in enumerate(sorted(self.m)):
            coef = self.m[name]
            if coef < 0:
                out.append(("- " if k == 0 else " - ") + f"{-coef}*{name}")
            else:
                out.append(("" if k == 0 else " + ") + f"{coef}*{name}")
        if self.c < 0:
            out.append(f" - {-self.c}")
        elif self.c > 0:
            out.append(f" + {self.c}")
        return "".join(out)

    def __repr__(self) -> str:
        parts = []
        for name in sorted(self.m):
            parts.append(f"{_lit(self.m[name])!r},{name!r}")
        if self.c != 0:
            parts.append(repr(_lit(self.c)))
        return f"{type(self).__name__}({','.join(parts)})"

    # -- arithmetic ----------------------------------------------------------------------------
    def __iadd__(self, other) -> "LinExpr":
        if isinstance(other, LinExpr):
            for name, coef in list(other.m.items()):
                self._accumulate(name, coef)
            self.c += other.c
        else:
            self.c += Fraction(other)
        return self

    def __isub__(self, other) -> "LinExpr":
        if isinstance(other, LinExpr):
            for name, coef in list(other.m.items()):
                self._accumulate(name, -coef)
            self.c -= other.c
        else:
            self.c -= Fraction(other)
        return self

    def __neg__(self) -> "LinExpr":
        out = LinExpr()
        out.m = {name: -coef for name, coef in self.m.items()}
        out.c = -self.c
        return out

    def __pos__(self) -> "LinExpr":
        return self

    def __add__(self, other) -> "LinExpr":
        out = self.copy()
        out += other
        return out

    __radd__ = __add__

    def __sub__(self, other) -> "LinExpr":
        out = self.copy()
        out -= other
        return out

    def __rsub__(self, other) -> "LinExpr":
        return -(self - other)

    # -- constraints ---------------------------------------------------------------------------
    def _constraint(self, other, comp: str) -> "LinCon":
        rhs = other if isinstance(other, LinExpr) else LinExpr(other)
        return LinCon(self, comp, rhs)

    def constrantEq(self, a) -> "LinCon":   # sic: the reference's spelling (linprog.py:180)
        return self._constraint(a, "==")

    def constraintLeq(self, a) -> "LinCon":
        return self._constraint(a, "<=")

    def constraintGeq(self, a) -> "LinCon":
        return self._constraint(a, ">=")

    # -- evaluation ----------------------------------------------------------------------------
    def evaluate(self, vars: Dict[str, Any]) -> Fraction:
        return self.c + sum(coef * Fraction(vars[name]) for name, coef in self.m.items())

    def substitute(self, vars: Dict[str, Any]) -> "LinExpr":
        """Replace variables by expressions or numbers; unknown variables stay (linprog.py:198-222)."""
        out = LinExpr(self.c)
        for name, coef in self.m.items():
            if name not in vars:
                out._accumulate(name, coef)
                continue
            repl = vars[name]
            if isinstance(repl, LinExpr):
                for inner, icoef in repl.m.items():
                    out._accumulate(inner, coef * icoef)
                out.c += coef * repl.c
            else:
                out.c += coef * Fraction(repl)
        return out


class LinCon:
    """Two expressions related by '==', '<=' or '>=' (linprog.py:224-309)."""

    def __init__(self, left, comp: str, right):
        # reference quirk kept (linprog.py:235-238): `right` is wrapped only when `left` was not
        # already a LinExpr; LinExpr.__eq__ accepting scalars is what makes LinCon(e,'==',-4) usable.
        wrap = not isinstance(left, LinExpr)
        self.left = LinExpr(left) if wrap else left
        self.comp = comp
        self.right = LinExpr(right) if wrap else right

    def __eq__(self, other) -> bool:
        return isinstance(other, LinCon) and self.left == other.left and self.comp == other.comp \
            and self.right == other.right

    __hash__ = None

    def copy(self) -> "LinCon":
        return LinCon(self.left.copy(), self.comp, self.right.copy())

    def __str__(self) -> str:
        return f"{self.left} {self.comp} {self.right}"

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self.left!r},{self.comp!r},{self.right!r})"

    def reverse(self) -> "LinCon":
        return LinCon(self.right, _FLIP[self.comp], self.left)

    def simplify(self) -> "LinCon":
        """Variables to the left, constant to the right (linprog.py:261-268)."""
        if self.left.c == 0 and not self.right.m:
            return self
        diff = self.left - self.right
        const, diff.c = diff.c, Fraction(0)
        return LinCon(diff, self.comp, LinExpr(-const))

    def evaluate(self, vars: Dict[str, Any]) -> bool:
        lhs, rhs = self.left.evaluate(vars), self.right.evaluate(vars)
        if self.comp == "<=":
            return lhs <= rhs
        if self.comp == ">=":
            return lhs >= rhs
        return lhs == rhs

    def addLeft(self, a):
        self.left += a

    def addRight(self, a):
        self.right += a

    def subLeft(self, a):
        self.left -= a

    def subRight(self, a):
        self.right -= a


class LinVar:
    """A variable with optional bounds; integer variables keep integer bounds (linprog.py:311-381)."""

    def __init__(self, x: str, integral: bool = False, lb: Any = None, ub: Any = None):
        _check_name(x)
        self.x = x
        self.isint = integral
        self.lb: Optional[Fraction] = None
        self.ub: Optional[Fraction] = None
        if lb is not None:
            self.lb = self._round_lb(Fraction(lb))
        if ub is not None:
            self.ub = self._round_ub(Fraction(ub))

    def _round_lb(self, v: Fraction) -> Fraction:
        return Fraction(math.ceil(v)) if self.isint else v

    def _round_ub(self, v: Fraction) -> Fraction:
        return Fraction(math.floor(v)) if self.isint else v

    def copy(self) -> "LinVar":
        return LinVar(self.x, self.isint, self.lb, self.ub)

    def getBounds(self) -> Tuple[Optional[Fraction], Optional[Fraction]]:
        return (self.lb, self.ub)

    def boundAbove(self, ub: Any):
        ub = Fraction(ub)
        if self.ub is None or ub < self.ub:
            self.ub = self._round_ub(ub)

    def boundBelow(self, lb: Any):
        lb = Fraction(lb)
        if self.lb is None or lb > self.lb:
            self.lb = self._round_lb(lb)

    def isFeasible(self) -> bool:
        return self.lb is None or self.ub is None or self.lb <= self.ub

    def __str__(self) -> str:
        lo = "-inf" if self.lb is None else self.lb
        hi = "+inf" if self.ub is None else self.ub
        return f"{self.x}@{'Z' if self.isint else 'R'}[{lo},{hi}]"

    def __repr__(self) -> str:
        lo = None if self.lb is None else _lit(self.lb)
        hi = None if self.ub is None else _lit(self.ub)
        return f"{type(self).__name__}({self.x!r},{self.isint!r},{lo!r},{hi!r})"

    def __eq__(self, other) -> bool:
        return isinstance(other, LinVar) and (self.x, self.isint, self.lb, self.ub) == \
            (other.x, other.isint, other.lb, other.ub)

    __hash__ = None
