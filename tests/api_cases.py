"""Programs against the lpsol class surface, shared by tools/make_golden_api.py (which records the reference's
answers into tests/golden/api.json) and tests/test_api_golden.py (which replays them against this package).

A program is (name, setup source, [expression source, ...]).  Names available to a program:
  LinExpr, LinCon, LinVar, Simplex, Fraction / F, and
  mk(c, A, b, z=0, names=None, marks=None) -> Tableau built with the public setters
  T(m, n)                                   -> empty Tableau
  snap(x)                                   -> deep copy (the reference's getters hand out live lists, ours snapshots)
Every expression is recorded as repr(value), or "!ExcType: message" when it raises.
"""
from __future__ import annotations

import copy
from fractions import Fraction


def run_case(pkg, setup: str, exprs, make_tableau):
    def mk(c, A, b, z=0, names=None, marks=None):
        t = make_tableau(len(A), len(c))
        t.setC(list(c))
        t.setB(list(b))
        t.setA([list(r) for r in A])
        t.setZ(z)
        if names is not None:
            t.setVarNames(list(names))
        if marks is not None:
            t.setVarMarks(list(marks))
        return t

    ns = {"LinExpr": pkg.LinExpr, "LinCon": pkg.LinCon, "LinVar": pkg.LinVar, "Simplex": pkg.Simplex,
          "Fraction": Fraction, "F": Fraction, "mk": mk, "T": make_tableau, "snap": copy.deepcopy}
    out = []
    try:
        exec(setup, ns)
    except Exception as e:  # a failing setup is itself the recorded answer
        return [f"!setup {type(e).__name__}: {e}"]
    for src in exprs:
        try:
            out.append(repr(eval(src, ns)))
        except Exception as e:
            out.append(f"!{type(e).__name__}: {e}")
    return out


# -------------------------------------------------------------------------------------------------
# host programs: model layer and tableau bookkeeping, no device work
# -------------------------------------------------------------------------------------------------
HOST = [
    ("linexpr_ctor_forms", """
a = LinExpr()
b = LinExpr(5)
c = LinExpr(2, 'x')
d = LinExpr(2, 'x', '-1/3', 'y', 7)
e = LinExpr(1, 'x', -1, 'x', 4, 'z')
f = LinExpr('2/4', 'w', 0.5, 'w')
""", ["str(a)", "repr(a)", "str(b)", "repr(b)", "str(c)", "repr(c)", "str(d)", "repr(d)", "str(e)", "repr(e)",
      "str(f)", "repr(f)", "d.getConstant()", "d.getCoefficient('y')", "d.getCoefficient('nope')",
      "eval(repr(d)) == d", "a == 0", "b == 5", "b == '10/2'", "c == 2", "e == LinExpr(4, 'z')"]),
    ("linexpr_ctor_errors", "", [
        "LinExpr(1, 2)", "LinExpr(1, '3x')", "LinExpr(1, 'x y')", "LinExpr(1, '')", "LinExpr('a', 'x')",
        "LinExpr(1, '_ok9')", "LinExpr(1, 'x', 2, None)"]),
    # __eq__ coerces a scalar BEFORE looking at the variables (linprog.py:64-68): a bad scalar raises, it is not "unequal"
    ("linexpr_eq_coerces_first", """
a = LinExpr(2, 'x', 1)
b = LinExpr(3)
""", ["a == 'abc'", "a == None", "a == [1]", "b == 'abc'", "b == '6/2'", "b == 3.0", "a == 1", "a != 1", "b != 3",
      "a == LinExpr(2, 'x', 1)", "a.constraintLeq(1) == 1", "LinVar('x') == 'x'"]),
    ("linexpr_str_signs", """
a = LinExpr(-1, 'b', 1, 'a', -3)
b = LinExpr(-2, 'a', -5, 'c')
c = LinExpr('-7/2')
d = LinExpr('3/9', 'q', '-2/8', 'p', '1/2')
""", ["str(a)", "str(b)", "str(c)", "str(d)", "repr(a)", "repr(b)", "repr(c)", "repr(d)", "str(-a)", "str(+a)",
      "(+a) is a", "(-a) == LinExpr(1, 'b', -1, 'a', 3)"]),
    ("linexpr_arithmetic", """
x = LinExpr(1, 'x')
y = LinExpr(1, 'y')
s = x + y
t = x - y
u = 3 + x
v = 3 - x
w = x + '1/2'
z = (x + y) - (x + y)
k = x.copy(); k += y; k += 4
l = x.copy(); l -= x
n = x.copy(); n -= LinExpr(2, 'y', 1)
""", ["str(s)", "str(t)", "str(u)", "str(v)", "str(w)", "str(z)", "z == 0", "str(k)", "str(l)", "str(n)", "str(x)",
      "x + 1 == 1 + x", "str(x - 1)", "str(1 - x - y)", "str(F(2, 3) + y)"]),
    ("linexpr_evaluate_substitute", """
e = LinExpr(2, 'x', -3, 'y', 1)
""", ["e.evaluate({'x': 1, 'y': 2})", "e.evaluate({'x': '1/2', 'y': 0, 'unused': 9})", "e.evaluate({'x': 1})",
      "str(e.substitute({'x': LinExpr(1, 'a', 1, 'b')}))", "str(e.substitute({'x': 5}))",
      "str(e.substitute({'y': LinExpr(1, 'y', 1)}))", "str(e.substitute({'x': LinExpr(3, 'y', 2), 'y': LinExpr(2, 'x')}))",
      "str(e.substitute({}))", "str(e.substitute({'x': '3/2', 'y': '1/3'}))", "e.substitute({'x': 1, 'y': 1}) == 0",
      "str(e)"]),
    ("linexpr_constraints", """
e = LinExpr(1, 'x', 2, 'y')
a = e.constrantEq(4)
b = e.constraintLeq(LinExpr(1, 'z', 1))
c = e.constraintGeq('1/2')
""", ["str(a)", "repr(a)", "str(b)", "repr(b)", "str(c)", "repr(c)", "hasattr(e, 'constraintEq')"]),
    ("lincon_basics", """
a = LinCon(LinExpr(1, 'x', 1), '<=', LinExpr(2, 'y'))
b = LinCon(LinExpr(1, 'x'), '==', -4)
c = LinCon(3, '>=', 5)
d = a.copy()
""", ["str(a)", "repr(a)", "str(b)", "repr(b)", "str(c)", "repr(c)", "a == d", "a is d", "str(a.reverse())",
      "str(c.reverse())", "str(b.reverse())", "str(a.simplify())", "str(b.simplify())", "str(c.simplify())",
      "eval(repr(a)) == a", "a == b", "LinCon(3, '>=', LinExpr(1, 'x'))", "str(LinCon('1/2', '<=', '3/4'))"]),
    ("lincon_bad_comp", "", ["LinCon(LinExpr(1, 'x'), '<', 3)", "LinCon(LinExpr(1, 'x'), '=', 3)",
                             "LinCon(LinExpr(1, 'x'), '!=', 3)"]),
    ("lincon_evaluate", """
a = LinCon(LinExpr(1, 'x', 1, 'y'), '<=', LinExpr(4))
b = LinCon(LinExpr(2, 'x'), '==', LinExpr(1, 'y', 1))
c = LinCon(LinExpr(1, 'x'), '>=', LinExpr(1, 'y'))
""", ["a.evaluate({'x': 1, 'y': 3})", "a.evaluate({'x': 2, 'y': 3})", "b.evaluate({'x': 1, 'y': 1})",
      "b.evaluate({'x': 1, 'y': 2})", "c.evaluate({'x': '1/2', 'y': '1/2'})", "c.evaluate({'x': 0, 'y': 1})",
      "a.evaluate({'x': 1})"]),
    ("lincon_inplace", """
a = LinCon(LinExpr(1, 'x'), '<=', LinExpr(3))
a.addLeft(LinExpr(1, 'y'))
s1 = str(a)
a.addRight(2)
s2 = str(a)
a.subLeft(LinExpr(1, 'x', 1))
s3 = str(a)
a.subRight(LinExpr(1, 'z'))
s4 = str(a)
a.addLeft('1/2')
s5 = str(a)
""", ["s1", "s2", "s3", "s4", "s5", "str(a.simplify())", "repr(a)"]),
    ("linvar_bounds", """
a = LinVar('x')
b = LinVar('y', True, '1/2', '7/2')
c = LinVar('z', False, '1/2', '7/2')
d = LinVar('w', True, -2.5, None)
e = LinVar('v', True, None, '-1/3')
f = LinVar('u', True, 3, 3)
g = LinVar('t', True, '1/3', '2/3')
""", ["str(a)", "repr(a)", "str(b)", "repr(b)", "str(c)", "repr(c)", "str(d)", "repr(d)", "str(e)", "repr(e)",
      "str(f)", "str(g)", "a.getBounds()", "b.getBounds()", "c.getBounds()", "d.getBounds()", "e.getBounds()",
      "a.isFeasible()", "b.isFeasible()", "f.isFeasible()", "g.isFeasible()", "eval(repr(b)) == b", "b == c",
      "b.copy() == b", "b.copy() is b"]),
    ("linvar_tighten", """
a = LinVar('x', True, 0, 10)
a.boundAbove('9/2')
s1 = str(a)
a.boundAbove(7)
s2 = str(a)
a.boundBelow('1/3')
s3 = str(a)
a.boundBelow(-5)
s4 = str(a)
a.boundBelow(5)
s5 = str(a)
b = LinVar('y', False)
b.boundAbove('9/2')
b.boundBelow('19/4')
""", ["s1", "s2", "s3", "s4", "s5", "a.isFeasible()", "str(b)", "b.isFeasible()", "LinVar('3bad')", "LinVar('')"]),
    ("tableau_ctor_and_getters", """
t = mk([-3, -5, 0, 0], [[1, 0, 1, 0], [0, 2, 0, 1]], [4, 12], z=7, names=['x1', 'x2', 's1', 's2'],
       marks=[False, False, True, True])
""", ["t.getNumCons()", "t.getNumVars()", "t.getTableauSize()", "t.getZ()", "t.getC()", "t.getCj(1)", "t.getB()",
      "t.getBi(1)", "t.getA()", "t.getAij(1, 3)", "t.getVarNames()", "t.getVarName(2)", "t.getVarMarks()",
      "t.getVarMark(0)", "t.getCj(4)", "t.getBi(2)", "t.getAij(2, 0)", "t.getAij(0, 4)", "t.getVarName(9)",
      "t.getVarMark(4)", "T(0, 3)", "T(2, 0)", "T(-1, -1)"]),
    ("tableau_setters", """
t = T(2, 3)
t.setZ('3/2')
t.setCj(0, '-1/2')
t.setBi(1, 0.25)
t.setAij(1, 2, '7/3')
t.setVarName(1, 'mid')
t.setVarMark(2, True)
t.toggleVarMark(0)
t.toggleVarMark(2)
""", ["t.getZ()", "t.getC()", "t.getB()", "t.getA()", "t.getVarNames()", "t.getVarMarks()", "t.setC([1, 2])",
      "t.setB([1, 2, 3])", "t.setA([[1, 2, 3]])", "t.setA([[1, 2], [3, 4]])", "t.setVarNames(['a'])",
      "t.setVarMarks([True])", "t.setCj(3, 1)", "t.setBi(2, 1)", "t.setAij(0, 3, 1)"]),
    # the bulk setters assign element by element (tableau.py:131-153): what precedes a bad or missing element stays
    ("tableau_setters_stop_at_the_bad_element", """
t = T(2, 3)
""", ["t.setC([5, 'abc', 7])", "t.getC()", "t.setB(['1/2'])", "t.getB()", "t.setA([[1, 2, 3], [4, None, 6]])",
      "t.getA()", "t.setA([[9, 8]])", "t.getA()", "t.setC([0.5])", "t.getC()", "t.setVarNames(['p', 'q'])",
      "t.getVarNames()", "t.setVarMarks([True])", "t.getVarMarks()"]),
    ("tableau_resize", """
t = mk([1, 2], [[3, 4], [5, 6]], [7, 8], names=['a', 'b'])
t.addVar('c')
s1 = snap((t.getTableauSize(), t.getC(), t.getA(), t.getVarNames(), t.getVarMarks()))
t.addVars(['d', 'e'])
s2 = snap((t.getTableauSize(), t.getC(), t.getA(), t.getVarNames()))
t.addCon()
s3 = snap((t.getTableauSize(), t.getB(), t.getA()))
t.addCons(2)
s4 = snap((t.getTableauSize(), t.getB(), t.getA()))
""", ["s1", "s2", "s3", "s4"]),
    ("tableau_permute_copy_eq", """
t = mk([1, 2, 3], [[1, 2, 3], [4, 5, 6]], [7, 8], z=-1, names=['a', 'b', 'c'], marks=[True, False, False])
u = t.copy()
e1 = (u == t, u is t)
u.permuteCols([2, 0, 1])
s1 = snap((u.getC(), u.getA(), u.getVarNames(), u.getVarMarks(), u.getB(), u.getZ()))
u.permuteRows([1, 0])
s2 = snap((u.getA(), u.getB()))
e2 = (u == t)
""", ["e1", "s1", "s2", "e2", "t.permuteRows([0, 0])", "t.permuteRows([0])", "t.permuteCols([0, 1, 1])",
      "t.permuteCols([0, 1, 3])", "t == 5", "t.getA()"]),
    ("tableau_json", """
t = mk(['-1/2', 3], [[1, '2/3'], [0, -4]], ['5/7', 6], z='9/4', names=['x', 'y'], marks=[False, True])
j = t.saveJson()
u = T(1, 1)
u.loadJson(j)
""", ["sorted(j.keys())", "j", "u == t", "u.getTableauSize()", "u.getZ()", "u.getVarNames()", "u.getVarMarks()"]),
    ("tableau_print_formats", """
t = mk(['-1/2', 3, 0], [[1, '2/3', 0], [0, -4, 1]], ['5/7', 6], z='9/4', names=['x', 'yy', 's'],
       marks=[False, True, False])
""", ["t.printGrid()", "t.printGrid(False, False)", "t.printGrid(True, False, '[', ']')", "t.printText()",
      "t.printText(False)", "t.printText(True, True, 4, True)", "t.printLatex()", "t.printLatex(False, True)",
      "t.printCSV()", "t.printCSV(False, True)", "str(t)", "repr(t)[:9]"]),
]

# -------------------------------------------------------------------------------------------------
# device programs: row operations, pivots, form checks, Simplex
# -------------------------------------------------------------------------------------------------
_LP3 = "mk([-3, -2, -4, 0, 0, 0], [[1, 1, 2, 1, 0, 0], [2, 0, 3, 0, 1, 0], [2, 1, 3, 0, 0, 1]], [4, 5, 7])"

DEVICE = [
    ("tableau_row_ops", """
t = mk([1, -2, 0], [[2, 4, 6], [1, 1, 1]], [8, 3], z=5)
t.rowMult(0, '1/2')
s1 = snap((t.getA(), t.getB()))
t.rowDiv(1, 4)
s2 = snap((t.getA(), t.getB()))
t.rowAdd(0, 1, 4)
s3 = snap((t.getA(), t.getB()))
t.rowSub(1, 0)
s4 = snap((t.getA(), t.getB()))
t.rowAddToObj(0, 2)
s5 = snap((t.getC(), t.getZ()))
t.rowSubFromObj(1, F(1, 3))
s6 = snap((t.getC(), t.getZ()))
t.rowAdd(0, 1, 0)
t.rowMult(1, 1)
s7 = snap((t.getA(), t.getB()))
""", ["s1", "s2", "s3", "s4", "s5", "s6", "s7", "t.rowDiv(0, 0)", "t.rowMult(2, 1)", "t.rowMult(2, 2)",
      "t.rowAdd(0, 2, 1)", "t.rowAddToObj(2, 1)", "t.rowSubFromObj(1, '1/3')", "t.rowSub(0, 1, '1/3')"]),
    ("tableau_pivot_sequence", f"""
t = {_LP3}
t.pivot(0, 2)
s1 = snap((t.getZ(), t.getC(), t.getB(), t.getA(), t.getVarMarks()))
t.pivot(1, 0)
s2 = snap((t.getZ(), t.getC(), t.getB(), t.getA()))
t.pivot(2, 1)
s3 = snap((t.getZ(), t.getC(), t.getB(), t.getA()))
""", ["s1", "s2", "s3", "t.pivot(0, 5)", "t.pivot(3, 0)", "t.pivot(0, 6)", "t.isOptimal()"]),
    ("tableau_form_checks", f"""
t = {_LP3}
b1 = [-9, -9, -9]
r1 = t.isCanonical(b1)
u = mk([0, 0, -1], [[1, 0, 1], [0, 1, 1]], [2, 3])
b2 = [-9, -9]
r2 = u.isCanonical(b2)
v = mk([0, 1, -1], [[1, 1, 1], [1, 0, 1]], [-2, 3])
b3 = [-9, -9]
r3 = v.isCanonical(b3)
w = mk([0, 0, 0, 0], [[1, 1, 2, 0], [0, 0, 0, 1]], [1, 1])
b4 = [-9, -9]
r4 = w.isCanonical(b4)
x = mk([0, 0], [[2, 0], [0, 1]], [1, 1])
b5 = [-9, -9]
r5 = x.isCanonical(b5)
""", ["(r1, b1)", "(r2, b2)", "(r3, b3)", "(r4, b4)", "(r5, b5)", "t.isOptimal()", "u.isOptimal()",
      "mk([0, 1], [[1, 1]], [1]).isOptimal()", "t.isCanonical()"]),
    ("simplex_rules_and_accessors", f"""
t = {_LP3}
t.setVarNames(['x1', 'x2', 'x3', 's1', 's2', 's3'])
s = Simplex(t)
q0 = (list(s.getBasicSequence()), list(s.getBasicSequenceNames()), dict(s.getBFS()), dict(s.getBFSNames()), s.getObjValue())
p_std = s.findPivotStandard()
p_min = s.findPivotMinIndex()
p_max = s.findPivotMaxIncrease()
p_all = s.findPivotAll()
m0 = list(t.getVarMarks())
done = s.findPivotStandard(True)
q1 = (list(s.getBasicSequence()), dict(s.getBFS()), s.getObjValue(), list(t.getVarMarks()))
done2 = s.findPivotMaxIncrease(True)
q2 = (list(s.getBasicSequence()), dict(s.getBFS()), s.getObjValue())
done3 = s.findPivotMinIndex(True)
q3 = (list(s.getBasicSequence()), dict(s.getBFS()), s.getObjValue())
""", ["q0", "p_std", "p_min", "p_max", "p_all", "m0", "done", "q1", "done2", "q2", "done3", "q3",
      "s.findPivotStandard()", "s.findPivotAll()", "str(s)", "s.pivot(0, 1)", "s.pivot(1, 3)"]),
    ("simplex_solve_text_lp", f"""
t = {_LP3}
s = Simplex(t)
s.solve()
""", ["s.getBasicSequence()", "s.getBFS()", "s.getObjValue()", "t.isOptimal()", "s.findPivotStandard()",
      "s.findPivotMinIndex()", "s.findPivotMaxIncrease()", "t.getC()", "t.getB()", "t.getA()", "t.getVarMarks()",
      "str(s)"]),
    ("simplex_phase1_equalities", """
t = mk([2, 3, 1, 0], [[1, 1, 1, 0], [1, -1, 0, 1], [0, 1, 2, 0]], [6, -1, 5], names=['a', 'b', 'c', 'd'])
s = Simplex(t)
q0 = (list(s.getBasicSequence()), dict(s.getBFS()), s.getObjValue(), t.getTableauSize(), list(t.getVarNames()), list(t.getVarMarks()))
s.solve()
""", ["q0", "s.getBasicSequence()", "s.getBFSNames()", "s.getObjValue()", "t.getC()", "t.getB()", "t.getA()"]),
    ("simplex_special_cases", """
inf = mk([0, 0], [[1, 1], [1, 1]], [1, 2])
unb = mk([-1, 0], [[-1, 1]], [1])
su = Simplex(unb)
deg = mk([-1, -1, 0, 0], [[1, 1, 1, 0], [1, 0, 0, 1]], [0, 1])
sd = Simplex(deg)
sd.solve()
""", ["Simplex(inf)", "inf.getTableauSize()", "su.findPivotStandard()", "su.findPivotMinIndex()",
      "su.findPivotMaxIncrease()", "su.solve()", "sd.getBasicSequence()", "sd.getBFS()", "sd.getObjValue()",
      "deg.isDegenerate()"]),
    ("tableau_host_predicates", """
a = mk([-1, 2], [[-1, 1], [0, 1]], [1, 2])
b = mk([1, 2], [[-1, -1], [1, 1]], [1, 2])
c = mk([1, 2], [[1, 1], [1, 1]], [0, 2])
""", ["a.isUnbounded()", "b.isUnbounded()", "a.isInfeasible()", "b.isInfeasible()", "c.isInfeasible()",
      "a.isDegenerate()", "c.isDegenerate()"]),
]
