"""B200-native simplex pivot engine behind the `lpsol` API.

Same public names as the reference package (/root/reference/lpsol/__init__.py:5-17):
`Tableau`, `Simplex`, `LinExpr`, `LinCon`, `LinVar`, `LinProg`.  The numbers of a tableau live
on the GPU; `Tableau.pivot`, pricing, the ratio test and the whole `Simplex.solve` loop run in
hand-written sm_100a kernels behind the C-ABI of include/lps_b200.h.  Importing the package
does not touch the GPU; the first device operation raises `NativeUnavailable` if the CUDA
library or a device is missing (there is no CPU fallback).
"""
from ._native import NativeError, NativeUnavailable
from .model import LinCon, LinExpr, LinVar
from .tableau import Tableau
from .simplex import Simplex
from .linprog import LinProg

__all__ = ["Tableau", "Simplex", "LinExpr", "LinCon", "LinVar", "LinProg", "NativeUnavailable", "NativeError"]
