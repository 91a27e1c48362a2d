"""Exact-mode device engine: a rational tableau as a fraction-free integer grid on the GPU.

The reference computes in `fractions.Fraction` (tableau.py:2).  Here a tableau T of Fractions is
held on the device as an integer matrix N and one common denominator D, as residues modulo a set of
31-bit primes sized from a Hadamard bound (include/lps_b200.h, "exact mode"):

    T[i][j] == Fraction(N[i][j], D * sigma[i])

`sigma[i]` is the positive integer that cleared the denominators of row i at upload time; a row
loses it the first time it is a pivot row (the pivot row of a simplex tableau is scale free).  The
host side of this module only *encodes* (Fractions -> integer limbs, prime set) and *decodes*
(mixed-radix digits -> Python ints -> Fractions); every arithmetic step of a pivot, pricing, ratio
test and the solve loop runs in the CUDA library.  There is no CPU fallback.
"""
from __future__ import annotations

import math
from ctypes import byref, c_int, c_longlong, c_size_t, c_void_p
from fractions import Fraction
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N

# ------------------------------------------------------------------------------------------------
# prime set
# ------------------------------------------------------------------------------------------------
_MR_BASES = (2, 3, 5, 7)            # deterministic Miller-Rabin below 3,215,031,751


def _is_prime(q: int) -> bool:
    if q < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if q % p == 0:
            return q == p
    d, s = q - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in _MR_BASES:
        x = pow(a, d, q)
        if x in (1, q - 1):
            continue
        for _ in range(s - 1):
            x = x * x % q
            if x == q - 1:
                break
        else:
            return False
    return True


_PRIMES: List[int] = []


def primes_below_2_31(count: int, skip: Sequence[int] = ()) -> List[int]:
    """The `count` largest primes below 2^31 (descending), leaving out those in `skip`."""
    q = (_PRIMES[-1] - 2) if _PRIMES else (1 << 31) - 1
    while len(_PRIMES) < count + len(skip):
        if _is_prime(q):
            _PRIMES.append(q)
        q -= 2
    skip = set(skip)
    out = [p for p in _PRIMES if p not in skip]
    return out[:count]


def hadamard_bits(rows: List[List[int]]) -> int:
    """Bits of a bound on |minor| over all square submatrices of the integer matrix `rows`:
    min(product of row norms, product of the min(#rows, #cols) largest column norms), each norm
    clamped to >= 1 (so that sub-minors are covered too)."""
    nr, nc = len(rows), len(rows[0])
    row_sq = [max(1, sum(x * x for x in r)) for r in rows]
    col_sq = [0] * nc
    for r in rows:
        for j, x in enumerate(r):
            if x:
                col_sq[j] += x * x
    col_sq = [max(1, c) for c in col_sq]
    order = min(nr, nc)
    by_rows = sum(math.log2(v) for v in sorted(row_sq, reverse=True)[:order]) / 2
    by_cols = sum(math.log2(v) for v in sorted(col_sq, reverse=True)[:order]) / 2
    return int(math.ceil(min(by_rows, by_cols))) + 2


def hadamard_bits_np(G: np.ndarray) -> int:
    """Same bound for an int64 matrix with small entries (float64 norms are accurate enough; the +2
    bits of `hadamard_bits` plus one more absorb their rounding)."""
    A = G.astype(np.float64)
    row_sq = np.maximum(1.0, (A * A).sum(axis=1))
    col_sq = np.maximum(1.0, (A * A).sum(axis=0))
    order = min(A.shape)
    by_rows = np.log2(np.sort(row_sq)[::-1][:order]).sum() / 2
    by_cols = np.log2(np.sort(col_sq)[::-1][:order]).sum() / 2
    return int(math.ceil(min(by_rows, by_cols))) + 3


def column_weights(G) -> List[int]:
    """w[j] >= 256 * log2 of the 2-norm of column j (all m+1 rows) of the integer matrix G (ndarray or nested
    lists), 0 for a zero column: the weights of `lps_exact_set_weights`."""
    import math
    if isinstance(G, np.ndarray):
        sq = (G.astype(np.float64) ** 2).sum(axis=0) * (1.0 + 1e-9)
        out = np.zeros(G.shape[1], dtype=np.int64)
        nzc = sq > 0
        out[nzc] = np.ceil(128.0 * np.log2(sq[nzc])).astype(np.int64) + 1
        return [int(max(v, 0)) for v in out]
    cols = len(G[0])
    sums = [0] * cols
    for row in G:
        for j, v in enumerate(row):
            sums[j] += v * v
    return [0 if s == 0 else max(0, math.ceil(128.0 * math.log2(s) * (1.0 + 1e-12)) + 1) for s in sums]


def weights_from_squares(sq: Sequence[int]) -> List[int]:
    """w >= 256 * log2(sqrt(s)) for exact squared column norms s (Python ints), 0 for s = 0."""
    return [0 if s == 0 else max(0, math.ceil(128.0 * math.log2(s) * (1.0 + 1e-12)) + 1) for s in sq]


def bound_bits_from_squares(row_sq: Sequence[float], col_sq: Sequence[float]) -> int:
    """`hadamard_bits` from squared row and column norms (floats; one more guard bit for their rounding)."""
    order = min(len(row_sq), len(col_sq))
    by_rows = sum(math.log2(max(1.0, v)) for v in sorted(row_sq, reverse=True)[:order]) / 2
    by_cols = sum(math.log2(max(1.0, v)) for v in sorted(col_sq, reverse=True)[:order]) / 2
    return int(math.ceil(min(by_rows, by_cols))) + 3


def row_combine_coefficients(s_d: int, s_s: Optional[int], mult: Fraction) -> Tuple[int, int, int]:
    """(a, b, scale) of a row operation in the fraction-free form.  Entries are N / (D s) with s the row's scale:
    row_d + (p/q) row_s = (q s_s N_d + p s_d N_s) / (D s_d q s_s);  s_s None: (p/q) row_d = p N_d / (D s_d q).
    Common factors of the three integers are removed."""
    p, q = mult.numerator, mult.denominator
    if s_s is None:
        a, b, scale = p, 0, s_d * q
    else:
        a, b, scale = q * s_s, p * s_d, s_d * q * s_s
    g = math.gcd(a, b, scale)
    return a // g, b // g, scale // g


def primes_for_bits(bits: int, skip: Sequence[int] = ()) -> List[int]:
    """Enough primes that their product exceeds 2^(bits+2): symmetric range plus one guard bit."""
    need = bits + 2
    count = max(2, int(math.ceil(need / 30.999999)))
    while True:
        ps = primes_below_2_31(count, skip)
        if sum(math.log2(p) for p in ps) > need:
            return ps
        count += 1


# ------------------------------------------------------------------------------------------------
# encode / decode
# ------------------------------------------------------------------------------------------------
def clear_denominators(grid) -> Tuple[List[List[int]], List[int]]:
    """Fractions -> (integer rows, per-row scale sigma) with sigma_i = lcm of row i's denominators."""
    ints, sigma = [], []
    for row in grid:
        s = 1
        for x in row:
            d = x.denominator
            if d != 1:
                s = s * d // math.gcd(s, d)
        sigma.append(s)
        ints.append([int(x * s) for x in row] if s != 1 else [int(x) for x in row])
    return ints, sigma


def to_limbs(values: Sequence[int], nlimbs: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray, int]:
    """Signed Python ints -> (uint32 limbs [len, nlimbs] little endian, uint8 negative flags, nlimbs)."""
    mags = [abs(int(v)) for v in values]
    top = max(mags, default=0)
    need = max(1, (top.bit_length() + 31) // 32)
    nlimbs = max(need, nlimbs or 1)
    neg = np.fromiter((1 if int(v) < 0 else 0 for v in values), dtype=np.uint8, count=len(mags))
    if nlimbs <= 2:
        arr = np.fromiter(mags, dtype=np.uint64, count=len(mags))
        limbs = np.empty((len(mags), nlimbs), dtype=np.uint32)
        limbs[:, 0] = (arr & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        if nlimbs == 2:
            limbs[:, 1] = (arr >> np.uint64(32)).astype(np.uint32)
        return limbs, neg, nlimbs
    buf = b"".join(v.to_bytes(4 * nlimbs, "little") for v in mags)
    return np.frombuffer(buf, dtype="<u4").reshape(len(mags), nlimbs).copy(), neg, nlimbs


class MixedRadix:
    """Digits (p_0, p_1, ...) -> Python ints, vectorised over numbers with numpy object arrays."""

    def __init__(self, primes: Sequence[int]):
        self.primes = [int(p) for p in primes]

    def decode(self, sign: np.ndarray, digits: np.ndarray) -> List[int]:
        """sign: int8 [count]; digits: uint32 [P, count] (digit k of |x|).  Horner from the highest
        digit that is nonzero anywhere."""
        count = sign.shape[0]
        nz = np.nonzero(digits.any(axis=1))[0]
        acc = np.zeros(count, dtype=object)
        if nz.size:
            for k in range(int(nz[-1]), -1, -1):
                acc = acc * self.primes[k] + digits[k].astype(object)
        out = [int(v) for v in acc]
        return [-v if s < 0 else v for v, s in zip(out, sign.tolist())]


# ------------------------------------------------------------------------------------------------
# engine
# ------------------------------------------------------------------------------------------------
# Closed handles wait here for the next tableau of the same shape and prime set: creating and destroying a handle
# costs several milliseconds of driver calls (allocation, stream, captured graph) -- more than a 50 x 100 solve --
# and every upload resets the whole device state anyway.  Bounded; handles whose shape was changed (`phase2`) or that
# were switched to per-launch profiling are destroyed instead.
_HANDLE_POOL: dict = {}
_HANDLE_POOL_MAX = 8
_HANDLE_POOL_BYTES = 1 << 30


def _pool_size() -> int:
    return sum(len(v) for v in _HANDLE_POOL.values())


def drain_handle_pool():
    """Destroy every pooled handle (tests, memory-sensitive callers)."""
    L = N.lib()
    for lst in _HANDLE_POOL.values():
        for h in lst:
            L.lps_destroy(h)
    _HANDLE_POOL.clear()


class ExactEngine:
    """Exact rational tableau on one GPU (same method surface as `engine.F64Engine`)."""

    mode = "exact"

    def __init__(self, m: int, n: int, device: Optional[int] = None, primes: Optional[Sequence[int]] = None):
        import torch
        self.L = N.lib()
        self.device = int(torch.cuda.current_device() if device is None else device)
        self.m, self.n = int(m), int(n)
        self.h = None
        self.primes: List[int] = []
        self.forced_primes = [int(p) for p in primes] if primes is not None else None
        self.banned: List[int] = []
        self.reencodes = 0
        self.last_bits = 0
        self.adaptive = True          # False: every kernel runs on all the primes the static bound asks for

    # -- lifetime ------------------------------------------------------------------------------
    def _create(self, primes: Sequence[int]):
        self.close()
        key = (self.device, self.m, self.n, tuple(int(p) for p in primes))
        pooled = _HANDLE_POOL.get(key)
        if pooled:
            h = pooled.pop()
        else:
            arr = np.asarray(primes, dtype=np.uint32)
            h = c_void_p()
            N.check(self.L.lps_create_exact(self.m, self.n, len(primes), arr.ctypes.data, self.device, None, byref(h)))
        self.h = h
        self._pool_key = key
        self.primes = [int(p) for p in primes]
        self.codec = MixedRadix(self.primes)

    def close(self):
        if getattr(self, "h", None):
            key = getattr(self, "_pool_key", None)
            small = (self.m + 1) * (self.n + 32) * 4 * len(self.primes) <= _HANDLE_POOL_BYTES   # bound on its residue planes
            if key is not None and small and _pool_size() < _HANDLE_POOL_MAX:
                _HANDLE_POOL.setdefault(key, []).append(self.h)
            else:
                self.L.lps_destroy(self.h)
        self.h = None
        self._pool_key = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        N.check(self.L.lps_sync(self.h))

    # -- transfers -----------------------------------------------------------------------------
    def upload(self, grid, cost_hint=None):
        """grid: (m+1) x (n+1) Fractions / ints in device layout, or an int64 numpy array.

        cost_hint (phase 1, simplex.py:50-51): the original objective -- n0 <= n reduced costs followed by the _z
        cell -- that `phase2` will install on the device once the artificial problem is solved.  The prime set is
        then sized for the matrix [constraint rows ; that objective] as well."""
        self.phase2_info = None
        if isinstance(grid, np.ndarray) and grid.dtype.kind in "iu":
            G = np.ascontiguousarray(grid, dtype=np.int64)
            ints = None
            bits = hadamard_bits_np(G)
            sigma = [1] * (self.m + 1)
            mags = np.abs(G).astype(np.uint64)
            nl = 1 if int(mags.max(initial=0)) < (1 << 32) else 2
            limbs = np.empty(G.shape + (nl,), dtype=np.uint32)
            limbs[..., 0] = (mags & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            if nl == 2:
                limbs[..., 1] = (mags >> np.uint64(32)).astype(np.uint32)
            neg = (G < 0).astype(np.uint8)
        else:
            G = None
            rows = [[x if isinstance(x, Fraction) else Fraction(x) for x in row] for row in grid]
            ints, sigma = clear_denominators(rows)
            bits = hadamard_bits(ints)
            flat = [v for row in ints for v in row]
            limbs, neg, nl = to_limbs(flat)
        sigma_bits = max(s.bit_length() for s in sigma)
        if cost_hint is not None:
            hint_bits, self.phase2_info = self._prepare_phase2(cost_hint, G if G is not None else ints, sigma)
            bits = max(bits, hint_bits)
            sigma_bits = max(sigma_bits, self.phase2_info["sigma_m"].bit_length())
        bits += sigma_bits                               # "== 1" tests compare with D * sigma_i
        self.last_bits = bits
        primes = self.forced_primes if self.forced_primes is not None else primes_for_bits(bits, self.banned)
        if self.h is None or primes != self.primes:
            self._create(primes)
        sl, _, snl = to_limbs(sigma)
        limbs = np.ascontiguousarray(limbs)
        neg = np.ascontiguousarray(neg)
        self.sigma = sigma
        self.op_bits = 0                                 # growth allowed for by row operations since this upload
        self.wmax_bits = 0
        if self.adaptive and self.forced_primes is None:
            # working prime set: sized by the norms of the current basis columns, extended on the device on demand;
            # only its residue planes are allocated and encoded now
            wcol = np.ascontiguousarray(column_weights(G if G is not None else ints), dtype=np.int32)
            self.wmax_bits = (int(wcol.max(initial=0)) + 255) // 256
            N.check(self.L.lps_upload_exact_weighted(self.h, limbs.ctypes.data, nl, neg.ctypes.data, self.n + 1,
                                                     sl.ctypes.data, snl, wcol.ctypes.data, sigma_bits + 3))
        else:
            N.check(self.L.lps_upload_exact(self.h, limbs.ctypes.data, nl, neg.ctypes.data, self.n + 1, sl.ctypes.data, snl))

    # -- leaving phase 1 on the device ----------------------------------------------------------------
    def _prepare_phase2(self, cost_hint, ints, sigma):
        """The integer objective row C = sigma_m * cost that `phase2` installs, the squared column norms of the
        uploaded constraint rows, and the bits of the Hadamard bound of the uploaded matrix with its objective row
        replaced, entry by entry, by the larger of the two objective rows in magnitude (covers both phases)."""
        m, n = self.m, self.n
        n0 = len(cost_hint) - 1
        assert 0 < n0 <= n
        cq = [v if isinstance(v, Fraction) else Fraction(v) for v in cost_hint]
        sigma_m = math.lcm(*[v.denominator for v in cq], *[int(s) for s in sigma[:m]])
        C = [int(v * sigma_m) for v in cq]
        wide = [0] * (n + 1)                              # C in the uploaded layout: artificial columns cost 0
        wide[:n0] = C[:n0]
        wide[n] = C[n0]
        if isinstance(ints, np.ndarray):
            A = ints[:m].astype(np.float64)
            col_sq_a = [int(v) for v in (ints[:m].astype(object) ** 2).sum(axis=0)] if m * (n + 1) <= 1 << 16 else \
                [int(math.ceil(v * (1.0 + 1e-9))) for v in (A * A).sum(axis=0)]
            row_sq = [float(v) for v in (A * A).sum(axis=1)]
            obj = [int(v) for v in ints[m]]
        else:
            col_sq_a = [0] * (n + 1)
            row_sq = []
            for row in ints[:m]:
                acc = 0
                for j, v in enumerate(row):
                    if v:
                        vv = v * v
                        col_sq_a[j] += vv
                        acc += vv
                row_sq.append(float(acc))
            obj = ints[m]
        both = [max(abs(a), abs(b)) for a, b in zip(obj, wide)]
        row_sq.append(float(sum(v * v for v in both)))
        col_sq = [float(col_sq_a[j] + both[j] * both[j]) for j in range(n + 1)]
        info = {"n0": n0, "C": C, "sigma_m": sigma_m, "col_sq_a": col_sq_a, "n_up": n}
        return bound_bits_from_squares(row_sq, col_sq), info

    def phase2(self, n_new: int, basis: Sequence[int]) -> bool:
        """Simplex._find_bfs simplex.py:95-105 on the device (`lps_phase2_exact`): cut the artificial columns, install
        the original objective priced out against `basis`.  False (nothing done) when the upload carried no cost hint
        for this shape or the tableau has been re-encoded since."""
        info = getattr(self, "phase2_info", None)
        if info is None or self.h is None or info["n0"] != n_new or info["n_up"] != self.n or len(basis) != self.m:
            return False
        if self.forced_primes is not None:               # a caller-chosen prime set was not sized for the new objective
            return False
        m = self.m
        C, sigma_m = info["C"], info["sigma_m"]
        alive = self._alive()
        G = []
        for i in range(m):
            q, rem = divmod(C[basis[i]], self.sigma[i] if alive[i] else 1)
            assert rem == 0
            G.append(q)
        sq = [info["col_sq_a"][j] + C[j] * C[j] for j in range(n_new)] + [info["col_sq_a"][self.n] + C[n_new] * C[n_new]]
        wcol = np.ascontiguousarray(weights_from_squares(sq), dtype=np.int32)
        self.wmax_bits = (int(wcol.max(initial=0)) + 255) // 256
        sig = self.sigma[:m] + [sigma_m]
        guard = max(s.bit_length() for s in sig) + 3 + self.op_bits
        cl, cn, cnl = to_limbs(C)
        gl, gn, gnl = to_limbs(G)
        sl, _, snl = to_limbs([sigma_m])
        cl, cn, gl, gn = (np.ascontiguousarray(a) for a in (cl, cn, gl, gn))
        bas = np.ascontiguousarray(basis, dtype=np.int32)
        N.check(self.L.lps_phase2_exact(self.h, n_new, cl.ctypes.data, cnl, cn.ctypes.data, gl.ctypes.data, gnl, gn.ctypes.data,
                                        sl.ctypes.data, snl, bas.ctypes.data, wcol.ctypes.data, guard))
        self.n = n_new
        self.sigma = sig
        self.phase2_info = None
        self._pool_key = None                            # the handle's width changed: not reusable for the old shape
        return True

    def _read(self, what: int, index: int, count: int) -> List[int]:
        sign = np.empty(count, dtype=np.int8)
        digits = np.zeros((len(self.primes), count), dtype=np.uint32)     # rows beyond the working set stay 0
        N.check(self.L.lps_read_exact(self.h, what, index, sign.ctypes.data, digits.ctypes.data))
        return self.codec.decode(sign, digits)

    def denominator(self) -> int:
        return self._read(2, 0, 1)[0]

    def _alive(self) -> np.ndarray:
        out = np.empty(self.m + 1, dtype=np.int32)
        N.check(self.L.lps_read_alive(self.h, out.ctypes.data))
        return out

    def _row_den(self, D: int, alive, i: int) -> int:
        return D * self.sigma[i] if alive[i] else D

    def read_row(self, i: int) -> List[Fraction]:
        D, alive = self.denominator(), self._alive()
        den = self._row_den(D, alive, i)
        return [Fraction(v, den) for v in self._read(0, i, self.n + 1)]

    def read_col(self, j: int) -> List[Fraction]:
        D, alive = self.denominator(), self._alive()
        vals = self._read(1, j, self.m + 1)
        return [Fraction(v, self._row_den(D, alive, i)) for i, v in enumerate(vals)]

    def download(self) -> List[List[Fraction]]:
        D, alive = self.denominator(), self._alive()
        out = []
        for i in range(self.m + 1):
            den = self._row_den(D, alive, i)
            out.append([Fraction(v, den) for v in self._read(0, i, self.n + 1)])
        return out

    def integer_state(self) -> Tuple[List[List[int]], int]:
        """(N, D) as Python ints -- what tests compare with the oracle's Bareiss restatement."""
        return [self._read(0, i, self.n + 1) for i in range(self.m + 1)], self.denominator()

    def info(self):
        p, bad, nbytes = c_int(0), c_int(-1), c_size_t(0)
        N.check(self.L.lps_exact_info(self.h, byref(p), byref(bad), byref(nbytes)))
        w, a, g = c_int(0), c_int(0), c_int(0)
        N.check(self.L.lps_exact_working(self.h, byref(w), byref(a), byref(g)))
        return {"primes": p.value, "bad_prime": bad.value, "residue_bytes": nbytes.value, "bound_bits": self.last_bits,
                "allocated_primes": a.value, "extensions": g.value}

    # -- re-encoding when a prime divides a pivot element -----------------------------------------
    def _reencode(self, drop_prime: bool = True):
        """Download the (still intact) tableau and upload it again: without the offending prime (LPS_BAD_PRIME), or
        -- after row operations outgrew the handle's primes (LPS_NEED_PRIMES) -- as a new matrix with its own bound."""
        if drop_prime:
            bad = self.info()["bad_prime"]
            if self.forced_primes is not None:
                if len(self.forced_primes) <= 1:
                    raise N.NativeError("prime set exhausted while re-encoding")
                self.forced_primes = [p for k, p in enumerate(self.forced_primes) if k != bad]
            else:
                self.banned.append(self.primes[bad])
        grid = self.download()
        self.reencodes += 1
        self.upload(grid)

    # -- hot path ------------------------------------------------------------------------------
    def pivot(self, r: int, c: int) -> int:
        while True:
            rc = N.check(self.L.lps_pivot(self.h, r, c))
            if rc not in (N.BAD_PRIME, N.NEED_PRIMES):
                return rc
            self._reencode(drop_prime=(rc == N.BAD_PRIME))

    def find_pivot(self, rule: int):
        r, c = c_int(-1), c_int(-1)
        rc = N.check(self.L.lps_find_pivot(self.h, rule, byref(r), byref(c)))
        if rc == N.OPTIMAL:
            return "optimal"
        if rc == N.UNBOUNDED:
            return "unbounded"
        return r.value, c.value

    # -- row operations outside a pivot (tableau.py:254-293) ------------------------------------------------
    def row_combine(self, rd: int, rs: Optional[int], mult: Fraction) -> bool:
        """row rd <- row rd + mult * row rs, or (rs None) row rd <- mult * row rd, on the device residues
        (`lps_row_combine_exact`); rd == m addresses the objective row.  In the fraction-free form, with s the scale a
        row carries (1 once pivoted) and mult = p/q:   N_rd <- (q s_rs) N_rd + (p s_rd) N_rs,  scale s_rd q s_rs
        (common factors removed).  False, with nothing changed, when this handle cannot do it (fixed prime set, or the
        grown magnitude bound does not fit its primes): the caller edits a host copy instead."""
        if self.h is None or not self.adaptive or self.forced_primes is not None:
            return False
        alive = self._alive()
        s_d = self.sigma[rd] if alive[rd] else 1
        s_s = None if rs is None else (self.sigma[rs] if alive[rs] else 1)
        a, b, new_sigma = row_combine_coefficients(s_d, s_s, mult)
        # magnitudes grow by at most |a| + |b|; a row addition may also pair entries with the basis of the other row
        grow = (abs(a) + abs(b) - 1).bit_length() + (self.wmax_bits + 1 if rs is not None else 0)
        sig = self.sigma[:]
        sig[rd] = new_sigma
        guard = max(s.bit_length() for s in sig) + 3 + self.op_bits + grow
        al, _, anl = to_limbs([a])
        bl, _, bnl = to_limbs([b])
        sl, _, snl = to_limbs([new_sigma])
        rc = N.check(self.L.lps_row_combine_exact(self.h, rd, -1 if rs is None else rs, al.ctypes.data, anl, 1 if a < 0 else 0,
                                                  bl.ctypes.data, bnl, 1 if b < 0 else 0, sl.ctypes.data, snl, guard))
        if rc == N.NEED_PRIMES:
            return False
        self.sigma = sig
        self.op_bits += grow
        self.phase2_info = None                          # the uploaded matrix is no longer what the cost hint was sized for
        return True

    # -- whole-tableau rules (simplex.py:286-360) ---------------------------------------------------------
    def read_cells(self, rows: Sequence[int], cols: Sequence[int]) -> List[Fraction]:
        """Entries (rows[e], cols[e]) as Fractions (one reconstruction pass on the device)."""
        count = len(rows)
        if count == 0:
            return []
        out: List[Fraction] = []
        D, alive = self.denominator(), self._alive()
        step = max(self.n, self.m) + 1
        for lo in range(0, count, step):
            rr = np.ascontiguousarray(rows[lo:lo + step], dtype=np.int32)
            cc = np.ascontiguousarray(cols[lo:lo + step], dtype=np.int32)
            k = rr.size
            sign = np.empty(k, dtype=np.int8)
            digits = np.zeros((len(self.primes), k), dtype=np.uint32)
            N.check(self.L.lps_read_exact_cells(self.h, k, rr.ctypes.data, cc.ctypes.data, sign.ctypes.data, digits.ctypes.data))
            vals = self.codec.decode(sign, digits)
            out += [Fraction(v, self._row_den(D, alive, int(i))) for v, i in zip(vals, rr)]
        return out

    def col_ratios(self, only_negative: bool, want_ties: bool = False):
        """Minimum-ratio row of every column (`lps_exact_col_ratios`): (minrow [n], tie lists or None)."""
        minrow = np.empty(self.n, dtype=np.int32)
        if not want_ties:
            N.check(self.L.lps_exact_col_ratios(self.h, 1 if only_negative else 0, minrow.ctypes.data, None, None, 0))
            return minrow, None
        off = np.zeros(self.n + 1, dtype=np.int64)
        cap = self.n + self.m
        while True:
            ties = np.empty(max(cap, 1), dtype=np.int32)
            N.check(self.L.lps_exact_col_ratios(self.h, 1 if only_negative else 0, minrow.ctypes.data, off.ctypes.data,
                                                ties.ctypes.data, cap))
            if off[self.n] <= cap:
                return minrow, [ties[off[j]:off[j + 1]].tolist() for j in range(self.n)]
            cap = int(off[self.n])

    def has_unbounded_column(self) -> bool:
        """Tableau.isUnbounded (tableau.py:504-508): a column with c_j < 0 and no positive entry."""
        minrow, _ = self.col_ratios(True)
        return bool((minrow == -1).any())

    def find_pivot_all(self):
        """Simplex.findPivotAll (simplex.py:330-360): every (row, column) of minimum ratio, column by column."""
        _, ties = self.col_ratios(False, want_ties=True)
        return [(i, j) for j, rows in enumerate(ties) for i in rows]

    def find_pivot_max_increase(self):
        """Simplex.findPivotMaxIncrease (simplex.py:286-328).  The m x n part -- the ratio test of every column with
        c_j < 0 -- runs on the device; what is left is one exact product per candidate column, -c_j * b_i / a_ij,
        compared in column order with the reference's strict ">" (first column among equals)."""
        minrow, _ = self.col_ratios(True)
        cand = [j for j in range(self.n) if minrow[j] != -2]
        if not cand:
            return "optimal"
        if any(minrow[j] == -1 for j in cand):
            return "unbounded"
        rows = [int(minrow[j]) for j in cand]
        cost = self.read_row(self.m)
        rhs = self.read_col(self.n)
        piv = self.read_cells(rows, cand)
        best, si, sj = None, -1, -1
        for j, i, a in zip(cand, rows, piv):
            inc = -cost[j] * (rhs[i] / a)
            if best is None or inc > best:
                best, si, sj = inc, i, j
        return si, sj

    def is_canonical(self, want_cols: bool = True, cols: Optional[np.ndarray] = None):
        """(canonical?, bcols).  The library leaves bcols untouched when some b_i < 0 (tableau.py:476-477): pass
        `cols` pre-filled to tell."""
        if cols is None:
            cols = np.full(self.m, -1, dtype=np.int32)
        flag = c_int(0)
        N.check(self.L.lps_is_canonical(self.h, cols.ctypes.data if want_cols else None, byref(flag)))
        return bool(flag.value), cols

    def is_optimal(self) -> bool:
        flag = c_int(0)
        N.check(self.L.lps_is_optimal(self.h, byref(flag)))
        return bool(flag.value)

    def is_infeasible(self) -> bool:
        """Tableau.isInfeasible (tableau.py:510-514) on the device."""
        flag = c_int(0)
        N.check(self.L.lps_is_infeasible(self.h, byref(flag)))
        return bool(flag.value)

    def solve(self, basis: np.ndarray, marks: Optional[np.ndarray] = None, max_pivots: int = -1, trace_cap: int = 0):
        """`Simplex.solve` on the device; basis / marks are int32 arrays updated in place."""
        assert basis.dtype == np.int32 and basis.size == self.m
        total, pieces = 0, []
        left = max_pivots
        while True:
            trace = np.zeros((max(trace_cap - total, 0), 3), dtype=np.int32) if trace_cap > 0 else None
            npiv = c_longlong(0)
            rc = N.check(self.L.lps_solve(self.h, left, basis.ctypes.data, marks.ctypes.data if marks is not None else None,
                                          trace.ctypes.data if trace is not None and trace.size else None,
                                          trace.shape[0] if trace is not None else 0, byref(npiv)))
            k = npiv.value
            if trace is not None:
                pieces.append(trace[: min(k, trace.shape[0])])
            total += k
            if rc not in (N.BAD_PRIME, N.NEED_PRIMES):
                break
            # a prime divides the next pivot element (probability ~2^-31 per prime and pivot), or row operations
            # outgrew the handle's primes: re-encode and continue the loop exactly where it stopped
            rule, stuck, moved, pend = c_int(0), c_longlong(0), c_int(0), c_int(0)
            N.check(self.L.lps_exact_loop_state(self.h, 0, byref(rule), byref(stuck), byref(moved), byref(pend)))
            self._reencode(drop_prime=(rc == N.BAD_PRIME))
            N.check(self.L.lps_exact_loop_state(self.h, 1, byref(rule), byref(stuck), byref(moved), byref(pend)))
            if left >= 0:
                left -= k
        tr = np.concatenate(pieces) if pieces else None
        return rc, total, tr

    # -- profiling -----------------------------------------------------------------------------
    def prof_reset(self):
        self._pool_key = None                            # per-launch profiling stays on for the handle's life
        N.check(self.L.lps_prof_reset(self.h))

    def prof_read(self):
        from ctypes import c_double
        ms, nu, na = c_double(0), c_longlong(0), c_longlong(0)
        N.check(self.L.lps_prof_read(self.h, byref(ms), byref(nu), byref(na)))
        return ms.value, nu.value, na.value
