"""Host-side pieces of the exact mode (no GPU): prime set, Hadamard bound, limb encoding, mixed-radix
decoding, and the fraction-free integer form the kernels implement, checked against the oracle."""
import math
import random
from fractions import Fraction

import numpy as np
import pytest

from linear_program_solver_b200 import exact as X
from linear_program_solver_b200 import problems
from oracle import lpsol_oracle as O


def test_prime_set_is_descending_31_bit_primes():
    sympy = pytest.importorskip("sympy")
    ps = X.primes_below_2_31(40)
    assert ps[0] == 2147483647 and ps == sorted(ps, reverse=True) and len(set(ps)) == 40
    assert all(sympy.isprime(p) and (1 << 30) < p < (1 << 31) for p in ps)
    # no prime skipped between the first and the last
    assert sum(1 for q in range(ps[-1], ps[0] + 1, 2) if sympy.isprime(q)) == 40
    assert X.primes_below_2_31(3, skip=[ps[1]]) == [ps[0], ps[2], ps[3]]


def test_primes_for_bits_cover_the_symmetric_range():
    for bits in (1, 30, 31, 62, 400, 533, 15970):
        ps = X.primes_for_bits(bits)
        assert math.prod(ps) > (1 << (bits + 2))
        assert math.prod(ps[:-1]) <= (1 << (bits + 2)) or len(ps) == 2


@pytest.mark.parametrize("m,seed,hi", [(6, 0, 9), (12, 1, 9), (20, 2, 127)])
def test_hadamard_bound_covers_every_bareiss_integer_of_a_solve(m, seed, hi):
    """|N_ij| and |D| along a whole reference solve stay below the bound that sizes the prime set."""
    c, A, b = problems.dense_lp(m, seed, 1, hi)
    grid = O.make_grid(c, A, b)
    ints = [[int(x) for x in row] for row in grid]
    bits = X.hadamard_bits(ints)
    assert abs(X.hadamard_bits_np(np.array(ints, dtype=np.int64)) - bits) <= 1
    N, D = [row[:] for row in ints], 1
    basis = list(range(m, 2 * m))
    trace = []
    O.solve(grid, basis, trace)
    for r, cc, _ in trace:
        D = O.bareiss_pivot(N, D, r, cc)
        assert abs(D).bit_length() <= bits
        assert max(abs(x).bit_length() for row in N for x in row) <= bits
    # and the integer form is the reference's tableau: entry == N_ij / D everywhere
    assert all(Fraction(N[i][j], D) == grid[i][j] for i in range(m + 1) for j in range(2 * m + 1))


def test_clear_denominators_and_limbs_roundtrip():
    rng = random.Random(5)
    grid = [[Fraction(rng.randint(-50, 50), rng.randint(1, 12)) for _ in range(7)] for _ in range(4)]
    ints, sigma = X.clear_denominators(grid)
    assert all(Fraction(ints[i][j], sigma[i]) == grid[i][j] for i in range(4) for j in range(7))
    assert all(s == math.lcm(*[x.denominator for x in row]) for s, row in zip(sigma, grid))
    vals = [0, 1, -1, (1 << 32) - 1, -(1 << 32), (1 << 64) - 1, -(1 << 200) + 12345, 1 << 1000]
    limbs, neg, nl = X.to_limbs(vals)
    back = [int.from_bytes(limbs[i].astype("<u4").tobytes(), "little") * (-1 if neg[i] else 1) for i in range(len(vals))]
    assert back == vals and nl == 32
    small, neg2, nl2 = X.to_limbs([5, -7, 0, (1 << 40) + 3])
    assert nl2 == 2 and small[3].tolist() == [3, 256] and neg2.tolist() == [0, 1, 0, 0]


def test_mixed_radix_decode():
    ps = X.primes_below_2_31(12)
    codec = X.MixedRadix(ps)
    rng = random.Random(1)
    M = math.prod(ps)
    vals = [0, 1, -1, ps[0], -ps[0] * ps[1] - 5, M // 2, -(M // 2)] + [rng.randrange(-M // 2, M // 2) for _ in range(20)]
    digits = np.zeros((len(ps), len(vals)), dtype=np.uint32)
    sign = np.zeros(len(vals), dtype=np.int8)
    for e, v in enumerate(vals):
        sign[e] = (v > 0) - (v < 0)
        a = abs(v)
        for k, p in enumerate(ps):
            digits[k, e] = a % p
            a //= p
    assert codec.decode(sign, digits) == vals
