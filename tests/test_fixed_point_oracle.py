"""The fixed-point definition of the integer statistics path (oracle/fixed_point.py), on the CPU:
digits reconstruct q, the recombination is the integer sum, and the result is within
n * R * 2^-55 * 2^e_c of the exact rational sum - closer than a float64 accumulation in general."""
from fractions import Fraction
import numpy as np
import pytest

from oracle import fixed_point as FP


def test_digits_round_trip_and_range():
    rs = np.random.RandomState(0)
    for q in [0, 1, -1, 2 ** 54, -2 ** 54, 127, 128, -128, -129] + [int(v) for v in rs.randint(-2 ** 62, 2 ** 62, size=2000) >> 8]:
        d = FP.digits(q)
        assert len(d) == 7 and all(-128 <= x <= 127 for x in d)
        assert sum(x * 256 ** s for s, x in enumerate(d)) == q


def test_column_exponent_is_tight():
    for mx in [1.0, 0.999, 2.0, 3.7e-5, 1e300, 5e-324]:
        e = FP.column_exponent(np.array([0.0, -mx, mx / 3]))
        assert mx < 2.0 ** e if e < 1024 else True
        assert e <= -1073 or mx >= 2.0 ** (e - 1)
    assert FP.column_exponent(np.zeros(3)) == 0


@pytest.mark.parametrize('K,n,R', [(3, 40, 3), (5, 200, 2), (8, 64, 127)])
def test_product_block_error_bound(K, n, R):
    rs = np.random.RandomState(K * 100 + n)
    F = rs.normal(size=(n, K)) * np.exp(rs.normal(size=(1, K)))
    counts = rs.randint(0, R + 1, size=(4, n))
    got = FP.product_block(F, counts, check_digits=True)
    exact = FP.exact_product_block(F, counts)
    naive = np.zeros_like(got)
    c = 0
    for k1 in range(K):
        for k2 in range(k1 + 1):
            z = F[:, k1] * F[:, k2]                      # the product itself is one float64 rounding on every path
            e = FP.column_exponent(z)
            zx = [Fraction(float(v)) for v in z]
            for m in range(counts.shape[0]):
                want = sum(int(cc) * v for cc, v in zip(counts[m], zx) if cc)   # exact sum of the ROUNDED products
                err = abs(Fraction(float(got[m, c])) - want)
                bound = Fraction(int(counts[m].sum()), 2 ** 55) * Fraction(2) ** e + abs(want) * Fraction(1, 2 ** 52)
                assert err <= bound, (m, c, float(err), float(bound))
                naive[m, c] = float(np.dot(counts[m].astype(float), z))
            c += 1
    # and it agrees with the plain float64 contraction to accumulation noise
    np.testing.assert_allclose(got, naive, rtol=0, atol=1e-13 * np.abs(naive).max())
    assert len(exact) == counts.shape[0]


def test_fast_product_block_equals_the_definition():
    rs = np.random.RandomState(5)
    F = rs.normal(size=(90, 5)) * np.exp(2 * rs.normal(size=(1, 5)))
    counts = rs.randint(0, 4, size=(7, 90))
    assert np.array_equal(FP.product_block_fast(F, counts), FP.product_block(F, counts))
