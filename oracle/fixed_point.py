"""CPU restatement of the exact fixed-point contraction of the integer statistics path.  TEST INFRASTRUCTURE.

The engine (functionalmf_b200/csrc/stats_i8.cu + i8gemm.cu) computes the product block of the
sufficient statistics

    out[m, c] = sum_k cnt[m, k] * Z[k, c],      Z[k, (k1, k2)] = F[k, k1] * F[k, k2]   (k2 <= k1, packed)

(the `c * v v^T` sums of factor.py:351-356 and 396-400 with the integer replicate counts as weights)
by writing every column of Z as a fixed-point number against a power-of-two column scale,

    e_c = smallest integer with max_k |Z[k, c]| < 2^e_c,     q[k, c] = rint(Z[k, c] * 2^(54 - e_c)),

splitting q into eight signed base-128 digits, contracting the digit planes with the counts in exact
int32 arithmetic on the int8 tensor cores and recombining them with ONE rounding:

    out[m, c] = RN( (sum_k cnt[m, k] * q[k, c]) * 2^(e_c - 54) ).

This module states that definition in Python integers, so the device result can be compared bit for bit
(tests/test_gpu_i8.py) and its distance to the exact rational sum bounded (tests/test_fixed_point_oracle.py).
"""
from fractions import Fraction
import numpy as np

FIXBITS = 54
NPLANES = 8
DIGIT_OFFSET = sum(64 * 128 ** s for s in range(NPLANES))      # 0x0081020408102040


def column_exponent(z):
    """smallest e with max|z| < 2^e (0 for an all-zero column)"""
    mx = float(np.abs(z).max()) if len(z) else 0.0
    if mx == 0.0:
        return 0
    m, e = np.frexp(mx)            # mx = m * 2^e, 0.5 <= m < 1  ->  mx < 2^e
    return int(e)


def quantise(z, e):
    """q = rint(z * 2^(54 - e)) as Python integers (ties to even, like the device's cvt.rni)"""
    return [int(np.rint(np.ldexp(float(v), FIXBITS - e))) for v in z]


def digits(q):
    """the eight signed base-128 digits of q, carry-free as on the device: fields of q + sum_s 64 * 128^s, minus 64"""
    qq = q + DIGIT_OFFSET
    assert 0 < qq < 128 ** NPLANES
    return [((qq >> (7 * s)) & 127) - 64 for s in range(NPLANES)]


def recombine(planes):
    """two int64 Horner sums (each below 2^53 in magnitude), one rounding: fma(hi, 2^28, lo)"""
    hi = ((planes[7] * 128 + planes[6]) * 128 + planes[5]) * 128 + planes[4]
    lo = ((planes[3] * 128 + planes[2]) * 128 + planes[1]) * 128 + planes[0]
    assert abs(hi) < 2 ** 53 and abs(lo) < 2 ** 53
    return hi * 2 ** 28 + lo          # exact integer; the device rounds it once when converting


def product_block(F, counts, check_digits=False):
    """counts [m, k] (non-negative integers), F [k, K]  ->  out [m, L] float64, L = K (K + 1) / 2 packed (k1 >= k2).
    check_digits: also go through the eight digit planes and their recombination (slow; the CPU test does)."""
    F = np.asarray(F, dtype=np.float64)
    counts = np.asarray(counts)
    K = F.shape[1]
    out = np.zeros((counts.shape[0], K * (K + 1) // 2))
    c = 0
    for k1 in range(K):
        for k2 in range(k1 + 1):
            z = F[:, k1] * F[:, k2]
            e = column_exponent(z)
            q = quantise(z, e)
            dig = [digits(v) for v in q] if check_digits else None
            for m in range(counts.shape[0]):
                tot = sum(int(cc) * v for cc, v in zip(counts[m], q) if cc)
                if check_digits:
                    planes = [sum(int(cc) * d[s] for cc, d in zip(counts[m], dig) if cc) for s in range(NPLANES)]
                    assert recombine(planes) == tot
                # one rounding: integer -> nearest double (Python's int -> float conversion rounds to nearest even),
                # then an exact power-of-two scaling
                out[m, c] = float(np.ldexp(np.float64(float(tot)), e - FIXBITS))
            c += 1
    return out


def exact_product_block(F, counts):
    """the same sums in exact rational arithmetic (list of lists of Fraction)"""
    F = np.asarray(F, dtype=np.float64)
    K = F.shape[1]
    rows = []
    for m in range(len(counts)):
        row = []
        for k1 in range(K):
            for k2 in range(k1 + 1):
                row.append(sum(int(cc) * Fraction(float(a)) * Fraction(float(b))
                               for cc, a, b in zip(counts[m], F[:, k1], F[:, k2]) if cc))
        rows.append(row)
    return rows
