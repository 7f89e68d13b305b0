#!/usr/bin/env python3
"""Groundwork for cic_comb_fir_kernel's output scale (csrc/decimator.cu scale_out): RN(sum / 375) without the division.

scale_out() computes (float)((double)sum / 24576000.0) (rtlsdr_ft8d.c:197-198) as __fdiv_rn(sum, 375.0f) * 2^-16 -- eight IEEE
divisions per thread, ~14 instructions each, 11 % of the kernel's instructions (profiles/ncu_lines_cic_comb_fir_kernel_r2y.txt).
With c = RN(1/375) the three-instruction sequence

    q0 = RN(sum * c);   r = fma(-375, q0, sum)  (exact);   q1 = fma(r, c, q0)

is the correctly rounded quotient for EVERY float in the normal range: this script proves it with exact integer arithmetic over all
2^23 mantissas (the identity is binade-independent: scaling by powers of two is exact while nothing under- or overflows, and
scale_out already sends |sum| < 1e-20 down the FP64 path).  Not applied to the kernel in round 2: the GPU budget was spent when it
was found; tests/test_oracle_golden.py::test_division_by_375_in_three_instructions keeps the proof."""
import numpy as np


def round_to_24_bits(num, shift):
    """RN-even of num / 2^shift to 24 significant bits for positive int64 num: -> (mantissa in [2^23, 2^24), exponent e) with value
    mantissa * 2^(e), where the unrounded value is num * 2^-shift."""
    num = num.astype(np.int64)
    bl = np.floor(np.log2(num.astype(np.float64))).astype(np.int64) + 1          # bit length (exact: num < 2^62, checked below)
    bl = np.where((np.int64(1) << (bl - 1).clip(0, 62)) > num, bl - 1, bl)
    bl = np.where((np.int64(1) << bl.clip(0, 62)) <= num, bl + 1, bl)
    drop = bl - 24                                                                # low bits to round away (may be <= 0)
    d = drop.clip(1, 62)
    low = num & ((np.int64(1) << d) - 1)
    half = np.int64(1) << (d - 1)
    q = num >> d
    up = (low > half) | ((low == half) & ((q & 1) == 1))
    q = np.where(drop > 0, q + up, num << (-drop).clip(0, 62))
    e = np.where(drop > 0, drop, drop) - shift
    carry = q >= (1 << 24)
    q = np.where(carry, q >> 1, q)
    e = np.where(carry, e + 1, e)
    return q, e


def prove():
    C = int(round(2 ** 32 / 375))                     # c = RN(1/375) = C * 2^-32, C has 24 significant bits
    assert 2 ** 23 <= C < 2 ** 24 and abs(C / 2 ** 32 - 1 / 375) * 375 < 2 ** -24
    assert np.float32(1.0) / np.float32(375.0) == np.float32(C * 2.0 ** -32)
    m = np.arange(2 ** 23, 2 ** 24, dtype=np.int64)   # sum = m (binade 2^23..2^24; every other binade is a power-of-two scaling)
    # q0 = RN(m * c)
    q0, e0 = round_to_24_bits(m * C, 32)              # q0 * 2^e0
    # r = m - 375 * q0 * 2^e0, exact; in units of 2^e0 (e0 < 0 here)
    assert np.all(e0 <= 0)
    R = (m << (-e0)) - 375 * q0                       # r = R * 2^e0
    assert np.all(np.abs(R) < 2 ** 24), "the residual is a float (so the fma that forms it is exact)"
    # q1 = RN(q0 + r * c) = 2^e0 * RN(q0 + R * C * 2^-32)
    T = (q0 << 32) + R * C                            # (q0 + R c) in units of 2^(e0 - 32); positive, < 2^57
    assert np.all(T > 0) and np.all(T < 2 ** 61)
    q1, e1 = round_to_24_bits(T, 32)
    e1 = e1 + e0
    # reference: RN(m / 375), exact: scale so that the quotient has >= 26 bits, keep the remainder for the sticky bit
    K = 34
    num = m << K
    quo, rem = num // 375, num % 375
    sticky = (quo << 1) | (rem != 0)                  # one extra low bit that is 1 iff the division is inexact: ties stay ties only if exact
    qr, er = round_to_24_bits(sticky, K + 1)
    ok = (q1 == qr) & (e1 == er)
    return int(np.count_nonzero(~ok)), int(m.size)


if __name__ == "__main__":
    bad, n = prove()
    print("mantissas checked: %d, mismatches: %d" % (n, bad))
