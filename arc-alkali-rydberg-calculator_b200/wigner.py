"""Angular-momentum algebra for the hot path (host side, exact-integer Racah sums).

Mirrors the public functions of the reference's arc/wigner.py -- Wigner3j (wigner.py:55-196),
Wigner6j (:199-371), CG (:497-510), WignerDmatrix (:513-590) -- with the same
argument conventions.  Written from the textbook Racah formulas with Python big
integers (no precomputed tables, no sympy), memoised on doubled quantum numbers.
"""
from functools import lru_cache
from math import cos, factorial, sin, sqrt

import numpy as np

__all__ = ["Wigner3j", "Wigner6j", "CG", "WignerDmatrix", "wigner_small_d"]


def _x2(v):
    r = round(2 * v)
    if abs(2 * v - r) > 1e-9:
        raise ValueError("All arguments must be integers or half-integers.")
    return int(r)


def _f(n2):
    """factorial of n2/2 for even non-negative doubled argument."""
    return factorial(n2 // 2)


@lru_cache(maxsize=None)
def _w3j(j1, j2, j3, m1, m2, m3):
    # all arguments doubled integers
    if m1 + m2 + m3 != 0:
        return 0.0
    if (j1 - m1) % 2 or (j2 - m2) % 2 or (j3 - m3) % 2:
        raise ValueError("2*j and 2*m must have the same parity")
    if j3 > j1 + j2 or j3 < abs(j1 - j2) or (j1 + j2 + j3) % 2:
        raise ValueError("j3 is out of bounds.")
    if abs(m1) > j1 or abs(m2) > j2 or abs(m3) > j3:
        raise ValueError("m is out of bounds.")
    t1 = (j2 - m1 - j3)
    t2 = (j1 + m2 - j3)
    t3 = (j1 + j2 - j3)
    t4 = (j1 - m1)
    t5 = (j2 + m2)
    tmin = max(0, t1, t2)
    tmax = min(t3, t4, t5)
    # Racah sum with exact rationals: sum_t (-1)^t / prod(factorials)
    from fractions import Fraction
    acc = Fraction(0)
    for t in range(tmin, tmax + 1, 2):
        den = (_f(t) * _f(t - t1) * _f(t - t2) * _f(t3 - t) * _f(t4 - t) * _f(t5 - t))
        acc += Fraction((-1) ** (t // 2), den)
    pref = Fraction(_f(j1 + j2 - j3) * _f(j1 - j2 + j3) * _f(-j1 + j2 + j3),
                    _f(j1 + j2 + j3 + 2))
    pref *= (_f(j1 + m1) * _f(j1 - m1) * _f(j2 + m2) * _f(j2 - m2)
             * _f(j3 + m3) * _f(j3 - m3))
    sign = -1.0 if ((j1 - j2 - m3) // 2) % 2 else 1.0
    val = acc * acc * pref           # exact square; take signed sqrt
    s = 1.0 if acc >= 0 else -1.0
    return sign * s * sqrt(float(val))


def Wigner3j(j1, j2, j3, m1, m2, m3):
    """Wigner 3-j symbol (wigner.py:55-196)."""
    return _w3j(_x2(j1), _x2(j2), _x2(j3), _x2(m1), _x2(m2), _x2(m3))


def CG(j1, m1, j2, m2, j3, m3):
    """Clebsch-Gordan <j1 m1 j2 m2 | j3 m3> (wigner.py:497-510)."""
    a, b, c = _x2(j1), _x2(j2), _x2(j3)
    ma, mb, mc = _x2(m1), _x2(m2), _x2(m3)
    if ma + mb != mc or c > a + b or c < abs(a - b) or (a + b + c) % 2 \
            or abs(ma) > a or abs(mb) > b or abs(mc) > c:
        return 0.0
    sign = -1.0 if ((a - b + mc) // 2) % 2 else 1.0
    return _w3j(a, b, c, ma, mb, -mc) * sqrt(c + 1.0) * sign


def _tri(a, b, c):
    from fractions import Fraction
    return Fraction(_f(a + b - c) * _f(a - b + c) * _f(-a + b + c), _f(a + b + c + 2))


@lru_cache(maxsize=None)
def _w6j(a, b, c, d, e, f):
    # { a b c ; d e f } doubled integers; triads (a,b,c) (a,e,f) (d,b,f) (d,e,c)
    for (x, y, z) in ((a, b, c), (a, e, f), (d, b, f), (d, e, c)):
        if z > x + y or z < abs(x - y) or (x + y + z) % 2:
            return 0.0
    from fractions import Fraction
    tmin = max(a + b + c, a + e + f, d + b + f, d + e + c)
    tmax = min(a + b + d + e, b + c + e + f, a + c + d + f)
    acc = Fraction(0)
    for t in range(tmin, tmax + 1, 2):
        den = (_f(t - a - b - c) * _f(t - a - e - f) * _f(t - d - b - f) * _f(t - d - e - c)
               * _f(a + b + d + e - t) * _f(b + c + e + f - t) * _f(a + c + d + f - t))
        acc += Fraction((-1) ** (t // 2) * _f(t + 2), den)
    pref = _tri(a, b, c) * _tri(a, e, f) * _tri(d, b, f) * _tri(d, e, c)
    s = 1.0 if acc >= 0 else -1.0
    return s * sqrt(float(acc * acc * pref))


def Wigner6j(j1, j2, j3, J1, J2, J3):
    """Wigner 6-j symbol {j1 j2 j3; J1 J2 J3} (wigner.py:199-371)."""
    return _w6j(_x2(j1), _x2(j2), _x2(j3), _x2(J1), _x2(J2), _x2(J3))


@lru_cache(maxsize=None)
def _small_d_coeffs(j2, m2, n2):
    """terms (coef, power_cos, power_sin) of d^j_{m n}(beta), z-y-z convention."""
    jpm, jmm = (j2 + m2) // 2, (j2 - m2) // 2
    jpn, jmn = (j2 + n2) // 2, (j2 - n2) // 2
    mmn = (m2 - n2) // 2
    pref = sqrt(float(factorial(jpm) * factorial(jmm) * factorial(jpn) * factorial(jmn)))
    out = []
    for s in range(max(0, -mmn), min(jpn, jmm) + 1):
        den = factorial(jpn - s) * factorial(s) * factorial(mmn + s) * factorial(jmm - s)
        out.append(((-1.0) ** (mmn + s) * pref / den,
                    (2 * j2 + n2 - m2) // 2 - 2 * s, mmn + 2 * s))
    return tuple(out)


def wigner_small_d(j, m, n, beta):
    """d^j_{m n}(beta) = <j m| exp(-i beta J_y) |j n>."""
    j2, m2, n2 = _x2(j), _x2(m), _x2(n)
    c, s = cos(0.5 * beta), sin(0.5 * beta)
    tot = 0.0
    for coef, pc, ps in _small_d_coeffs(j2, m2, n2):
        tot += coef * (c ** pc) * (s ** ps)
    return tot


class WignerDmatrix:
    """D^j(phi, theta, gamma) with rows/cols ordered m = -j..j (wigner.py:513-590).

    `get(j)` returns a dense complex128 (2j+1)x(2j+1) ndarray,
    D[m, n] = exp(-i m phi) d^j_{m n}(theta) exp(-i n gamma).
    The reference switches to a Bessel-function approximation when j > m+10 and
    j > n+10 (wigner.py:419-421); that quirk is reproduced when
    `reference_bessel_quirk` is True (default) so results match the reference.
    """

    def __init__(self, theta, phi, gamma=0.0, reference_bessel_quirk=True):
        self.theta, self.phi, self.gamma = theta, phi, gamma
        self.trivial = abs(theta) < 1e-5 and abs(phi) < 1e-5 and abs(gamma) < 1e-5
        self._quirk = reference_bessel_quirk
        self._cache = {}

    def get(self, j):
        j2 = _x2(j)
        if j2 in self._cache:
            return self._cache[j2]
        dim = j2 + 1
        if self.trivial:
            mat = np.eye(dim, dtype=np.complex128)
        else:
            mat = np.zeros((dim, dim), dtype=np.complex128)
            ms = [(-j2 + 2 * k) / 2.0 for k in range(dim)]
            for a, m in enumerate(ms):
                for b, n in enumerate(ms):
                    if self._quirk and (j > m + 10) and (j > n + 10):
                        from scipy.special import jv
                        d = float(jv(m - n, j * self.theta))
                    else:
                        d = wigner_small_d(j, m, n, self.theta)
                    mat[a, b] = np.exp(-1j * m * self.phi) * d * np.exp(-1j * n * self.gamma)
        self._cache[j2] = mat
        return mat
