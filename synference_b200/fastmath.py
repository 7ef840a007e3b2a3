"""Tables for the weight builder's float64 special functions (normal-CDF tail, log, exp).

``weights_kernel`` is instruction-issue bound, and most of its instructions came from the general-purpose
``erfc`` / ``log`` / ``exp`` of the CUDA math library (branchy range reduction, constants built from pairs of 32-bit
immediates).  The kernel only needs three narrow functions, so it evaluates them from small tables built here:

* ``log_tab``  [256][2]   ln(m0), 1/m0 at the centres m0 of 256 mantissa intervals of [1, 2)
* ``exp_tab``  [64]       2**(j/64)
* ``tail_tab`` [n][8]     degree-7 local polynomials of g(u) = exp(u*u/2) * Q(u), Q(u) = erfc(u/sqrt 2)/2 the normal tail,
                          on intervals of width ``tail_w``; Q(u) = exp(-u*u/2) * g(u)

The ``ref_*`` functions restate the device arithmetic in numpy float64 (same operation order) so the accuracy can be
checked on the CPU (tests/test_fastmath.py): relative error < 2e-13 for log and exp, < 1e-12 for the tail.
"""

from __future__ import annotations

import numpy as np
from scipy import special

TAIL_W = 0.25
TAIL_UMAX = 40.0          # Q(40) = 3.7e-350: zero in float64
LN2_64_HI = float(np.ldexp(np.round(np.ldexp(np.log(2.0) / 64.0, 40)), -40))   # 2**-6 ln2, low 40+ bits cleared
LN2_64_LO = float(np.log(2.0) / 64.0 - LN2_64_HI)


def build_tables():
    i = np.arange(256)
    m0 = 1.0 + (i + 0.5) / 256.0
    log_tab = np.stack([np.log(m0), 1.0 / m0], 1)
    exp_tab = np.exp2(np.arange(64) / 64.0)
    n = int(round(TAIL_UMAX / TAIL_W))
    tail = np.zeros((n, 8))
    # Chebyshev interpolation of g on each interval (scipy's erfcx is accurate to ~1 ulp), converted to a
    # polynomial in s = u - centre
    k = np.arange(16)
    nodes = np.cos(np.pi * (k + 0.5) / 16)
    for j in range(n):
        c = (j + 0.5) * TAIL_W
        u = c + 0.5 * TAIL_W * nodes
        g = 0.5 * special.erfcx(u / np.sqrt(2.0))
        cheb = np.polynomial.chebyshev.chebfit(nodes, g, 7)
        mono = np.polynomial.chebyshev.cheb2poly(cheb)            # in x = s / (w/2)
        tail[j, :mono.size] = mono / (0.5 * TAIL_W) ** np.arange(mono.size)
    return dict(log_tab=np.ascontiguousarray(log_tab), exp_tab=np.ascontiguousarray(exp_tab),
                tail_tab=np.ascontiguousarray(tail), tail_w=TAIL_W, tail_n=n)


# ---- numpy restatements of the device code (same operation order) --------------------------------------

def ref_log(x, t):
    x = np.asarray(x, dtype=np.float64)
    bits = x.view(np.int64)
    e = (bits >> 52) - 1023
    idx = (bits >> 44) & 0xFF
    m = ((bits & 0x000FFFFFFFFFFFFF) | 0x3FF0000000000000).view(np.float64)
    lm0, inv = t["log_tab"][idx, 0], t["log_tab"][idx, 1]
    r = m * inv - 1.0                                 # device: fma(m, inv, -1)
    p = 0.2
    p = p * r - 0.25
    p = p * r + 1.0 / 3.0
    p = p * r - 0.5
    p = p * r + 1.0
    return e * 0.6931471805599453 + (lm0 + r * p)


def ref_exp(y, t):
    y = np.asarray(y, dtype=np.float64)
    k = np.rint(y * (64.0 / 0.6931471805599453))
    r = (y - k * LN2_64_HI) - k * LN2_64_LO
    ki = k.astype(np.int64)
    j, e = ki & 63, ki >> 6
    p = 1.0 / 120.0
    p = p * r + 1.0 / 24.0
    p = p * r + 1.0 / 6.0
    p = p * r + 0.5
    p = p * r + 1.0
    p = p * r + 1.0
    v = t["exp_tab"][j] * p
    out = np.ldexp(v, e.astype(np.int64))
    return np.where(e < -1021, 0.0, out)


def ref_tail(u, t):
    """Q(u) = 1 - Phi(u) for u >= 0."""
    u = np.asarray(u, dtype=np.float64)
    j = np.minimum((u / t["tail_w"]).astype(np.int64), t["tail_n"] - 1)
    s = u - (j + 0.5) * t["tail_w"]
    c = t["tail_tab"][j]
    g = c[:, 7]
    for d in range(6, -1, -1):
        g = g * s + c[:, d]
    q = ref_exp(-0.5 * u * u, t) * g
    return np.where(u >= TAIL_UMAX, 0.0, q)
