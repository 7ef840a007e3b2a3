"""Inoue et al. (2014) IGM attenuation: coefficient table and device tables.

The reference applies ``synthesizer.emission_models.attenuation.Inoue14`` in both
entry points (``library.py:2462``, ``library.py:2604``, ``library.py:5765``); the
class and its coefficient files live in the third-party package.
The 39-row Lyman-series table below (lambda_j, A^LAF_{1..3}, A^DLA_{1..2}) is
reproduced from Inoue+14 Table 2 as distributed with that implementation; it is
a *data input* shared by the CUDA path and the oracle (both receive the arrays
through their arguments), so parity does not depend on the digits.

Two things are built here for the CUDA path:
  * :func:`device_tables` - per-wavelength-bin power tables and per-line prefix
    sums that make tau(z, lambda_i (1+z)) a short sum of separable
    ``bin_power[i] * (1+z)**p`` products (no ``pow`` in the kernel);
  * :func:`transmission` - a vectorised host evaluation used by the slow
    single-galaxy utilities (plotting-style helpers), never by the hot path.
"""

from __future__ import annotations

import numpy as np

__all__ = ["INOUE14_LAF", "INOUE14_DLA", "LAM_L", "transmission", "device_tables", "Inoue14"]

LAM_L = 911.8  # Lyman limit used by the reference implementation [Angstrom]

# j, lambda_j [A], A_LAF1, A_LAF2, A_LAF3
INOUE14_LAF = np.array([
    [2, 1215.670, 1.68976e-02, 2.35379e-03, 1.02611e-04],
    [3, 1025.720, 4.69229e-03, 6.53625e-04, 2.84940e-05],
    [4, 972.537, 2.23898e-03, 3.11884e-04, 1.35962e-05],
    [5, 949.743, 1.31901e-03, 1.83735e-04, 8.00974e-06],
    [6, 937.803, 8.70656e-04, 1.21280e-04, 5.28707e-06],
    [7, 930.748, 6.17843e-04, 8.60640e-05, 3.75186e-06],
    [8, 926.226, 4.60924e-04, 6.42055e-05, 2.79897e-06],
    [9, 923.150, 3.56887e-04, 4.97135e-05, 2.16720e-06],
    [10, 920.963, 2.84278e-04, 3.95992e-05, 1.72628e-06],
    [11, 919.352, 2.31771e-04, 3.22851e-05, 1.40743e-06],
    [12, 918.129, 1.92348e-04, 2.67936e-05, 1.16804e-06],
    [13, 917.181, 1.62155e-04, 2.25878e-05, 9.84689e-07],
    [14, 916.429, 1.38498e-04, 1.92925e-05, 8.41033e-07],
    [15, 915.824, 1.19611e-04, 1.66615e-05, 7.26340e-07],
    [16, 915.329, 1.04314e-04, 1.45306e-05, 6.33446e-07],
    [17, 914.919, 9.17397e-05, 1.27791e-05, 5.57091e-07],
    [18, 914.576, 8.12784e-05, 1.13219e-05, 4.93564e-07],
    [19, 914.286, 7.25069e-05, 1.01000e-05, 4.40299e-07],
    [20, 914.039, 6.50549e-05, 9.06198e-06, 3.95047e-07],
    [21, 913.826, 5.86816e-05, 8.17421e-06, 3.56345e-07],
    [22, 913.641, 5.31918e-05, 7.40949e-06, 3.23008e-07],
    [23, 913.480, 4.84261e-05, 6.74563e-06, 2.94068e-07],
    [24, 913.339, 4.42740e-05, 6.16726e-06, 2.68854e-07],
    [25, 913.215, 4.06311e-05, 5.65981e-06, 2.46733e-07],
    [26, 913.104, 3.73821e-05, 5.20723e-06, 2.27003e-07],
    [27, 913.006, 3.45377e-05, 4.81102e-06, 2.09731e-07],
    [28, 912.918, 3.19891e-05, 4.45601e-06, 1.94255e-07],
    [29, 912.839, 2.97110e-05, 4.13867e-06, 1.80421e-07],
    [30, 912.768, 2.76635e-05, 3.85346e-06, 1.67987e-07],
    [31, 912.703, 2.58178e-05, 3.59636e-06, 1.56779e-07],
    [32, 912.645, 2.41479e-05, 3.36374e-06, 1.46638e-07],
    [33, 912.592, 2.26347e-05, 3.15296e-06, 1.37450e-07],
    [34, 912.543, 2.12567e-05, 2.96100e-06, 1.29081e-07],
    [35, 912.499, 1.99967e-05, 2.78549e-06, 1.21430e-07],
    [36, 912.458, 1.88476e-05, 2.62543e-06, 1.14452e-07],
    [37, 912.420, 1.77928e-05, 2.47850e-06, 1.08047e-07],
    [38, 912.385, 1.68222e-05, 2.34330e-06, 1.02153e-07],
    [39, 912.353, 1.59286e-05, 2.21882e-06, 9.67268e-08],
    [40, 912.324, 1.50996e-05, 2.10334e-06, 9.16925e-08],
])

# j, lambda_j [A], A_DLA1, A_DLA2
INOUE14_DLA = np.array([
    [2, 1215.670, 1.61698e-04, 5.38995e-05],
    [3, 1025.720, 1.54539e-04, 5.15129e-05],
    [4, 972.537, 1.49767e-04, 4.99222e-05],
    [5, 949.743, 1.46031e-04, 4.86769e-05],
    [6, 937.803, 1.42893e-04, 4.76312e-05],
    [7, 930.748, 1.40159e-04, 4.67196e-05],
    [8, 926.226, 1.37714e-04, 4.59048e-05],
    [9, 923.150, 1.35495e-04, 4.51650e-05],
    [10, 920.963, 1.33452e-04, 4.44841e-05],
    [11, 919.352, 1.31561e-04, 4.38536e-05],
    [12, 918.129, 1.29785e-04, 4.32617e-05],
    [13, 917.181, 1.28117e-04, 4.27056e-05],
    [14, 916.429, 1.26540e-04, 4.21799e-05],
    [15, 915.824, 1.25041e-04, 4.16804e-05],
    [16, 915.329, 1.23614e-04, 4.12046e-05],
    [17, 914.919, 1.22248e-04, 4.07494e-05],
    [18, 914.576, 1.20938e-04, 4.03127e-05],
    [19, 914.286, 1.19681e-04, 3.98938e-05],
    [20, 914.039, 1.18469e-04, 3.94896e-05],
    [21, 913.826, 1.17298e-04, 3.90995e-05],
    [22, 913.641, 1.16167e-04, 3.87225e-05],
    [23, 913.480, 1.15071e-04, 3.83572e-05],
    [24, 913.339, 1.14011e-04, 3.80037e-05],
    [25, 913.215, 1.12983e-04, 3.76609e-05],
    [26, 913.104, 1.11972e-04, 3.73241e-05],
    [27, 913.006, 1.11002e-04, 3.70005e-05],
    [28, 912.918, 1.10051e-04, 3.66836e-05],
    [29, 912.839, 1.09125e-04, 3.63749e-05],
    [30, 912.768, 1.08220e-04, 3.60734e-05],
    [31, 912.703, 1.07337e-04, 3.57789e-05],
    [32, 912.645, 1.06473e-04, 3.54909e-05],
    [33, 912.592, 1.05629e-04, 3.52096e-05],
    [34, 912.543, 1.04802e-04, 3.49340e-05],
    [35, 912.499, 1.03991e-04, 3.46636e-05],
    [36, 912.458, 1.03198e-04, 3.43994e-05],
    [37, 912.420, 1.02420e-04, 3.41402e-05],
    [38, 912.385, 1.01657e-04, 3.38856e-05],
    [39, 912.353, 1.00908e-04, 3.36359e-05],
    [40, 912.324, 1.00168e-04, 3.33895e-05],
])

Z1_LAF, Z2_LAF, Z1_DLA = 1.2, 4.7, 2.0

# exponents of the per-bin power table, in the order the kernel indexes them
BIN_POWERS = (1.2, 2.1, 3.7, 5.5, -0.3, 2.0, 3.0, 1.0)
# exponents of the per-galaxy (1+z) power vector
Z_POWERS = (1.2, 2.1, 3.7, 5.5, -0.3, 2.0, 3.0, -0.9, 1.6, 3.4, 2.3, 3.3)


def transmission(z, lam_obs, laf=INOUE14_LAF, dla=INOUE14_DLA):
    """exp(-tau) at observed wavelengths ``lam_obs`` [A] for a source at ``z`` (vectorised)."""
    lobs = np.asarray(lam_obs, dtype=float)
    zp1 = 1.0 + float(z)
    lj = laf[:, 1][:, None]
    u = lobs[None, :] / lj
    on = lobs[None, :] < lj * zp1
    r1 = on & (u < 1 + Z1_LAF)
    r2 = on & (u >= 1 + Z1_LAF) & (u < 1 + Z2_LAF)
    r3 = on & (u >= 1 + Z2_LAF)
    tau = (np.where(r1, laf[:, 2][:, None] * u**1.2, 0.0)
           + np.where(r2, laf[:, 3][:, None] * u**3.7, 0.0)
           + np.where(r3, laf[:, 4][:, None] * u**5.5, 0.0)).sum(0)
    d1 = on & (u < 1 + Z1_DLA)
    d2 = on & (u >= 1 + Z1_DLA)
    tau += (np.where(d1, dla[:, 2][:, None] * u**2, 0.0)
            + np.where(d2, dla[:, 3][:, None] * u**3, 0.0)).sum(0)
    x = lobs / LAM_L
    lc = lobs < LAM_L * zp1
    with np.errstate(all="ignore"):
        if z < Z1_LAF:
            laf_lc = 0.3248 * (x**1.2 - zp1**-0.9 * x**2.1)
        elif z < Z2_LAF:
            laf_lc = np.where(x >= 1 + Z1_LAF,
                              2.545e-2 * (zp1**1.6 * x**2.1 - x**3.7),
                              2.545e-2 * zp1**1.6 * x**2.1 + 0.3248 * x**1.2 - 0.2496 * x**2.1)
        else:
            laf_lc = np.where(
                x > 1 + Z2_LAF, 5.221e-4 * (zp1**3.4 * x**2.1 - x**5.5),
                np.where((x >= 1 + Z1_LAF) & (x < 1 + Z2_LAF),
                         5.221e-4 * zp1**3.4 * x**2.1 + 0.2182 * x**2.1 - 2.545e-2 * x**3.7,
                         np.where(x < 1 + Z1_LAF,
                                  5.221e-4 * zp1**3.4 * x**2.1 + 0.3248 * x**1.2 - 3.140e-2 * x**2.1,
                                  0.0)))
        if z < Z1_DLA:
            dla_lc = 0.2113 * zp1**2 - 0.07661 * zp1**2.3 * x**-0.3 - 0.1347 * x**2
        else:
            dla_lc = np.where(
                x >= 1 + Z1_DLA,
                0.04696 * zp1**3 - 0.01779 * zp1**3.3 * x**-0.3 - 0.02916 * x**3,
                0.6340 + 0.04696 * zp1**3 - 0.01779 * zp1**3.3 * x**-0.3 - 0.1347 * x**2
                - 0.2905 * x**-0.3)
    tau += np.where(lc, laf_lc + dla_lc, 0.0)
    return np.exp(-tau)


class Inoue14:
    """Name-compatible holder (the reference passes the class itself as ``igm=Inoue14``)."""

    name = "Inoue14"

    @staticmethod
    def get_transmission(redshift, lam_obs):
        return transmission(redshift, lam_obs)


def device_tables(lam, laf=INOUE14_LAF, dla=INOUE14_DLA):
    """Tables consumed by the weights/IGM kernel.

    With x = lam_i (1+z) / LAM_L every term of tau is ``coef * x**p`` (times a
    power of (1+z) for the Lyman-continuum part), and ``x**p`` factorises into
    ``(lam_i/LAM_L)**p * (1+z)**p``.  Lines are sorted by decreasing wavelength,
    so "lines with lam_j > lam_i" and "lines in regime k" are index prefixes and
    per-regime sums become differences of prefix sums.

    Returns a dict of float64/int32 arrays:
      n_blue            number of bins with lam_i < lam_Lyalpha (others have T=1)
      bin_pow[8, n_blue] (lam_i/LAM_L)**p for p in BIN_POWERS
      nline[n_blue]     J_i = #{j : lam_j > lam_i}
      lc_on[n_blue]     1 where lam_i < LAM_L (Lyman continuum applies)
      thr[3, 64]        regime thresholds c*lam_j/LAM_L for c in (2.2, 5.7, 3.0), padded with -1
      pre[5, 40]        prefix sums over lines of A_jk (LAM_L/lam_j)**p_k for
                        (LAF1,1.2) (LAF2,3.7) (LAF3,5.5) (DLA1,2) (DLA2,3)
    """
    lam = np.asarray(lam, dtype=float)
    lj = laf[:, 1]
    assert np.all(np.diff(lj) < 0) and np.allclose(lj, dla[:, 1])
    nl = len(lj)
    n_blue = int(np.searchsorted(lam, lj[0], side="left"))  # lam_i < 1215.67
    lb = lam[:n_blue]
    bin_pow = np.stack([(lb / LAM_L) ** p for p in BIN_POWERS]) if n_blue else np.zeros((len(BIN_POWERS), 0))
    nline = (lj[None, :] > lb[:, None]).sum(1).astype(np.int32)
    lc_on = (lb < LAM_L).astype(np.int32)
    thr = -np.ones((3, 64))
    for r, c in enumerate((1 + Z1_LAF, 1 + Z2_LAF, 1 + Z1_DLA)):
        thr[r, :nl] = c * lj / LAM_L
    pre = np.zeros((5, nl + 1))
    for r, (coef, p) in enumerate(((laf[:, 2], 1.2), (laf[:, 3], 3.7), (laf[:, 4], 5.5),
                                   (dla[:, 2], 2.0), (dla[:, 3], 3.0))):
        pre[r, 1:] = np.cumsum(coef * (LAM_L / lj) ** p)
    return dict(n_blue=n_blue, bin_pow=np.ascontiguousarray(bin_pow), nline=nline, lc_on=lc_on,
                thr=thr, pre=pre, n_lines=nl)
