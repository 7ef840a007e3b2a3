"""Accuracy of the weight builder's table-driven float64 functions, through their numpy restatements
(synference_b200/fastmath.py ref_* mirror the device code in prep_kernel.cuh operation by operation)."""

import numpy as np
from scipy import special

from synference_b200 import fastmath as F


def test_log_exp_tail_tables_accuracy():
    t = F.build_tables()
    assert t["log_tab"].shape == (256, 2) and t["exp_tab"].shape == (64,) and t["tail_tab"].shape == (t["tail_n"], 8)
    rng = np.random.default_rng(0)
    x = np.exp(rng.uniform(np.log(1e-300), np.log(1e300), 400000))
    assert np.max(np.abs(F.ref_log(x, t) - np.log(x))) < 2e-13
    x = np.exp(rng.uniform(np.log(1e2), np.log(2e10), 400000))        # lookback times in years
    assert np.max(np.abs(F.ref_log(x, t) - np.log(x))) < 8e-15
    y = -np.exp(rng.uniform(np.log(1e-8), np.log(700.0), 400000))
    assert np.max(np.abs(F.ref_exp(y, t) / np.exp(y) - 1)) < 5e-14
    assert np.all(F.ref_exp(np.array([-760.0, -1e4]), t) == 0.0) and F.ref_exp(np.array([0.0]), t)[0] == 1.0
    u = np.concatenate([rng.uniform(0, 37.0, 400000), np.linspace(0, 37.0, 100001), np.arange(0, 148) * 0.25])
    q = 0.5 * special.erfc(u / np.sqrt(2.0))
    assert np.max(np.abs(F.ref_tail(u, t) / q - 1)) < 2e-12
    assert np.all(F.ref_tail(np.array([40.0, 55.0]), t) == 0.0)


def test_cody_waite_constants():
    assert F.LN2_64_HI + F.LN2_64_LO == np.log(2.0) / 64.0
    # hi has few enough significant bits that k * hi is exact for every k the kernel can form (|k| < 2**17)
    assert float(np.ldexp(F.LN2_64_HI, 40)).is_integer()
    assert F.LN2_64_HI == 0.010830424695996044 and abs(F.LN2_64_LO - 2.531013593154441e-13) < 1e-27   # literals in prep_kernel.cuh
