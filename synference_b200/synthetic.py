"""Synthetic SPS grid and filter set of the same *shape* as the reference's inputs.

BPASS/Cloudy grids and SVO filter curves are downloaded at test time by the
reference (``tests/conftest.py:56-87``) and cannot be fetched offline, so parity
and benchmarks run on these seeded stand-ins: a 51 x 13 (age, Z)
grid with ``incident / transmitted / nebular_continuum / linecont`` components
and smooth-edged top-hat filters at the real JWST pivot wavelengths and widths.
They are *inputs*: both the CUDA path and the oracle receive the arrays.
"""

from __future__ import annotations

import numpy as np

from .parametric import Filter, FilterCollection, Grid
from .units import Angstrom, Quantity

BPASS_LOG10AGES = np.round(np.arange(6.0, 11.0 + 1e-9, 0.1), 1)
BPASS_METALLICITIES = np.array([1e-5, 1e-4, 1e-3, 2e-3, 3e-3, 4e-3, 6e-3, 8e-3, 1e-2, 1.4e-2,
                                2e-2, 3e-2, 4e-2])

# pivot wavelength [um], bandwidth [um] of the named JWST bands
JWST_BANDS = {
    "JWST/NIRCam.F070W": (0.704, 0.128), "JWST/NIRCam.F090W": (0.901, 0.194),
    "JWST/NIRCam.F115W": (1.154, 0.225), "JWST/NIRCam.F150W": (1.501, 0.318),
    "JWST/NIRCam.F200W": (1.990, 0.461), "JWST/NIRCam.F277W": (2.786, 0.672),
    "JWST/NIRCam.F356W": (3.563, 0.787), "JWST/NIRCam.F444W": (4.421, 1.024),
    "JWST/NIRCam.F300M": (2.996, 0.318), "JWST/NIRCam.F335M": (3.365, 0.347),
    "JWST/NIRCam.F410M": (4.092, 0.436),
    "JWST/MIRI.F560W": (5.6, 1.2), "JWST/MIRI.F770W": (7.7, 2.2), "JWST/MIRI.F1000W": (10.0, 2.0),
    "JWST/MIRI.F1130W": (11.3, 0.7), "JWST/MIRI.F1280W": (12.8, 2.4), "JWST/MIRI.F1500W": (15.0, 3.0),
    "JWST/MIRI.F1800W": (18.0, 3.0), "JWST/MIRI.F2100W": (21.0, 5.0), "JWST/MIRI.F2550W": (25.5, 4.0),
}
NIRCAM_WIDE8 = [f"JWST/NIRCam.{b}" for b in
                ("F070W", "F090W", "F115W", "F150W", "F200W", "F277W", "F356W", "F444W")]
NIRCAM_MIRI20 = NIRCAM_WIDE8 + [f"JWST/NIRCam.{b}" for b in ("F300M", "F335M", "F410M")] + \
    [f"JWST/MIRI.{b}" for b in ("F560W", "F770W", "F1000W", "F1130W", "F1280W", "F1500W",
                                "F1800W", "F2100W", "F2550W")]


def synthetic_filter(code: str, n_native: int = 1000) -> Filter:
    """tanh-edged top-hat on its own 1000-point axis, truncated at T ~ 3e-4 like an SVO table."""
    if code not in JWST_BANDS:
        raise ValueError(f"No synthetic curve for filter '{code}' (offline: SVO is unreachable)")
    piv, bw = JWST_BANDS[code]
    lam = np.linspace(piv - 0.56 * bw, piv + 0.56 * bw, n_native) * 1.0e4
    edge = 0.015 * bw * 1.0e4
    lo, hi = (piv - 0.5 * bw) * 1.0e4, (piv + 0.5 * bw) * 1.0e4
    t = 0.25 * (1.0 + np.tanh((lam - lo) / edge)) * (1.0 - np.tanh((lam - hi) / edge))
    # mild wavelength dependence so the curve is not symmetric
    t *= 0.55 + 0.35 * (lam - lam[0]) / (lam[-1] - lam[0])
    return Filter(code, lam, t)


def synthetic_filters(codes, new_lam=None) -> FilterCollection:
    return FilterCollection(filters=[synthetic_filter(c) for c in codes], new_lam=new_lam)


def synthetic_grid(lam, log10ages=BPASS_LOG10AGES, metallicities=BPASS_METALLICITIES, seed=42,
                   grid_name="synthetic-bpass-shaped") -> Grid:
    """Smooth positive stellar continua with breaks, nebular continuum and single-bin lines."""
    lam = np.asarray(getattr(lam, "value", lam), dtype=float)
    rng = np.random.default_rng(seed)
    na, nz = len(log10ages), len(metallicities)
    age = 10.0 ** np.asarray(log10ages)[:, None, None]
    zmet = np.asarray(metallicities)[None, :, None]
    lm = lam[None, None, :]

    teff = np.clip(4.5e4 * (age / 1.0e6) ** -0.27 * (zmet / 0.02) ** -0.05, 3.2e3, 5.5e4)
    x = 1.43877688e8 / (lm * teff)  # hc / (lambda k T), lambda in Angstrom
    planck = x**3 / np.expm1(np.minimum(x, 600.0))
    planck /= planck.max(axis=-1, keepdims=True)
    lbol = 2.0e21 * (age / 1.0e6) ** -0.85  # erg/s/Hz per Msun at the spectral peak
    balmer = 1.0 / (1.0 + 0.9 / (1.0 + np.exp(-(np.log10(age) - 8.4) / 0.25)) * (lm < 3646.0))
    blanket = np.where(lm < 3000.0, (zmet / 0.02) ** (-0.12 * np.clip(3000.0 / lm - 1.0, 0.0, 4.0)), 1.0)
    lyc = np.where(lm < 911.8, 0.4 * np.exp(-(np.log10(age) - 6.0) * 3.0) * (lm / 911.8) ** 2 + 1e-9, 1.0)
    jitter = rng.uniform(0.5, 1.5, size=(na, nz, 1))
    incident = lbol * planck * balmer * blanket * lyc * jitter

    transmitted = incident * np.where(lm < 911.8, 1.0e-6, 1.0)
    # ionising output proxy drives the nebular components
    q_ion = np.where(lm < 911.8, incident, 0.0).sum(-1, keepdims=True) * 1.0e-2
    neb_shape = np.where(lm > 911.8, (np.minimum(lm, 3646.0) / 3646.0) ** 1.5 *
                         np.where(lm > 3646.0, 0.35 * (3646.0 / lm) ** 0.3, 1.0), 0.0)
    nebular_continuum = 0.6 * q_ion * neb_shape * rng.uniform(0.5, 1.5, size=(na, nz, 1))
    linecont = np.zeros_like(incident)
    for lam0, strength in ((1215.67, 40.0), (3727.0, 6.0), (4861.3, 5.0), (5006.8, 12.0),
                           (6562.8, 15.0), (18751.0, 2.0)):
        if lam[0] < lam0 < lam[-1]:
            j = int(np.argmin(np.abs(lam - lam0)))
            linecont[..., j] += (strength * q_ion * (zmet / 0.02) ** 0.2
                                 * rng.uniform(0.5, 1.5, size=(na, nz, 1)))[..., 0]
    spectra = dict(incident=incident, transmitted=transmitted,
                   nebular_continuum=nebular_continuum, linecont=linecont)
    return Grid(grid_name, grid_dir="<synthetic>", log10ages=np.asarray(log10ages),
                metallicity=np.asarray(metallicities), lam=Quantity(lam, Angstrom), spectra=spectra)
