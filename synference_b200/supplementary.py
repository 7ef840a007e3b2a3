"""Supplementary parameters as by-products of the SFZH weights.

The reference attaches per-galaxy callbacks to the Synthesizer pipeline (``library.py:2579-2601``: every keyword of
``create_mock_library(..., name=func | (func, *args))`` becomes ``pipeline.add_analysis_func(func, "supp_<name>", *args)``)
and collects their results into ``Grid/SupplementaryParameters``.  The callbacks take Synthesizer ``Galaxy`` objects, which
the batched path never builds; the ones that only read the star-formation / metal-enrichment history are provided here under
the reference's own names and evaluated for a whole batch from the weight builder's output:

====================================  =========================================================  ==================
name (``library.py``)                 definition                                                 units
====================================  =========================================================  ==================
``calculate_mass_weighted_age`` :238  sum(sf_hist * ages) / sum(sf_hist)                         Myr
``calculate_sfr`` :223                mass formed at ages in [0, timescale] / timescale          Msun/yr (scales)
``calculate_burstiness`` :427         SFR(10 Myr) / SFR(100 Myr)                                 dimensionless
``calculate_sfh_quantile`` :468       age by which a fraction of the mass has formed (as coded)  Myr | dimensionless
``calculate_surviving_mass`` :512     log10 sum(w * grid.stellar_fraction)                       log10_Msun (scales)
``calculate_muv`` :172                f_nu through the rest-frame 1500 +- 50 A top-hat            nJy (scales)
``calculate_MUV`` :199                L_nu through the same top-hat                              erg/s/Hz (scales)
====================================  =========================================================  ==================

The two UV ones read the synthesised spectrum of the emission key being processed (the reference returns one value per saved
spectrum; the library keeps the requested key's): the band average is reduced on the device, ``SynthEngine.rest_band_flux``.

``sf_hist`` is the mass per age bin of A2 (bin edges at the mid-points of the grid ages, first edge 0); within a bin the mass
is taken as uniform in age when a timescale cuts through it (pin: the third-party ``Stars.calculate_average_sfr`` is not
available to check against).  Callbacks built on third-party ``Sed`` measurements or lines (D4000, beta, equivalent widths)
are not provided and raise ``NotImplementedError`` when requested.
"""

import numpy as np

from .units import strip_units

__all__ = ["calculate_muv", "calculate_MUV", "calculate_mass_weighted_age", "calculate_sfr", "calculate_burstiness", "calculate_sfh_quantile",
           "calculate_surviving_mass", "evaluate"]


class _Context:
    """What a batch offers: ``sf_hist`` (N, n_age) and ``sfzh`` (N, n_age, n_z) in Msun at the base mass, ages [yr]."""

    def __init__(self, sfzh, log10ages, redshift, cosmo, band_flux=None):
        self.band_flux = band_flux        # callable (lam_lo, lam_hi) -> (N,) observed-frame f_nu [nJy] in a rest-frame top-hat
        self.sfzh = sfzh
        self.sf_hist = sfzh.sum(axis=2)
        self.ages = 10.0 ** np.asarray(log10ages, dtype=np.float64)
        self.redshift = np.asarray(redshift, dtype=np.float64)
        self.cosmo = cosmo
        e = np.empty(self.ages.size + 1)
        e[0], e[1:-1], e[-1] = 0.0, 0.5 * (self.ages[1:] + self.ages[:-1]), self.ages[-1]
        self.edges = e            # A2: the last age bin is empty, its upper edge is never used

    def mass_younger_than(self, t_yr):
        lo, hi = self.edges[:-1], self.edges[1:]
        with np.errstate(divide="ignore", invalid="ignore"):
            frac = np.clip((float(t_yr) - lo) / (hi - lo), 0.0, 1.0)
        frac[~np.isfinite(frac)] = 0.0
        return self.sf_hist @ frac


def _batched(units):
    def deco(fn):
        fn._sb2_batched = True
        fn._sb2_units = units
        return fn
    return deco


MUV_BAND = (1450.0, 1550.0)      # tophats = {"MUV": {"lam_eff": 1500 A, "lam_fwhm": 100 A}}  (library.py:100-104)


@_batched("nJy")
def calculate_muv(ctx, cosmo=None):
    if ctx.band_flux is None:
        raise ValueError("calculate_muv needs the synthesised spectrum (run it through create_mock_library)")
    return ctx.band_flux(*MUV_BAND)


@_batched("erg/s/Hz")
def calculate_MUV(ctx, cosmo=None):
    """Rest-frame L_nu: the observed flux density with the distance factor of A7 taken out again."""
    c = cosmo or ctx.cosmo
    dl_cm = np.asarray(strip_units(c.luminosity_distance(ctx.redshift), "cm"), dtype=float)
    return calculate_muv(ctx) * 1e-32 * 4.0 * np.pi * dl_cm**2 / (1.0 + ctx.redshift)


@_batched("Myr")
def calculate_mass_weighted_age(ctx):
    return (ctx.sf_hist @ ctx.ages) / ctx.sf_hist.sum(axis=1) / 1e6


@_batched("Msun/yr")
def calculate_sfr(ctx, timescale=1e7):
    t = float(strip_units(timescale, "yr"))
    return ctx.mass_younger_than(t) / t


@_batched("dimensionless")
def calculate_burstiness(ctx):
    with np.errstate(divide="ignore", invalid="ignore"):
        return (ctx.mass_younger_than(1e7) / 1e7) / (ctx.mass_younger_than(1e8) / 1e8)


def calculate_sfh_quantile(ctx, quantile=0.5, norm=False, cosmo=None):
    """``library.py:468-509`` as written: cumulative mass from the OLDEST bin down, first bin at which it reaches
    ``quantile`` of the total, that bin's age in Myr (``norm``: as a fraction of the age of the universe at the redshift)."""
    assert 0 <= quantile <= 1, "quantile must be between 0 and 1."
    cum = np.cumsum(ctx.sf_hist[:, ::-1], axis=1)
    target = quantile * cum[:, -1]
    idx = np.array([np.searchsorted(c, t) for c, t in zip(cum, target)])
    look = ctx.ages[::-1][np.minimum(idx, ctx.ages.size - 1)] / 1e6
    if norm:
        look = look / np.asarray(strip_units((cosmo or ctx.cosmo).age(ctx.redshift), "Myr"), dtype=float)
    return look


calculate_sfh_quantile._sb2_batched = True
calculate_sfh_quantile._sb2_units = lambda quantile=0.5, norm=False, cosmo=None: "dimensionless" if norm else "Myr"


def calculate_surviving_mass(ctx, grid):
    frac = getattr(grid, "stellar_fraction", None)
    if frac is None:
        raise ValueError("calculate_surviving_mass needs grid.stellar_fraction (N_age, N_Z)")
    frac = np.asarray(frac, dtype=np.float64)
    if frac.shape != ctx.sfzh.shape[1:]:
        raise ValueError(f"grid.stellar_fraction has shape {frac.shape}, expected {ctx.sfzh.shape[1:]}")
    return np.log10(np.einsum("naz,az->n", ctx.sfzh, frac))


calculate_surviving_mass._sb2_batched = True
calculate_surviving_mass._sb2_units = "log10_Msun"


def scales_with_mass(units: str) -> str:
    """How create_full_library rescales a supplementary column from the base mass to the galaxy's mass
    (``utils.py:929-988`` check_scaling / check_log_scaling on the unit string): 'linear', 'log' or 'none'."""
    if "log" in units:
        return "log"
    if any(k in units for k in ("Msun", "Jy", "erg")):        # mass, mass / time, flux densities, luminosities
        return "linear"
    return "none"


def split(spec):
    """``func`` or ``(func, *args)`` as the reference accepts them (``library.py:2593-2601``)."""
    if isinstance(spec, (tuple, list)):
        return spec[0], tuple(spec[1:])
    return spec, ()


def check_supported(extra_analysis_functions):
    for name, spec in extra_analysis_functions.items():
        fn, _ = split(spec)
        if not getattr(fn, "_sb2_batched", False):
            raise NotImplementedError(
                f"supplementary analysis '{name}': {getattr(fn, '__name__', fn)!r} operates on Synthesizer objects; the "
                "batched path evaluates the history-based ones of synference_b200.supplementary (same names as the "
                "reference) and has no spectrum / line callbacks yet")


def evaluate(extra_analysis_functions, sfzh, log10ages, redshift, cosmo, band_flux=None):
    """``{name: (values (N,), unit string)}`` for one batch (unit strings as unyt would print them)."""
    ctx = _Context(sfzh, log10ages, redshift, cosmo, band_flux)
    out = {}
    for name, spec in extra_analysis_functions.items():
        fn, args = split(spec)
        units = fn._sb2_units(*args) if callable(fn._sb2_units) else fn._sb2_units
        out[name] = (np.asarray(fn(ctx, *args), dtype=np.float64), str(units))
    return out
