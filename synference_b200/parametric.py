"""Parameter-holder stand-ins for the Synthesizer objects the reference passes around.

synference builds one ``SFH.*`` / ``ZDist.*`` / ``Stars`` / ``Galaxy`` object per
galaxy and hands lists of them to the third-party ``synthesizer`` package
(``library.py:1372-1379``, ``library.py:1422``; SURVEY 8c lists every call site).
That package is not available here and, more to the point, per-galaxy Python
objects are exactly what makes the reference path slow.  These classes keep the
reference's spelling (``SFH.LogNormal(tau=..., peak_age=..., max_age=...)``,
``ZDist.DeltaConstant(log10metallicity=...)``, ``Grid``, ``FilterCollection``,
``Instrument``, ``PacmanEmission`` ...) but are only *descriptions*: they are
lowered once to a struct-of-arrays parameter block (``pack_sfh`` / ``pack_zdist``)
which the CUDA weight builder consumes.  Array containers (``SFHArray``,
``ZDistArray``) hold a whole population without creating per-galaxy objects.

Functional forms follow SURVEY Appendix A2/A3/A5/A6 (the pin list).
"""

from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from .units import Angstrom, Myr, Quantity, has_units, strip_units, yr

# --------------------------------------------------------------------------
# Star formation histories
# --------------------------------------------------------------------------

SFH_CONSTANT, SFH_GAUSSIAN, SFH_EXPONENTIAL, SFH_DECLINING_EXP, SFH_DELAYED_EXP, \
    SFH_LOGNORMAL, SFH_DOUBLE_POWERLAW, SFH_CONTINUITY = range(8)
SFH_MAX_PARAMS = 24  # device row width: [min_age, max_age, p0, p1, ...]


def _yr(x):
    """Lookback ages are handled in years; bare numbers are taken to be years."""
    return strip_units(x, "yr") if has_units(x) else np.asarray(x, dtype=float)


class _SFHCommon:
    """Base of the SFH holders (``synthesizer.parametric.SFH.Common`` spelling)."""

    type_id = -1
    param_names: Sequence[str] = ()
    time_params: Sequence[str] = ()  # parameters that are ages (converted to yr)

    def __init__(self, max_age=None, min_age=0.0, **params):
        self.name = type(self).__name__
        self.parameters = dict(params)
        self.parameters["max_age"] = max_age
        self.parameters["min_age"] = min_age
        self.max_age = float(_yr(max_age))
        self.min_age = float(_yr(min_age))
        for k, v in params.items():
            setattr(self, k, float(_yr(v)) if k in self.time_params else float(strip_units(v)))

    def param_row(self) -> np.ndarray:
        row = np.zeros(SFH_MAX_PARAMS)
        row[0], row[1] = self.min_age, self.max_age
        for i, k in enumerate(self.param_names):
            row[2 + i] = getattr(self, k)
        return row

    # pointwise SFR(lookback age [yr]); used by the quad cross-check and plots only
    def _sfr(self, age):
        raise NotImplementedError

    def get_sfr(self, age):
        age = np.asarray(age, dtype=float)
        inside = (age >= self.min_age) & (age < self.max_age)
        with np.errstate(all="ignore"):
            return np.where(inside, self._sfr(np.where(inside, age, self.min_age)), 0.0)

    def __repr__(self):
        return f"SFH.{self.name}({self.parameters})"


class _Constant(_SFHCommon):
    type_id = SFH_CONSTANT

    def __init__(self, max_age=100 * yr, min_age=0.0, duration=None):
        if duration is not None:
            max_age = Quantity(float(_yr(min_age)) + float(_yr(duration)), yr)
        super().__init__(max_age=max_age, min_age=min_age)

    def _sfr(self, age):
        return np.ones_like(age)

    def get_sfr(self, age):  # closed on both ends, as in the reference form
        age = np.asarray(age, dtype=float)
        return ((age >= self.min_age) & (age <= self.max_age)).astype(float)


class _Gaussian(_SFHCommon):
    type_id = SFH_GAUSSIAN
    param_names = ("peak_age", "sigma")
    time_params = ("peak_age", "sigma")

    def __init__(self, peak_age, sigma, max_age, min_age=0.0):
        super().__init__(max_age=max_age, min_age=min_age, peak_age=peak_age, sigma=sigma)

    def _sfr(self, age):
        return np.exp(-0.5 * ((age - self.peak_age) / self.sigma) ** 2)


class _Exponential(_SFHCommon):
    type_id = SFH_EXPONENTIAL
    param_names = ("tau",)
    time_params = ("tau",)

    def __init__(self, tau, max_age, min_age=0.0):
        super().__init__(max_age=max_age, min_age=min_age, tau=tau)

    def _sfr(self, age):
        return np.exp((self.max_age - age) / self.tau)


class _DecliningExponential(_Exponential):
    type_id = SFH_DECLINING_EXP

    def _sfr(self, age):
        return np.exp(-(self.max_age - age) / self.tau)


class _DelayedExponential(_Exponential):
    type_id = SFH_DELAYED_EXP

    def _sfr(self, age):
        t = self.max_age - age
        return t * np.exp(-t / self.tau)


class _LogNormal(_SFHCommon):
    type_id = SFH_LOGNORMAL
    param_names = ("tau", "peak_age")
    time_params = ("peak_age",)

    def __init__(self, tau, peak_age, max_age, min_age=0.0):
        super().__init__(max_age=max_age, min_age=min_age, tau=tau, peak_age=peak_age)
        with np.errstate(all="ignore"):
            self.tpeak = self.max_age - self.peak_age
            self.t_0 = np.log(self.tpeak) + self.tau**2

    def _sfr(self, age):
        t = self.max_age - age
        return (1.0 / t) * np.exp(-((np.log(t) - self.t_0) ** 2) / 2.0 / self.tau**2)


class _DoublePowerLaw(_SFHCommon):
    type_id = SFH_DOUBLE_POWERLAW
    param_names = ("peak_age", "alpha", "beta")
    time_params = ("peak_age",)

    def __init__(self, peak_age, alpha, beta, max_age, min_age=0.0):
        super().__init__(max_age=max_age, min_age=min_age, peak_age=peak_age, alpha=alpha, beta=beta)

    def _sfr(self, age):
        x = age / self.peak_age
        return 1.0 / (x**self.alpha + x**self.beta)


class _Continuity(_SFHCommon):
    """Piecewise-constant SFR in lookback-age bins with log ratios between adjacent bins.

    ``agebins`` is ``(N_b, 2)`` in log10(yr) (Prospector convention, as built by
    ``continuity_agebins`` in ``final_library_generation_multinode.py:193-259``);
    ``log10(SFR_j / SFR_{j+1}) = logsfr_ratios[j]``.
    """

    type_id = SFH_CONTINUITY

    def __init__(self, logsfr_ratios, agebins, max_age=None, min_age=None):
        agebins = np.asarray(strip_units(agebins), dtype=float)
        self.edges = np.concatenate([10.0 ** agebins[:, 0], [10.0 ** agebins[-1, 1]]])
        if self.edges[0] <= 1.0:  # log10 age 0 means "from today"
            self.edges[0] = 0.0
        self.logsfr_ratios = np.asarray(strip_units(logsfr_ratios), dtype=float)
        nb = len(self.edges) - 1
        assert len(self.logsfr_ratios) == nb - 1, "need N_b - 1 log SFR ratios"
        assert 2 + (nb + 1) + (nb - 1) + 1 <= SFH_MAX_PARAMS, "too many continuity bins"
        self.name = "Continuity"
        self.parameters = {"logsfr_ratios": self.logsfr_ratios, "agebins": agebins}
        self.min_age, self.max_age = float(self.edges[0]), float(self.edges[-1])
        self.parameters["max_age"] = Quantity(self.max_age, yr)

    def bin_sfr(self):
        return 10.0 ** (-np.concatenate([[0.0], np.cumsum(self.logsfr_ratios)]))

    def param_row(self):
        row = np.zeros(SFH_MAX_PARAMS)
        nb = len(self.edges) - 1
        row[0], row[1], row[2] = self.min_age, self.max_age, nb
        row[3:3 + nb + 1] = self.edges
        row[3 + nb + 1:3 + 2 * nb] = self.logsfr_ratios
        return row

    def get_sfr(self, age):
        age = np.asarray(age, dtype=float)
        j = np.searchsorted(self.edges, age, side="right") - 1
        ok = (j >= 0) & (j < len(self.edges) - 1)
        return np.where(ok, self.bin_sfr()[np.clip(j, 0, len(self.edges) - 2)], 0.0)


class SFH:
    """Namespace mirroring ``synthesizer.parametric.SFH``."""

    Common = _SFHCommon
    Constant = _Constant
    Gaussian = _Gaussian
    Exponential = _Exponential
    DecliningExponential = _DecliningExponential
    DelayedExponential = _DelayedExponential
    LogNormal = _LogNormal
    DoublePowerLaw = _DoublePowerLaw
    Continuity = _Continuity


class SFHArray:
    """A whole population of one SFH type as a struct of arrays (no per-galaxy objects).

    Behaves like the object array ``generate_sfh_basis`` returns in the reference
    (``library.py:1334``): ``len()``, integer indexing (materialises one holder),
    slicing / boolean masks (another ``SFHArray``), ``.redshift`` per element.
    """

    def __init__(self, sfh_type, rows: np.ndarray, redshifts=None):
        self.sfh_type = sfh_type
        self.rows = np.ascontiguousarray(rows, dtype=np.float64)  # (N, SFH_MAX_PARAMS)
        self.redshifts = None if redshifts is None else np.asarray(redshifts, dtype=float)
        # leading columns that can be non-zero (min_age, max_age, the type's parameters); None: unknown, scan the rows
        names = getattr(sfh_type, "param_names", ())
        self.n_used = 2 + len(names) if names else None

    @property
    def type_id(self):
        return self.sfh_type.type_id

    @property
    def max_age(self):
        return self.rows[:, 1]

    def __len__(self):
        return self.rows.shape[0]

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            return self._materialise(int(idx))
        z = None if self.redshifts is None else self.redshifts[idx]
        return SFHArray(self.sfh_type, self.rows[idx], z)

    def __iter__(self):
        return (self._materialise(i) for i in range(len(self)))

    def _materialise(self, i):
        r = self.rows[i]
        t = self.sfh_type
        if t is _Continuity:
            nb = int(r[2])
            edges = r[3:3 + nb + 1]
            lo = np.log10(np.maximum(edges[:-1], 1.0))
            agebins = np.stack([lo, np.log10(edges[1:])], 1)
            obj = t(r[3 + nb + 1:3 + 2 * nb], agebins)
        elif t is _Constant:
            obj = t(max_age=Quantity(r[1], yr), min_age=Quantity(r[0], yr))
        else:
            kw = {k: (Quantity(r[2 + j], yr) if k in t.time_params else r[2 + j])
                  for j, k in enumerate(t.param_names)}
            obj = t(max_age=Quantity(r[1], yr), min_age=Quantity(r[0], yr), **kw)
        if self.redshifts is not None:
            obj.redshift = float(self.redshifts[i])
        return obj


def pack_sfh(sfhs) -> (int, np.ndarray):
    """Lower SFH holders to ``(type_id, rows[N, SFH_MAX_PARAMS])``."""
    if isinstance(sfhs, SFHArray):
        return sfhs.type_id, sfhs.rows
    if isinstance(sfhs, _SFHCommon):
        sfhs = [sfhs]
    sfhs = list(sfhs)
    tid = {s.type_id for s in sfhs}
    if len(tid) != 1:
        raise ValueError("All SFHs in one basis must be of the same type for the batched path")
    return tid.pop(), np.stack([s.param_row() for s in sfhs])


# --------------------------------------------------------------------------
# Metallicity distributions
# --------------------------------------------------------------------------

ZD_DELTA_LINEAR, ZD_DELTA_LOG10, ZD_NORMAL_LINEAR, ZD_NORMAL_LOG10 = range(4)


class _ZDistCommon:
    __slots__ = ()
    type_id = -1

    def __repr__(self):
        return f"ZDist.{type(self).__name__}({self.parameters})"


_PLAIN_NUMBERS = (float, int, np.float64, np.float32)


class _DeltaConstant(_ZDistCommon):
    """All mass at one metallicity (shared between the two bracketing grid points).

    The README flow builds one of these per galaxy in a list comprehension (README.md:113-114): the object holds two slots
    and derives the rest, so that a million of them cost what the comprehension itself costs."""

    __slots__ = ("type_id", "value")
    name = "DeltaConstant"
    sigma = 0.0

    def __init__(self, metallicity=None, log10metallicity=None):
        if metallicity is None:
            if log10metallicity is None:
                raise ValueError("Give exactly one of metallicity / log10metallicity")
            self.type_id, v = ZD_DELTA_LOG10, log10metallicity
        else:
            if log10metallicity is not None:
                raise ValueError("Give exactly one of metallicity / log10metallicity")
            self.type_id, v = ZD_DELTA_LINEAR, metallicity
        self.value = float(v) if type(v) in _PLAIN_NUMBERS else float(strip_units(v))

    @property
    def parameters(self):
        return {("metallicity" if self.type_id == ZD_DELTA_LINEAR else "log10metallicity"): self.value}

    def get_metallicity(self):
        return self.value if self.type_id == ZD_DELTA_LINEAR else 10.0**self.value


class _Normal(_ZDistCommon):
    """Gaussian in Z (or log10 Z) evaluated at the grid metallicities, normalised."""

    __slots__ = ("name", "type_id", "value", "sigma", "parameters")

    def __init__(self, mean, sigma, log10=True):
        self.name = "Normal"
        self.type_id = ZD_NORMAL_LOG10 if log10 else ZD_NORMAL_LINEAR
        self.value, self.sigma = float(strip_units(mean)), float(strip_units(sigma))
        self.parameters = {"mean": self.value, "sigma": self.sigma}


class ZDist:
    """Namespace mirroring ``synthesizer.parametric.ZDist``."""

    Common = _ZDistCommon
    DeltaConstant = _DeltaConstant
    Normal = _Normal


class ZDistArray:
    """A population of metallicity distributions as arrays."""

    def __init__(self, type_id: int, value, sigma=None):
        self.type_id = int(type_id)
        self.value = np.asarray(strip_units(value), dtype=np.float64)
        self.sigma = np.zeros_like(self.value) if sigma is None else \
            np.broadcast_to(np.asarray(strip_units(sigma), dtype=np.float64), self.value.shape).copy()

    @classmethod
    def delta(cls, metallicity=None, log10metallicity=None):
        if metallicity is not None:
            return cls(ZD_DELTA_LINEAR, metallicity)
        return cls(ZD_DELTA_LOG10, log10metallicity)

    @classmethod
    def normal(cls, mean, sigma, log10=True):
        return cls(ZD_NORMAL_LOG10 if log10 else ZD_NORMAL_LINEAR, mean, sigma)

    def __len__(self):
        return self.value.shape[0]

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            if self.type_id == ZD_DELTA_LINEAR:
                return _DeltaConstant(metallicity=self.value[idx])
            if self.type_id == ZD_DELTA_LOG10:
                return _DeltaConstant(log10metallicity=self.value[idx])
            return _Normal(self.value[idx], self.sigma[idx], log10=self.type_id == ZD_NORMAL_LOG10)
        return ZDistArray(self.type_id, self.value[idx], self.sigma[idx])


def pack_zdist(zd) -> (int, np.ndarray, np.ndarray):
    if isinstance(zd, ZDistArray):
        return zd.type_id, zd.value, zd.sigma
    if isinstance(zd, _ZDistCommon):
        zd = [zd]
    zd = list(zd)
    n = len(zd)
    tids = np.fromiter((d.type_id for d in zd), dtype=np.int64, count=n)
    if n == 0 or tids.min() != tids.max():
        raise ValueError("All metallicity distributions in one basis must share a type")
    value = np.fromiter((d.value for d in zd), dtype=np.float64, count=n)
    sigma = np.zeros(n) if int(tids[0]) in (ZD_DELTA_LINEAR, ZD_DELTA_LOG10) else \
        np.fromiter((d.sigma for d in zd), dtype=np.float64, count=n)
    return int(tids[0]), value, sigma


# --------------------------------------------------------------------------
# Dust curves
# --------------------------------------------------------------------------

class PowerLaw:
    def __init__(self, slope=-1.0):
        self.slope = float(slope)
        self.name = "PowerLaw"
        self.params = {"slope": self.slope}

    def get_tau(self, lam):
        lam = strip_units(lam, "Angstrom")
        return (lam / 5500.0) ** self.slope


class Calzetti2000:
    """Calzetti (2000) with the Noll+09 slope/bump modification.

    ``slope`` and/or ``ampl`` may name a per-galaxy emitter attribute instead of being numbers
    (``Calzetti2000(slope="slope", ampl="dust_bump_amplitude")``,
    ``final_library_generation_multinode.py:496``): the curve is then
    ``(K0 + ampl_g * D0) * (lam / 0.55um)**slope_g`` with the per-galaxy values supplied at run time
    (:meth:`components`).
    """

    def __init__(self, slope=0.0, cent_lam=0.2175, ampl=0.0, gamma=0.035):
        self.slope_name = slope if isinstance(slope, str) else None
        self.ampl_name = ampl if isinstance(ampl, str) else None
        self.slope = 0.0 if self.slope_name else float(strip_units(slope))
        self.ampl = 0.0 if self.ampl_name else float(strip_units(ampl))
        self.cent_lam = float(strip_units(cent_lam, "um")) if has_units(cent_lam) else float(cent_lam)
        self.gamma = float(strip_units(gamma, "um")) if has_units(gamma) else float(gamma)
        self.name = "Calzetti2000"
        self.params = {"slope": self.slope, "cent_lam": self.cent_lam, "ampl": self.ampl,
                       "gamma": self.gamma}

    @property
    def per_galaxy(self):
        return self.slope_name is not None or self.ampl_name is not None

    @staticmethod
    def _k(x):
        x = np.asarray(x, dtype=float)
        blue = -2.156 + 1.509 / x - 0.198 / x**2 + 0.011 / x**3
        red = -1.857 + 1.040 / x
        return 4.05 + 2.659 * np.where(x < 0.63, blue, red)

    @staticmethod
    def _interp_extrap(lam_um, x, helper):
        # linear interpolation with linear extrapolation beyond the helper range
        y = np.interp(lam_um, x, helper)
        lo = lam_um < x[0]
        hi = lam_um > x[-1]
        y = np.where(lo, helper[0] + (lam_um - x[0]) * (helper[1] - helper[0]) / (x[1] - x[0]), y)
        return np.where(hi, helper[-1] + (lam_um - x[-1]) * (helper[-1] - helper[-2]) / (x[-1] - x[-2]), y)

    def components(self, lam):
        """``(K0, D0, L2)``: curve at slope = 0, ampl = 0; bump profile per unit amplitude (both / k(0.55 um), on the
        helper grid, interpolated -- the helper curve is linear in the amplitude); log2(lam / 0.55 um)."""
        lam_um = strip_units(lam, "Angstrom") * 1.0e-4
        x = np.arange(0.12, 2.2, 0.001)
        k55 = self._k(0.55)
        bump1 = (x * self.gamma) ** 2 / ((x**2 - self.cent_lam**2) ** 2 + (x * self.gamma) ** 2)
        return (self._interp_extrap(lam_um, x, self._k(x) / k55), self._interp_extrap(lam_um, x, bump1 / k55),
                np.log2(lam_um / 0.55))

    def get_tau(self, lam, slope=None, ampl=None):
        """tau(lam)/tau_V for the global parameters, or for explicit ``slope`` / ``ampl`` scalars."""
        lam_um = strip_units(lam, "Angstrom") * 1.0e-4
        slope = self.slope if slope is None else float(slope)
        ampl = self.ampl if ampl is None else float(ampl)
        x = np.arange(0.12, 2.2, 0.001)
        k = self._k(x)
        bump = ampl * (x * self.gamma) ** 2 / ((x**2 - self.cent_lam**2) ** 2 + (x * self.gamma) ** 2)
        helper = (k + bump) / self._k(0.55)
        return self._interp_extrap(lam_um, x, helper) * (lam_um / 0.55) ** slope


# --------------------------------------------------------------------------
# SPS grid
# --------------------------------------------------------------------------

def rebin_flux_conserving(lam_old, spec, lam_new):
    """Flux-conserving rebin of ``spec[..., N_old]`` onto ``lam_new`` (spectres-like).

    Bin edges are midpoints; each new bin receives the mean of the old piecewise-constant
    spectrum over its extent (what ``Grid(..., new_lam=)`` does).
    """
    lam_old, lam_new = np.asarray(lam_old, float), np.asarray(lam_new, float)

    def edges(l):
        mid = 0.5 * (l[1:] + l[:-1])
        return np.concatenate([[l[0] - (mid[0] - l[0])], mid, [l[-1] + (l[-1] - mid[-1])]])

    eo, en = edges(lam_old), edges(lam_new)
    cum = np.concatenate([np.zeros(spec.shape[:-1] + (1,)), np.cumsum(spec * np.diff(eo), axis=-1)], -1)
    xi = np.clip(en, eo[0], eo[-1])
    j = np.clip(np.searchsorted(eo, xi, side="right") - 1, 0, len(lam_old) - 1)
    cum_at = cum[..., j] + spec[..., j] * (xi - eo[j])
    width = np.diff(xi)
    out = np.diff(cum_at, axis=-1) / np.where(width > 0, width, 1.0)
    return np.where(width > 0, out, 0.0)


class Grid:
    """SPS grid: ``log10ages``, ``metallicity``, ``lam`` and component spectra.

    ``spectra[name]`` is ``(N_age, N_Z, N_lam)`` in erg/s/Hz per Msun of initial
    mass.  Construct from arrays, or by name from an ``.npz`` (or, when
    ``h5py`` is importable, a Synthesizer HDF5 grid) in ``grid_dir``.
    """

    def __init__(self, grid_name="grid", grid_dir=None, new_lam=None, *, log10ages=None,
                 metallicity=None, lam=None, spectra: Optional[Dict[str, np.ndarray]] = None):
        self.grid_name = str(grid_name)
        self.grid_dir = grid_dir if grid_dir is not None else os.environ.get("SYNTHESIZER_GRID_DIR", ".")
        if spectra is None:
            log10ages, metallicity, lam, spectra = self._load()
        self.log10ages = np.asarray(log10ages, dtype=float)
        self.metallicity = np.asarray(metallicity, dtype=float)
        lam = strip_units(lam, "Angstrom")
        self.spectra = {k: np.asarray(v, dtype=float) for k, v in spectra.items()}
        if new_lam is not None:
            new_lam = strip_units(new_lam, "Angstrom")
            self.spectra = {k: rebin_flux_conserving(lam, v, new_lam) for k, v in self.spectra.items()}
            lam = new_lam
        self.lam = Quantity(lam, Angstrom)
        for k, v in self.spectra.items():
            assert v.shape == (self.log10ages.size, self.metallicity.size, lam.size), \
                f"component {k} has shape {v.shape}"

    # aliases used across the reference
    @property
    def log10age(self):
        return self.log10ages

    @property
    def metallicities(self):
        return self.metallicity

    @property
    def available_spectra(self):
        return list(self.spectra)

    def _load(self):
        base = os.path.join(self.grid_dir, self.grid_name)
        for cand in (base, base + ".npz"):
            if os.path.isfile(cand) and cand.endswith(".npz"):
                d = np.load(cand)
                spectra = {k[len("spectra/"):]: d[k] for k in d.files if k.startswith("spectra/")}
                return d["log10ages"], d["metallicity"], d["lam"], spectra
        for cand in (base, base + ".hdf5"):
            if os.path.isfile(cand) and cand.endswith(".hdf5"):
                import h5py  # noqa: F401  (only when the user has it)
                with h5py.File(cand, "r") as f:
                    spectra = {k: f["spectra"][k][()] for k in f["spectra"] if k != "wavelength"}
                    return (f["axes/log10ages"][()], f["axes/metallicities"][()],
                            f["spectra/wavelength"][()], spectra)
        raise FileNotFoundError(f"No grid '{self.grid_name}' (.npz/.hdf5) in {self.grid_dir}")

    def save(self, path):
        np.savez(path, log10ages=self.log10ages, metallicity=self.metallicity,
                 lam=np.asarray(self.lam), **{f"spectra/{k}": v for k, v in self.spectra.items()})


# --------------------------------------------------------------------------
# Filters / instrument
# --------------------------------------------------------------------------

class Filter:
    def __init__(self, filter_code, lam, transmission):
        self.filter_code = filter_code
        self.lam = strip_units(lam, "Angstrom")
        self.t = np.asarray(transmission, dtype=float)
        assert self.lam.shape == self.t.shape

    def pivwv(self):
        if getattr(self, "_pivwv", None) is None:       # curves are immutable once built; callers ask per simulate() call
            self._pivwv = np.sqrt(np.trapezoid(self.t * self.lam, self.lam) / np.trapezoid(self.t / self.lam, self.lam))
        return self._pivwv


class FilterCollection:
    """A set of transmission curves, optionally resampled on one shared wavelength axis."""

    def __init__(self, filter_codes: Optional[List[str]] = None, filters: Optional[List[Filter]] = None,
                 new_lam=None, filter_dir: Optional[str] = None):
        if filters is None:
            filters = [self._lookup(c, filter_dir) for c in (filter_codes or [])]
        self.filters = list(filters)
        self.filter_codes = [f.filter_code for f in self.filters]
        self.lam = None
        if new_lam is not None:
            self.resample_filters(new_lam=new_lam)

    @staticmethod
    def _lookup(code, filter_dir):
        if filter_dir is not None:
            path = os.path.join(filter_dir, code.replace("/", "_") + ".dat")
            if os.path.isfile(path):
                lam, t = np.loadtxt(path, unpack=True)
                return Filter(code, lam, t)
        from .synthetic import synthetic_filter  # SVO is unreachable offline
        return synthetic_filter(code)

    def resample_filters(self, new_lam):
        """Linear interpolation of every curve onto ``new_lam`` with 0 fill (``min_example.py:39``)."""
        new_lam = strip_units(new_lam, "Angstrom")
        self.filters = [Filter(f.filter_code, new_lam, np.interp(new_lam, f.lam, f.t, left=0.0, right=0.0))
                        for f in self.filters]
        self.lam = Quantity(new_lam, Angstrom)
        return self

    def get_non_zero_lam_lims(self):
        lo = min(f.lam[f.t > 0].min() for f in self.filters)
        hi = max(f.lam[f.t > 0].max() for f in self.filters)
        return Quantity(lo, Angstrom), Quantity(hi, Angstrom)

    @property
    def pivot_lams(self):
        return Quantity([f.pivwv() for f in self.filters], Angstrom)

    def __len__(self):
        return len(self.filters)

    def __iter__(self):
        return iter(self.filters)


class Instrument:
    def __init__(self, label, filters: Optional[FilterCollection] = None, **kwargs):
        self.label = label
        self.filters = filters

    @property
    def can_do_photometry(self):
        return self.filters is not None and len(self.filters) > 0


# --------------------------------------------------------------------------
# Dust emission generators (min_example.py:110, generate_library_basic.py:195)
# --------------------------------------------------------------------------
_H_OVER_K = 4.799243073366221e-11      # h / k_B  [K s]
_C_ANGSTROM = 2.99792458e18            # c [A / s]


class Greybody:
    """Optically thin modified blackbody ``nu^emissivity B_nu(T)``, normalised to unit integral over ALL frequencies
    (energy balance: the emission model scales it by the energy the dust absorbed).  ``shape(lam)`` is in 1/Hz."""

    def __init__(self, temperature, emissivity=1.5, **kwargs):
        self.temperature = float(strip_units(temperature))
        self.emissivity = float(strip_units(emissivity))
        if not self.temperature > 0:
            raise ValueError("dust temperature must be positive")

    def shape(self, lam):
        from scipy.special import gamma, zeta
        nu = _C_ANGSTROM / np.asarray(lam, dtype=np.float64)
        x = _H_OVER_K * nu / self.temperature
        p = 3.0 + self.emissivity
        with np.errstate(over="ignore", under="ignore"):
            f = x ** p / np.expm1(x)                       # -> 0 in the Wien tail (expm1 overflows to inf)
        return f / (gamma(p + 1.0) * zeta(p + 1.0)) * (_H_OVER_K / self.temperature)     # d nu = (kT/h) dx

    def same_as(self, other):
        return type(other) is type(self) and other.temperature == self.temperature and other.emissivity == self.emissivity


class Blackbody(Greybody):
    def __init__(self, temperature, **kwargs):
        super().__init__(temperature, emissivity=0.0)


# --------------------------------------------------------------------------
# Emission models
# --------------------------------------------------------------------------

LYA = 1215.67
KEYS = ("incident", "transmitted", "nebular", "reprocessed", "escaped", "intrinsic",
        "attenuated", "emergent", "total")


class EmissionModel:
    """Premade stellar emission-model tree reduced to what the batched path needs.

    ``recipe(key)`` returns two ``(N_age, N_Z, N_lam)`` float64 grids: the part of
    spectrum ``key`` that passes through the dust screen and the part that does
    not, already combined over grid components with the *global* ``fesc`` /
    ``fesc_ly_alpha``.  Per-galaxy ``tau_v`` is an emitter parameter supplied at run time.

    ``fesc`` may instead name a per-galaxy emitter attribute (``fesc="fesc"``, the string
    convention of the reference's emission models, SURVEY A5;
    ``docs/source/library_gen/complex_library_generation.ipynb:357,398``): the two grids are then
    built for fesc = 0 and fesc enters as per-galaxy coefficients ``(1 - fesc, fesc)`` on them
    (:meth:`coefficients`) -- the contraction kernel's two-component form
    ``c_att * dust(att) + c_un * un``.
    """

    label = "emission"
    available = KEYS

    def __init__(self, grid: Grid, fesc=0.0, fesc_ly_alpha=1.0, dust_curve=None, tau_v="tau_v",
                 dust_emission=None, **kwargs):
        self.grid = grid
        self.fesc = fesc
        self.fesc_ly_alpha = fesc_ly_alpha
        self.dust_curve = dust_curve
        self.tau_v = tau_v
        self.dust_emission = dust_emission
        self.per_particle = False
        self.saved_spectra = None
        self.fesc_per_galaxy = isinstance(fesc, str)
        self.fesc_name = fesc if self.fesc_per_galaxy else None
        # fesc_ly_alpha="fesc_lya" (final_library_generation_multinode.py:499): the line-continuum value of the ONE bin
        # nearest 1216 A is scaled per galaxy; grids are lowered without it and lya_line() hands the kernel that value
        self.lya_per_galaxy = isinstance(fesc_ly_alpha, str)
        self.lya_name = fesc_ly_alpha if self.lya_per_galaxy else None
        if dust_emission is not None and not hasattr(dust_emission, "shape"):
            raise NotImplementedError(f"dust emission generator {type(dust_emission).__name__}: only Greybody / Blackbody "
                                      "(a fixed spectral shape scaled by energy balance) are in the batched path")

    # reference hooks (library.py:2506, 2512) - bookkeeping only
    def set_per_particle(self, flag):
        self.per_particle = bool(flag)

    def save_spectra(self, *keys):
        self.saved_spectra = list(keys)

    def _component(self, name):
        sp = self.grid.spectra
        if name in sp:
            return sp[name]
        if name in ("nebular_continuum", "linecont"):
            return np.zeros_like(next(iter(sp.values())))
        if name == "transmitted":
            return sp["incident"]
        raise KeyError(f"grid has no '{name}' spectra")

    # which per-galaxy factor multiplies the (dust-screened, unscreened) grid of each spectrum when fesc is per galaxy
    _PER_GALAXY = {"incident": (None, "one"), "transmitted": (None, "1-f"), "nebular": (None, "1-f"),
                   "reprocessed": (None, "1-f"), "escaped": (None, "f"), "intrinsic": ("1-f", "f"),
                   "attenuated": ("1-f", None), "emergent": ("1-f", "f"), "total": ("1-f", "f")}

    def has_dust_emission(self, key):
        """'total' = 'emergent' + the generator's spectrum scaled to the energy the screen(s) removed (A5)."""
        return key == "total" and self.dust_emission is not None and self.dust_curve is not None

    def dust_free(self, key):
        """True when the first grid of ``recipe(key)`` must NOT be attenuated (per-galaxy fesc, 'intrinsic':
        reprocessed and escaped light need different coefficients but neither sees the screen)."""
        return self.fesc_per_galaxy and key == "intrinsic"

    def coefficients(self, key, fesc_values):
        """Per-galaxy factors ``(c_first, c_second)`` on the two grids of ``recipe(key)``; ``None`` = 1."""
        if not self.fesc_per_galaxy:
            return None, None
        f = np.asarray(fesc_values, dtype=np.float64)
        if np.any((f < 0) | (f > 1)) or not np.all(np.isfinite(f)):
            raise ValueError("fesc must lie in [0, 1]")
        pick = {"1-f": 1.0 - f, "f": f, "one": None, None: None}
        a, b = self._PER_GALAXY[key]
        return pick[a], pick[b]

    def lya_line(self, key):
        """``(values (N_age, N_Z), bin)`` of the Lyman-alpha line term that a per-galaxy ``fesc_ly_alpha`` multiplies in
        spectrum ``key`` (it always lives in the first grid of ``recipe(key)``), or ``None``."""
        if not self.lya_per_galaxy or key in ("incident", "transmitted", "escaped"):
            return None
        lam = np.asarray(self.grid.lam)
        i = int(np.argmin(np.abs(lam - LYA)))
        scale = 1.0 if self.fesc_per_galaxy else 1.0 - float(self.fesc)
        return scale * self._component("linecont")[..., i], i

    def recipe(self, key):
        if key not in self.available:
            raise ValueError(f"Emission model {type(self).__name__} has no spectrum '{key}'")
        lam = np.asarray(self.grid.lam)
        flya = 0.0 if self.lya_per_galaxy else float(self.fesc_ly_alpha)
        if self.fesc_per_galaxy:
            inc = self._component("incident")
            zero = np.zeros_like(inc)
            line = self._component("linecont").copy()
            line[..., int(np.argmin(np.abs(lam - LYA)))] *= flya
            trans, neb = self._component("transmitted"), line + self._component("nebular_continuum")
            repro = trans + neb
            table = {"incident": (zero, inc), "transmitted": (zero, trans), "nebular": (zero, neb),
                     "reprocessed": (zero, repro), "escaped": (zero, inc), "intrinsic": (repro, inc),
                     "attenuated": (repro, zero), "emergent": (repro, inc), "total": (repro, inc)}
            return table[key]
        fesc = float(self.fesc)
        inc = self._component("incident")
        zero = np.zeros_like(inc)
        if key == "incident":
            return zero, inc
        line = self._component("linecont").copy()
        line[..., int(np.argmin(np.abs(lam - LYA)))] *= flya
        trans = (1.0 - fesc) * self._component("transmitted")
        neb = (1.0 - fesc) * (line + self._component("nebular_continuum"))
        repro = trans + neb
        esc = fesc * inc
        table = {"transmitted": (zero, trans), "nebular": (zero, neb), "reprocessed": (zero, repro),
                 "escaped": (zero, esc), "intrinsic": (zero, repro + esc), "attenuated": (repro, zero),
                 "emergent": (repro, esc), "total": (repro, esc)}
        return table[key]


class IncidentEmission(EmissionModel):
    available = ("incident",)


class IntrinsicEmission(EmissionModel):
    available = ("incident", "transmitted", "nebular", "reprocessed", "escaped", "intrinsic")


class PacmanEmission(EmissionModel):
    pass


class BimodalPacmanEmission(EmissionModel):
    """Charlot & Fall style two-screen attenuation (``generate_library_full.py:221-231``,
    ``generate_spectral_library.py:329-340``): stars younger than ``age_pivot`` (log10 yr, ``log10ages < age_pivot``) sit
    behind their birth cloud AND the ISM, older stars behind the ISM only.  ``tau_v_ism`` / ``tau_v_birth`` name per-galaxy
    emitter attributes.  Lowered as two grid components (young, old reprocessed light) that are BOTH attenuated
    (``two_screens``); an escape fraction would need a third, unscreened component and is not supported."""

    def __init__(self, grid, tau_v_ism="tau_v_ism", tau_v_birth="tau_v_birth", dust_curve_ism=None, dust_curve_birth=None,
                 age_pivot=7.0, dust_emission_ism=None, dust_emission_birth=None, fesc=0.0, fesc_ly_alpha=1.0, **kwargs):
        if dust_curve_ism is None or dust_curve_birth is None:
            raise ValueError("BimodalPacmanEmission needs dust_curve_ism and dust_curve_birth")
        if isinstance(fesc, str) or float(fesc) != 0.0:
            raise NotImplementedError("BimodalPacmanEmission with a non-zero escape fraction needs a third (unscreened) "
                                      "component; not in the batched path yet")
        for c in (dust_curve_ism, dust_curve_birth):
            if getattr(c, "per_galaxy", False):
                raise NotImplementedError("per-galaxy dust-curve shape together with two screens is not in the batched path yet")
        if (dust_emission_ism is None) != (dust_emission_birth is None) or (
                dust_emission_ism is not None and not getattr(dust_emission_ism, "same_as", lambda o: False)(dust_emission_birth)):
            raise NotImplementedError("two dust screens re-emitting with DIFFERENT generators need two absorbed-energy sums; "
                                      "the batched path has one (every reference script passes the same generator twice)")
        super().__init__(grid, fesc=0.0, fesc_ly_alpha=fesc_ly_alpha, dust_curve=dust_curve_ism, tau_v=tau_v_ism,
                         dust_emission=dust_emission_ism, **kwargs)
        self.dust_curve_birth = dust_curve_birth
        self.tau_v_ism_name, self.tau_v_birth_name = tau_v_ism, tau_v_birth
        self.age_pivot = float(strip_units(age_pivot))

    def two_screens(self, key):
        return key in ("attenuated", "emergent", "total")

    def recipe(self, key):
        att, un = super().recipe(key)
        if not self.two_screens(key):
            return att, un
        young = (np.asarray(self.grid.log10ages) < self.age_pivot)[:, None, None]
        return np.where(young, att, 0.0), np.where(young, 0.0, att)       # (young, old), both behind dust


class TotalEmission(EmissionModel):
    def __init__(self, grid, dust_curve=None, tau_v="tau_v", dust_emission_model=None, **kw):
        super().__init__(grid, dust_curve=dust_curve, tau_v=tau_v, dust_emission=dust_emission_model, **kw)


class EmergentEmission(EmissionModel):
    pass
