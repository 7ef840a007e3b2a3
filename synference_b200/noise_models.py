"""Uncertainty-model classes with the reference's names and semantics.

Mirrors ``src/synference/noise_models.py``: ``UncertaintyModel`` (:26-73), ``DepthUncertaintyModel``
(:76-208), ``SpectralUncertaintyModel`` (:211-259), ``EmpiricalUncertaintyModel`` (:262-440),
``AsinhEmpiricalUncertaintyModel`` (:443-635), ``GeneralEmpiricalUncertaintyModel`` (:638-1099),
the HDF5 registry (:1106-1156).

Where the work is: training rows (millions of them) are scattered on the GPU by
``sb2_depth_noise_features`` (:meth:`DepthUncertaintyModel.apply_noise_device`,
:func:`synference_b200.features.create_feature_array_from_raw_photometry`).  The per-object methods
below act on the handful of values ``GalaxySimulator._scatter`` passes (``library.py:5975-5993``) and
on model *construction* (binning, interpolators), which is set-up work on the host exactly as in the
reference; they are not a substitute for the CUDA path and the device entry points raise without it.
The empirical models' row-wise application on device is SURVEY 8f-3 ("next").
"""

from __future__ import annotations

import warnings
from abc import ABC, abstractmethod
from typing import Any, Dict, Optional, Tuple, Union

import numpy as np
from scipy import stats
from scipy.interpolate import interp1d

from .units import Jy, Quantity, Unit, has_units, strip_units
from .utils import f_jy_err_to_asinh, f_jy_to_asinh, read_container, write_container

_LN10 = np.log(10)


def _to_jy(flux, units=None):
    """Plain float array in Jy from a quantity, or from bare numbers in ``units`` (default Jy)."""
    if has_units(flux):
        return np.asarray(strip_units(flux, "Jy"), dtype=float)
    f = np.asarray(flux, dtype=float)
    if units is None:
        return f
    return f * Unit(str(units)).factor


def _from_jy(values, units):
    return np.asarray(values, dtype=float) / Unit(str(units)).factor


class UncertaintyModel(ABC):
    """Common interface plus the static photometric converters (``noise_models.py:55-73``)."""

    def __init__(self, return_noise: bool = False, **kwargs: Any) -> None:
        self.return_noise = return_noise

    @abstractmethod
    def apply_noise(self, flux):
        ...

    @abstractmethod
    def serialize_to_hdf5(self, hdf5_group):
        ...

    @classmethod
    @abstractmethod
    def _from_hdf5_group(cls, hdf5_group):
        ...

    @staticmethod
    def ab_to_jy(magnitude):
        return Quantity(10 ** (-0.4 * (np.asarray(magnitude, dtype=float) - 8.90)), Jy)

    @staticmethod
    def jy_to_ab(flux):
        with np.errstate(all="ignore"):
            return -2.5 * np.log10(_to_jy(flux)) + 8.90

    @staticmethod
    def ab_err_to_jy(magnitude_err, flux_jy):
        return Quantity((_to_jy(flux_jy) * magnitude_err * _LN10) / 2.5, Jy)

    @staticmethod
    def jy_err_to_ab(flux_err_jy, flux_jy):
        with np.errstate(all="ignore"):
            return np.abs((2.5 / _LN10) * (_to_jy(flux_err_jy) / _to_jy(flux_jy)))


class DepthUncertaintyModel(UncertaintyModel):
    """Gaussian noise from a fixed survey depth: ``sigma = ab_to_jy(depth) / n_sigma``."""

    def __init__(self, depth_ab: float, depth_sigma_level=5.0, min_flux_error: Optional[float] = None,
                 max_flux_error: Optional[float] = None, **kwargs: Any):
        super().__init__(**kwargs)
        self.depth_ab = depth_ab
        self.depth_sigma_level = depth_sigma_level
        self.sigma = Quantity(np.asarray(self.ab_to_jy(depth_ab)) / depth_sigma_level, Jy)
        self.min_flux_error = min_flux_error if min_flux_error is not None else 0.0
        self.max_flux_error = max_flux_error if max_flux_error is not None else np.inf
        assert not np.isnan(float(self.sigma.value)), "sigma must not be NaN"

    def _true_flux_jy(self, flux, true_flux_units):
        if true_flux_units is not None:
            if true_flux_units == "AB":
                return np.asarray(self.ab_to_jy(flux))
            if has_units(flux):
                assert Unit(str(true_flux_units)) == flux.units, \
                    "If true_flux_units is specified, flux must be a unyt_array with the same units."
            return _to_jy(strip_units(flux), true_flux_units)
        return _to_jy(flux)

    def apply_noise(self, flux, true_flux_units=None, out_units=None, **kwargs):
        """``noisy = flux_Jy + N(0, sigma)`` (``noise_models.py:113-166``); AB output of a negative noisy
        flux is NaN, as in the reference."""
        if kwargs:
            print(f"WARNING {kwargs} arguments will have no effect with this model.")
        flux_jy = self._true_flux_jy(flux, true_flux_units)
        sigma = float(self.sigma.value)
        noisy = flux_jy + np.random.normal(loc=0.0, scale=sigma, size=np.shape(flux_jy))
        unc = np.ones_like(noisy) * sigma
        noisy_out, unc_out = Quantity(noisy, Jy), Quantity(unc, Jy)
        if out_units is not None:
            if out_units == "AB":
                unc_out = self.jy_err_to_ab(unc, noisy)
                noisy_out = self.jy_to_ab(noisy)
            else:
                noisy_out, unc_out = _from_jy(noisy, out_units), _from_jy(unc, out_units)
        unc_out = np.clip(unc_out, self.min_flux_error, self.max_flux_error)
        return (noisy_out, unc_out) if self.return_noise else noisy_out

    def apply_noise_device(self, flux_njy, n_scatter=1, normals=None, seed=0, epoch=0, device=0):
        """Row-wise scatter of a ``(n_gal, 1)`` flux column on the GPU (torch tensors out)."""
        from .engine import depth_noise_features
        sigma_njy = np.array([float(self.sigma.value) * 1e9])
        return depth_noise_features(flux_njy, sigma_njy, n_scatter=n_scatter, normals=normals, seed=seed,
                                    epoch=epoch, want_features=False, device=device)

    def apply_scalings(self, flux, error, flux_units: str, out_units: str):
        if flux_units == out_units:
            return flux, error
        if flux_units == "AB":
            flux_jy = np.asarray(self.ab_to_jy(flux))
            error_jy = np.asarray(self.ab_err_to_jy(error, flux_jy))
        else:
            flux_jy, error_jy = _to_jy(strip_units(flux), flux_units), _to_jy(strip_units(error), flux_units)
        error_jy = np.clip(error_jy, self.min_flux_error, self.max_flux_error)
        if out_units == "AB":
            return self.jy_to_ab(flux_jy), self.jy_err_to_ab(error_jy, flux_jy)
        return _from_jy(flux_jy, out_units), _from_jy(error_jy, out_units)

    def serialize_to_hdf5(self, hdf5_group):
        a = hdf5_group.attrs
        a["__class__"] = self.__class__.__name__
        a["depth_ab"], a["depth_sigma_level"] = float(self.depth_ab), float(self.depth_sigma_level)
        a["return_noise"] = bool(self.return_noise)
        a["min_flux_error"], a["max_flux_error"] = float(self.min_flux_error), float(self.max_flux_error)

    @classmethod
    def _from_hdf5_group(cls, hdf5_group):
        a = hdf5_group.attrs
        return cls(depth_ab=a["depth_ab"], depth_sigma_level=a["depth_sigma_level"],
                   return_noise=a["return_noise"], min_flux_error=a.get("min_flux_error", 0.0),
                   max_flux_error=a.get("max_flux_error", np.inf))


class SpectralUncertaintyModel(UncertaintyModel):
    """Per-pixel Gaussian noise from a fixed error kernel (``noise_models.py:211-259``)."""

    def __init__(self, error_kernel: np.ndarray, **kwargs: Any):
        super().__init__(**kwargs)
        self.error_kernel = np.asarray(error_kernel, dtype=float)

    def apply_noise(self, flux, **kwargs):
        if kwargs:
            print(f"WARNING {kwargs} arguments will have no effect with this model.")
        flux = np.asarray(flux, dtype=float)
        if flux.shape != self.error_kernel.shape:
            raise ValueError("Input flux shape must match the error kernel shape.")
        noisy = flux + np.random.normal(loc=0.0, scale=self.error_kernel, size=flux.shape)
        return (noisy, self.error_kernel) if self.return_noise else noisy

    def serialize_to_hdf5(self, hdf5_group):
        hdf5_group.attrs["__class__"] = self.__class__.__name__
        hdf5_group.attrs["return_noise"] = bool(self.return_noise)
        hdf5_group.create_dataset("error_kernel", data=self.error_kernel)

    @classmethod
    def _from_hdf5_group(cls, hdf5_group):
        return cls(error_kernel=np.asarray(hdf5_group["error_kernel"]), return_noise=hdf5_group.attrs["return_noise"])


class EmpiricalUncertaintyModel(UncertaintyModel, ABC):
    """sigma(flux) learnt from a catalogue: binned median / std -> linear interpolators ->
    truncated-normal sigma draw (``noise_models.py:262-440``)."""

    def __init__(self, extrapolate: bool = False, min_samples_per_bin: int = 10, num_bins: int = 20,
                 log_bins: bool = True, **kwargs: Any):
        super().__init__(**kwargs)
        self.extrapolate = extrapolate
        self._min_samples_per_bin, self._num_bins, self._log_bins = min_samples_per_bin, num_bins, log_bins
        self.bin_centers = self.median_error_in_bin = self.std_error_in_bin = None
        self._mu_sigma_interpolator = self._sigma_sigma_interpolator = None

    def _compute_bins_from_data(self, fluxes, errors, precomputed_bins=None):
        fluxes, errors = np.asarray(fluxes, dtype=float), np.asarray(errors, dtype=float)
        if precomputed_bins is not None:
            bins = precomputed_bins
        else:
            valid = np.isfinite(fluxes)
            if not np.any(valid):
                raise ValueError("No valid finite data to build bins.")
            fb = fluxes[valid]
            if self._log_bins:
                pos = fb > 0
                if not np.any(pos):
                    raise ValueError("Log-binning requires positive flux values.")
                bins = np.logspace(np.log10(fb[pos].min()), np.log10(fb.max()), self._num_bins + 1)
            else:
                bins = np.linspace(fb.min(), fb.max(), self._num_bins + 1)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            med, edges, _ = stats.binned_statistic(fluxes, errors, "median", bins=bins)
            std, _, _ = stats.binned_statistic(fluxes, errors, np.std, bins=bins)
            cnt, _, _ = stats.binned_statistic(fluxes, fluxes, "count", bins=bins)
        centers = (edges[:-1] + edges[1:]) / 2.0
        ok = cnt >= self._min_samples_per_bin
        if ok.sum() < 2:
            raise ValueError("Could not create enough valid bins for interpolation.")
        self.bin_centers, self.median_error_in_bin, self.std_error_in_bin = centers[ok], med[ok], std[ok]

    def _create_interpolators(self):
        if self.bin_centers is None or len(self.bin_centers) < 2:
            raise AttributeError("Binned data not found. Cannot create interpolators.")
        ext = getattr(self, "extrapolate", False)
        fm = "extrapolate" if ext else (self.median_error_in_bin[0], self.median_error_in_bin[-1])
        fs = "extrapolate" if ext else (self.std_error_in_bin[0], self.std_error_in_bin[-1])
        self._mu_sigma_interpolator = interp1d(self.bin_centers, self.median_error_in_bin, kind="linear",
                                               bounds_error=False, fill_value=fm)
        self._sigma_sigma_base_interpolator = interp1d(self.bin_centers, self.std_error_in_bin, kind="linear",
                                                       bounds_error=False, fill_value=fs)
        self._sigma_sigma_interpolator = self._non_negative_sigma_wrapper

    def _non_negative_sigma_wrapper(self, flux_values):
        return np.maximum(0, self._sigma_sigma_base_interpolator(flux_values))

    def sample_uncertainty(self, flux_values):
        mu = self._mu_sigma_interpolator(flux_values)
        ss = self._sigma_sigma_interpolator(flux_values)
        a = (0 - mu) / np.where(ss > 1e-9, ss, 1)
        return stats.truncnorm.rvs(a=a, b=np.inf, loc=mu, scale=ss, size=len(flux_values))

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("_mu_sigma_interpolator", "_sigma_sigma_interpolator", "_sigma_sigma_base_interpolator"):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        if self.bin_centers is not None:
            self._create_interpolators()

    def serialize_to_hdf5(self, hdf5_group):
        a = hdf5_group.attrs
        a["__class__"] = self.__class__.__name__
        a["extrapolate"], a["min_samples_per_bin"] = bool(self.extrapolate), int(self._min_samples_per_bin)
        a["num_bins"], a["log_bins"] = int(self._num_bins), bool(self._log_bins)
        if self.bin_centers is not None:
            hdf5_group.create_dataset("bin_centers", data=self.bin_centers)
            hdf5_group.create_dataset("median_error_in_bin", data=self.median_error_in_bin)
            hdf5_group.create_dataset("std_error_in_bin", data=self.std_error_in_bin)

    @classmethod
    def _from_hdf5_group(cls, hdf5_group):
        a = hdf5_group.attrs
        inst = cls.__new__(cls)
        EmpiricalUncertaintyModel.__init__(inst, extrapolate=a.get("extrapolate", False),
                                           min_samples_per_bin=a.get("min_samples_per_bin", 10),
                                           num_bins=a.get("num_bins", 20), log_bins=a.get("log_bins", True))
        if "bin_centers" in hdf5_group:
            inst.bin_centers = np.asarray(hdf5_group["bin_centers"])
            inst.median_error_in_bin = np.asarray(hdf5_group["median_error_in_bin"])
            inst.std_error_in_bin = np.asarray(hdf5_group["std_error_in_bin"])
            inst._create_interpolators()
        return inst


class AsinhEmpiricalUncertaintyModel(EmpiricalUncertaintyModel):
    """Empirical model in asinh-magnitude space, ``b = k * median(sigma)`` (``noise_models.py:443-635``)."""

    def __init__(self, observed_phot_jy=None, observed_phot_errors_jy=None, asinh_b_factor: float = 5.0,
                 error_type: str = "empirical", min_flux_error=None, max_flux_error=None,
                 interpolation_flux_unit: str = "asinh", **kwargs: Any):
        super().__init__(**kwargs)
        self.error_type = error_type
        self.min_flux_error = min_flux_error if min_flux_error is not None else 0.0
        self.max_flux_error = max_flux_error if max_flux_error is not None else np.inf
        self.interpolation_flux_unit = interpolation_flux_unit
        self.b = None
        if observed_phot_jy is not None and observed_phot_errors_jy is not None:
            f, e = _to_jy(observed_phot_jy), _to_jy(observed_phot_errors_jy)
            ok = np.isfinite(f) & np.isfinite(e)
            f, e = f[ok], e[ok]
            self.b = Quantity(asinh_b_factor * np.median(e), Jy)
            if interpolation_flux_unit == "asinh":
                self._compute_bins_from_data(f_jy_to_asinh(f, self.b), f_jy_err_to_asinh(f, e, self.b))
            else:
                self._compute_bins_from_data(_from_jy(f, interpolation_flux_unit), _from_jy(e, interpolation_flux_unit))
            self._create_interpolators()

    def apply_noise(self, flux, true_flux_units: Optional[str] = None, **kwargs):
        if true_flux_units == "AB":
            f_jy = np.asarray(self.ab_to_jy(flux))
            warnings.warn("Using asinh model with AB input will not benefit from asinh scaling of neg fluxes.")
        elif true_flux_units is not None:
            f_jy = _to_jy(strip_units(flux), true_flux_units)
        else:
            f_jy = _to_jy(flux)
        b = float(self.b.value)
        if self.interpolation_flux_unit == "asinh":
            m_true = f_jy_to_asinh(f_jy, b)
            e_samp = self.sample_uncertainty(m_true)
            m_noisy = m_true + np.random.normal(loc=0.0, scale=e_samp)
            final = e_samp if self.error_type == "empirical" else self.sample_uncertainty(m_noisy)
        else:
            u = self.interpolation_flux_unit
            e_jy = _to_jy(self.sample_uncertainty(_from_jy(f_jy, u)), u)
            noisy_jy = f_jy + np.random.normal(loc=0.0, scale=e_jy)
            m_noisy = f_jy_to_asinh(noisy_jy, b)
            fe_jy = _to_jy(self.sample_uncertainty(_from_jy(noisy_jy, u)), u) if self.error_type == "empirical" else e_jy
            final = f_jy_err_to_asinh(noisy_jy, fe_jy, b)
        final = np.clip(final, self.min_flux_error, self.max_flux_error)
        return (m_noisy, final) if self.return_noise else m_noisy

    def device_model(self, true_flux_units=None, out_units=None):
        """This model as the C ABI's ``sb2_empirical_model`` (asinh modes; ``out_units`` has no effect: asinh magnitudes)."""
        from . import _capi
        m = _capi.EmpiricalModel()
        nb = len(self.bin_centers)
        if nb > _capi.EMP_MAX_BINS:
            raise ValueError(f"empirical model has {nb} bins; the device table holds {_capi.EMP_MAX_BINS}")
        m.n_bins, m.extrapolate = nb, int(bool(self.extrapolate))
        for i in range(nb):
            m.centers[i], m.median[i], m.stdev[i] = self.bin_centers[i], self.median_error_in_bin[i], self.std_error_in_bin[i]
        u = "Jy" if true_flux_units is None else true_flux_units
        m.in_is_ab, m.in_to_jy = (1, 1.0) if str(u) == "AB" else (0, float(Unit(str(u)).factor))
        m.out_is_ab, m.out_to_jy, m.internal_is_ab = 0, 1.0, 0
        if self.interpolation_flux_unit == "asinh":
            m.asinh_mode, m.internal_to_jy = 1, 1.0
        else:
            m.asinh_mode, m.internal_to_jy = 2, float(Unit(str(self.interpolation_flux_unit)).factor)
        m.asinh_b = float(self.b.value)
        m.observed_error = int(self.error_type != "empirical")
        m.sigma_clip, m.upper_limits, m.ul_active, m.ul_scatter_std = -1.0, 0, 0, -1.0
        m.min_err, m.max_err = float(self.min_flux_error), float(self.max_flux_error)
        return m

    def apply_scalings(self, flux, error, **kwargs):
        if kwargs:
            print(f"WARNING {kwargs} arguments will have no effect with this model. Input must be in Jy.")
        f, e = _to_jy(flux), _to_jy(error)
        b = float(self.b.value)
        return f_jy_to_asinh(f, b), np.clip(f_jy_err_to_asinh(f, e, b), self.min_flux_error, self.max_flux_error)

    def serialize_to_hdf5(self, hdf5_group):
        super().serialize_to_hdf5(hdf5_group)
        a = hdf5_group.attrs
        a["error_type"], a["b_value"], a["b_units"] = self.error_type, float(self.b.value), "Jy"
        a["return_noise"] = bool(self.return_noise)
        a["min_flux_error"], a["max_flux_error"] = float(self.min_flux_error), float(self.max_flux_error)
        a["interpolation_flux_unit"] = self.interpolation_flux_unit

    @classmethod
    def _from_hdf5_group(cls, hdf5_group):
        inst = super(AsinhEmpiricalUncertaintyModel, cls)._from_hdf5_group(hdf5_group)
        a = hdf5_group.attrs
        inst.error_type = a["error_type"]
        inst.b = Quantity(a["b_value"], a["b_units"])
        inst.return_noise = a["return_noise"]
        inst.min_flux_error, inst.max_flux_error = a["min_flux_error"], a["max_flux_error"]
        inst.interpolation_flux_unit = a["interpolation_flux_unit"]
        return inst


class GeneralEmpiricalUncertaintyModel(EmpiricalUncertaintyModel):
    """Empirical sigma(flux) in AB or physical units with optional upper-limit handling
    (``noise_models.py:638-1099``)."""

    def __init__(self, observed_fluxes, observed_errors, flux_unit: str = "AB", interpolation_flux_unit=None,
                 already_binned: bool = False, bin_median_errors=None, bin_std_errors=None, flux_bins=None,
                 min_flux_for_binning=None, sigma_clip: float = None, min_flux_error: float = 0.0,
                 max_flux_error: float = np.inf, error_type: str = "empirical", upper_limits: bool = False,
                 treat_as_upper_limits_below=None, upper_limit_flux_behaviour="scatter_limit",
                 upper_limit_flux_err_behaviour: str = "flux", **kwargs: Any):
        super().__init__(**kwargs)
        self.flux_unit = flux_unit
        self.interpolation_flux_unit = interpolation_flux_unit if interpolation_flux_unit else flux_unit
        self.sigma_clip = sigma_clip
        self.min_flux_error, self.max_flux_error = min_flux_error, max_flux_error
        self.error_type = error_type
        self.upper_limits = upper_limits
        self.treat_as_upper_limits_below = treat_as_upper_limits_below
        self.upper_limit_flux_behaviour = upper_limit_flux_behaviour
        self.upper_limit_flux_err_behaviour = upper_limit_flux_err_behaviour
        self.log_snr_interpolator = None
        self.upper_limit_value = None
        if already_binned:
            self.bin_centers = np.asarray(observed_fluxes, dtype=float)
            self.median_error_in_bin = np.asarray(bin_median_errors, dtype=float)
            self.std_error_in_bin = np.asarray(bin_std_errors, dtype=float)
            self._create_interpolators()
            return
        f, e = self._convert_units(observed_fluxes, observed_errors)
        ok = np.isfinite(f) & np.isfinite(e) & (e > 0)
        if min_flux_for_binning is not None:
            ok &= f > min_flux_for_binning
        self._compute_bins_from_data(f[ok], e[ok], precomputed_bins=flux_bins)
        if self.upper_limits:
            self._setup_upper_limit_interpolator(f[ok], e[ok])
        self._create_interpolators()

    # unit plumbing: internal unit <-> AB / physical
    def _convert_units(self, fluxes, errors, fluxes_unit=None):
        fluxes_unit = self.flux_unit if fluxes_unit is None else fluxes_unit
        f, e = np.asarray(strip_units(fluxes), dtype=float), np.asarray(strip_units(errors), dtype=float)
        iu = self.interpolation_flux_unit
        if str(iu) == str(fluxes_unit):
            return f, e
        if fluxes_unit == "AB":
            fj = np.asarray(self.ab_to_jy(f))
            return _from_jy(fj, iu), _from_jy(np.asarray(self.ab_err_to_jy(e, fj)), iu)
        if iu == "AB":
            fj, ej = _to_jy(f, fluxes_unit), _to_jy(e, fluxes_unit)
            return self.jy_to_ab(fj), self.jy_err_to_ab(ej, fj)
        conv = Unit(str(fluxes_unit)).factor / Unit(str(iu)).factor
        return f * conv, e * conv

    def _convert_units_inverse(self, fluxes, errors, out_unit=None):
        out_unit = self.flux_unit if out_unit is None else out_unit
        iu = self.interpolation_flux_unit
        if str(iu) == str(out_unit):
            return fluxes, errors
        if iu == "AB":
            fj = np.asarray(self.ab_to_jy(fluxes))
            return _from_jy(fj, out_unit), _from_jy(np.asarray(self.ab_err_to_jy(errors, fj)), out_unit)
        if out_unit == "AB":
            fj, ej = _to_jy(fluxes, iu), _to_jy(errors, iu)
            return self.jy_to_ab(fj), self.jy_err_to_ab(ej, fj)
        conv = Unit(str(iu)).factor / Unit(str(out_unit)).factor
        return fluxes * conv, errors * conv

    def _internal_to_jy(self, fluxes, errors):
        if self.interpolation_flux_unit == "AB":
            fj = np.asarray(self.ab_to_jy(fluxes))
            return fj, np.asarray(self.ab_err_to_jy(errors, fj))
        return _to_jy(fluxes, self.interpolation_flux_unit), _to_jy(errors, self.interpolation_flux_unit)

    def _setup_upper_limit_interpolator(self, fluxes, errors):
        fj, ej = self._internal_to_jy(fluxes, errors)
        with np.errstate(divide="ignore", invalid="ignore"):
            snr = fj / ej
        ok = np.isfinite(snr) & (snr > 0) & np.isfinite(fj) & (fj > 0)
        if ok.sum() < 2:
            return
        order = np.argsort(snr[ok])
        self._snr_x_data, self._snr_y_data = np.log10(snr[ok][order]), np.log10(fj[ok][order])
        self.log_snr_interpolator = interp1d(self._snr_x_data, self._snr_y_data, kind="linear",
                                             bounds_error=False, fill_value="extrapolate")
        ul = 10 ** self.log_snr_interpolator(np.log10(self.treat_as_upper_limits_below))
        self.upper_limit_value = float(self.jy_to_ab(ul)) if self.interpolation_flux_unit == "AB" else \
            float(_from_jy(ul, self.interpolation_flux_unit))

    def _get_snr_mask(self, fluxes, errors):
        fj, ej = self._internal_to_jy(fluxes, errors)
        with np.errstate(divide="ignore", invalid="ignore"):
            snr = fj / ej
        return ~np.isfinite(snr) | (snr < self.treat_as_upper_limits_below)

    def _apply_flux_behaviour(self, fluxes, mask, scatter: bool):
        if self.upper_limit_flux_behaviour == "scatter_limit":
            if scatter:
                std = self._sigma_sigma_interpolator(self.upper_limit_value)
                fluxes[mask] = self.upper_limit_value + stats.truncnorm.rvs(-3, 3, loc=0, scale=std, size=mask.sum())
            else:
                fluxes[mask] = self.upper_limit_value
        elif self.upper_limit_flux_behaviour == "upper_limit":
            fluxes[mask] = self.upper_limit_value
        else:
            fluxes[mask] = float(self.upper_limit_flux_behaviour)
        return fluxes

    def _apply_error_behaviour(self, errors, mask):
        b = self.upper_limit_flux_err_behaviour
        if b == "flux":
            errors[mask] = self._mu_sigma_interpolator(self.upper_limit_value)
        elif b == "upper_limit":
            errors[mask] = self.upper_limit_value
        elif b == "max":
            errors[mask] = self.max_flux_error
        elif b.startswith("sig_"):
            sig = float(b.split("_")[1])
            if self.interpolation_flux_unit == "AB":
                errors[mask] = (2.5 / _LN10) / sig
            else:
                if self.log_snr_interpolator is None:
                    raise ValueError("SNR interpolator is not available for 'sig_X' error behaviour in flux space.")
                f_at = _from_jy(10 ** self.log_snr_interpolator(np.log10(sig)), self.interpolation_flux_unit)
                errors[mask] = self._mu_sigma_interpolator(f_at)
        return errors

    def apply_noise(self, flux, true_flux_units: str = None, out_units=None):
        """Sampled sigma, Gaussian (or sigma-clipped) scatter, SNR-based upper limits before and after
        the scatter, unit round trip and the final clip (``noise_models.py:818-880``)."""
        flux = np.asarray(strip_units(flux), dtype=float)
        f_int, _ = self._convert_units(flux, np.zeros_like(flux), true_flux_units)
        sig = self.sample_uncertainty(f_int)
        noisy, final = np.copy(f_int), np.copy(sig)
        init_mask = self._get_snr_mask(f_int, sig) if self.upper_limits else np.zeros_like(f_int, dtype=bool)
        apply = ~init_mask
        if np.any(apply):
            if self.sigma_clip is not None:
                noise = stats.truncnorm.rvs(-self.sigma_clip, self.sigma_clip, 0, sig[apply])
            else:
                noise = np.random.normal(loc=0.0, scale=sig[apply])
            noisy[apply] += noise
        if self.error_type == "observed":
            final = self.sample_uncertainty(noisy)
        if self.upper_limits and self.upper_limit_value is not None:
            mask = init_mask | self._get_snr_mask(noisy, final)
            if np.any(mask):
                noisy = self._apply_flux_behaviour(noisy, mask, scatter=True)
                final = self._apply_error_behaviour(final, mask)
        out_f, out_s = self._convert_units_inverse(noisy, final, out_units)
        out_s = np.clip(out_s, self.min_flux_error, self.max_flux_error)
        return (out_f, out_s) if self.return_noise else out_f

    def device_model(self, true_flux_units=None, out_units=None):
        """This model as the C ABI's ``sb2_empirical_model`` (``include/synference_b200.h``): the interpolation tables,
        the three units, and the upper-limit rules with their replacement values resolved here (they are constants:
        ``noise_models.py:882-957``)."""
        from . import _capi
        m = _capi.EmpiricalModel()
        nb = len(self.bin_centers)
        if nb > _capi.EMP_MAX_BINS:
            raise ValueError(f"empirical model has {nb} bins; the device table holds {_capi.EMP_MAX_BINS}")
        m.n_bins, m.extrapolate = nb, int(bool(self.extrapolate))
        for i in range(nb):
            m.centers[i], m.median[i], m.stdev[i] = self.bin_centers[i], self.median_error_in_bin[i], self.std_error_in_bin[i]

        def unit(u):
            return (1, 1.0) if str(u) == "AB" else (0, float(Unit(str(u)).factor))
        m.internal_is_ab, m.internal_to_jy = unit(self.interpolation_flux_unit)
        m.in_is_ab, m.in_to_jy = unit(self.flux_unit if true_flux_units is None else true_flux_units)
        m.out_is_ab, m.out_to_jy = unit(self.flux_unit if out_units is None else out_units)
        m.sigma_clip = -1.0 if self.sigma_clip is None else float(self.sigma_clip)
        m.observed_error = int(self.error_type == "observed")
        m.upper_limits = int(bool(self.upper_limits))
        m.ul_active = int(bool(self.upper_limits) and self.upper_limit_value is not None)
        m.snr_threshold = float(self.treat_as_upper_limits_below) if self.treat_as_upper_limits_below is not None else 0.0
        m.ul_flux, m.ul_scatter_std, m.ul_err = 0.0, -1.0, 0.0
        if m.ul_active:
            if self.upper_limit_flux_behaviour == "scatter_limit":
                m.ul_flux = float(self.upper_limit_value)
                m.ul_scatter_std = float(self._sigma_sigma_interpolator(self.upper_limit_value))
            elif self.upper_limit_flux_behaviour == "upper_limit":
                m.ul_flux = float(self.upper_limit_value)
            else:
                m.ul_flux = float(self.upper_limit_flux_behaviour)
            m.ul_err = float(self._apply_error_behaviour(np.zeros(1), np.ones(1, dtype=bool))[0])
        m.min_err, m.max_err = float(self.min_flux_error), float(self.max_flux_error)
        return m

    def apply_scalings(self, flux, error, flux_units=None, out_units=None):
        """Deterministic part only: unit conversion, upper-limit replacement, clip."""
        f, e = self._convert_units(flux, error, flux_units)
        f, e = np.array(f, dtype=float), np.array(e, dtype=float)
        if self.upper_limits and self.upper_limit_value is not None:
            mask = self._get_snr_mask(f, e)
            if np.any(mask):
                f = self._apply_flux_behaviour(f, mask, scatter=False)
                e = self._apply_error_behaviour(e, mask)
        f, e = self._convert_units_inverse(f, e, out_units)
        return f, np.clip(e, self.min_flux_error, self.max_flux_error)

    def serialize_to_hdf5(self, hdf5_group):
        a = hdf5_group.attrs
        a["__class__"] = self.__class__.__name__
        if self.bin_centers is not None:
            hdf5_group.create_dataset("bin_centers", data=self.bin_centers)
            hdf5_group.create_dataset("median_error_in_bin", data=self.median_error_in_bin)
            hdf5_group.create_dataset("std_error_in_bin", data=self.std_error_in_bin)
        if self.log_snr_interpolator is not None:
            hdf5_group.create_dataset("snr_x_data", data=self._snr_x_data)
            hdf5_group.create_dataset("snr_y_data", data=self._snr_y_data)
        for k in ("flux_unit", "interpolation_flux_unit", "error_type", "upper_limit_flux_behaviour",
                  "upper_limit_flux_err_behaviour"):
            a[k] = str(getattr(self, k))
        for k in ("min_flux_error", "max_flux_error"):
            a[k] = float(getattr(self, k))
        a["sigma_clip"] = -1.0 if self.sigma_clip is None else float(self.sigma_clip)
        a["upper_limits"], a["return_noise"], a["extrapolate"] = bool(self.upper_limits), bool(self.return_noise), bool(self.extrapolate)
        a["treat_as_upper_limits_below"] = -1.0 if self.treat_as_upper_limits_below is None else float(self.treat_as_upper_limits_below)
        a["upper_limit_value"] = float("nan") if self.upper_limit_value is None else float(self.upper_limit_value)

    @classmethod
    def _from_hdf5_group(cls, hdf5_group):
        a = hdf5_group.attrs
        ul = a.get("treat_as_upper_limits_below", -1.0)
        beh = a.get("upper_limit_flux_behaviour", "scatter_limit")
        try:
            beh = float(beh)
        except ValueError:
            pass
        inst = cls(observed_fluxes=np.asarray(hdf5_group["bin_centers"]), observed_errors=None,
                   flux_unit=a["flux_unit"], interpolation_flux_unit=a["interpolation_flux_unit"],
                   already_binned=True, bin_median_errors=np.asarray(hdf5_group["median_error_in_bin"]),
                   bin_std_errors=np.asarray(hdf5_group["std_error_in_bin"]),
                   sigma_clip=None if a.get("sigma_clip", -1.0) < 0 else a["sigma_clip"],
                   min_flux_error=a.get("min_flux_error", 0.0), max_flux_error=a.get("max_flux_error", np.inf),
                   error_type=a.get("error_type", "empirical"), upper_limits=a.get("upper_limits", False),
                   treat_as_upper_limits_below=None if ul < 0 else ul, upper_limit_flux_behaviour=beh,
                   upper_limit_flux_err_behaviour=a.get("upper_limit_flux_err_behaviour", "flux"),
                   return_noise=a.get("return_noise", False), extrapolate=a.get("extrapolate", False))
        if "snr_x_data" in hdf5_group:
            inst._snr_x_data, inst._snr_y_data = np.asarray(hdf5_group["snr_x_data"]), np.asarray(hdf5_group["snr_y_data"])
            inst.log_snr_interpolator = interp1d(inst._snr_x_data, inst._snr_y_data, kind="linear",
                                                 bounds_error=False, fill_value="extrapolate")
        v = a.get("upper_limit_value", float("nan"))
        inst.upper_limit_value = None if not np.isfinite(v) else float(v)
        return inst


# ---- (de)serialisation registry (noise_models.py:1106-1156) --------------------------------------

_MODEL_REGISTRY = {c.__name__: c for c in (DepthUncertaintyModel, SpectralUncertaintyModel,
                                           AsinhEmpiricalUncertaintyModel, GeneralEmpiricalUncertaintyModel)}


class _Group:
    """Just enough of an ``h5py.Group`` for ``serialize_to_hdf5``: ``.attrs``, ``create_dataset``,
    ``in`` and item access."""

    def __init__(self):
        self.attrs: Dict[str, Any] = {}
        self.data: Dict[str, np.ndarray] = {}

    def create_dataset(self, name, data=None, **kw):
        self.data[name] = np.asarray(data)

    def __contains__(self, k):
        return k in self.data

    def __getitem__(self, k):
        return self.data[k]


def save_unc_model_to_hdf5(model: UncertaintyModel, filepath: str, group_name: str, overwrite: bool = False):
    import os
    datasets, attrs = ({}, {})
    if os.path.exists(filepath):
        datasets, attrs = read_container(filepath)
    prefix = f"{group_name}/"
    if any(k.startswith(prefix) for k in list(datasets) + list(attrs)):
        if not overwrite:
            raise ValueError(f"Group '{group_name}' already exists in {filepath}. Use overwrite=True.")
        datasets = {k: v for k, v in datasets.items() if not k.startswith(prefix)}
        attrs = {k: v for k, v in attrs.items() if not k.startswith(prefix)}
    g = _Group()
    model.serialize_to_hdf5(g)
    datasets.update({prefix + k: v for k, v in g.data.items()})
    attrs.update({prefix + k: v for k, v in g.attrs.items()})
    write_container(filepath, datasets, attrs)


def load_unc_model_from_hdf5(filepath: str, group_name: str) -> UncertaintyModel:
    datasets, attrs = read_container(filepath)
    prefix = f"{group_name}/"
    g = _Group()
    g.data = {k[len(prefix):]: v for k, v in datasets.items() if k.startswith(prefix)}
    g.attrs = {k[len(prefix):]: v for k, v in attrs.items() if k.startswith(prefix)}
    if "__class__" not in g.attrs:
        raise KeyError(f"Group '{group_name}' not found in {filepath}")
    name = g.attrs["__class__"]
    if name not in _MODEL_REGISTRY:
        raise TypeError(f"Unknown model class '{name}' in HDF5 file.")
    return _MODEL_REGISTRY[name]._from_hdf5_group(g)


def create_uncertainty_models_from_EPOCHS_cat(file, bands, new_band_names=None, plot=False, old=False, hdu=0, save=False,
                                              save_path=None, model_class="general", **kwargs):
    """One uncertainty model per band from an EPOCHS-style catalogue (``noise_models.py:1159-1330``).

    ``file``: a table with the columns ``MAG_APER_<band>_aper_corr``, ``FLUX_APER_<band>_aper_corr_Jy`` and
    ``loc_depth_<band>`` -- an astropy ``Table``, a pandas ``DataFrame``, a numpy structured array or a dict of arrays; a path
    is read with ``astropy.table.Table.read`` (needs astropy).  ``model_class``: ``"general"`` (AB magnitudes, SNR < 1 treated
    as upper limits), ``"depth"`` (median local depth) or ``"asinh"``.  Flux errors are ``ab_to_jy(loc_depth) / 5``."""
    if plot:
        raise NotImplementedError("plotting helpers are outside the hot path; build the models with plot=False")
    if isinstance(bands, str):
        bands = [bands]
    if isinstance(file, (str, bytes)) or hasattr(file, "__fspath__"):
        try:
            from astropy.table import Table
        except ImportError as err:
            raise ImportError("reading a catalogue file needs astropy; pass the table's columns as a dict / DataFrame / "
                              "structured array instead") from err
        file = Table.read(file, hdu=hdu)
    names = list(getattr(file, "colnames", None) or getattr(getattr(file, "dtype", None), "names", None) or
                 getattr(file, "columns", None) or file.keys())
    col = lambda k: np.asarray(getattr(file[k], "filled", lambda v: file[k])(np.nan), dtype=float)   # noqa: E731
    if new_band_names is not None:
        assert len(new_band_names) == len(bands), \
            f"new_band_names length {len(new_band_names)} does not match bands length {len(bands)}. Cannot create uncertainty models."
    else:
        new_band_names = bands
    models = {}
    for band, new_name in zip(bands, new_band_names):
        if f"loc_depth_{band}" not in names:
            raise ValueError(f"Column loc_depth_{band} not found in the table.")
        mag = col(f"MAG_APER_{band}_aper_corr")
        flux = col(f"FLUX_APER_{band}_aper_corr_Jy")
        loc_depth = col(f"loc_depth_{band}")
        flux_err = UncertaintyModel.ab_to_jy(loc_depth) / 5
        if old:
            mag, flux = mag[:, 0], flux[:, 0]
            flux_err = flux          # (sic) the reference's old-format branch, noise_models.py:1236
        with np.errstate(all="ignore"):
            mag_err = (2.5 * flux_err) / (flux * np.log(10))
        mask = (mag != -99) & np.isfinite(mag) & (mag_err >= 0)
        base = {"return_noise": True, "error_type": "observed", "num_bins": 20}
        if model_class == "general":
            kw = dict(log_bins=False, upper_limits=True, treat_as_upper_limits_below=1, upper_limit_flux_behaviour=40,
                      upper_limit_flux_err_behaviour="sig_1")
            kw.update(base)
            kw.update(kwargs)
            model = GeneralEmpiricalUncertaintyModel(mag[mask], mag_err[mask], **kw)
        elif model_class == "depth":
            base.update(kwargs)
            model = DepthUncertaintyModel(float(np.nanmedian(loc_depth)), depth_sigma_level=5.0, **base)
        elif model_class == "asinh":
            base.update(kwargs)
            base["interpolation_flux_unit"] = "asinh"
            base["log_bins"] = True
            model = AsinhEmpiricalUncertaintyModel(flux, flux_err, **base)
        else:
            raise ValueError(f"Unknown model_class: {model_class}. Supported: 'general', 'depth', 'asinh'.")
        models[new_name] = model
    return models

