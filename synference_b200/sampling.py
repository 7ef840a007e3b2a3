"""Prior sampling and SFH-basis construction (host side of the path, kept API-compatible).

``draw_from_hypercube``  <- ``src/synference/library.py:1021-1115``
``load_hypercube_from_npy`` <- ``library.py:1118-1134``
``generate_sfh_basis``   <- ``library.py:1137-1334``
``generate_metallicity_distribution`` <- ``library.py:876-928``

The reference instantiates (and deep-copies) one SFH object per galaxy in a Python loop; here
the same inputs are lowered with vectorised numpy to an :class:`SFHArray`, which indexes like
the reference's object array but never creates per-galaxy objects unless asked to.
"""

from __future__ import annotations

import inspect
from typing import List, Optional

import numpy as np
from scipy.stats import qmc

from .cosmology import Planck18
from .parametric import SFH, SFH_MAX_PARAMS, SFHArray, ZDist, ZDistArray
from .units import Myr, Quantity, Unit, has_units, strip_units

__all__ = ["draw_from_hypercube", "load_hypercube_from_npy", "generate_sfh_basis",
           "generate_metallicity_distribution"]


def draw_from_hypercube(param_ranges, N: int = 1e6, model=qmc.LatinHypercube, rng=None,
                        unlog_keys: Optional[List[str]] = None):
    """Draw ``N`` samples from the hypercube spanned by ``param_ranges`` (dict name -> (lo, hi)).

    Output arrays are float32 (``library.py:1098``); keys in ``unlog_keys`` are raised to the power
    of ten and lose their ``log_`` prefix; ranges given as quantities keep their unit.
    """
    unlog_keys = unlog_keys or []
    sig = inspect.signature(model).parameters
    if "rng" in sig:
        kw = "rng"
    elif "seed" in sig:
        kw = "seed"
    else:
        raise ValueError("The model must accept either 'rng' or 'seed' as an argument.")
    sampler = model(d=len(param_ranges), **({kw: rng} if rng is not None else {}))
    sample = sampler.random(int(N))
    lows, highs, units = [], [], []
    for key, rng_pair in param_ranges.items():
        lo, hi = rng_pair[0], rng_pair[1]
        unit = None
        if has_units(rng_pair):            # (lo, hi) * Myr  -> one array quantity
            unit, lo, hi = rng_pair.units, float(rng_pair.value[0]), float(rng_pair.value[1])
        elif has_units(lo):
            unit, lo, hi = lo.units, float(lo.value), float(hi.value)
        assert lo < hi, f"Parameter range {lo} must be less than {hi}"
        lows.append(lo)
        highs.append(hi)
        units.append(unit)
    scaled = qmc.scale(sample, np.array(lows, dtype=float), np.array(highs, dtype=float))
    out = {}
    for i, key in enumerate(param_ranges.keys()):
        samples = scaled[:, i].astype(np.float32)
        if key in unlog_keys:
            samples = 10**samples
            key = key.replace("log_", "")
        if units[i] is not None:
            samples = Quantity(samples, units[i])
        if np.any(~np.isfinite(samples)):
            raise ValueError(f"Non-finite values found in samples for parameter '{key}'. "
                             "Check the parameter ranges and ensure they are valid.")
        out[key] = samples
    return out


def load_hypercube_from_npy(file_path: str):
    return np.load(file_path).astype(np.float32)


def generate_sfh_basis(sfh_type, sfh_param_names: List[str], sfh_param_arrays, redshifts,
                       sfh_param_units=None, max_redshift: float = 20, calculate_min_age: bool = False,
                       min_age_frac=0.001, cosmo=Planck18, iterate_redshifts: bool = False):
    """Build one SFH per galaxy; returns ``(SFHArray, redshifts)``.

    ``max_age = age(z) - age(max_redshift)`` (``library.py:1206``); parameters named ``*_norm`` are
    fractions of that age (``library.py:1287-1289``); callables receive ``max_age`` in Myr;
    ``sfh_timescale`` becomes ``max_age = min_age + sfh_timescale``; an explicit ``max_age`` is capped.
    """
    if isinstance(redshifts, dict):
        redshifts = redshifts["prior"].rvs(size=int(redshifts["size"]), loc=redshifts["min"],
                                           scale=redshifts["max"] - redshifts["min"])
    elif isinstance(redshifts, (float, int)):
        redshifts = np.array([redshifts], dtype=float)
        if not iterate_redshifts:
            redshifts = np.full(len(sfh_param_arrays[0]), redshifts[0])
    elif isinstance(redshifts, np.ndarray):
        pass
    else:
        raise ValueError("redshifts must be a dictionary, float/int, or numpy array")
    redshifts = np.asarray(strip_units(redshifts), dtype=np.float64)
    max_ages = np.asarray((cosmo.age(redshifts) - cosmo.age(max_redshift)).to("Myr").value, dtype=np.float64)

    if sfh_param_units is None:
        sfh_param_units = [None] * len(sfh_param_names)
    sfh_param_units = list(sfh_param_units)
    if isinstance(sfh_param_arrays, tuple):
        sfh_param_arrays = list(sfh_param_arrays)
    if isinstance(sfh_param_arrays, np.ndarray) and sfh_param_arrays.ndim == 2 \
            and sfh_param_arrays.shape[1] == len(sfh_param_names):
        sfh_param_arrays = [sfh_param_arrays[:, j] for j in range(sfh_param_arrays.shape[1])]
    cols = []
    for pos, param in enumerate(sfh_param_arrays):
        if has_units(param):
            sfh_param_units[pos] = Unit(str(param.units))
            param = param.value
        assert isinstance(sfh_param_units[pos], (Unit, type(None)))
        cols.append(param)
    n_rows = len(cols[0])

    if iterate_redshifts:  # every redshift x every parameter row
        z_idx = np.repeat(np.arange(len(redshifts)), n_rows)
        row_idx = np.tile(np.arange(n_rows), len(redshifts))
    else:
        assert len(redshifts) == n_rows, \
            "If iterate_redshifts is False, len(redshifts) must equal len(sfh_param_arrays)"
        z_idx = row_idx = np.arange(n_rows)
    mx_myr = max_ages[z_idx]
    values = {}
    for name, col, unit in zip(sfh_param_names, cols, sfh_param_units):
        if len(col) and callable(np.asarray(col, dtype=object).flat[0]):
            v = np.array([f(m) for f, m in zip(np.asarray(col, dtype=object)[row_idx], mx_myr)], dtype=float)
            v_unit = unit
        else:
            v = np.asarray(col, dtype=np.float64)[row_idx]
            v_unit = unit
            if name.endswith("_norm") and not iterate_redshifts:
                v, v_unit = v * mx_myr, Myr
                if unit is not None:
                    raise ValueError(f"'{name}' is a fraction of max_age and cannot carry a unit")
        values[name.replace("_norm", "") if not iterate_redshifts else name] = (v, v_unit)

    def in_yr(key):
        v, u = values[key]
        return v * (u.factor if u is not None and u.dimensions == "time" else 1.0)

    max_age_yr = mx_myr * 1.0e6
    if "sfh_timescale" in values:
        max_age_yr = in_yr("min_age") + in_yr("sfh_timescale")
        values.pop("sfh_timescale")
    if "max_age" in values:
        max_age_yr = np.minimum(mx_myr * 1.0e6, in_yr("max_age"))
        values.pop("max_age")
    rows = np.zeros((len(z_idx), SFH_MAX_PARAMS))
    rows[:, 1] = max_age_yr
    if "min_age" in values:
        rows[:, 0] = in_yr("min_age")
        values.pop("min_age")
    if sfh_type is SFH.Continuity:
        raise ValueError("Build Continuity populations with SFHArray / continuity_sfh_array")
    for key in values:
        if key not in sfh_type.param_names:
            raise TypeError(f"{sfh_type.__name__}() got an unexpected parameter '{key}'")
    for j, key in enumerate(sfh_type.param_names):
        if key not in values:
            raise TypeError(f"{sfh_type.__name__}() missing required parameter '{key}'")
        rows[:, 2 + j] = in_yr(key) if key in sfh_type.time_params else values[key][0]
    out_z = redshifts[z_idx]
    return SFHArray(sfh_type, rows, out_z), redshifts


def continuity_sfh_array(logsfr_ratios, agebins, redshifts=None):
    """Continuity (piecewise-constant) SFHs for a population.

    logsfr_ratios ``(N, n_b - 1)``; agebins ``(N, n_b, 2)`` or ``(n_b, 2)`` in log10(yr), as built by
    ``continuity_agebins`` in ``final_library_generation_multinode.py:193-259``.
    """
    r = np.asarray(logsfr_ratios, dtype=np.float64)
    ab = np.asarray(agebins, dtype=np.float64)
    n = r.shape[0]
    if ab.ndim == 2:
        ab = np.broadcast_to(ab, (n,) + ab.shape)
    nb = ab.shape[1]
    assert r.shape[1] == nb - 1 and 3 + 2 * nb <= SFH_MAX_PARAMS
    edges = np.concatenate([10.0 ** ab[:, :, 0], 10.0 ** ab[:, -1:, 1]], axis=1)
    edges[:, 0] = np.where(edges[:, 0] <= 1.0, 0.0, edges[:, 0])
    rows = np.zeros((n, SFH_MAX_PARAMS))
    rows[:, 0], rows[:, 1], rows[:, 2] = edges[:, 0], edges[:, -1], nb
    rows[:, 3:3 + nb + 1] = edges
    rows[:, 3 + nb + 1:3 + 2 * nb] = r
    return SFHArray(SFH.Continuity, rows, redshifts)


def generate_metallicity_distribution(zmet_dist=ZDist.DeltaConstant, zmet=None, **kwargs):
    """Population of metallicity distributions from parameter arrays (``library.py:876-928``)."""
    if zmet_dist is ZDist.DeltaConstant:
        if zmet is not None and "log10metallicity" not in kwargs and "metallicity" not in kwargs:
            kwargs["log10metallicity"] = zmet
        return ZDistArray.delta(metallicity=kwargs.get("metallicity"),
                                log10metallicity=kwargs.get("log10metallicity"))
    if zmet_dist is ZDist.Normal:
        return ZDistArray.normal(kwargs["mean"], kwargs["sigma"], log10=kwargs.get("log10", True))
    raise ValueError(f"Unsupported metallicity distribution {zmet_dist}")


def generate_sfh_grid(sfh_type, sfh_priors, redshift, max_redshift: float = 15, cosmo=Planck18):
    """Every combination of drawn redshifts and SFH-parameter samples (``library.py:742-873``).

    ``sfh_priors[name]`` = ``{"prior": scipy.stats distribution, "min", "max", "size", ["units"], ["name"],
    ["depends_on": "max_redshift"]}``; a parameter that depends on redshift is drawn per redshift with its upper bound capped
    at the age available there (``size // n_redshifts`` samples each).  Returns ``(SFHArray, param_combinations)`` with
    ``param_combinations[:, 0]`` the redshift -- the reference returns a list of SFH objects; the array materialises the
    same objects on indexing.  ``max_age = age(z) - age(max_redshift)`` as in ``generate_sfh_basis``."""
    if isinstance(redshift, dict):
        redshifts = redshift["prior"].rvs(size=int(redshift["size"]), loc=redshift["min"], scale=redshift["max"] - redshift["min"])
    else:
        redshifts = np.array([redshift], dtype=float)
    redshifts = np.asarray(redshifts, dtype=np.float64)
    max_ages = np.asarray((cosmo.age(redshifts) - cosmo.age(max_redshift)).to("Myr").value, dtype=np.float64)
    param_arrays, param_names, units = [redshifts], ["redshift"], {}
    for key, pd in sfh_priors.items():
        size, lo, hi = int(pd["size"]), pd["min"], pd["max"]
        unit = pd.get("units")
        if pd.get("depends_on") == "max_redshift":
            vals = []
            for mx in max_ages:
                top = min(hi, mx)
                vals.append(pd["prior"].rvs(size=size // len(redshifts), loc=lo, scale=top - lo))
            values = np.concatenate(vals)
        else:
            values = pd["prior"].rvs(size=size, loc=lo, scale=hi - lo)
        param_arrays.append(np.asarray(values, dtype=np.float64))
        param_names.append(pd.get("name", key))
        units[pd.get("name", key)] = unit
    mesh = np.meshgrid(*param_arrays, indexing="ij")
    combos = np.stack([m.reshape(-1) for m in mesh], axis=1)
    z = combos[:, 0]
    mx_myr = np.asarray((cosmo.age(z) - cosmo.age(max_redshift)).to("Myr").value, dtype=np.float64)
    rows = np.zeros((len(z), SFH_MAX_PARAMS))
    rows[:, 1] = mx_myr * 1.0e6
    cols = {name: combos[:, j + 1] for j, name in enumerate(param_names[1:])}
    for name in cols:
        if name not in sfh_type.param_names and name != "min_age":
            raise TypeError(f"{sfh_type.__name__}() got an unexpected parameter '{name}'")

    def in_yr(name):
        u = units.get(name)
        return cols[name] * (Unit(str(u)).factor if u is not None else 1.0)

    if "min_age" in cols:
        rows[:, 0] = in_yr("min_age")
    for j, name in enumerate(sfh_type.param_names):
        if name not in cols:
            raise TypeError(f"{sfh_type.__name__}() missing required parameter '{name}'")
        rows[:, 2 + j] = in_yr(name) if name in sfh_type.time_params else cols[name]
    return SFHArray(sfh_type, rows, z), combos


def generate_emission_models(emission_model, varying_params: dict, grid, fixed_params: dict = None):
    """One emission model per combination of drawn parameter values (``library.py:931-1018``): returns
    ``(models, {name: values per model})``.  ``varying_params[name]`` = ``{"prior", "min", "max", "size", ["units"]}``."""
    arrays = []
    for key, pd in varying_params.items():
        args = {k: v for k, v in pd.items() if k not in ("units", "name", "prior", "min", "max")}
        if "min" in pd:
            args["loc"] = pd["min"]
        if "max" in pd:
            args["scale"] = pd["max"] - pd["min"]
        arrays.append(np.asarray(pd["prior"].rvs(**args), dtype=np.float64))
    mesh = np.meshgrid(*arrays, indexing="ij") if arrays else []
    combos = np.stack([m.reshape(-1) for m in mesh], axis=1) if arrays else np.zeros((1, 0))
    fixed_params = fixed_params or {}
    models, out_params = [], {k: [] for k in varying_params}
    for row in combos:
        kw = {}
        for j, key in enumerate(varying_params):
            unit = varying_params[key].get("units")
            kw[key] = row[j] * unit if unit is not None else float(row[j])
            out_params[key].append(kw[key])
        kw.update(fixed_params)
        models.append(emission_model(grid=grid, **kw))
    return models, out_params

