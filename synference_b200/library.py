"""Library builders and on-the-fly simulator with the reference's public names and file contract.

Mirrors (paths relative to the reference repository, ``src/synference/library.py``):
  ``create_galaxy``                       :1340-1424
  ``GalaxyBasis``                         :1497-3183  (``process_galaxies`` :2447, ``create_mock_library`` :3022)
  ``CombinedBasis``                       :3185-4919  (``process_bases`` :3291, ``load_bases`` :3385,
                                          ``create_full_library`` :4435, ``save_library`` :4031,
                                          ``_validate_library`` :3976, ``create_spectral_grid`` :4887)
  ``GalaxySimulator``                     :4922-6001  (``simulate`` :5553, ``_scatter`` :5906, ``_normalize`` :5866)

What is different by design: no per-galaxy Python objects and no Synthesizer ``Pipeline``.  A basis
is lowered once to a struct-of-arrays :class:`GalaxyParams`; every batch goes through one call of the
CUDA path (:class:`synference_b200.engine.SynthEngine`); the per-galaxy Python loops of
``create_full_library`` are array expressions.  Galaxies are sharded contiguously over ranks with the
reference's rule (``library.py:3127-3138``).
"""

from __future__ import annotations

import copy
import os
from concurrent.futures import ThreadPoolExecutor
from datetime import datetime
from typing import Dict, List, Optional, Union

import numpy as np

from . import distributed as _dist
from . import supplementary as _supp
from .cosmology import Planck18
from .engine import GalaxyParams, SynthEngine
from .igm import Inoue14
from .parametric import (EmissionModel, Grid, Instrument, SFHArray, ZDistArray, _SFHCommon, _ZDistCommon,
                         pack_sfh, pack_zdist)
from .sampling import draw_from_hypercube, generate_sfh_basis  # noqa: F401  (re-exported)
from .units import Quantity, has_units, strip_units
from .utils import logger, read_container, write_container

library_folder = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libraries")

UNIT_DICT = {
    "log10metallicity": "log10(Zmet)", "metallicity": "Zmet", "av": "mag", "tau_v": "mag",
    "tau_v_ism": "mag", "tau_v_birth": "mag", "weight_fraction": "dimensionless",
    "log_sfr": "log10(Msun/yr)", "sfr": "Msun/yr", "log_stellar_mass": "log10(Msun)",
    "log_surviving_mass": "log10(Msun)", "stellar_mass": "Msun",
}

# emitter parameters that change the spectrum and are understood by the CUDA path
_SPECTRAL_EMITTER_PARAMS = ("tau_v", "fesc")
# emitter parameters of the reference scripts that would change the spectrum but are not implemented
_UNSUPPORTED_EMITTER_PARAMS = ("slope", "fesc_lya", "fesc_ly_alpha", "dust_bump_amplitude",
                               "tau_v_ism", "tau_v_birth")   # ... unless the emission model / dust curve names them


def create_galaxy(sfh, redshift, metal_dist, grid, log_stellar_masses=9, **galaxy_kwargs):
    """A single-galaxy description (``library.py:1340-1424``); a plain record, not a Synthesizer object."""
    assert not has_units(log_stellar_masses), \
        "log_stellar_masses must be a float or list of floats, not a unyt array"
    return {"sfh": sfh, "redshift": float(redshift), "metal_dist": metal_dist,
            "log_stellar_mass": log_stellar_masses, "params": dict(galaxy_kwargs)}


def _rank_size(multi_node):
    return _dist.rank_world() if multi_node else (0, 1)


class GalaxyBasis:
    """A population of model galaxies sharing one grid / emission model / instrument."""

    def __init__(self, model_name: str, redshifts, grid: Grid, emission_model: EmissionModel, sfhs,
                 metal_dists, log_stellar_masses=None, galaxy_params: dict = None,
                 alt_parametrizations: Dict[str, tuple] = None, cosmo=Planck18, instrument: Instrument = None,
                 redshift_dependent_sfh: bool = False, params_to_ignore: List[str] = None,
                 build_library: bool = False) -> None:
        galaxy_params = {} if galaxy_params is None else galaxy_params
        self.model_name = model_name
        self.grid = grid
        self.emission_model = emission_model
        self.galaxy_params = galaxy_params
        self.alt_parametrizations = alt_parametrizations or {}
        self.cosmo = cosmo
        self.instrument = instrument
        self.redshift_dependent_sfh = redshift_dependent_sfh
        self.log_stellar_masses = log_stellar_masses
        self.params_to_ignore = params_to_ignore or []
        self.build_library = build_library
        self.galaxies = []
        self.per_particle = False
        if isinstance(sfhs, _SFHCommon):
            sfhs = [sfhs]
        if isinstance(metal_dists, _ZDistCommon):
            metal_dists = [metal_dists]
        if isinstance(redshifts, (float, int)) and not build_library:
            redshifts = np.full(len(sfhs), redshifts)
        self.sfhs, self.metal_dists = sfhs, metal_dists
        self.redshifts = np.asarray(strip_units(redshifts), dtype=np.float64)
        for key in list(galaxy_params.keys()):
            value = galaxy_params[key]
            if isinstance(value, dict):
                if key in ("bh", "gas"):
                    raise NotImplementedError("black-hole / gas emitters are outside the stellar hot path")
                self.galaxy_params[key] = self.process_priors(value)
        dust = getattr(emission_model, "dust_curve", None)
        dust_names = {getattr(dust, "slope_name", None), getattr(dust, "ampl_name", None),
                      getattr(emission_model, "lya_name", None), getattr(emission_model, "tau_v_ism_name", None),
                      getattr(emission_model, "tau_v_birth_name", None)} - {None}
        for name in dust_names:
            if name not in galaxy_params:
                raise ValueError(f"the emission model reads '{name}' per galaxy, but galaxy_params does not provide it")
        for key in galaxy_params:
            if key in _UNSUPPORTED_EMITTER_PARAMS and key not in dust_names:
                raise NotImplementedError(
                    f"per-galaxy emitter parameter '{key}' is not implemented in the CUDA path yet "
                    "(SURVEY 8f-1); 'tau_v', 'fesc' and the dust curve's named slope / bump amplitude vary per galaxy")
        per_gal = getattr(emission_model, "fesc_per_galaxy", False)
        if per_gal and emission_model.fesc_name not in galaxy_params:
            raise ValueError(f"the emission model reads fesc from the per-galaxy parameter "
                             f"'{emission_model.fesc_name}', which galaxy_params does not provide")
        if not per_gal and "fesc" in galaxy_params:
            raise ValueError("galaxy_params has 'fesc' but the emission model uses a global fesc; construct it with "
                             "fesc='fesc' (the reference's string convention for per-emitter parameters)")
        if not build_library:
            logger.info("Generating library directly from provided parameter samples.")
        elif redshift_dependent_sfh:
            for sfh in self.sfhs:
                if not hasattr(sfh, "redshift"):
                    raise ValueError("SFH must have a redshift attr if redshift_dependent_sfh==True")
                if sfh.redshift not in self.redshifts:
                    raise ValueError(f"SFH redshift {sfh.redshift} not in redshifts array")
        self.params: Optional[GalaxyParams] = None
        self.all_parameters: Dict[str, np.ndarray] = {}
        self.varying_param_names: List[str] = []
        self.fixed_param_names: List[str] = []
        self.fixed_param_values: list = []
        self.fixed_param_units: List[str] = []
        self._engines: Dict[str, SynthEngine] = {}

    # ---------------------------------------------------------------------------------------
    def process_priors(self, prior_dict):
        """Draw a per-galaxy parameter from a scipy.stats prior description (``library.py:1653-1692``)."""
        assert isinstance(prior_dict, dict) and "prior" in prior_dict and "size" in prior_dict
        kw = {k: v for k, v in prior_dict.items() if k not in ("prior", "size", "units", "name")}
        values = prior_dict["prior"].rvs(size=int(prior_dict["size"]), **kw)
        if prior_dict.get("units") is not None:
            values = Quantity(values, str(prior_dict["units"]))
        return values

    @staticmethod
    def _sfh_param_table(sfhs, n):
        """``{name: (N,) array}`` of the SFH constructor parameters (what ``sfh.parameters`` holds)."""
        sfh_type, rows = pack_sfh(sfhs)
        if rows.shape[0] == 1 and n > 1:
            rows = np.repeat(rows, n, axis=0)
        cls = sfhs.sfh_type if isinstance(sfhs, SFHArray) else type(list(sfhs)[0])
        out = {"max_age": rows[:, 1]}
        if np.any(rows[:, 0] != 0):
            out["min_age"] = rows[:, 0]
        for j, name in enumerate(getattr(cls, "param_names", ())):
            out[name] = rows[:, 2 + j]
        if cls.__name__ == "_Continuity":
            nb = int(rows[0, 2])
            for j in range(nb - 1):
                out[f"logsfr_ratio_{j}"] = rows[:, 3 + nb + 1 + j]
        return out

    def _finalise_parameters(self, table: Dict[str, np.ndarray]):
        """Split parameters into varying / fixed exactly like ``library.py:2382-2441``."""
        for key, (new_key, func) in self.alt_parametrizations.items():
            if key in table:
                if isinstance(new_key, str):
                    table[new_key] = np.asarray(func(table))
                else:
                    for k in new_key:
                        table[k] = np.asarray(func(k, table))
                table.pop(key)
        self.all_parameters = table
        self.varying_param_names, self.fixed_param_names = [], []
        self.fixed_param_values, self.fixed_param_units = [], []
        for key, value in table.items():
            if key in self.params_to_ignore:
                continue
            v = np.asarray(strip_units(value), dtype=float)
            if v.size and bool((v == v.flat[0]).all()):   # one distinct value (np.unique sorts: 10x the time for 1M rows)
                self.fixed_param_names.append(key)
                self.fixed_param_values.append(float(v.flat[0]))
                self.fixed_param_units.append(str(value.units) if has_units(value) else "")
            else:
                self.varying_param_names.append(key)

    def _lower(self, redshifts, sfhs, metal_dists, galaxy_params: Dict[str, np.ndarray], log_mass=None):
        n = len(redshifts)
        tau_v = galaxy_params.get(getattr(self.emission_model, "tau_v_ism_name", None) or "tau_v")
        self.params = GalaxyParams.from_objects(redshifts, sfhs, metal_dists, log_mass=log_mass, tau_v=tau_v)
        table = {k: np.broadcast_to(np.asarray(strip_units(v), dtype=float), (n,)).copy()
                 for k, v in galaxy_params.items()}
        table["redshift"] = np.asarray(redshifts, dtype=float)
        table.update(self._sfh_param_table(sfhs, n))
        zt, zv, zs = pack_zdist(metal_dists)
        zv = np.broadcast_to(zv, (n,)).copy()
        if zt == 0:
            table["metallicity"] = zv
        elif zt == 1:
            table["log10metallicity"] = zv
        else:
            table["mean"], table["sigma"] = zv, np.broadcast_to(zs, (n,)).copy()
        self._finalise_parameters(table)
        self.galaxies = _LazyGalaxyList(self)
        return self.galaxies

    def _create_matched_galaxies(self, log_base_masses=9, galaxies_mask=None, n_proc: int = 1):
        """One galaxy per (SFH, redshift, metallicity distribution) triple (``library.py:2263-2445``)."""
        n = len(self.sfhs)
        if len(self.metal_dists) == 1 and n > 1:
            md = self.metal_dists if isinstance(self.metal_dists, ZDistArray) else \
                [self.metal_dists[0]] * n
            self.metal_dists = md if not isinstance(md, ZDistArray) else ZDistArray(
                md.type_id, np.repeat(md.value, n), np.repeat(md.sigma, n))
        assert n == len(self.redshifts), \
            f"If iterate_redshifts is False, sfhs and redshifts must be the same length, got {n} and {len(self.redshifts)}"
        assert n == len(self.metal_dists), \
            f"sfhs and metal_dists must be the same length, got {n} and {len(self.metal_dists)}"
        gp = {}
        for k, v in self.galaxy_params.items():
            if isinstance(v, (list, np.ndarray)) or has_units(v) and np.ndim(v) > 0:
                assert len(v) == n, f"All varying parameters must be the same length, got {n} and {len(v)}"
            gp[k] = v
        sel = slice(None) if galaxies_mask is None else np.asarray(galaxies_mask, dtype=bool)
        if galaxies_mask is not None:
            assert len(galaxies_mask) == n, "galaxies_mask must be the same length as sfhs"
            gp = {k: (np.asarray(strip_units(v))[sel] if np.ndim(v) > 0 else v) for k, v in gp.items()}
        sfhs = self.sfhs[sel] if isinstance(self.sfhs, (SFHArray, np.ndarray)) else \
            [s for s, m in zip(self.sfhs, np.ones(n, bool) if galaxies_mask is None else sel) if m]
        zds = self.metal_dists[sel] if isinstance(self.metal_dists, (ZDistArray, np.ndarray)) else \
            [s for s, m in zip(self.metal_dists, np.ones(n, bool) if galaxies_mask is None else sel) if m]
        return self._lower(self.redshifts[sel], sfhs, zds, gp)

    def _create_galaxies(self, log_base_masses=9):
        """Every combination of redshift x SFH x metallicity distribution x varying galaxy parameters,
        in the loop order of ``library.py:1694-1873``."""
        if not self.build_library:
            raise ValueError("You probably meant to call_create_matched_galaxies instead.")
        varying = {k: np.asarray(strip_units(v), dtype=float) for k, v in self.galaxy_params.items()
                   if isinstance(v, (list, np.ndarray))}
        fixed = {k: v for k, v in self.galaxy_params.items() if k not in varying}
        if varying:
            combos = np.array(np.meshgrid(*varying.values())).T.reshape(-1, len(varying))
        else:
            combos = np.zeros((1, 0))
        sfh_list = list(self.sfhs)
        zd_list = list(self.metal_dists)
        zs, sf, zd, cm = [], [], [], []
        for z in self.redshifts:
            for s in (sfh_list if not self.redshift_dependent_sfh else
                      [x for x in sfh_list if x.redshift == z]):
                for d in zd_list:
                    for row in combos:
                        zs.append(z); sf.append(s); zd.append(d); cm.append(row)
        cm = np.array(cm).reshape(len(zs), len(varying))
        gp = {k: cm[:, j] for j, k in enumerate(varying)}
        gp.update(fixed)
        out = self._lower(np.array(zs), sf, zd, gp)
        # varying-parameter combinations must be unique (library.py:1842-1858)
        if self.varying_param_names:
            mat = np.stack([np.asarray(strip_units(self.all_parameters[k]), float) for k in self.varying_param_names], 1)
            if np.unique(mat, axis=0).shape[0] != mat.shape[0]:
                raise ValueError("Varying parameters are not unique across galaxies. Check your input parameters.")
        return out

    def create_galaxies_optimized(self, galaxies_mask=None, varying_param_names=None, log_base_masses=9,
                                  fixed_params=None, n_proc=1):
        return self._create_matched_galaxies(log_base_masses, galaxies_mask, n_proc)

    # ---------------------------------------------------------------------------------------
    _MAX_ENGINES = 3     # device models kept alive per basis (one per emission key in use; least recently used is closed)

    def _engine(self, emission_key, igm=Inoue14, max_batch=40_000) -> SynthEngine:
        tag = f"{emission_key}|{bool(igm)}|{max_batch}"
        if tag in self._engines:
            self._engines[tag] = self._engines.pop(tag)          # most recently used last
            return self._engines[tag]
        while len(self._engines) >= self._MAX_ENGINES:
            self._engines.pop(next(iter(self._engines))).close()
        self._engines[tag] = SynthEngine(self.grid, self.emission_model, emission_key,
                                         self.instrument.filters, cosmo=self.cosmo, igm=bool(igm),
                                         max_batch=max_batch, device=_dist.local_device())
        return self._engines[tag]

    def process_galaxies(self, galaxies=None, out_name: str = "auto", out_dir: str = "internal", n_proc: int = 4,
                         verbose: int = 1, save: bool = True, emission_model_keys=None,
                         batch_galaxies: bool = True, batch_size: int = 40_000, overwrite: bool = False,
                         multi_node: bool = False, spectra_to_save: list = None, em_lines_to_save: list = None,
                         igm=Inoue14, **extra_analysis_functions):
        """Synthesise the basis in batches and write the pipeline files (``library.py:2447-2694``).

        Returns the dict ``{"photometry": {key: (N, n_filt)}, "spectra": {...}}`` of what was computed.
        Each batch is one pass of the CUDA path at the base mass; per-batch files are the resume unit:
        an existing ``<out_name>_<i>.hdf5`` is skipped unless ``overwrite``.
        """
        _supp.check_supported(extra_analysis_functions)     # history-based ones only; others raise NotImplementedError
        if em_lines_to_save:
            raise NotImplementedError("emission-line outputs are outside the hot path")
        if self.params is None:
            raise ValueError("create the galaxies first (_create_matched_galaxies / _create_galaxies)")
        self.emission_model.set_per_particle(self.per_particle)
        keys = list(emission_model_keys) if emission_model_keys is not None else ["total"]
        if emission_model_keys is not None:
            self.emission_model.save_spectra(*keys, *(spectra_to_save or []))
        n_gal = len(self.params)
        if not batch_galaxies:
            batch_size = n_gal
        n_batches = int(np.ceil(n_gal / batch_size))
        if out_dir == "internal":
            out_dir = library_folder
        if out_name == "auto":
            out_name = self.model_name
        if not out_name.endswith(".hdf5"):
            out_name += ".hdf5"
        fullpath = os.path.join(out_dir, out_name)
        if save:
            os.makedirs(out_dir, exist_ok=True)
        spectra_keys = set(spectra_to_save or [])
        results = {"photometry": {k: [] for k in keys}, "spectra": {k: [] for k in keys if k in spectra_keys}}
        rank, size = _rank_size(multi_node)
        # pipeline files are written by a background thread (the npz / HDF5 write releases the GIL in its I/O), so the
        # next batch's kernels run while the previous batch's file goes to disk; what was written is also kept in
        # memory for CombinedBasis.load_bases, which otherwise re-reads every file it has just produced
        writer = ThreadPoolExecutor(max_workers=self._WRITER_THREADS) if save else None
        pending = []
        self._pipeline_cache = {"out_dir": os.path.abspath(out_dir), "parts": [], "complete": True}
        full_props = {name: np.asarray(strip_units(arr), dtype=float) for name, arr in self.all_parameters.items()}
        # create_mock_library's single-base build: the kernels also write the mass-scaled float64 block of every batch into
        # its columns of the library's (n_filters, n_galaxies) matrix (sb2_params.scaled_ld), so that create_full_library has
        # no cast / multiply / transpose left to do on the host
        sink = getattr(self, "_library_sink", None) if len(keys) == 1 else None
        for batch_i in range(n_batches):
            sl = slice(batch_i * batch_size, min(n_gal, (batch_i + 1) * batch_size))
            final = fullpath if n_batches == 1 else fullpath.replace(".hdf5", f"_{batch_i + 1}.hdf5")
            if multi_node and size > 1:
                final = final.replace(".hdf5", f"_rank{rank}.hdf5")
            if save and os.path.exists(final) and not overwrite:
                logger.warning(f"Skipping batch {batch_i + 1} as {final} already exists.")
                # keep the returned arrays aligned with the parameters: read the rows this file holds back
                self._pipeline_cache["complete"] = False
                old, _ = read_container(final)
                label = self.instrument.label
                for key in keys:
                    cols = [old.get(f"Galaxies/Stars/Photometry/Fluxes/{key}/{label}/{c}") for c in self.instrument.filters.filter_codes]
                    if all(c is not None for c in cols):
                        results["photometry"][key].append(np.stack(cols, 1).astype(np.float32))
                    skey = f"Galaxies/Stars/Spectra/SpectralFluxDensities/{key}"
                    if key in results["spectra"] and skey in old:
                        results["spectra"][key].append(old[skey])
                continue
            start = datetime.now()
            supp_units = {}
            p = self.params.slice(sl)
            # the pipeline always runs at the base mass (library.py:3217): no mass scaling here
            p.log_mass = None
            datasets = {}
            phot_blocks = []
            for key in keys:
                eng = self._engine(key, igm=igm, max_batch=max(batch_size, 1))
                if getattr(self.emission_model, "fesc_per_galaxy", False):
                    fesc = np.asarray(strip_units(self.all_parameters[self.emission_model.fesc_name]), dtype=float)[sl]
                    p.coef_att, p.coef_unatt = self.emission_model.coefficients(key, fesc)
                dust = getattr(self.emission_model, "dust_curve", None)
                for attr, name in (("dust_slope", getattr(dust, "slope_name", None)), ("dust_ampl", getattr(dust, "ampl_name", None))):
                    if name is not None:
                        setattr(p, attr, np.asarray(strip_units(self.all_parameters[name]), dtype=float)[sl])
                birth = getattr(self.emission_model, "tau_v_birth_name", None)
                p.tau_v_birth = None
                if birth is not None and self.emission_model.two_screens(key):
                    p.tau_v_birth = np.asarray(strip_units(self.all_parameters[birth]), dtype=float)[sl]
                lya_name = getattr(self.emission_model, "lya_name", None)
                p.fesc_lya = None
                if lya_name is not None and self.emission_model.lya_line(key) is not None:
                    p.fesc_lya = np.asarray(strip_units(self.all_parameters[lya_name]), dtype=float)[sl]
                if sink is not None and not eng.general:
                    p.log_mass = sink["log_mass"][sl]
                    flux = eng.photometry(p, scaled=False, library_out=(sink["matrix"], sl.start))
                    p.log_mass = None
                    sink["filled"] += sl.stop - sl.start
                else:
                    sink = None
                    flux = eng.photometry(p, scaled=False)
                results["photometry"][key].append(flux)
                if extra_analysis_functions and key == keys[0]:
                    # by-products of the weights (library.py:2593-2601 stores callback results as supp_<name>)
                    supp = _supp.evaluate(extra_analysis_functions, eng.sfzh(p) * 10.0 ** 9, self.grid.log10ages,
                                          p.redshift, self.cosmo,
                                          band_flux=lambda lo, hi, eng=eng, p=p: eng.rest_band_flux(p, lo, hi))
                    for name, (vals, units) in supp.items():
                        datasets[f"Galaxies/supp_{name}"] = vals
                        supp_units[name] = units
                label = self.instrument.label
                # the per-filter float64 columns of the pipeline file are cut from this matrix by the writer thread
                phot_blocks.append((f"Galaxies/Stars/Photometry/Fluxes/{key}/{label}/", list(eng.filter_codes), flux))
                if key in spectra_keys:
                    spec = eng.spectra(p)
                    results["spectra"][key].append(spec)
                    datasets[f"Galaxies/Stars/Spectra/SpectralFluxDensities/{key}"] = spec
            elapsed = datetime.now() - start
            logger.info(f"Pipeline (CUDA) took {elapsed} for {sl.stop - sl.start} galaxies.")
            if save:
                for name, arr in full_props.items():
                    datasets[f"Galaxies/{name}"] = arr[sl]
                datasets["Galaxies/mass"] = np.full(sl.stop - sl.start, 10.0 ** 9)
                datasets["Wavelengths"] = np.asarray(self.grid.lam)
                attrs = {"varying_param_names": list(self.varying_param_names),
                         "fixed_param_names": list(self.fixed_param_names),
                         "fixed_param_values": list(self.fixed_param_values),
                         "fixed_param_units": list(self.fixed_param_units),
                         "model_name": self.model_name, "grid_name": self.grid.grid_name,
                         "grid_dir": str(self.grid.grid_dir), "date_created": str(datetime.now()),
                         "pipeline_time": str(elapsed), "WavelengthUnits": "Angstrom",
                         "FilterCodes": list(self.instrument.filters.filter_codes),
                         "InstrumentLabel": self.instrument.label, "batch": batch_i + 1, "n_batches": n_batches,
                         "rank": rank, "world_size": size, "galaxy_start": sl.start, "galaxy_stop": sl.stop,
                         "supp_names": list(supp_units), "supp_units": [supp_units[k] for k in supp_units]}
                self._pipeline_cache["parts"].append((os.path.basename(final), datasets, attrs, phot_blocks))
                self._pipeline_cache.setdefault("phot", {}).setdefault(keys[0], []).append(results["photometry"][keys[0]][-1])
                pending.append(writer.submit(_write_pipeline_part, final, datasets, attrs, phot_blocks))
                logger.info(f"Writing pipeline to disk at {final}.")
        # the files keep being written while the caller goes on (create_mock_library compiles the library from the in-memory
        # copy meanwhile); wait_for_writes() joins them and re-raises a failed write
        if sink is not None and save and self._pipeline_cache["complete"] and sink["filled"] == sink["matrix"].shape[1]:
            self._pipeline_cache["scaled_matrix"] = {keys[0]: (sink["matrix"], sink["log_mass"])}
        self._pipeline_cache["full_props"] = full_props
        self._pending_writes = (writer, pending)
        if not getattr(self, "_defer_write_join", False):
            self.wait_for_writes()
        else:
            return None      # create_mock_library's internal call: nobody reads the concatenated copies
        return {k: {kk: (np.concatenate(vv) if vv else None) for kk, vv in v.items()} for k, v in results.items()}

    _WRITER_THREADS = 4   # file writes release the GIL (crc32 and write of whole arrays): batches go to disk side by side

    def wait_for_writes(self):
        """Block until the pipeline files of the last process_galaxies call are on disk."""
        writer, pending = getattr(self, "_pending_writes", (None, []))
        self._pending_writes = (None, [])
        for fut in pending:
            fut.result()
        if writer is not None:
            writer.shutdown()

    def process_base(self, out_name, log_stellar_masses=9, emission_model_key="total", out_dir=library_folder,
                     n_proc=6, overwrite=False, verbose=False, batch_size=40_000, multi_node=False, **kw):
        """Create the galaxies of this basis and run them (``library.py:2939-3020``)."""
        if self.build_library:
            self._create_galaxies(log_stellar_masses)
        else:
            self._create_matched_galaxies(log_stellar_masses)
        return self.process_galaxies(out_name=out_name, out_dir=out_dir, overwrite=overwrite,
                                     emission_model_keys=[emission_model_key], batch_size=batch_size,
                                     multi_node=multi_node, **kw)

    def create_mock_library(self, out_name, log_stellar_masses=None, emission_model_key: str = "total",
                            out_dir: str = library_folder, n_proc: int = 6, overwrite=False, verbose=False,
                            batch_size: int = 40_000, parameter_transforms_to_save=None, cat_type="photometry",
                            compile_grid: bool = True, multi_node: bool = False, spectra_to_save=None,
                            em_lines_to_save=None, **extra_analysis_functions):
        """Convenience wrapper: basis -> CombinedBasis -> process -> compile -> store model
        (``library.py:3022-3183``)."""
        if log_stellar_masses is None:
            assert self.log_stellar_masses is not None, \
                "log_stellar_masses must be provided or set in the GalaxyBasis"
            log_stellar_masses = self.log_stellar_masses
        assert not has_units(log_stellar_masses), "log_stellar_masses must be not be a unyt_array"
        combined = CombinedBasis(bases=[self], log_stellar_masses=log_stellar_masses, redshifts=self.redshifts,
                                 base_emission_model_keys=[emission_model_key], combination_weights=None,
                                 out_name=out_name, out_dir=out_dir, draw_parameter_combinations=False)
        galaxy_mask = None
        if multi_node:
            rank, size = _dist.rank_world()
            total = len(combined.redshifts)
            start, end = _dist.shard_bounds(total, rank, size)
            galaxy_mask = np.zeros(total, dtype=bool)
            galaxy_mask[start:end] = True
            logger.info(f"Node {rank} processing galaxies from {start} to {end}.")
        if cat_type not in ("photometry", "spectra"):
            raise ValueError(f"Unknown catalog type: {cat_type}. Use 'photometry' or 'spectra'.")
        if cat_type == "spectra" and not spectra_to_save:
            spectra_to_save = [emission_model_key]
        self._defer_write_join = bool(compile_grid)      # the library is compiled from memory while the files are written
        self._library_sink = None
        lm = np.asarray(log_stellar_masses, dtype=float)
        if compile_grid and cat_type == "photometry" and lm.ndim == 1 and lm.size == len(combined.redshifts) and not self.build_library:
            if galaxy_mask is not None:
                lm = lm[galaxy_mask]
            self._library_sink = {"matrix": np.empty((len(self.instrument.filters.filter_codes), lm.size)),
                                  "log_mass": np.ascontiguousarray(lm), "filled": 0}
        try:
            combined.process_bases(n_proc=n_proc, overwrite=overwrite, verbose=verbose, batch_size=batch_size,
                                   multi_node=multi_node, galaxies_mask=galaxy_mask, spectra_to_save=spectra_to_save,
                                   em_lines_to_save=em_lines_to_save, **extra_analysis_functions)
        finally:
            self._defer_write_join = False
            self._library_sink = None
        if not compile_grid:
            self.wait_for_writes()
        if compile_grid:
            logger.info("Compiling the library after processing bases.")
            # the Model/ block (library.py:2017-2132) travels with the library's single write
            combined._extra_datasets, combined._extra_attrs = self._model_block(
                {"emission_model_key": emission_model_key, "timestamp": datetime.now().isoformat(), "cat_type": cat_type},
                parameter_transforms_to_save)
            combined._save_executor = ThreadPoolExecutor(max_workers=1)   # the library file is written beside the pipeline files
            try:
                if cat_type == "photometry":
                    combined.create_library(overwrite=overwrite)
                else:
                    combined.create_spectral_grid(overwrite=overwrite)
            finally:
                ex, combined._save_executor = combined._save_executor, None
                pending_save, combined._pending_save = getattr(combined, "_pending_save", None), None
                try:
                    if pending_save is not None:
                        pending_save.result()
                finally:
                    ex.shutdown()
            self.wait_for_writes()
            logger.info("Processed the bases and saved the output.")
            return combined

    def _store_model(self, model_path, other_info=None, parameter_transforms_to_save=None):
        """Record what is needed to rebuild the simulator next to the library (``library.py:2017-2132``).  (Rewrites an
        existing library file; ``create_mock_library`` hands the same block to ``save_library`` instead, so that the
        library is written once.)"""
        m_data, m_attrs = self._model_block(other_info, parameter_transforms_to_save)
        if os.path.exists(model_path):
            data, attrs = read_container(model_path)
            data = {k: v for k, v in data.items() if not k.startswith("Model/")}
            attrs = {k: v for k, v in attrs.items() if not k.startswith("Model")}
            data.update(m_data)
            attrs.update(m_attrs)
            write_container(model_path, data, attrs, compress=_library_compression())
        else:
            write_container(model_path, m_data, m_attrs, compress=False)
        return True

    def _model_block(self, other_info=None, parameter_transforms_to_save=None):
        """The ``Model`` group of a library file in the reference's layout (``library.py:2017-2132``) as
        ``(datasets, attributes)`` for :func:`write_container` (``"Model/EmissionModel@name"`` is the attribute ``name`` of
        the group ``Model/EmissionModel``): grid name / directory, emission model class + fixed parameters + dust law +
        dust emission, cosmology, the instrument's filters on their wavelength axis, SFH / metallicity-distribution class
        names, emitter parameters, varying / fixed parameter bookkeeping, the source of saved parameter transforms."""
        from inspect import getsource
        em = self.emission_model
        dust = em.dust_curve
        A, D = {}, {}
        A["Model@grid_name"] = self.grid.grid_name
        A["Model@grid_dir"] = str(self.grid.grid_dir)
        A["Model/EmissionModel@name"] = type(em).__name__
        keys, vals = ["fesc", "fesc_ly_alpha"], [em.fesc, em.fesc_ly_alpha]
        if getattr(em, "tau_v_birth_name", None) is not None:       # two screens: the names the emitter parameters travel under
            keys += ["tau_v_ism", "tau_v_birth", "age_pivot"]
            vals += [em.tau_v_ism_name, em.tau_v_birth_name, em.age_pivot]
        else:
            keys.append("tau_v")
            vals.append(em.tau_v)
        A["Model/EmissionModel@parameter_keys"] = keys
        A["Model/EmissionModel@parameter_values"] = [str(v) for v in vals]     # a name (per-galaxy parameter) or a number
        A["Model/EmissionModel@parameter_units"] = [""] * len(keys)
        if dust is not None:
            A["Model/EmissionModel@dust_law"] = dust.name
            dp = dict(dust.params)
            for nm, attr in (("slope", "slope_name"), ("ampl", "ampl_name")):
                if getattr(dust, attr, None) is not None:
                    dp[nm] = getattr(dust, attr)
            A["Model/EmissionModel@dust_attenuation_keys"] = list(dp)
            A["Model/EmissionModel@dust_attenuation_values"] = [str(v) for v in dp.values()]
            A["Model/EmissionModel@dust_attenuation_units"] = ["um" if k in ("cent_lam", "gamma") else "" for k in dp]
            birth = getattr(em, "dust_curve_birth", None)
            if birth is not None:
                A["Model/EmissionModel@dust_law_birth"] = birth.name
                A["Model/EmissionModel@dust_attenuation_birth_keys"] = list(birth.params)
                A["Model/EmissionModel@dust_attenuation_birth_values"] = [str(v) for v in birth.params.values()]
        if em.dust_emission is not None:
            A["Model/EmissionModel@dust_emission"] = type(em.dust_emission).__name__
            A["Model/EmissionModel@dust_emission_keys"] = ["temperature", "emissivity"]
            A["Model/EmissionModel@dust_emission_values"] = [em.dust_emission.temperature, em.dust_emission.emissivity]
            A["Model/EmissionModel@dust_emission_units"] = ["K", ""]
        A["Model@cosmology"] = repr(self.cosmo)
        A["Model@instrument"] = self.instrument.label if self.instrument else "None"
        if self.instrument:
            A["Model@filters"] = list(self.instrument.filters.filter_codes)
            A["Model/Instrument@label"] = self.instrument.label
            fc = self.instrument.filters
            lam0 = fc.lam if fc.lam is not None else fc.filters[0].lam
            D["Model/Instrument/Filters/Header/Wavelengths"] = np.asarray(strip_units(lam0), dtype=np.float64)
            A["Model/Instrument/Filters/Header@filter_codes"] = list(fc.filter_codes)
            for f in fc.filters:
                D[f"Model/Instrument/Filters/{f.filter_code}/Transmission"] = np.asarray(f.t, dtype=np.float64)
                if fc.lam is None:
                    D[f"Model/Instrument/Filters/{f.filter_code}/Wavelengths"] = np.asarray(f.lam, dtype=np.float64)

        def cls_name(objs):
            if isinstance(objs, SFHArray):
                c = objs.sfh_type
            elif isinstance(objs, ZDistArray):
                c = type(objs[0])
            else:
                c = type(list(objs)[0])
            return c.__name__.lstrip("_")
        A["Model@sfh_class"] = cls_name(self.sfhs)
        A["Model@metallicity_distribution_class"] = cls_name(self.metal_dists)
        A["Model@model_name"] = self.model_name
        for key, value in (other_info or {}).items():
            if isinstance(value, (list, np.ndarray)) or (has_units(value) and np.ndim(value) > 0):
                D[f"Model/{key}"] = np.asarray(strip_units(value))
                if has_units(value):
                    A[f"Model/{key}@units"] = str(value.units)
            else:
                A[f"Model@{key}"] = value
        A["Model@stellar_params"] = list(self.galaxy_params.keys())
        for key, value in (parameter_transforms_to_save or {}).items():
            if isinstance(value, tuple) or callable(value):
                if callable(value):
                    value = (key, value)
                name = key if isinstance(key, str) else "+".join(key)
                D[f"Model/Transforms/{name}"] = np.frombuffer(getsource(value[1]).encode("utf-8"), dtype=np.uint8)
                A[f"Model/Transforms/{name}@new_parameter_name"] = value[0] if isinstance(value[0], str) else list(value[0])
        A["Model@varying_param_names"] = list(self.varying_param_names)
        A["Model@fixed_param_names"] = list(self.fixed_param_names)
        A["Model@fixed_param_values"] = list(self.fixed_param_values)
        A["Model@fixed_param_units"] = list(self.fixed_param_units)
        return D, A

    def _model_info(self, other_info=None, parameter_transforms_to_save=None):
        """Attribute half of :meth:`_model_block` (kept for callers that only need the bookkeeping)."""
        return self._model_block(other_info, parameter_transforms_to_save)[1]

    def plot_galaxy(self, *a, **k):
        raise NotImplementedError("plotting helpers are outside the hot path (SURVEY 2 row 13)")


def _photometry_columns(phot_blocks):
    """``{dataset path: (n,) float64}`` of a batch's per-filter columns (the reference's pipeline layout,
    ``library.py:3437-3565``) from the kernels' float32 ``(n, n_filt)`` matrices: one transposing cast per block."""
    out = {}
    for prefix, codes, flux in phot_blocks:
        cols = np.ascontiguousarray(flux.T, dtype=np.float64)
        for j, code in enumerate(codes):
            out[prefix + code] = cols[j]
    return out


def _write_pipeline_part(path, datasets, attrs, phot_blocks):
    """Writer-thread body of process_galaxies: cut the photometry columns, then write the container."""
    data = _photometry_columns(phot_blocks)
    data.update(datasets)
    return write_container(path, data, attrs, False)


def _library_compression():
    """Deflate level for library files: ``SYNFERENCE_B200_COMPRESS`` = 0 (default: none -- float photometry deflates by a
    quarter at 25 MB/s on one core, which would be most of a library build; SURVEY: "optional compression off"), 1 ... 9."""
    try:
        return int(os.environ.get("SYNFERENCE_B200_COMPRESS", "0"))
    except ValueError:
        return 0


class _LazyColumns(dict):
    """``{filter code: column}`` whose columns are built on first access."""

    def __init__(self, codes, make):
        super().__init__()
        self._codes, self._make = list(codes), make

    def __missing__(self, code):
        if code not in self._codes:
            raise KeyError(code)
        self[code] = self._make(code)
        return self[code]

    def __contains__(self, code):
        return code in self._codes

    def __iter__(self):
        return iter(self._codes)

    def __len__(self):
        return len(self._codes)

    def keys(self):
        return list(self._codes)

    def items(self):
        return [(c, self[c]) for c in self._codes]

    def values(self):
        return [self[c] for c in self._codes]


class _LazyGalaxyList:
    """``basis.galaxies``: indexable records without materialising N Python objects."""

    def __init__(self, basis: GalaxyBasis):
        self._b = basis

    def __len__(self):
        return len(self._b.params)

    def __getitem__(self, i):
        b = self._b
        return {"redshift": float(b.params.redshift[i]),
                "all_params": {k: np.asarray(strip_units(v))[i] for k, v in b.all_parameters.items()}}


class CombinedBasis:
    """Turns processed bases into the photometry library the trainer reads."""

    def __init__(self, bases: List[GalaxyBasis], log_stellar_masses, redshifts, base_emission_model_keys: List[str],
                 combination_weights, out_name: str = "combined_basis", out_dir: str = library_folder,
                 log_base_masses=9, draw_parameter_combinations: bool = False) -> None:
        self.bases = bases
        self.log_stellar_masses = log_stellar_masses
        self.redshifts = redshifts
        self.combination_weights = combination_weights
        self.out_name, self.out_dir = out_name, out_dir
        self.log_base_masses = log_base_masses
        self.base_emission_model_keys = base_emission_model_keys
        self.draw_parameter_combinations = draw_parameter_combinations
        if isinstance(redshifts, (int, float)):
            self.redshifts = np.full(len(self.log_stellar_masses), redshifts)
        if self.combination_weights is None:
            assert len(self.bases) == 1
            self.combination_weights = np.ones((len(self.redshifts), 1))
        self._mask = None
        self._multi_node = False

    def process_bases(self, n_proc=6, overwrite: Union[bool, List[bool]] = False, verbose=False,
                      batch_size=40_000, multi_node=False, galaxies_mask=None, spectra_to_save=None,
                      em_lines_to_save=None, **extra_analysis_functions):
        """Create and run every basis (``library.py:3291-3383``)."""
        if not isinstance(overwrite, (list, tuple, np.ndarray)):
            overwrite = [overwrite] * len(self.bases)
        self._mask, self._multi_node = galaxies_mask, multi_node
        for i, base in enumerate(self.bases):
            if base.build_library:
                base._create_galaxies(self.log_base_masses)
            else:
                base._create_matched_galaxies(self.log_base_masses, galaxies_mask=galaxies_mask)
            base.process_galaxies(out_name=base.model_name, out_dir=self.out_dir, n_proc=n_proc, verbose=verbose,
                                  save=True, overwrite=overwrite[i], multi_node=multi_node,
                                  emission_model_keys=[self.base_emission_model_keys[i]], batch_size=batch_size,
                                  spectra_to_save=spectra_to_save, em_lines_to_save=em_lines_to_save,
                                  **extra_analysis_functions)
        if multi_node:
            _dist.barrier()

    def _join_writes(self):
        for base in self.bases:
            base.wait_for_writes()

    def load_bases(self, load_spectra=False) -> dict:
        """Read the pipeline files back and concatenate batches / rank shards (``library.py:3385-3642``)."""
        import re
        out = {}
        rank, size = _rank_size(self._multi_node)
        for i, base in enumerate(self.bases):
            stem = os.path.join(self.out_dir, base.model_name)
            # <model>.hdf5 | <model>_<batch>.hdf5 | ..._rank<r>.hdf5 -- nothing else that merely starts with the name.
            # With multi_node every rank compiles ITS OWN shard: only this rank's files are loaded (the ranks share out_dir).
            pat = re.compile(r"^" + re.escape(base.model_name) + r"(_\d+)?(_rank(\d+))?\.hdf5$")
            cache = getattr(base, "_pipeline_cache", None)
            parts = []
            from_cache = bool(cache and cache["complete"] and cache["parts"] and cache["out_dir"] == os.path.abspath(self.out_dir))
            if from_cache:
                for fname, data, attrs, blocks in cache["parts"]:      # just written by process_galaxies: no need to read them back
                    parts.append((int(attrs.get("galaxy_start", 0)), int(attrs.get("rank", 0)), data, attrs, blocks))
            else:
                base.wait_for_writes()
                files = []
                for f in sorted(os.listdir(self.out_dir)):
                    mt = pat.match(f)
                    if mt and (size == 1 or mt.group(3) is None or int(mt.group(3)) == rank):
                        files.append(f)
                if not files:
                    raise FileNotFoundError(f"No pipeline output for base {base.model_name} in {self.out_dir}")
                for f in files:
                    data, attrs = read_container(os.path.join(self.out_dir, f))
                    if size > 1 and int(attrs.get("world_size", 1)) > 1 and int(attrs.get("rank", 0)) != rank:
                        continue
                    parts.append((int(attrs.get("galaxy_start", 0)), int(attrs.get("rank", 0)), data, attrs))
            parts.sort(key=lambda t: (t[1], t[0]))
            key = self.base_emission_model_keys[i]
            label = parts[0][3].get("InstrumentLabel", base.instrument.label)
            codes = list(parts[0][3].get("FilterCodes", base.instrument.filters.filter_codes))
            props, supp_props = {}, {}
            s_units = dict(zip(parts[0][3].get("supp_names", []), parts[0][3].get("supp_units", [])))
            # every batch of this process is in the cache, in order: a parameter column is the basis' own full array
            whole = cache.get("full_props", {}) if from_cache else {}
            for name in parts[0][2]:
                if name.startswith("Galaxies/") and name.count("/") == 1:
                    short = name.split("/", 1)[1]
                    col = whole[short] if short in whole else np.concatenate([p[2][name] for p in parts])
                    if short.startswith("supp_"):      # library.py:3446-3448
                        supp_props[short[5:]] = (col, s_units.get(short[5:], "dimensionless"))
                    else:
                        props[short] = col
            pkey = f"Galaxies/Stars/Photometry/Fluxes/{key}/{label}/"
            if from_cache:       # columns are cut from the kernels' matrices only if somebody asks for them
                phot = _LazyColumns(codes, lambda c, parts=parts, pkey=pkey: np.concatenate(
                    [_photometry_columns(p[4])[pkey + c] for p in parts]))
            else:
                phot = {c: np.concatenate([p[2][pkey + c] for p in parts]) for c in codes}
            entry = {"properties": props, "observed_photometry": phot, "supp_properties": supp_props,
                     "wavelengths": parts[0][2]["Wavelengths"], "filter_codes": codes, "stem": stem}
            if cache and cache["complete"] and cache.get("phot", {}).get(key) and from_cache:
                # the (N, n_filt) float32 matrix the kernels produced, in batch order (= sorted by galaxy_start)
                if key in cache.get("scaled_matrix", {}):
                    # (n_filt, N) float64 already scaled to the stellar masses it was built for, and those masses
                    entry["scaled_matrix"], entry["scaled_log_mass"] = cache["scaled_matrix"][key]
                    entry["photometry_blocks"] = cache["phot"][key]
                else:
                    entry["photometry_matrix"] = np.concatenate(cache["phot"][key]) if len(cache["phot"][key]) > 1 else cache["phot"][key][0]
            skey = f"Galaxies/Stars/Spectra/SpectralFluxDensities/{key}"
            if load_spectra:
                if skey not in parts[0][2]:
                    raise KeyError(f"No spectra stored for {base.model_name}; pass spectra_to_save")
                entry["observed_spectra"] = np.concatenate([p[2][skey] for p in parts])
            out[base.model_name] = entry
        return out

    # ---------------------------------------------------------------------------------------
    def create_library(self, override_instrument=None, save=True, overload_out_name="", overwrite=False):
        if not self.draw_parameter_combinations:
            return self.create_full_library(override_instrument, overwrite=overwrite, save=save,
                                            overload_out_name=overload_out_name)
        return self._create_combination_library(override_instrument, save=save, overload_out_name=overload_out_name,
                                                overwrite=overwrite)

    def _create_combination_library(self, override_instrument=None, save=True, overload_out_name="", overwrite=False):
        """``draw_parameter_combinations=True`` (``library.py:3644-3974``): for every redshift x total mass x combination
        weight, every combination of the bases' galaxies AT that redshift (``np.meshgrid(..., indexing="ij")`` order), base
        photometry (cast to float32 first) scaled to ``weight * total mass`` and summed over the bases; supplementary
        parameters that scale with mass are scaled and summed the same way.  Vectorised over the combinations."""
        out_name = overload_out_name or self.out_name
        path = os.path.join(self.out_dir, out_name)
        if not path.endswith(".hdf5"):
            path += ".hdf5"
        if os.path.exists(path) and not overwrite:
            logger.warning(f"File {path} already exists. Skipping.")
            return self.load_library_from_file(path)
        outputs = self.load_bases()
        base_filters = outputs[self.bases[0].model_name]["filter_codes"]
        for b in self.bases[1:]:
            if outputs[b.model_name]["filter_codes"] != base_filters:
                raise ValueError("All bases must share the same filters")
        if override_instrument is not None:
            for code in override_instrument.filters.filter_codes:
                if code not in base_filters:
                    raise ValueError(f"Filter {code} not found in base filters. Cannot override instrument.")
            filter_codes = list(override_instrument.filters.filter_codes)
        else:
            filter_codes = list(base_filters)
        multi = len(self.bases) > 1
        names_per_base = {b.model_name: [(f"{b.model_name}/{n}" if multi else n, n) for n in b.varying_param_names
                                         if n != "redshift"] for b in self.bases}
        param_columns = ["redshift", "log_mass"] + (["weight_fraction"] if multi else [])
        param_units = ["dimensionless", "log10(Mstar/Msun)"] + (["dimensionless"] if multi else [])
        for b in self.bases:
            for col, short in names_per_base[b.model_name]:
                param_columns.append(col)
                src = b.all_parameters.get(short)
                param_units.append(UNIT_DICT.get(short.lower(), str(src.units) if has_units(src) else "dimensionless"))
        supp_keys = list(outputs[self.bases[0].model_name]["supp_properties"].keys())
        for b in self.bases:
            if list(outputs[b.model_name]["supp_properties"].keys()) != supp_keys:
                raise AssertionError("Not all bases have the same supplementary parameters.")
        supp_units = [outputs[self.bases[0].model_name]["supp_properties"][k][1] for k in supp_keys]
        weights = np.asarray(self.combination_weights, dtype=float).reshape(-1, len(self.bases))
        redshifts = np.atleast_1d(np.asarray(strip_units(self.redshifts), dtype=float))     # looped as given, like the reference
        all_out, all_par, all_supp = [], [], []
        for z in redshifts:
            per_base = []
            for b in self.bases:
                o = outputs[b.model_name]
                mask = np.asarray(o["properties"]["redshift"], dtype=float) == z
                phot = np.stack([np.asarray(o["observed_photometry"][c])[mask] for c in filter_codes], 0).astype(np.float32)
                per_base.append((mask, phot, np.asarray(o["properties"]["mass"], dtype=float)[mask]))
            counts = [pb[1].shape[1] for pb in per_base]
            if min(counts) == 0:
                continue
            combos = np.array(np.meshgrid(*[np.arange(c) for c in counts], indexing="ij")).T.reshape(-1, len(counts))
            for log_total_mass in np.atleast_1d(self.log_stellar_masses):
                total_mass = 10.0 ** float(log_total_mass)
                for comb in weights:
                    dim = combos.shape[0]
                    out = np.zeros((len(filter_codes), dim))
                    rows = [np.full(dim, z), np.full(dim, float(log_total_mass))] + ([np.full(dim, comb[0])] if multi else [])
                    supp = np.zeros((len(supp_keys), dim))
                    for j, b in enumerate(self.bases):
                        mask, phot, mass = per_base[j]
                        scale = comb[j] * total_mass / mass                          # per galaxy of this base
                        out += (phot * scale)[:, combos[:, j]]
                        o = outputs[b.model_name]
                        for col, short in names_per_base[b.model_name]:
                            rows.append(np.asarray(o["properties"][short], dtype=float)[mask][combos[:, j]])
                        for k, key in enumerate(supp_keys):
                            vals, units = o["supp_properties"][key]
                            vals = np.asarray(vals, dtype=float)[mask]
                            # (the reference scales every unit-carrying column here, library.py:3866-3881)
                            if units != "dimensionless":
                                vals = vals * scale
                            supp[k] += vals[combos[:, j]]
                    all_out.append(out)
                    all_par.append(np.stack(rows, 0))
                    all_supp.append(supp)
        if not all_out:
            raise ValueError("no galaxies found at the requested redshifts")
        combined_outputs, combined_params = np.hstack(all_out), np.hstack(all_par)
        combined_supp = np.hstack(all_supp)
        out = {"photometry": combined_outputs, "parameters": combined_params, "parameter_names": param_columns,
               "filter_codes": filter_codes, "supplementary_parameters": combined_supp,
               "supplementary_parameter_names": supp_keys, "supplementary_parameter_units": supp_units,
               "parameter_units": param_units}
        self.library_photometry, self.library_parameters = combined_outputs, combined_params
        self.library_parameter_names, self.library_filter_codes = param_columns, filter_codes
        self.library_parameter_units = param_units
        self.library_supplementary_parameters = combined_supp
        self.library_supplementary_parameter_names = supp_keys
        self.library_supplementary_parameter_units = supp_units
        if save:
            self.save_library(out, overload_out_name=overload_out_name, overwrite=overwrite)
        return out

    def create_spectral_grid(self, override_instrument=None, save=True, overload_out_name="", overwrite=False):
        return self.create_full_library(override_instrument, save=save, overload_out_name=overload_out_name,
                                        overwrite=overwrite, spectral_mode=True)

    def create_full_library(self, override_instrument=None, save: bool = True, overload_out_name: str = "",
                            overwrite: bool = False, spectral_mode=False):
        """Scale base-mass outputs to each galaxy's stellar mass and assemble the parameter table
        (``library.py:4435-4885``).  Photometry is cast to float32 before the float64 mass ratio is
        applied (``:4588-4609``); multi-base libraries add the weighted contributions (``:4739-4742``)."""
        if self.draw_parameter_combinations:
            raise AssertionError("Cannot create full grid with draw_parameter_combinations set to True. "
                                 "Set to False to create full grid.")
        outputs = self.load_bases(load_spectra=spectral_mode)
        base_filters = outputs[self.bases[0].model_name]["filter_codes"]
        for b in self.bases[1:]:
            if outputs[b.model_name]["filter_codes"] != base_filters:
                raise ValueError("All bases must share the same filters")
        if override_instrument is not None:
            for code in override_instrument.filters.filter_codes:
                if code not in base_filters:
                    raise ValueError(f"Filter {code} not found in base filters. Cannot override instrument.")
            filter_codes = list(override_instrument.filters.filter_codes)
        else:
            filter_codes = list(base_filters)
        sel = slice(None) if self._mask is None else np.asarray(self._mask, dtype=bool)
        redshift = np.broadcast_to(np.asarray(strip_units(self.redshifts), dtype=float),
                                   (len(self.log_stellar_masses),))[sel]
        log_mass = np.asarray(self.log_stellar_masses, dtype=float)[sel]
        weights = np.asarray(self.combination_weights, dtype=float)
        weights = weights.reshape(len(weights), -1)[sel]
        multi = len(self.bases) > 1
        param_columns = ["redshift", "log_mass"] + (["weight_fraction"] if multi else [])
        param_units = ["dimensionless", "log10_Msun"] + (["dimensionless"] if multi else [])
        rows = [redshift, log_mass] + ([weights[:, 0]] if multi else [])
        total = None
        for i, base in enumerate(self.bases):
            o = outputs[base.model_name]
            mass = o["properties"]["mass"]
            if len(mass) != len(log_mass):
                raise ValueError(f"base {base.model_name} has {len(mass)} galaxies, expected {len(log_mass)}")
            scale = (weights[:, i] if multi else 1.0) * 10.0 ** log_mass / mass
            if spectral_mode:
                contrib = o["observed_spectra"].astype(np.float32) * scale[:, None]
            elif (not multi and "scaled_matrix" in o and filter_codes == list(o["filter_codes"]) and np.all(mass == 1e9)
                  and o["scaled_matrix"].shape[1] == len(log_mass) and np.array_equal(o["scaled_log_mass"], log_mass)):
                # the kernels wrote float32(base) x 10^log_mass / base_mass into this matrix batch by batch (library_out)
                contrib = o["scaled_matrix"]
            else:
                if "photometry_blocks" in o and "photometry_matrix" not in o:
                    o["photometry_matrix"] = np.concatenate(o["photometry_blocks"])
                if "photometry_matrix" in o and filter_codes == list(o["filter_codes"]):
                    phot = o["photometry_matrix"]                     # float32 (N, n_filt) straight from the kernels
                else:
                    phot = np.stack([o["observed_photometry"][c] for c in filter_codes], 1).astype(np.float32)
                # float32(base) x float64 mass ratio (library.py:4588-4609), formed directly in the library's
                # (n_filters, n_galaxies) layout
                contrib = phot.T.astype(np.float64) * scale[None, :]
            total = contrib if total is None else total + contrib
            for name in base.varying_param_names:
                if name == "redshift":
                    continue
                col = f"{base.model_name}/{name}" if multi else name
                param_columns.append(col)
                rows.append(np.asarray(o["properties"][name], dtype=float))
                short = name.lower()
                src = base.all_parameters.get(name)
                param_units.append(UNIT_DICT.get(short, str(src.units) if has_units(src) else "dimensionless"))
        combined_outputs = np.ascontiguousarray(total.T if spectral_mode else total)          # (n_filters | n_lam, n_gal)
        combined_params = np.stack(rows, 0)                        # (n_params, n_gal)
        # supplementary parameters: rescaled from the base mass like the photometry (library.py:4631-4656)
        supp_names, supp_units_l, supp_rows = [], [], []
        for i, base in enumerate(self.bases):
            o = outputs[base.model_name]
            if not o["supp_properties"]:
                continue
            # every base's by-products are rescaled by ITS share of the total mass (weight x mass / base mass), and in a
            # multi-base library they are named <model_name>/<name>, like the parameters
            scale = (weights[:, i] if multi else 1.0) * 10.0 ** log_mass / o["properties"]["mass"]
            for name, (vals, units) in o["supp_properties"].items():
                how = _supp.scales_with_mass(units)
                vals = np.asarray(vals, dtype=float)
                with np.errstate(divide="ignore"):
                    supp_rows.append(vals * scale if how == "linear" else vals + np.log10(scale) if how == "log" else vals)
                supp_names.append(f"{base.model_name}/{name}" if multi else name)
                supp_units_l.append(units)
        supp = np.stack(supp_rows, 0) if supp_rows else np.zeros((0, combined_params.shape[1]))
        out = {"parameters": combined_params, "parameter_names": param_columns,
               "supplementary_parameters": supp, "supplementary_parameter_names": supp_names,
               "supplementary_parameter_units": supp_units_l, "parameter_units": param_units}
        if spectral_mode:
            out["spectra"] = combined_outputs
            self.library_spectra = combined_outputs
            self.library_filter_codes = (outputs[self.bases[0].model_name]["wavelengths"] * 1e-4).tolist()
        else:
            out["photometry"] = combined_outputs
            self.library_photometry = combined_outputs
            self.library_filter_codes = filter_codes
        out["filter_codes"] = self.library_filter_codes
        self.library_parameters = combined_params
        self.library_parameter_names = param_columns
        self.library_parameter_units = param_units
        self.library_supplementary_parameters = supp
        self.library_supplementary_parameter_names = supp_names
        self.library_supplementary_parameter_units = supp_units_l
        logger.info(f"Combined outputs shape: {combined_outputs.shape}; parameters {combined_params.shape}")
        if save:
            self.save_library(out, overload_out_name=overload_out_name, overwrite=overwrite)
        return out

    def _validate_library(self, library_dict, check_type="photometry"):
        """NaN / Inf / shape validation (``library.py:3976-4029``)."""
        for req in (check_type, "parameters", "parameter_names", "filter_codes"):
            if req not in library_dict:
                raise ValueError(f"library dictionary is missing '{req}'")
        phot, par = library_dict[check_type], library_dict["parameters"]
        if phot.ndim != 2 or par.ndim != 2 or phot.shape[1] != par.shape[1]:
            raise ValueError(f"{check_type} {phot.shape} and parameters {par.shape} must share their last axis")
        if len(library_dict["parameter_names"]) != par.shape[0]:
            raise ValueError("parameter_names does not match the parameter array")
        for name, arr in ((check_type, phot), ("parameters", par)):
            if np.isfinite(arr).all():       # one pass over the (large) array in the usual case
                continue
            if np.isnan(arr).any():
                raise ValueError(f"{name} array contains NaN values.")
            if np.isinf(arr).any():
                raise ValueError(f"{name} array contains infinite values.")
        return True

    def save_library(self, library_dict: dict, overload_out_name: str = "", overwrite: bool = False,
                     library_params_to_save=("model_name",)) -> None:
        """Write ``Grid/Photometry`` (or ``Grid/Spectra``), ``Grid/Parameters``,
        ``Grid/SupplementaryParameters`` and the attribute block (``library.py:4031-4153``)."""
        check_type = "photometry" if "photometry" in library_dict else "spectra"
        self._validate_library(library_dict, check_type=check_type)
        os.makedirs(self.out_dir, exist_ok=True)
        out_name = overload_out_name or self.out_name
        if not out_name.endswith(".hdf5"):
            out_name = f"{out_name}.hdf5"
        rank, size = _rank_size(self._multi_node)
        if size > 1:
            out_name = out_name.replace(".hdf5", f"_{rank}.hdf5")  # one shard per rank (utils.py:2288-2299)
        path = os.path.join(self.out_dir, out_name)
        if os.path.exists(path) and not overwrite:
            logger.warning(f"File {path} already exists. Skipping.")
            return
        if os.path.exists(path):
            logger.warning(f"File {path} already exists. Overwriting.")
            os.remove(path)
        datasets = {"Grid/Parameters": library_dict["parameters"]}
        if "photometry" in library_dict:
            datasets["Grid/Photometry"] = library_dict["photometry"]
        if "spectra" in library_dict:
            datasets["Grid/Spectra"] = library_dict["spectra"]
        if "supplementary_parameters" in library_dict:
            datasets["Grid/SupplementaryParameters"] = library_dict["supplementary_parameters"]
        attrs = {"ParameterNames": list(library_dict["parameter_names"]),
                 "FilterCodes": list(library_dict["filter_codes"]), "PhotometryUnits": "nJy",
                 "SupplementaryParameterNames": list(library_dict.get("supplementary_parameter_names", [])),
                 "SupplementaryParameterUnits": list(library_dict.get("supplementary_parameter_units", [])),
                 "ParameterUnits": list(library_dict.get("parameter_units", [])),
                 "Grids": [b.grid.grid_name for b in self.bases],
                 "CreationDT": datetime.now().strftime("%Y%m%d_%H%M%S"), "rank": rank, "world_size": size}
        for param in library_params_to_save:
            attrs[param] = [str(getattr(b, param)) for b in self.bases]
        attrs.update(getattr(self, "_extra_attrs", None) or {})
        datasets.update(getattr(self, "_extra_datasets", None) or {})
        ex = getattr(self, "_save_executor", None)
        if ex is not None:       # create_mock_library joins this write before it returns
            self._pending_save = ex.submit(write_container, path, datasets, attrs, _library_compression())
        else:
            write_container(path, datasets, attrs, compress=_library_compression())
        self.library_path = path

    def load_library_from_file(self, file_path: str):
        from .utils import load_library_from_hdf5
        lib = load_library_from_hdf5(file_path)
        self.library_photometry = lib.get("photometry")
        self.library_parameters = lib["parameters"]
        self.library_parameter_names = lib["parameter_names"]
        self.library_filter_codes = lib["filter_codes"]
        return lib


class GalaxySimulator:
    """On-the-fly simulator: parameter vector(s) -> photometry (``library.py:4922-6001``).

    Unlike the reference (one galaxy per call, ~ms-100 ms of Python/unyt overhead each), ``simulate``
    also accepts a 2-D array / tensor of N parameter vectors and synthesises them in one CUDA pass.
    """

    def __init__(self, sfh_model, zdist_model, grid: Grid, instrument: Instrument, emission_model: EmissionModel,
                 emission_model_key: str, emitter_params: dict = None, cosmo=Planck18, param_order=None,
                 param_units: dict = None, param_transforms: dict = None, out_flux_unit: str = "nJy",
                 required_keys=("redshift", "log_mass"), extra_functions=None, normalize_method=None,
                 output_type="photo_fnu", include_phot_errors: bool = False, depths=None, depth_sigma: int = 5,
                 noise_models=None, fixed_params: dict = None, photometry_to_remove=None, ignore_params=None,
                 ignore_scatter: bool = False, return_type: str = "array", device="cpu", max_batch=1 << 16) -> None:
        assert isinstance(grid, Grid), f"Grid must be a subclass of Grid. Got {type(grid)} instead."
        assert isinstance(instrument, Instrument), "Instrument must be an Instrument"
        assert isinstance(emission_model, EmissionModel), "Emission model must be an EmissionModel"
        assert return_type in ("array", "tensor")
        if extra_functions:
            raise NotImplementedError("extra_functions operate on Synthesizer objects; not in the batched path")
        self.sfh_model, self.zdist_model = sfh_model, zdist_model
        self.grid, self.instrument, self.emission_model = grid, instrument, emission_model
        self.emission_model_key = emission_model_key
        self.emitter_params = emitter_params or {"stellar": ["tau_v"], "galaxy": []}
        self.cosmo = cosmo
        self.param_order = list(param_order) if param_order is not None else None
        self.param_units = param_units or {}
        self.param_transforms = param_transforms or {}
        self.out_flux_unit = out_flux_unit
        self.required_keys = list(required_keys)
        self.normalize_method = normalize_method
        self.output_type = [output_type] if isinstance(output_type, str) else list(output_type)
        self.include_phot_errors = include_phot_errors
        self.depths, self.depth_sigma = depths, depth_sigma
        self.noise_models = noise_models
        self.fixed_params = fixed_params or {}
        self.ignore_params = ignore_params or []
        self.ignore_scatter = ignore_scatter
        self.return_type, self.device = return_type, device
        self.unused_params, self.reported_unused = [], False
        if photometry_to_remove:
            self.update_photo_filters(photometry_to_remove=photometry_to_remove)
        if depths is not None:
            assert len(depths) == len(self.instrument.filters.filter_codes), \
                "depths must have one entry per filter"
        if noise_models is not None:
            assert isinstance(noise_models, dict), "noise_models must be a dict keyed by filter code"
            missing = [c for c in self.instrument.filters.filter_codes if c not in noise_models]
            assert not missing, f"no noise model for filters {missing}"
        import inspect
        self.sfh_params = [p for p in inspect.signature(sfh_model.__init__).parameters
                           if p not in ("self", "min_age")]
        self.optional_sfh_params = ["min_age"]
        self.zdist_params = list(getattr(zdist_model, "required", ())) or self._zdist_names(zdist_model)
        self.optional_zdist_params = []
        self.total_possible_keys = set(self.sfh_params + self.zdist_params + self.required_keys + ["min_age"])
        self._max_batch = int(max_batch)
        self._engine: Optional[SynthEngine] = None

    @staticmethod
    def _zdist_names(zdist_model):
        name = zdist_model.__name__
        return ["log10metallicity"] if name == "_DeltaConstant" else ["mean", "sigma"]

    @classmethod
    def from_library(cls, library_path: str, override_synthesizer_grid_dir: Union[None, str, bool] = True,
                     override_emission_model: Optional[EmissionModel] = None, **kwargs):
        """Rebuild the simulator that produced a library from the ``Model`` group stored in it
        (``library.py:5219-5551``): grid by name / directory (``override_synthesizer_grid_dir`` or ``SYNTHESIZER_GRID_DIR``
        when the recorded directory does not exist here; a ready ``grid=`` may be passed instead), instrument from the
        stored filter curves, SFH / metallicity-distribution classes, emission model with its dust law and dust emission,
        parameter order and units from the library's ``ParameterNames`` / ``ParameterUnits``, fixed parameters and saved
        parameter transforms.  Extra keyword arguments go to the constructor, as in the reference."""
        import inspect
        import json
        from . import parametric as P
        from .parametric import SFH, ZDist, Filter, FilterCollection
        from .units import Unit
        if not os.path.exists(library_path):
            raise FileNotFoundError(f"Library path {library_path} does not exist. Cannot create GalaxySimulator.")
        data, attrs = read_container(library_path)
        if not any(k.startswith("Model@") or k.startswith("Model/") for k in list(attrs) + list(data)):
            raise ValueError(f"Library file {library_path} does not contain 'Model' group. Cannot create GalaxySimulator.")
        g = lambda k, d=None: attrs.get(f"Model@{k}", d)                                        # noqa: E731
        e = lambda k, d=None: attrs.get(f"Model/EmissionModel@{k}", d)                          # noqa: E731
        lam = np.asarray(data["Model/Instrument/Filters/Header/Wavelengths"], dtype=float)
        # Step 1: grid
        grid = kwargs.pop("grid", None)
        if grid is None:
            grid_name, grid_dir = g("grid_name"), g("grid_dir")
            if override_synthesizer_grid_dir is not None and not os.path.exists(str(grid_dir)):
                if isinstance(override_synthesizer_grid_dir, str) and os.path.exists(override_synthesizer_grid_dir):
                    grid_dir = override_synthesizer_grid_dir
                else:
                    grid_dir = os.getenv("SYNTHESIZER_GRID_DIR", None) or "."
            if isinstance(grid_dir, str) and grid_dir.endswith((".hdf5", ".h5", ".npz")):
                logger.info("Overriding internal library name from provided file path.")
                grid_name = os.path.basename(grid_dir).replace(".hdf5", "").replace(".h5", "").replace(".npz", "")
                grid_dir = os.path.dirname(grid_dir)
            grid = Grid(grid_name, grid_dir)
        if np.asarray(grid.lam).shape != lam.shape or np.max(np.abs(np.asarray(grid.lam) / lam - 1.0)) > 1e-12:
            grid = Grid(grid.grid_name, grid.grid_dir, new_lam=lam, log10ages=grid.log10ages, metallicity=grid.metallicity,
                        lam=grid.lam, spectra=grid.spectra)
        # Step 2: instrument
        codes = list(attrs.get("Model/Instrument/Filters/Header@filter_codes", g("filters", [])))
        filters = []
        for code in codes:
            t = np.asarray(data[f"Model/Instrument/Filters/{code}/Transmission"], dtype=float)
            fl = data.get(f"Model/Instrument/Filters/{code}/Wavelengths")
            filters.append(Filter(code, lam if fl is None else np.asarray(fl, dtype=float), t))
        fc = FilterCollection(filters=filters)
        if all(f.lam.shape == lam.shape for f in filters):
            fc.lam = Quantity(lam, "Angstrom")
        instrument = Instrument(attrs.get("Model/Instrument@label", g("instrument", "instrument")), filters=fc)
        # Step 3: cosmology (the reference falls back to Planck18 when the stored one cannot be rebuilt)
        cosmo = Planck18
        if "Planck18" not in str(g("cosmology", "Planck18")):
            logger.warning("Failed to load cosmology from the library file. Using Planck18 instead.")
        # Step 4: SFH / metallicity-distribution classes
        sfh_name, zd_name = g("sfh_class"), g("metallicity_distribution_class")
        sfh_model = getattr(SFH, str(sfh_name), None)
        if sfh_model is None:
            raise ValueError(f"SFH model {sfh_name} not found in SFH module. Cannot create GalaxySimulator.")
        zdist_model = getattr(ZDist, str(zd_name), None)
        if zdist_model is None:
            raise ValueError(f"ZDist model {zd_name} not found in ZDist module. Cannot create GalaxySimulator.")
        emission_model_key = kwargs.pop("emission_model_key", g("emission_model_key", "total"))

        def value(v, unit=""):
            if unit not in ("", None):
                return Quantity(float(v), str(unit))
            if isinstance(v, str):
                try:
                    return float(v)
                except ValueError:
                    return v                     # the name of a per-galaxy emitter parameter
            return v

        def build(prefix_keys, prefix_vals, prefix_units=None):
            ks, vs = list(e(prefix_keys, [])), list(e(prefix_vals, []))
            us = list(e(prefix_units, [""] * len(ks))) if prefix_units else [""] * len(ks)
            return {k: value(v, u) for k, v, u in zip(ks, vs, us)}

        if override_emission_model is not None:
            emission_model = override_emission_model
        else:
            em_name = e("name")
            em_cls = getattr(P, str(em_name), None)
            if em_cls is None or not (inspect.isclass(em_cls) and issubclass(em_cls, EmissionModel)):
                raise ValueError(f"Emission model {em_name} not found in synference_b200.parametric. Cannot create GalaxySimulator.")
            params = build("parameter_keys", "parameter_values", "parameter_units")
            dust_model = None
            if e("dust_law") is not None:
                dcls = getattr(P, str(e("dust_law")), None)
                if dcls is None:
                    raise ValueError(f"Dust model {e('dust_law')} not found. Cannot create GalaxySimulator.")
                dust_model = dcls(**build("dust_attenuation_keys", "dust_attenuation_values", "dust_attenuation_units"))
            dust_emission = None
            if e("dust_emission") is not None:
                gcls = getattr(P, str(e("dust_emission")), None)
                if gcls is None:
                    raise ValueError(f"Dust emission model {e('dust_emission')} not found. Cannot create from_library.")
                gp = build("dust_emission_keys", "dust_emission_values", "dust_emission_units")
                gp.pop("cmb_factor", None)
                gp.pop("temperature_z", None)
                if gcls.__name__ == "Blackbody":
                    gp.pop("emissivity", None)
                dust_emission = gcls(**{k: float(strip_units(v)) for k, v in gp.items()})
            if e("dust_law_birth") is not None:       # two screens
                bcls = getattr(P, str(e("dust_law_birth")))
                birth = bcls(**build("dust_attenuation_birth_keys", "dust_attenuation_birth_values"))
                emission_model = em_cls(grid, dust_curve_ism=dust_model, dust_curve_birth=birth,
                                        dust_emission_ism=dust_emission, dust_emission_birth=dust_emission, **params)
            else:
                sig = inspect.signature(em_cls.__init__).parameters
                if dust_model is not None:
                    params["dust_curve"] = dust_model
                if dust_emission is not None:
                    params["dust_emission_model" if "dust_emission_model" in sig else "dust_emission"] = dust_emission
                emission_model = em_cls(grid=grid, **params)
        # Steps 5-9: emitter parameters, order / units, fixed parameters, transforms
        emitter_params = {"stellar": list(g("stellar_params", [])), "galaxy": {}}
        param_order = list(attrs.get("ParameterNames", []))
        units = list(attrs.get("ParameterUnits", []))
        param_units = {}
        for pname, u in zip(param_order, units):
            if not any(tag in u for tag in ("dimensionless", "log", "mag")):
                try:
                    param_units[pname] = Unit(u)
                except Exception:
                    pass
        fixed_params = {}
        for name, val, unit in zip(g("fixed_param_names", []), g("fixed_param_values", []), g("fixed_param_units", [])):
            fixed_params[name] = Quantity(val, unit) if unit not in ("", None) else val
        param_transforms = {}
        for key in [k for k in data if k.startswith("Model/Transforms/")]:
            name = key[len("Model/Transforms/"):]
            code = inspect.cleandoc("\n" + bytes(np.asarray(data[key], dtype=np.uint8)).decode("utf-8") + "\n")
            scope = {}
            try:
                exec(code, {"np": np}, scope)
                func = scope[code.split("def ")[-1].split("(")[0]]
            except Exception as err:
                logger.error(f"Error evaluating transform function for {name}: {err}")
                continue
            new_key = attrs.get(f"{key}@new_parameter_name")
            tkey = tuple(name.split("+")) if "+" in name else name
            param_transforms[tkey] = (new_key if isinstance(new_key, str) else tuple(new_key), func) if new_key is not None else func
        dict_create = dict(sfh_model=sfh_model, zdist_model=zdist_model, grid=grid, emission_model=emission_model,
                           emission_model_key=emission_model_key, instrument=instrument, cosmo=cosmo,
                           emitter_params=emitter_params, param_order=param_order, param_units=param_units,
                           param_transforms=param_transforms, fixed_params=fixed_params)
        dict_create.update(kwargs)
        return cls(**dict_create)

    def update_photo_filters(self, photometry_to_remove=None, photometry_to_add=None):
        """Restrict / extend the filter set (``library.py:5180-5216``); rebuilds the device tables lazily."""
        from .parametric import FilterCollection
        filters = list(self.instrument.filters.filters)
        if photometry_to_remove:
            filters = [f for f in filters if f.filter_code not in set(photometry_to_remove)]
        have = {f.filter_code for f in filters}
        new_codes = [c for c in (photometry_to_add or []) if c not in have]
        if new_codes:
            # the reference re-fetches every curve by code (FilterCollection(filter_codes=...)); here the added curves are
            # looked up the same way and resampled onto the axis the others already share
            added = FilterCollection(filter_codes=new_codes, new_lam=self.instrument.filters.lam
                                     if self.instrument.filters.lam is not None else self.grid.lam)
            filters = filters + list(added.filters)
        fc = FilterCollection(filters=filters)
        fc.lam = self.instrument.filters.lam
        self.instrument = Instrument(self.instrument.label, filters=fc)
        for name in ("_engine", "_engine_rest"):
            if getattr(self, name, None) is not None:
                getattr(self, name).close()
                setattr(self, name, None)

    def _get_engine(self, rest_frame=False):
        if rest_frame:
            if getattr(self, "_engine_rest", None) is None:
                self._engine_rest = SynthEngine(self.grid, self.emission_model, self.emission_model_key,
                                                self.instrument.filters, cosmo=self.cosmo, igm=False, rest_frame=True,
                                                max_batch=self._max_batch, device=_dist.local_device())
            return self._engine_rest
        if self._engine is None:
            self._engine = SynthEngine(self.grid, self.emission_model, self.emission_model_key,
                                       self.instrument.filters, cosmo=self.cosmo, igm=True,
                                       max_batch=self._max_batch, device=_dist.local_device())
        return self._engine

    # ---------------------------------------------------------------------------------------
    def _params_to_dict(self, params):
        try:
            import torch
            if isinstance(params, torch.Tensor):
                params = params.detach().cpu().numpy()
        except ImportError:
            pass
        params = copy.deepcopy(params)
        if not isinstance(params, dict):
            if self.param_order is None:
                raise ValueError("simulate() input requires a dictionary unless param_order is set. "
                                 "Cannot create photometry.")
            arr = np.asarray(params, dtype=float)
            if arr.ndim == 2 and arr.shape[0] == 1 and arr.shape[1] == len(self.param_order):
                arr = arr[0]
            assert arr.shape[-1] == len(self.param_order), \
                f"Parameter array length {arr.shape[-1]} does not match parameter order length " \
                f"{len(self.param_order)}. Cannot create photometry."
            params = {k: arr[..., j] for j, k in enumerate(self.param_order)}
        params.update(self.fixed_params)
        for key in self.required_keys:
            if key not in params:
                raise ValueError(f"Missing required parameter {key}. Cannot create photometry.")
        for key in params:
            if key in self.param_units:
                params[key] = params[key] * self.param_units[key]
        for key, value in self.param_transforms.items():
            if isinstance(key, tuple):
                name, func = value
                params[name] = func(**{k: params[k] for k in key if k in params})
            elif isinstance(value, tuple):
                name, func = value
                params[name] = func(params[key]) if key in params else func(params)
            elif callable(value):
                params[key] = value(params[key]) if key in params else value(params)
        for key in self.sfh_params + self.zdist_params:
            if key not in params:
                raise ValueError(f"Missing required parameter {key} for SFH or ZDist. Cannot create photometry.")
        return params

    def _lower(self, params: dict) -> GalaxyParams:
        n = int(np.max([np.size(strip_units(v)) for v in params.values()]))
        bc = lambda v: np.broadcast_to(np.asarray(strip_units(v), dtype=float).reshape(-1), (n,))  # noqa: E731
        from .parametric import SFH_MAX_PARAMS, ZD_DELTA_LINEAR, ZD_DELTA_LOG10, ZD_NORMAL_LOG10
        cls = self.sfh_model
        # (only the columns this SFH family uses: a 100 k x 24 float64 block was most of the host time per call)
        rows = np.zeros((n, min(SFH_MAX_PARAMS, max(4, 2 + len(cls.param_names)))))
        to_yr = lambda v: bc(v) * (v.units.factor if has_units(v) else 1.0)  # noqa: E731
        rows[:, 1] = to_yr(params["max_age"])
        if "min_age" in params:
            rows[:, 0] = to_yr(params["min_age"])
        for j, name in enumerate(cls.param_names):
            rows[:, 2 + j] = to_yr(params[name]) if name in cls.time_params else bc(params[name])
        if self.zdist_model.__name__ == "_DeltaConstant":
            if "log10metallicity" in params:
                zd = ZDistArray(ZD_DELTA_LOG10, bc(params["log10metallicity"]))
            else:
                zd = ZDistArray(ZD_DELTA_LINEAR, bc(params["metallicity"]))
        else:
            zd = ZDistArray(ZD_NORMAL_LOG10, bc(params["mean"]), bc(params["sigma"]))
        ism_name = getattr(self.emission_model, "tau_v_ism_name", None) or "tau_v"
        birth_name = getattr(self.emission_model, "tau_v_birth_name", None)
        tau_v = bc(params[ism_name]) if ism_name in params else None
        fesc_name = getattr(self.emission_model, "fesc_name", None)
        dust = getattr(self.emission_model, "dust_curve", None)
        slope_name, ampl_name = getattr(dust, "slope_name", None), getattr(dust, "ampl_name", None)
        lya_name = getattr(self.emission_model, "lya_name", None)
        used = [k for k in params if k not in self.total_possible_keys and k not in (ism_name, birth_name, fesc_name, slope_name, ampl_name, lya_name)
                and k not in self.ignore_params and k not in cls.param_names]
        for k in used:
            if k not in self.unused_params:
                self.unused_params.append(k)
        if self.unused_params and not self.reported_unused:
            logger.warning(f"The following parameters are not used by the simulator: {self.unused_params}")
            self.reported_unused = True
        gp = GalaxyParams.from_objects(bc(params["redshift"]), SFHArray(cls, rows), zd,
                                       log_mass=bc(params["log_mass"]), tau_v=tau_v)
        if fesc_name is not None:   # per-galaxy escape fraction (emission model built with fesc="<name>")
            if fesc_name not in params:
                raise ValueError(f"Missing required parameter '{fesc_name}' (per-galaxy escape fraction of the emission model)")
            gp.coef_att, gp.coef_unatt = self.emission_model.coefficients(self.emission_model_key, bc(params[fesc_name]))
        for attr, name in (("dust_slope", slope_name), ("dust_ampl", ampl_name)):
            if name is not None:
                if name not in params:
                    raise ValueError(f"Missing required parameter '{name}' (read per galaxy by the dust curve)")
                setattr(gp, attr, np.array(bc(params[name]), dtype=float))
        if lya_name is not None and self.emission_model.lya_line(self.emission_model_key) is not None:
            if lya_name not in params:
                raise ValueError(f"Missing required parameter '{lya_name}' (per-galaxy Lyman-alpha escape fraction)")
            gp.fesc_lya = np.array(bc(params[lya_name]), dtype=float)
        if birth_name is not None and self.emission_model.two_screens(self.emission_model_key):
            for name in (ism_name, birth_name):
                if name not in params:
                    raise ValueError(f"Missing required parameter '{name}' (optical depth of one of the two dust screens)")
            gp.tau_v_birth = np.array(bc(params[birth_name]), dtype=float)
        return gp

    def simulate(self, params):
        """Photometry (and/or spectra) for one parameter vector or a batch of them."""
        batched = not isinstance(params, dict) and np.ndim(params) == 2 and np.shape(params)[0] > 1
        p = self._lower(self._params_to_dict(params))
        eng = self._get_engine()
        outputs = {}
        if "photo_fnu" in self.output_type:
            outputs["photo_fnu"] = eng.photometry(p, scaled=True)             # nJy
            outputs["photo_wav"] = self.instrument.filters.pivot_lams
        if "fnu" in self.output_type:
            spec = eng.spectra(p).astype(np.float64) * (10.0 ** p.log_mass / eng.base_mass)[:, None]
            outputs["fnu"] = spec
            outputs["fnu_wav"] = np.asarray(self.grid.lam)[None, :] * (1.0 + p.redshift)[:, None]
        if "sfh" in self.output_type:
            # mass formed per age bin / bin width (library.py:5736-5750): Msun / yr on the grid's ages
            ages = 10.0 ** np.asarray(self.grid.log10ages, dtype=float)
            sfh = eng.sfzh(p).sum(-1) * (10.0 ** p.log_mass)[:, None] / np.diff(ages, prepend=0.0)[None, :]
            time_myr = ages / 1.0e6
            outputs["sfh"] = Quantity(sfh if batched else sfh[0], "Msun/yr")
            outputs["sfh_time"] = Quantity(time_myr, "Myr")
            outputs["redshift"] = p.redshift if batched else float(p.redshift[0])
            t_abs = np.asarray(self.cosmo.age(p.redshift).to("Myr").value, dtype=float)[:, None] - time_myr[None, :]
            outputs["sfh_time_abs"] = Quantity(t_abs if batched else t_abs[0], "Myr")
        if "photo_lnu" in self.output_type:
            # rest-frame luminosities through the filters (library.py:5756-5761): no redshift, no IGM, no distance
            outputs["photo_lnu"] = self._get_engine(rest_frame=True).photometry(p, scaled=True)    # erg / s / Hz
        if "lnu" in self.output_type:
            # rest-frame luminosity spectrum on the grid's axis (library.py:5752-5754): the rest-frame engine's spectra
            er = self._get_engine(rest_frame=True)
            outputs["lnu"] = er.spectra(p).astype(np.float64) * (10.0 ** p.log_mass / er.base_mass)[:, None]    # erg / s / Hz
            outputs["lnu_wav"] = np.asarray(self.grid.lam)
        for t in self.output_type:
            if t not in ("photo_fnu", "fnu", "sfh", "photo_lnu", "lnu"):
                raise NotImplementedError(f"output_type '{t}' is not available in the batched path")
        conv = {"nJy": 1.0, "uJy": 1e-3, "mJy": 1e-6, "Jy": 1e-9}
        for k in ("photo_fnu", "fnu"):
            if k not in outputs:
                continue
            if self.out_flux_unit == "AB":
                with np.errstate(all="ignore"):
                    outputs[k] = -2.5 * np.log10(outputs[k] * 1e-9) + 8.9
                if k == "fnu":
                    outputs[k][np.isinf(outputs[k])] = 99
            elif self.out_flux_unit == "asinh":
                raise NotImplementedError("asinh fluxes not implemented yet. Please use AB or Jy units.")
            elif self.out_flux_unit in conv:
                outputs[k] = outputs[k] * conv[self.out_flux_unit]
            else:
                raise ValueError(f"unknown out_flux_unit {self.out_flux_unit}")
        if len(self.output_type) > 1:
            outputs["filters"] = self.instrument.filters
            return outputs
        if self.output_type[0] == "sfh":
            return outputs["sfh"]
        fluxes = outputs[self.output_type[0]]
        # scatter / normalise / append errors for all rows at once: the functions below work along the last axis (the
        # reference handles one galaxy per call; a per-row Python loop here cost 100x the GPU time for 100 k rows)
        f, errors = self._scatter(fluxes, flux_units=self.out_flux_unit)
        if self.normalize_method is not None:
            f = self._normalize(f, method=self.normalize_method, norm_unit=self.out_flux_unit)
        if self.include_phot_errors:
            if errors is None:
                raise ValueError("include_phot_errors needs depths or noise_models (and ignore_scatter=False)")
            f = np.concatenate((f, np.broadcast_to(errors, fluxes.shape)), axis=-1)
        out = f if batched else f[0]
        if self.return_type == "tensor":
            import torch
            out = torch.tensor(np.atleast_2d(out), device=self.device)
        return out

    def _normalize(self, fluxes, method=None, norm_unit="AB", add_norm_pos=-1):
        """``library.py:5866-5904`` along the last axis (one galaxy ``(n_filt,)`` or a batch ``(n, n_filt)``)."""
        if method is None:
            return fluxes
        fluxes = np.asarray(fluxes)
        func = np.subtract if norm_unit == "AB" else np.divide
        if isinstance(method, str):
            codes = self.instrument.filters.filter_codes
            if method not in codes:
                raise ValueError(f"Filter {method} not found in filter codes. Cannot normalize photometry.")
            norm = fluxes[..., codes.index(method)]
        elif has_units(method):
            norm = -2.5 * np.log10(float(strip_units(method, "Jy"))) + 8.9 if norm_unit == "AB" else \
                float(strip_units(method, norm_unit))
            norm = np.full(fluxes.shape[:-1], norm)
        elif callable(method):
            norm = method(fluxes) if fluxes.ndim == 1 else np.array([method(row) for row in fluxes])
        else:
            norm = np.full(fluxes.shape[:-1], method, dtype=float)
        norm = np.asarray(norm, dtype=float)
        fluxes = func(fluxes, norm[..., None])
        if add_norm_pos is not None:
            pos = fluxes.shape[-1] if add_norm_pos == -1 else add_norm_pos
            fluxes = np.insert(fluxes, pos, norm, axis=-1)
        return fluxes

    def _scatter(self, fluxes: np.ndarray, flux_units: str = "nJy"):
        """Depth or noise-model scatter (``library.py:5906-5997``) along the last axis -- one galaxy ``(n_filt,)`` or a
        batch ``(n, n_filt)``; the depth branch consumes numpy's stream in the same order as a row-by-row loop would.
        Including the reference's quirks: depth errors are returned in uJy, and the AB-mode error carries a minus sign."""
        if self.ignore_scatter:
            return fluxes, None
        to_ujy = {"nJy": 1e-3, "uJy": 1.0, "mJy": 1e3, "Jy": 1e6}
        if self.depths is not None:
            depths = self.depths
            if flux_units == "AB":
                f_ujy = 10 ** ((fluxes - 23.9) / -2.5)
            else:
                f_ujy = fluxes * to_ujy[flux_units]
            if self.out_flux_unit == "AB" and not has_units(depths):
                depths_std = 10 ** ((np.asarray(depths, dtype=float) - 23.9) / -2.5) / self.depth_sigma
            else:
                depths_std = strip_units(depths, "uJy") / self.depth_sigma
            noisy = f_ujy + np.random.normal(loc=0, scale=depths_std, size=fluxes.shape)
            errors = depths_std
            if flux_units == "AB":
                with np.errstate(all="ignore"):
                    out = -2.5 * np.log10(noisy * 1e-6) + 8.9
                errors = -2.5 * depths_std / (np.log(10) * f_ujy)
            else:
                out = noisy / to_ujy[flux_units]
            return out, errors
        if self.noise_models is not None:
            scattered = np.zeros_like(fluxes, dtype=float)
            errors = np.zeros_like(fluxes, dtype=float)
            for i, code in enumerate(self.instrument.filters.filter_codes):
                model = self.noise_models.get(code)
                model.return_noise = True
                col = np.atleast_1d(fluxes[..., i])                    # one value, or this filter's column of a batch
                sf, sig = model.apply_noise(flux=col, true_flux_units=flux_units, out_units=self.out_flux_unit)
                scattered[..., i], errors[..., i] = np.asarray(sf).reshape(col.shape), np.asarray(sig).reshape(col.shape)
            return scattered, errors
        return fluxes, None

    def __call__(self, params):
        return self.simulate(params)
