"""Training-row builder: depth scatter + flux -> magnitude features on the GPU.

Mirrors the part of ``SBI_Fitter.create_feature_array_from_raw_photometry``
(``src/synference/sbi_runner.py:1429-2219``) that sits on the hot path: ``_apply_depths``
(:580-691), the AB conversion and error propagation (:1698-1716), the ``norm_mag_limit`` clip
(:1927-1932), optional normalisation by one band (:1781-1830), error columns, removal of
non-finite rows (:2083-2087), float32 ``(N_rows, N_feat)`` output (:2150) and the matching
``np.repeat`` of the parameter rows (``update_parameter_array``, :476-578).

The reference scatters once with ``N_scatters`` static replicas; :class:`ResampledFeatures` adds the
per-epoch resampling of BASELINE config 4 behind the same call (Philox counter = epoch), without the
table leaving HBM.
"""

from __future__ import annotations

from typing import List, Optional

import numpy as np

from .engine import depth_noise_features
from .units import Unit, has_units, strip_units

_TO_NJY = {"nJy": 1.0, "uJy": 1e3, "mJy": 1e6, "Jy": 1e9}


def depths_to_sigma_njy(depths, depth_sigma=5.0, n_filt=None):
    """sigma [nJy] per filter from 5-sigma depths given in AB magnitudes (bare numbers) or flux units."""
    if has_units(depths):
        d = np.asarray(strip_units(depths, "nJy"), dtype=np.float64)
    else:
        d = 10 ** ((np.asarray(depths, dtype=np.float64) - 23.9) / -2.5) * 1e3  # AB -> uJy -> nJy
    d = np.atleast_1d(d)
    if n_filt is not None and d.size == 1:
        d = np.full(n_filt, d[0])
    return d / depth_sigma


def create_feature_array_from_raw_photometry(
        phot_grid, raw_observation_names: List[str], raw_observation_units: str = "nJy",
        normalize_method: Optional[str] = None, normed_flux_units: str = "AB", scatter_fluxes=False,
        depths=None, depth_sigma: float = 5.0, include_errors_in_feature_array: bool = False,
        min_flux_pc_error: float = 0.0, norm_mag_limit: float = 50.0, remove_nan_inf: bool = True,
        photometry_to_remove: Optional[list] = None, drop_dropouts: bool = False,
        drop_dropout_fraction: float = 1.0, parameter_array=None, normals=None, seed: int = 0,
        epoch: int = 0, device: int = 0, return_torch: bool = False, empirical_noise_models=None,
        depth_indices=None, asinh_softening_parameters=None, normalization_unit: str = "AB",
        simulate_missing_fluxes: bool = False, missing_flux_value: float = 99.0, missing_flux_fraction: float = 0.0,
        missing_flux_options: Optional[list] = None, include_flags_in_feature_array: bool = False):
    """``(N_filters, N_gal)`` library photometry -> ``(feature_array (N_rows, N_feat) float32,
    feature_names, parameter_array (N_rows, N_par) | None)``.

    ``scatter_fluxes`` is the number of noisy replicas per galaxy (0/False: no scatter, one row per
    galaxy).  With ``normals`` (``(N_filters, N_rows)`` float64) the scatter is bit-exact with numpy's
    ``flux + np.random.normal(0, sigma)`` for the same draws; otherwise Philox4x32-10(seed, epoch).
    ``empirical_noise_models`` (``{filter name: model}``, instead of ``depths``) scatters with the per-filter empirical models
    on the device (``sbi_runner.py:1678-1692``): the models hand back magnitudes and their errors directly.
    2-D ``depths`` ``(k, N_filters)`` are k alternative depth sets: every (filter, scatter) picks one at random
    (``sbi_runner.py:626-647``; ``depth_indices (N_filters, scatter_fluxes)`` injects the pick, otherwise it is drawn from
    ``numpy.random.default_rng((seed, epoch))``).
    ``normed_flux_units="asinh"`` (``sbi_runner.py:1598-1625, 1660-1676, 1718-1730``): asinh magnitudes with the softening
    ``asinh_softening_parameters`` -- one flux per filter (a Quantity array, or a list / dict of quantities), or ``"SNR_x"``:
    x times the 1-sigma depth of each filter.  No magnitude limit is applied in this branch, as in the reference.
    Any other ``normed_flux_units`` (``sbi_runner.py:1734-1779``) is a flux unit (``"nJy"``, ``"uJy"``, ``"mJy"``, ``"Jy"``) or a
    scaling of one, ``"log10 nJy"`` / ``"log nJy"`` / ``"sqrt nJy"``, with the reference's error propagation; such rows are
    normalised by DIVISION -- by subtraction for the two logarithms, and, as in the reference, only when the rows carry errors.
    ``normalize_method`` = a filter name: that filter leaves the rows, the others are normalised by it, and the last column
    holds the filter's UNSCATTERED library flux in ``normalization_unit`` (``"AB"``, a flux unit, or ``"log10 <unit>"``;
    ``sbi_runner.py:1783-1831``), named ``norm_<filter>_<normalization_unit>`` (``:2024-2027``).  (The reference can only do
    this for at most one replica per galaxy -- its assignment of the column fails otherwise; here the value is repeated.)
    ``simulate_missing_fluxes`` (``sbi_runner.py:1976-2012``): every row gets a mask over its filters -- one of
    ``missing_flux_options`` (0/1 lists, 1 = missing) picked uniformly, or each band missing with probability
    ``missing_flux_fraction`` -- masked bands (and their errors) become ``missing_flux_value``, and
    ``include_flags_in_feature_array`` appends the mask as ``flag_<filter>`` columns after the errors.  The reference draws
    the masks from numpy's global stream; here they come from a device generator keyed by ``(seed, epoch)``.
    """
    import torch
    asinh = normed_flux_units == "asinh"
    other_scaling, other_unit = None, None
    if normed_flux_units not in ("AB", "asinh"):
        if str(normed_flux_units) in _TO_NJY:
            other_scaling, other_unit = "", str(normed_flux_units)
        else:
            try:
                other_scaling, other_unit = str(normed_flux_units).split(" ")
            except ValueError:
                raise ValueError("Don't understand normed_flux_units.If string, should be e.g. 'log10 nJy',Otherwise pass a Unit directly.")
            if other_scaling not in ("log10", "log", "", "sqrt"):
                raise ValueError(f'Scaling "{other_scaling}" not recognized. Use "log10", "log","", or "sqrt".')
            if other_unit not in _TO_NJY:
                raise ValueError(f"normed_flux_units: unknown flux unit '{other_unit}'")
        if empirical_noise_models is not None and depths is None:
            raise NotImplementedError("empirical noise models hand back AB (or asinh) magnitudes; other feature units take depths")
    if asinh:
        assert asinh_softening_parameters is not None, "asinh_softening_parameters must be provided for asinh normalization."
        if empirical_noise_models is not None and depths is None:
            raise NotImplementedError("asinh features with empirical noise models: use AsinhEmpiricalUncertaintyModel through "
                                      "apply_empirical_noise_models (it returns asinh magnitudes itself)")
    names = list(raw_observation_names)
    all_names = list(names)
    dev = torch.device("cuda", device)
    grid = phot_grid if isinstance(phot_grid, torch.Tensor) else torch.as_tensor(np.asarray(phot_grid, dtype=np.float64))
    grid = grid.to(dev, dtype=torch.float64)
    assert grid.shape[0] == len(names), "phot_grid must be (N_filters, N_gal)"
    if photometry_to_remove:
        keep = [i for i, n in enumerate(names) if n not in set(photometry_to_remove)]
        if len(keep) == len(names):
            raise ValueError(f"No matching photometry filters found in the raw photometry names: {photometry_to_remove}")
        if not keep:
            raise ValueError("No photometry filters left after removing the specified ones.")
        grid, names = grid[keep], [names[i] for i in keep]
    grid = grid * _TO_NJY[str(raw_observation_units)]
    n_filt, n_gal = grid.shape
    n_sc = int(scatter_fluxes) if scatter_fluxes else 1
    set_index = None
    empirical = bool(scatter_fluxes) and depths is None and empirical_noise_models is not None
    if empirical:
        sigma = None
    elif scatter_fluxes:
        assert depths is not None, "If scattering fluxes, depths or empirical noise models must be provided."
        if isinstance(depths, dict):  # keyed by filter name (sbi_runner.py:1639-1645)
            vals = [depths[n] for n in names]
            if has_units(vals[0]):
                sigma = np.array([float(strip_units(v, "nJy")) for v in vals]) / depth_sigma
            else:
                sigma = depths_to_sigma_njy(np.array(vals, dtype=float), depth_sigma)
        elif np.ndim(strip_units(depths)) == 2:
            d2 = np.asarray(strip_units(depths, "nJy") if has_units(depths)
                            else 10 ** ((np.asarray(depths, dtype=np.float64) - 23.9) / -2.5) * 1e3, dtype=np.float64)
            if d2.shape[1] != n_filt:
                raise ValueError(f"Mismatch in dimensions: photometry_array has {n_filt} rows but depths has "
                                 f"{d2.shape[1]} columns")
            sigma = d2 / depth_sigma
            set_index = (np.asarray(depth_indices) if depth_indices is not None
                         else np.random.default_rng((int(seed), int(epoch))).integers(0, d2.shape[0], size=(n_filt, n_sc)))
        else:
            sigma = depths_to_sigma_njy(depths, depth_sigma, n_filt)
        if sigma.shape[-1] != n_filt:
            raise ValueError(f"Mismatch in dimensions: photometry_array has {n_filt} rows but depths has "
                             f"{sigma.shape[0]} elements")
    else:
        sigma = np.zeros(n_filt)
        normals = torch.zeros((n_filt, n_gal), dtype=torch.float64, device=dev)
    if empirical:
        sub = {k: empirical_noise_models[k] for k in names} if all(k in empirical_noise_models for k in names) \
            else empirical_noise_models
        m, e = apply_empirical_noise_models(grid, names, sub, N_scatters=n_sc, flux_units="nJy", return_errors=True,
                                            normed_flux_units="AB", seed=seed, epoch=epoch, device=device)
        mags = torch.clamp(m.t(), max=norm_mag_limit).to(torch.float32)     # sbi_runner.py:1927-1932
        mags = torch.where(torch.isnan(m.t()), torch.full_like(mags, float("nan")), mags)
        errs = e.t().to(torch.float32)
    elif asinh:
        # softening per filter [Jy]
        sp = asinh_softening_parameters
        if isinstance(sp, str):
            assert sp.startswith("SNR_"), "If a string, asinh_softening_parameters must start with 'SNR_'."
            assert scatter_fluxes and sigma is not None and set_index is None, \
                "If setting asinh_softening_parameters from noise models, depths must be provided."
            b_jy = float(sp.split("_")[-1]) * np.asarray(sigma, dtype=np.float64) * 1e-9
        elif isinstance(sp, dict):
            b_jy = np.array([float(strip_units(sp[n], "Jy")) for n in names])
        elif has_units(sp):
            b_jy = np.atleast_1d(np.asarray(strip_units(sp, "Jy"), dtype=np.float64))
            if b_jy.size == len(all_names) and len(names) != len(all_names):
                b_jy = b_jy[[all_names.index(n) for n in names]]
        else:
            b_jy = np.array([float(strip_units(v, "Jy")) for v in sp])
            assert len(b_jy) == len(all_names), "asinh_softening_parameter must be a list of the same length as raw_observation_names"
            b_jy = b_jy[[all_names.index(n) for n in names]]
        if b_jy.size == 1:
            b_jy = np.full(n_filt, float(b_jy[0]))
        if b_jy.size != n_filt or not np.all(b_jy > 0):
            raise ValueError("asinh softening: one positive flux per filter")
        flux_gf = grid.t().contiguous()
        noisy, sig, _ = depth_noise_features(flux_gf, sigma, n_scatter=n_sc, normals=normals, seed=seed, epoch=epoch,
                                             norm_mag_limit=norm_mag_limit, min_flux_pc_error=min_flux_pc_error,
                                             want_flux=True, want_features=False, device=device, set_index=set_index)
        b = torch.as_tensor(b_jy, dtype=torch.float64, device=dev)[:, None]
        f_jy, e_jy = noisy * 1e-9, sig * 1e-9                              # (n_filt, n_rows)
        pog = 2.5 * np.log10(np.e)
        mags = (-pog * (torch.asinh(f_jy / (2 * b)) + torch.log(b / 3631.0))).t()       # utils.py:647-675
        errs = (pog * e_jy / torch.sqrt(f_jy * f_jy + (2 * b) ** 2)).t()                # utils.py:678-704
    elif other_unit is not None:
        # a flux unit, optionally scaled (sbi_runner.py:1734-1779): noisy fluxes and their sigmas from the kernel, then the
        # unit change, the scaling and the reference's error propagation
        flux_gf = grid.t().contiguous()
        noisy, sig, _ = depth_noise_features(flux_gf, sigma, n_scatter=n_sc, normals=normals, seed=seed, epoch=epoch,
                                             norm_mag_limit=norm_mag_limit, min_flux_pc_error=min_flux_pc_error,
                                             want_flux=True, want_features=False, device=device, set_index=set_index)
        ph, er = noisy.t() / _TO_NJY[other_unit], sig.t() / _TO_NJY[other_unit]
        if scatter_fluxes:
            if other_scaling == "log10":
                er = er / (ph * np.log(10.0))
            elif other_scaling == "log":
                er = er / ph
            elif other_scaling == "sqrt":
                er = er / (2.0 * torch.sqrt(ph))
        ph = {"log10": torch.log10, "log": torch.log, "sqrt": torch.sqrt, "": lambda x: x}[other_scaling](ph)
        mags, errs = ph, er
    else:
        flux_gf = grid.t().contiguous()                                   # (n_gal, n_filt), kernel layout
        _, _, feat = depth_noise_features(flux_gf, sigma, n_scatter=n_sc, normals=normals, seed=seed, epoch=epoch,
                                          norm_mag_limit=norm_mag_limit, min_flux_pc_error=min_flux_pc_error,
                                          want_flux=False, want_features=True, device=device, set_index=set_index)
        mags, errs = feat[:, :n_filt], feat[:, n_filt:]
    feature_names = list(names)
    zero_norm = None
    if normalize_method is not None:
        if normalize_method not in names:
            raise NotImplementedError("normalisation by a supplementary parameter is not part of the device path; "
                                      "use a filter name")
        j = names.index(normalize_method)
        norm = mags[:, j:j + 1]
        zero_norm = norm[:, 0] == 0.0                   # such rows are deleted (sbi_runner.py:1917-1925)
        others = [i for i in range(n_filt) if i != j]
        # magnitudes and logarithms are normalised by subtraction, fluxes by division; the reference switches the two
        # logarithmic scalings to subtraction only where it propagates errors, i.e. when the rows were scattered
        subtract = other_unit is None or (other_scaling in ("log10", "log") and bool(scatter_fluxes))
        mags = (mags[:, others] - norm) if subtract else (mags[:, others] / norm)
        if other_unit is None and not asinh:
            mags = torch.clamp(mags, max=norm_mag_limit)
        errs = errs[:, others]
        feature_names = [names[i] for i in others]
        # the last column: the filter's library flux BEFORE the scatter, in normalization_unit (sbi_runner.py:1788-1831)
        orig = grid[j].repeat_interleave(n_sc)
        nu = str(normalization_unit)
        nu_log = nu.startswith("log10")
        nu_unit = nu.split(" ")[1] if nu_log else nu
        if nu_unit == "AB":
            norm_col = -2.5 * torch.log10(orig * 1e-3) + 23.9
        elif nu_unit in _TO_NJY:
            norm_col = orig / _TO_NJY[nu_unit]
        else:
            raise ValueError(f"normalization_unit: unknown unit '{nu_unit}'")
        if nu_log:
            norm_col = torch.log10(norm_col)
            norm_col = torch.where(torch.isinf(norm_col), torch.zeros_like(norm_col), norm_col)
        last = [norm_col[:, None].to(mags.dtype)]
        tail = [f"norm_{normalize_method}_{normalization_unit}"]
    else:
        last, tail = [], []
    with_errs = bool(include_errors_in_feature_array and scatter_fluxes)
    band_names = list(feature_names)
    flags = []
    if simulate_missing_fluxes:
        nb_, n_rows_ = mags.shape[1], mags.shape[0]
        gen = torch.Generator(device=dev)
        gen.manual_seed((int(seed) * 1_000_003 + int(epoch) * 7919 + 0x6D697373) & 0x7FFFFFFFFFFFFFFF)
        if missing_flux_options is not None:
            opts = torch.as_tensor(np.asarray(missing_flux_options, dtype=np.float32), device=dev)
            if opts.ndim != 2 or opts.shape[1] != nb_:
                raise ValueError(f"missing_flux_options: every mask needs {nb_} entries (one per filter in the rows)")
            mask = opts[torch.randint(0, opts.shape[0], (n_rows_,), generator=gen, device=dev)]
        else:
            mask = (torch.rand((n_rows_, nb_), generator=gen, device=dev) < float(missing_flux_fraction)).to(torch.float32)
        miss = mask == 1.0
        mags = torch.where(miss, torch.full_like(mags, float(missing_flux_value)), mags)
        if with_errs:
            errs = torch.where(miss, torch.full_like(errs, float(missing_flux_value)), errs)
        if include_flags_in_feature_array:
            flags = [mask.to(mags.dtype)]
    elif include_flags_in_feature_array:
        flags = [torch.zeros_like(mags)]
    cols = [mags] + ([errs] if with_errs else []) + flags + last
    if with_errs:
        feature_names = feature_names + [f"unc_{n}" for n in band_names]
    if flags:
        feature_names = feature_names + [f"flag_{n}" for n in band_names]
    feature_names = feature_names + tail
    out = torch.cat(cols, 1) if len(cols) > 1 else cols[0]
    keep_rows = torch.ones(out.shape[0], dtype=torch.bool, device=dev)
    if zero_norm is not None:
        keep_rows &= ~zero_norm
    if remove_nan_inf:
        keep_rows &= torch.isfinite(out).all(1)
    if drop_dropouts:
        nb = mags.shape[1]
        keep_rows &= ~((out[:, :nb].abs() >= norm_mag_limit).sum(1) >= nb * drop_dropout_fraction)
    params = None
    if parameter_array is not None:
        params = torch.as_tensor(np.asarray(parameter_array, dtype=np.float32)).to(dev)
        params = params.repeat_interleave(n_sc, dim=0)                # update_parameter_array (np.repeat)
    if not bool(keep_rows.all()):
        out = out[keep_rows]
        params = params[keep_rows] if params is not None else None
    if out.shape[0] == 0:
        raise ValueError("All rows in the feature array were deleted. Please check the input parameters.")
    out = out.contiguous().to(torch.float32)
    if return_torch:
        return out, feature_names, params
    return out.cpu().numpy(), feature_names, (None if params is None else params.cpu().numpy())


def apply_empirical_noise_models(photometry_array, phot_names, empirical_noise_models, N_scatters: int = 5,
                                 min_flux_pc_error: float = 0.0, flux_units: str = "AB", return_errors: bool = False,
                                 normed_flux_units: str = "AB", draws=None, seed: int = 0, epoch: int = 0, device=0):
    """``SBI_Fitter._apply_empirical_noise_models`` (``sbi_runner.py:813-903``) in one CUDA launch for all filters.

    ``photometry_array`` is ``(N_f, N_gal)`` in ``flux_units``; each column is repeated ``N_scatters`` times
    (``np.repeat(..., axis=1)``) and every filter row goes through its model's ``apply_noise(flux, true_flux_units=flux_units,
    out_units=normed_flux_units)``.  numpy in -> numpy out; CUDA tensors in -> CUDA tensors out.  ``draws`` (4, N_f, N_rows)
    injects the random numbers (see ``include/synference_b200.h``); otherwise Philox keyed by ``(seed, epoch)``.
    ``min_flux_pc_error`` is accepted and unused, as in the reference."""
    import ctypes as C

    import torch

    from . import _capi
    from .noise_models import AsinhEmpiricalUncertaintyModel, GeneralEmpiricalUncertaintyModel
    if not isinstance(empirical_noise_models, dict):
        raise ValueError("empirical_noise_models must be a dictionary")
    for name in phot_names:
        if name not in empirical_noise_models:
            raise ValueError(f"No empirical noise model found for filter {name}. Please provide a valid model.")
    for name in empirical_noise_models:
        if name not in phot_names:
            raise ValueError(f"Filter {name} in empirical_noise_models is not in phot_names: {phot_names}.")
    lib = _capi.load()
    if lib.sb2_device_count() < 1:
        raise RuntimeError("synference_b200: no CUDA device visible; the noise path has no CPU fallback")
    models = (_capi.EmpiricalModel * len(phot_names))()
    for i, name in enumerate(phot_names):
        mod = empirical_noise_models[name]
        if not isinstance(mod, (GeneralEmpiricalUncertaintyModel, AsinhEmpiricalUncertaintyModel)):
            raise NotImplementedError(f"{type(mod).__name__} has no device form (General / Asinh empirical models do)")
        models[i] = mod.device_model(true_flux_units=flux_units, out_units=normed_flux_units)
    dev = torch.device("cuda", int(device))
    was_numpy = not isinstance(photometry_array, torch.Tensor)
    phot = torch.as_tensor(np.asarray(photometry_array, dtype=np.float64) if was_numpy else photometry_array, device=dev)
    phot = phot.to(torch.float64)
    if phot.ndim != 2 or phot.shape[0] != len(phot_names):
        raise ValueError("photometry_array must be (N_filters, N_galaxies)")
    rep = torch.repeat_interleave(phot, int(N_scatters), dim=1).contiguous()
    n = rep.shape[1]
    d_ptr = None
    if draws is not None:
        dr = torch.as_tensor(np.asarray(draws, dtype=np.float64) if not isinstance(draws, torch.Tensor) else draws,
                             device=dev).to(torch.float64).contiguous()
        if tuple(dr.shape) != (4, len(phot_names), n):
            raise ValueError(f"draws must have shape (4, {len(phot_names)}, {n})")
        d_ptr = C.c_void_p(dr.data_ptr())
    out_f, out_s = torch.empty_like(rep), torch.empty_like(rep)
    with torch.cuda.device(dev):      # the entry point launches on the current device
        st = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(lib.sb2_empirical_noise(C.c_void_p(rep.data_ptr()), n, len(phot_names), models, d_ptr, int(seed), int(epoch),
                                            C.c_void_p(out_f.data_ptr()), C.c_void_p(out_s.data_ptr()), C.c_void_p(st)),
                    "sb2_empirical_noise")
    if was_numpy:
        out_f, out_s = out_f.cpu().numpy(), out_s.cpu().numpy()
    return (out_f, out_s) if return_errors else out_f


class ResampledFeatures:
    """Library photometry resident in HBM; a fresh noise realisation of every row per epoch.

    ``epoch(e)`` returns ``(features (N_gal*n_scatter, N_feat) float32 CUDA tensor, params)`` for epoch
    ``e``; the Philox counter is (row, filter) and the key (seed, epoch), so epochs are reproducible and
    independent of the world size.
    """

    def __init__(self, phot_grid, names, depths, parameter_array=None, n_scatter=1, seed=42, device=0, **kw):
        import torch
        self.dev = device
        self.grid = torch.as_tensor(np.asarray(phot_grid, dtype=np.float64)).to(torch.device("cuda", device))
        self.names, self.depths, self.n_scatter, self.seed, self.kw = list(names), depths, n_scatter, seed, kw
        self.params = parameter_array

    def epoch(self, e: int):
        f, names, p = create_feature_array_from_raw_photometry(
            self.grid, self.names, scatter_fluxes=self.n_scatter, depths=self.depths, seed=self.seed, epoch=e,
            parameter_array=self.params, device=self.dev, return_torch=True, **self.kw)
        self.feature_names = names
        return f, p

    def __len__(self):
        """Training rows per epoch."""
        return int(self.grid.shape[1]) * int(self.n_scatter)

    def loader(self, batch_size: int, epochs: int, shuffle: bool = True, rank: int = 0, world_size: int = 1, start_epoch: int = 0):
        """Mini-batches ``(features, parameters)`` for ``epochs`` epochs: what a ``DataLoader`` over the static, pre-scattered
        array of the reference gives (``sbi_runner.py``: the feature array is built once with ``scatter_fluxes`` replicas),
        except that every epoch is a fresh noise realisation drawn on the device.  The row order of an epoch is a permutation
        keyed by ``(seed, epoch)``, the same on every rank; rank ``r`` of ``world_size`` takes every ``world_size``-th batch,
        so the ranks of a data-parallel trainer see disjoint batches of one global epoch."""
        import torch
        n = len(self)
        for e in range(int(start_epoch), int(start_epoch) + int(epochs)):
            feats, par = self.epoch(e)
            par_t = None if par is None else (par if isinstance(par, torch.Tensor) else torch.as_tensor(np.asarray(par))).to(feats.device)
            if shuffle:
                g = torch.Generator(device="cpu")
                g.manual_seed((int(self.seed) * 1_000_003 + e) & 0x7FFFFFFFFFFFFFFF)
                order = torch.randperm(n, generator=g).to(feats.device)
            else:
                order = torch.arange(n, device=feats.device)
            for b, lo in enumerate(range(0, n, int(batch_size))):
                if b % int(world_size) != int(rank):
                    continue
                idx = order[lo:lo + int(batch_size)]
                yield feats[idx], (None if par_t is None else par_t[idx])
