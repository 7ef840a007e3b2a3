"""Spectroscopic training path (cfg 5): library spectra -> instrument-frame pixel rows, on the GPU.

Mirrors, for whole libraries at once, what the reference does galaxy by galaxy in Python:

* :func:`transform_spectrum` -- ``utils.py:185-254`` (same signature, one spectrum per call);
* :class:`SpectrumResampler` -- the batched form behind it (one CUDA block per galaxy; ``csrc/resample_kernel.cuh``);
* :func:`create_feature_array_from_raw_spectra` -- the redshift / resample / unit / crop / extra-feature steps of
  ``SBI_Fitter.create_feature_array_from_raw_spectra`` (``sbi_runner.py:1180-1427``).

There is no CPU fallback: without the CUDA library these raise.
"""

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _capi
from .units import has_units, strip_units

__all__ = ["SpectrumResampler", "transform_spectrum", "create_feature_array_from_raw_spectra", "write_spectral_library"]


def _um(x):
    """Wavelengths as float64 microns (plain numbers are taken as microns, like the reference's ``.to("um").value``)."""
    if has_units(x):
        return np.asarray(x.to("um").value, dtype=np.float64)
    return np.asarray(x, dtype=np.float64)


class SpectrumResampler:
    """Device-resident plan for one (model axis, instrument pixel grid, resolution curve) triple."""

    def __init__(self, theory_wave, observed_wave, resolution_curve_wave, resolution_curve_r, theory_r=np.inf,
                 trunc_constant=4.0, fill=0.0, device=0):
        self.lib = _capi.load()
        if self.lib.sb2_device_count() < 1:
            raise RuntimeError("synference_b200: no CUDA device visible; the spectral path has no CPU fallback")
        self.theory_wave = np.ascontiguousarray(_um(theory_wave))
        self.observed_wave = np.ascontiguousarray(_um(observed_wave))
        rw = np.ascontiguousarray(_um(resolution_curve_wave))
        rr = np.ascontiguousarray(np.asarray(strip_units(resolution_curve_r), dtype=np.float64))
        if rw.shape != rr.shape or rw.ndim != 1:
            raise ValueError("resolution curve: wavelengths and R must be 1-D arrays of equal length")
        d = _capi.ResampleDesc()
        d.n_lam, d.n_px, d.n_res = self.theory_wave.size, self.observed_wave.size, rw.size
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        d.theory_wave, d.observed_wave, d.res_wave, d.res_r = dp(self.theory_wave), dp(self.observed_wave), dp(rw), dp(rr)
        tr = None
        if np.ndim(theory_r) == 0:
            d.theory_r, d.theory_r_scalar = None, float(theory_r)
        else:
            tr = np.ascontiguousarray(np.asarray(theory_r, dtype=np.float64))
            if tr.shape != self.theory_wave.shape:
                raise ValueError("theory_r must be a scalar or one value per model wavelength")
            d.theory_r, d.theory_r_scalar = dp(tr), float("inf")
        d.trunc, d.fill = float(trunc_constant), float(fill)
        h = C.c_void_p()
        _capi.check(self.lib.sb2_resampler_create(C.byref(d), int(device), C.byref(h)), "sb2_resampler_create")
        self._h = h
        self.device = int(device)
        self.n_lam, self.n_px = int(d.n_lam), int(d.n_px)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sb2_resampler_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform(self, spectra, redshift):
        """``spectra`` (n, n_lam) float32 -- one row per galaxy -- and ``redshift`` (n,) -> (n, n_px) float32.

        numpy in, numpy out (copies through the device); CUDA torch tensors in, CUDA tensor out (no copies)."""
        try:
            import torch
        except ImportError:  # pragma: no cover
            torch = None
        if torch is not None and isinstance(spectra, torch.Tensor):
            if not spectra.is_cuda:
                raise ValueError("torch input must live on the GPU; pass numpy arrays for host data")
            sp = spectra.to(torch.float32).contiguous()
            z = torch.as_tensor(redshift, device=sp.device).to(torch.float64).contiguous()
            if sp.ndim != 2 or sp.shape[1] != self.n_lam or z.shape != (sp.shape[0],):
                raise ValueError(f"expected spectra (n, {self.n_lam}) and redshift (n,)")
            out = torch.empty((sp.shape[0], self.n_px), dtype=torch.float32, device=sp.device)
            st = torch.cuda.current_stream(sp.device).cuda_stream
            _capi.check(self.lib.sb2_resample_spectra(self._h, C.c_void_p(sp.data_ptr()), C.c_void_p(z.data_ptr()),
                                                      sp.shape[0], C.c_void_p(out.data_ptr()), C.c_void_p(st)),
                        "sb2_resample_spectra")
            return out
        sp = np.ascontiguousarray(spectra, dtype=np.float32)
        z = np.ascontiguousarray(redshift, dtype=np.float64)
        if sp.ndim != 2 or sp.shape[1] != self.n_lam or z.shape != (sp.shape[0],):
            raise ValueError(f"expected spectra (n, {self.n_lam}) and redshift (n,)")
        out = np.empty((sp.shape[0], self.n_px), dtype=np.float32)
        _capi.check(self.lib.sb2_resample_spectra_host(self._h, sp.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p),
                                                       sp.shape[0], out.ctypes.data_as(C.c_void_p)),
                    "sb2_resample_spectra_host")
        return out

    def transform_into(self, spectra, redshift, out):
        """Device form without allocations: CUDA tensors ``spectra`` (n, n_lam) float32 and ``redshift`` (n,) float64 in,
        pixels written to the caller's CUDA tensor ``out`` (n, n_px) float32 on the current stream."""
        import torch
        assert spectra.is_cuda and out.is_cuda and redshift.is_cuda
        assert spectra.dtype == torch.float32 and out.dtype == torch.float32 and redshift.dtype == torch.float64
        assert spectra.is_contiguous() and out.is_contiguous() and redshift.is_contiguous()
        n = spectra.shape[0]
        assert spectra.shape == (n, self.n_lam) and out.shape == (n, self.n_px) and redshift.shape == (n,)
        st = torch.cuda.current_stream(spectra.device).cuda_stream
        _capi.check(self.lib.sb2_resample_spectra(self._h, C.c_void_p(spectra.data_ptr()), C.c_void_p(redshift.data_ptr()), n,
                                                  C.c_void_p(out.data_ptr()), C.c_void_p(st)), "sb2_resample_spectra")
        return out

    def last_ms(self) -> float:
        ms = C.c_float()
        _capi.check(self.lib.sb2_resample_last_ms(self._h, C.byref(ms)), "sb2_resample_last_ms")
        return float(ms.value)


_PLAN_CACHE = {}


def transform_spectrum(theory_wave, theory_flux, z, observed_wave, resolution_curve_wave, resolution_curve_r,
                       theory_r=np.inf, trunc_constant=4.0):
    """``utils.py:185-254`` for one spectrum: returns ``(observed_wave, flux on observed_wave)``.  The plan for a given
    set of axes is built once and reused (callers loop over galaxies with fixed axes, ``sbi_runner.py:1322-1334``)."""
    tw, ow = _um(theory_wave), _um(observed_wave)
    rw, rr = _um(resolution_curve_wave), np.asarray(strip_units(resolution_curve_r), dtype=np.float64)
    tr = float(theory_r) if np.ndim(theory_r) == 0 else np.asarray(theory_r, dtype=np.float64).tobytes()
    key = (tw.tobytes(), ow.tobytes(), rw.tobytes(), rr.tobytes(), tr, float(trunc_constant))
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        if len(_PLAN_CACHE) >= 4:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE))).close()
        plan = _PLAN_CACHE[key] = SpectrumResampler(tw, ow, rw, rr, theory_r=theory_r, trunc_constant=trunc_constant)
    flux = np.asarray(strip_units(theory_flux), dtype=np.float32).reshape(1, -1)
    return observed_wave, plan.transform(flux, np.array([float(z)]))[0]


def create_feature_array_from_raw_spectra(spectra, wavelengths, parameter_array, parameter_names: Sequence[str],
                                          extra_features=("redshift",), crop_wavelength_range=None,
                                          normed_flux_units: str = "AB", raw_units: str = "nJy",
                                          resample_wavelengths=None, inst_resolution_wavelengths=None,
                                          inst_resolution_r=None, theory_r=np.inf, min_flux_value=-np.inf,
                                          max_flux_value=np.inf, device=0):
    """Feature rows from a spectral library (``Grid/Spectra``), as ``sbi_runner.py:1180-1427`` builds them.

    ``spectra`` is ``(N_lam, N_gal)`` (the library's layout); ``parameter_array`` is ``(N_gal, N_par)``.  With ``"redshift"``
    among the parameters the spectra are moved to the observed frame, smoothed to the instrument's resolution and rebinned
    onto ``resample_wavelengths`` (cropped to ``crop_wavelength_range``, taken as observed-frame) by the CUDA kernel.
    Returns ``(feature_array (N_gal, N_px + len(extra_features)) float64, feature_names, wavelengths_um)``.
    Flux normalisation callbacks (``flux_norm_method``) are not part of the batched path.
    """
    spectra = np.asarray(strip_units(spectra))
    wavs = _um(wavelengths)
    names = list(parameter_names)
    params = np.asarray(parameter_array, dtype=np.float64)
    if spectra.ndim != 2 or spectra.shape[0] != wavs.size:
        raise ValueError("spectra must be (N_lam, N_gal) on `wavelengths`")
    if params.shape[0] != spectra.shape[1]:
        raise ValueError("parameter_array must have one row per galaxy")
    extra_features = list(extra_features or [])
    if "redshift" in names:
        assert resample_wavelengths is not None, "resample_wavelengths must be provided when transforming to observed frame."
        assert inst_resolution_wavelengths is not None, "inst_resolution_wavelengths must be provided for convolution."
        assert inst_resolution_r is not None, "inst_resolution_r must be provided for convolution."
        new_w = _um(resample_wavelengths)
        if crop_wavelength_range is not None:
            lo, hi = _um(crop_wavelength_range)
            new_w = new_w[(new_w >= lo) & (new_w <= hi)]
        plan = SpectrumResampler(wavs, new_w, inst_resolution_wavelengths, inst_resolution_r, theory_r=theory_r, device=device)
        grid = plan.transform(np.ascontiguousarray(spectra.T, dtype=np.float32), params[:, names.index("redshift")]).astype(np.float64)
        plan.close()
        wavs = new_w
    else:
        grid = np.asarray(spectra.T, dtype=np.float64)
        if crop_wavelength_range is not None:
            lo, hi = _um(crop_wavelength_range)
            keep = (wavs >= lo) & (wavs <= hi)
            grid, wavs = grid[:, keep], wavs[keep]
    to_jy = {"Jy": 1.0, "mJy": 1e-3, "uJy": 1e-6, "nJy": 1e-9}
    if raw_units not in to_jy:
        raise ValueError(f"raw_units must be one of {sorted(to_jy)}")
    with np.errstate(divide="ignore", invalid="ignore"):
        if normed_flux_units == "AB":
            feat = -2.5 * np.log10(grid * to_jy[raw_units]) + 8.90
        elif normed_flux_units.startswith("log10 "):
            feat = np.log10(grid * (to_jy[raw_units] / to_jy[normed_flux_units[6:]]))
        elif normed_flux_units in to_jy:
            feat = grid * (to_jy[raw_units] / to_jy[normed_flux_units])
        else:
            raise ValueError(f"normed_flux_units '{normed_flux_units}' is not supported")
    np.clip(feat, min_flux_value, max_flux_value, out=feat)
    cols = []
    for name in extra_features:
        if name not in names:
            raise ValueError(f"Feature {name} not found in parameter names.")
        cols.append(params[:, names.index(name)][:, None])
    feature_array = np.concatenate([feat] + cols, axis=1) if cols else feat
    return feature_array, ["spectra"] + extra_features, wavs


def write_spectral_library(engine, resampler, params, out_dir=None, name="spectral_library", batch_size=65536, device=None,
                           keep_in_memory=False):
    """cfg 5's write path (``library.py:4887-4919``, ``:4610-4617``): spectra + photometry of a population, batch by batch.

    Each batch runs the whole chain on the device -- contraction kernel with its full-wavelength output kept in HBM, then the
    instrument-resolution resampling (``resampler``) -- and only the ``n_px`` pixels and ``n_filt`` fluxes per galaxy leave it:
    they are copied into one of TWO pinned host buffers on a side stream while the next batch is being synthesised, and a
    writer thread stores each filled buffer as a pair of uncompressed shards ``<name>_<i>_start<row>.spectra.npy`` (n, n_px)
    float32 and ``....photometry.npy`` (n, n_filt) float32.  Returns ``dict(shards=[paths], seconds=..., galaxies_per_s=..., bytes=...,
    spectra=..., photometry=...)`` (the arrays only with ``keep_in_memory``; ``out_dir=None`` skips the files)."""
    import os
    import time
    from concurrent.futures import ThreadPoolExecutor
    import torch
    dev = torch.device("cuda", engine.device if device is None else device)
    n = len(params)
    bs = int(min(batch_size, engine.max_batch))
    n_px, n_filt = resampler.n_px, engine.n_filt
    # pinned and device buffers are kept on the engine between calls (pinning 2 x bs x n_px x 4 B costs more than a batch)
    key = (str(dev), bs, n_px, n_filt, engine.n_lam)
    ws = getattr(engine, "_spectral_ws", None)
    if ws is None or ws["key"] != key:
        ws = engine._spectral_ws = dict(
            key=key,
            pin_px=[torch.empty((bs, n_px), dtype=torch.float32).pin_memory() for _ in range(2)],
            pin_ph=[torch.empty((bs, n_filt), dtype=torch.float32).pin_memory() for _ in range(2)],
            spec=torch.empty((bs, engine.n_lam), dtype=torch.float32, device=dev),
            flux=[torch.empty((bs, n_filt), dtype=torch.float32, device=dev) for _ in range(2)],
            pix=[torch.empty((bs, n_px), dtype=torch.float32, device=dev) for _ in range(2)],
            side=torch.cuda.Stream(device=dev))
    pin_px, pin_ph, spec, flux, pix, side = ws["pin_px"], ws["pin_ph"], ws["spec"], ws["flux"], ws["pix"], ws["side"]
    done = [torch.cuda.Event(), torch.cuda.Event()]          # slot's device buffers are free again (copy-out finished)
    writer = ThreadPoolExecutor(max_workers=1) if out_dir is not None else None
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
    all_px = np.empty((n, n_px), dtype=np.float32) if keep_in_memory else None
    all_ph = np.empty((n, n_filt), dtype=np.float32) if keep_in_memory else None
    shards, pending, busy = [], [None, None], [False, False]

    def store(slot, start, cnt, index):
        px, ph = pin_px[slot][:cnt].numpy(), pin_ph[slot][:cnt].numpy()
        if all_px is not None:
            all_px[start:start + cnt], all_ph[start:start + cnt] = px, ph
        if out_dir is not None:
            # plain .npy files: an .npz member costs a CRC-32 pass (~1 GB/s on one core) on top of the write
            path = os.path.join(out_dir, f"{name}_{index:05d}_start{start}.spectra.npy")
            np.save(path, px)
            np.save(path.replace(".spectra.npy", ".photometry.npy"), ph)
            return path
        return None

    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    with torch.cuda.device(dev):
        for i, a in enumerate(range(0, n, bs)):
            b = min(n, a + bs)
            cnt, slot = b - a, i & 1
            if pending[slot] is not None:           # the pinned buffers of this slot must have been written out
                shards.append(pending[slot].result())
                pending[slot] = None
            if busy[slot]:
                torch.cuda.current_stream(dev).wait_event(done[slot])
            dpar = engine.to_device(params.slice(slice(a, b)))
            engine.photometry_device(dpar, flux_base=flux[slot][:cnt], spectra=spec[:cnt])
            resampler.transform_into(spec[:cnt], dpar.tensors["redshift"], pix[slot][:cnt])
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(side):
                side.wait_event(ready)
                pin_px[slot][:cnt].copy_(pix[slot][:cnt], non_blocking=True)
                pin_ph[slot][:cnt].copy_(flux[slot][:cnt], non_blocking=True)
                done[slot].record()
            busy[slot] = True
            ev = done[slot]

            def job(slot=slot, a=a, cnt=cnt, i=i, ev=ev):
                ev.synchronize()
                return store(slot, a, cnt, i)
            pending[slot] = (writer.submit(job) if writer is not None else _Immediate(job))
        for slot in (0, 1):
            if pending[slot] is not None:
                shards.append(pending[slot].result())
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    if writer is not None:
        writer.shutdown()
    shards = sorted(p for p in shards if p)
    return dict(shards=shards, seconds=dt, galaxies_per_s=n / dt, bytes=int(n) * 4 * (n_px + n_filt), spectra=all_px, photometry=all_ph)


class _Immediate:
    def __init__(self, fn):
        self._v = fn()

    def result(self):
        return self._v

