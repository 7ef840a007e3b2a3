"""Host side of the CUDA hot path: lowers a (grid, emission model, filters, cosmology) model to
device tables once, then runs batches of galaxy parameters through the C ABI.

What this replaces in the reference: the per-batch ``Pipeline(...)`` set-up and ``run()`` of
``GalaxyBasis.process_galaxies`` (``library.py:2571-2619``) and the single-galaxy chain of
``GalaxySimulator.simulate`` (``library.py:5711-5772``).  PyTorch is used only for device
memory and streams when the caller keeps data on the GPU; the host-buffer entry point needs
no torch at all.
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _capi, fastmath as _fm, igm as _igm
from .cosmology import Planck18
from .parametric import (EmissionModel, FilterCollection, Grid, SFH_MAX_PARAMS, ZD_NORMAL_LINEAR,
                         pack_sfh, pack_zdist)
from .units import strip_units

CHUNK_COLS = 256


def tf32_split(x64: np.ndarray):
    """(hi, lo) float32 arrays with hi = tf32_rna(x), lo = tf32_rna(x - hi)  (3xTF32 operands)."""
    def rna(a32):
        bits = a32.view(np.uint32)
        return ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    hi = rna(np.ascontiguousarray(x64, dtype=np.float32))
    lo = rna(np.ascontiguousarray(x64 - hi.astype(np.float64), dtype=np.float32))
    return hi, lo


@dataclass
class GalaxyParams:
    """Struct-of-arrays parameter block for a population (float64, one entry per galaxy)."""

    redshift: np.ndarray
    sfh_type: int
    sfh_rows: np.ndarray                 # (N, stride) [min_age, max_age, p0, ...] in yr
    zd_type: int
    zd_value: np.ndarray
    zd_sigma: Optional[np.ndarray] = None
    log_mass: Optional[np.ndarray] = None
    tau_v: Optional[np.ndarray] = None
    coef_att: Optional[np.ndarray] = None
    coef_unatt: Optional[np.ndarray] = None
    dust_slope: Optional[np.ndarray] = None     # per-galaxy dust-curve slope / bump amplitude (models built with
    dust_ampl: Optional[np.ndarray] = None      # Calzetti2000(slope="...", ampl="..."))
    fesc_lya: Optional[np.ndarray] = None       # per-galaxy Lyman-alpha escape fraction (fesc_ly_alpha="<name>" models)
    tau_v_birth: Optional[np.ndarray] = None    # birth-cloud optical depth (BimodalPacmanEmission; tau_v is then the ISM's)
    max_age_from_z: bool = False
    norm_mask: int = 0
    age_zmax_gyr: float = 0.0
    extra: dict = field(default_factory=dict)

    def __len__(self):
        return int(np.asarray(self.redshift).shape[0])

    @classmethod
    def from_objects(cls, redshifts, sfhs, metal_dists, log_mass=None, tau_v=None, **kw):
        """Lower the reference-style per-galaxy object lists once (``library.py:2168-2261``)."""
        sfh_type, rows = pack_sfh(sfhs)
        zd_type, zv, zs = pack_zdist(metal_dists)
        n = len(np.atleast_1d(redshifts))
        bc = lambda a: None if a is None else np.broadcast_to(  # noqa: E731
            np.asarray(strip_units(a), dtype=np.float64), (n,)).copy()
        if rows.shape[0] == 1 and n > 1:
            rows = np.repeat(rows, n, axis=0)
        used = getattr(sfhs, "n_used", None)       # an SFHArray knows its column count: no scan of the (N, 24) table
        if used is None:
            used = max(2, int(np.max(np.nonzero(np.abs(rows).sum(0))[0], initial=1)) + 1)
        return cls(redshift=bc(redshifts), sfh_type=sfh_type, sfh_rows=np.ascontiguousarray(rows[:, :used]),
                   zd_type=zd_type, zd_value=bc(zv), zd_sigma=bc(zs) if zd_type >= ZD_NORMAL_LINEAR else None,
                   log_mass=bc(log_mass), tau_v=bc(tau_v), **kw)

    def slice(self, sl):
        g = lambda a: None if a is None else a[sl]  # noqa: E731
        return GalaxyParams(self.redshift[sl], self.sfh_type, self.sfh_rows[sl], self.zd_type,
                            self.zd_value[sl], g(self.zd_sigma), g(self.log_mass), g(self.tau_v),
                            g(self.coef_att), g(self.coef_unatt), g(self.dust_slope), g(self.dust_ampl), g(self.fesc_lya),
                            g(self.tau_v_birth), self.max_age_from_z, self.norm_mask, self.age_zmax_gyr)


def geometric_ratio(lam):
    lam = np.asarray(lam, dtype=np.float64)
    r = lam[1:] / lam[:-1]
    q = float(np.exp(np.log(lam[-1] / lam[0]) / (len(lam) - 1)))
    if np.max(np.abs(r / q - 1.0)) > 1e-9:
        raise ValueError(
            "the CUDA path needs grid and filters on one shared constant-R wavelength axis "
            "(generate_constant_R, as every production script of the reference builds); resample the "
            "Grid and FilterCollection with new_lam=... first")
    return q


DUST_W0 = 1.0e14      # Hz: keeps the float32 trapezoid weights and the emission's shape near 1


def _dust_emission_tables(generator, lam, n_pad, uv, lo_l, hi_l, off_l, n_blue):
    """Energy balance: ``total = emergent + E_abs g(nu)`` with ``E_abs`` the trapezoid over frequency of what the
    screen removed.  The kernel forms ``E_abs`` per galaxy (``dust_wnu``); since the filter sums are linear, the emission's
    share of every filter numerator is ``E_abs`` times a function of the integer redshift shift m only, tabulated here in
    float64: ``dust_duv[m][f] = sum_i g_i (U_f, V_f)[i + m]``."""
    n_lam = lam.size
    nu = 2.99792458e18 / lam
    w = np.empty(n_lam)
    w[1:-1] = 0.5 * (nu[:-2] - nu[2:])
    w[0], w[-1] = 0.5 * (nu[0] - nu[1]), 0.5 * (nu[-2] - nu[-1])
    wnu = np.zeros(n_pad, dtype=np.float32)
    wnu[:n_lam] = w / DUST_W0
    g = np.asarray(generator.shape(lam), dtype=np.float64) * DUST_W0
    if n_blue > 0:
        if g[:n_blue].max() > 1e-30 * g.max():
            raise NotImplementedError("the dust emission is not negligible at wavelengths the IGM absorbs; the batched path "
                                      "adds it after the IGM step")
        g[:n_blue] = 0.0
    g32 = g.astype(np.float32).astype(np.float64)    # what the spectra path adds; the filter tables use the same values
    m_len = int(max(hi_l)) + 3
    n_filt = len(lo_l)
    ends = list(off_l[1:]) + [uv.shape[0]]
    pad = m_len + max(e - o for o, e in zip(off_l, ends))
    gpad = np.concatenate([np.zeros(pad), g32, np.zeros(pad)])
    duv = np.zeros((m_len, n_filt, 2))
    for f, (lo, o, e) in enumerate(zip(lo_l, off_l, ends)):
        win = np.lib.stride_tricks.sliding_window_view(gpad, e - o)
        rows = win[pad + lo - 2 - np.arange(m_len)]            # row m: g[lo-2-m+k], k = 0..len-1   (n = i + m)
        duv[:, f, :] = rows @ uv[o:e].astype(np.float64)
    return dict(dust_wnu=wnu, dust_g=np.ascontiguousarray(g32, dtype=np.float32),
                dust_duv=np.ascontiguousarray(duv, dtype=np.float32), dust_m_len=m_len)


X_BINS = 192          # pseudo-bins of the absorbed-energy sum: nodes in kappa (a multiple of 192 = lcm of the kernels' chunks)
X_DEGREE = 7          # Lagrange degree of the projection onto the nodes
X_H_TAU_MAX = 0.42    # node spacing x optical depth up to which the projection is good to < 1e-6 of E_abs (measured)


def _energy_pseudo_bins(kappa, wnu, comps, absorbing, n_bins=X_BINS, degree=X_DEGREE):
    """Project the absorbed-energy sum onto nodes in kappa.

    ``E_abs = sum_i wnu_i L_i (1 - exp(-tau kappa_i))`` (``L`` the light before the screen) sees a wavelength only
    through ``kappa_i``.  With Lagrange weights ``l_j`` of degree ``degree`` on ``n_bins`` uniform nodes over the curve's range,
    ``exp(-tau kappa_i) ~= sum_j l_j(kappa_i) exp(-tau node_j)`` and therefore
    ``E_abs ~= sum_j [sum_i wnu_i l_j(kappa_i) L_i] (1 - exp(-tau node_j))`` (the weights sum to 1, so tau = 0 stays exact):
    the bracket is linear in the grid, i.e. ``n_bins`` extra rows of it.  Returns ``(nodes, rows)`` with ``rows[c]`` of shape
    ``(n_bins, n_z, n_age)`` for every component of ``comps`` ((age, Z, lam) arrays; zero rows where ``absorbing[c]`` is False)."""
    lo, hi = float(kappa.min()), float(kappa.max())
    span = max(hi - lo, 1e-6)
    # (nodes as the float32 values the kernel's kappa table holds, so that the weights belong to exactly those nodes)
    nodes = np.linspace(lo - 1e-6 * span, hi + 1e-6 * span, n_bins).astype(np.float32).astype(np.float64)
    h = (nodes[-1] - nodes[0]) / (n_bins - 1)
    first = np.clip(np.floor((kappa - nodes[0]) / h).astype(np.int64) - (degree - 1) // 2, 0, n_bins - degree - 1)
    proj = np.zeros((kappa.size, n_bins))
    rows_i = np.arange(kappa.size)
    for a in range(degree + 1):
        wgt = np.ones_like(kappa)
        for b in range(degree + 1):
            if a != b:
                wgt = wgt * (kappa - nodes[first + b]) / (nodes[first + a] - nodes[first + b])
        proj[rows_i, first + a] = wgt
    proj *= wnu[:, None]
    rows = [np.einsum("azl,lj->jza", np.asarray(c, dtype=np.float64), proj) if on else None for c, on in zip(comps, absorbing)]
    return nodes, h, rows


def build_tables(grid: Grid, emission_model: EmissionModel, emission_key: str, filters: FilterCollection,
                 cosmo=Planck18, igm=True, variant="nu", z_table_max=100.0):
    """All float64 host-side derivations behind ``sb2_model_desc`` (kept as numpy arrays)."""
    lam = np.asarray(grid.lam, dtype=np.float64)
    n_lam = lam.size
    try:
        q = geometric_ratio(lam)
        general = None
    except ValueError:
        # not a constant-R axis (the reference's README / tests keep the grid's native one): the contraction kernel then
        # only writes spectra (a dummy full-axis "filter" drives it) and the filters are integrated from those spectra by
        # general_filter_kernel with the reference's general semantics -- slower, same results
        if np.any(np.diff(lam) <= 0):
            raise ValueError("the wavelength axis must increase")
        q = 1.0e30
        off = np.concatenate([[0], np.cumsum([len(f.lam) for f in filters])]).astype(np.int64)
        general = dict(off=off, lam=np.concatenate([np.asarray(f.lam, dtype=np.float64) for f in filters]),
                       t=np.concatenate([np.asarray(f.t, dtype=np.float64) for f in filters]), n_filt=len(filters))
    att, un = emission_model.recipe(emission_key)
    has_att, has_un = bool(np.any(att != 0)), bool(np.any(un != 0))
    dust = emission_model.dust_curve
    dust_free = bool(getattr(emission_model, "dust_free", lambda k: False)(emission_key))
    two_screens = bool(getattr(emission_model, "two_screens", lambda k: False)(emission_key))
    if has_att and dust is None and not dust_free:
        raise ValueError(f"spectrum '{emission_key}' is dust attenuated but the emission model has no dust_curve")
    screen = (lambda: np.zeros_like(lam)) if dust_free else (lambda: dust.get_tau(lam))
    if two_screens:           # (young, old) reprocessed light, both attenuated: always two components
        comps, kappa = [att, un], screen()
        has_att = has_un = True
    elif has_att and has_un:
        comps, kappa = [att, un], screen()
    elif has_att:
        comps, kappa = [att], screen()
    else:
        comps, kappa = [un], None
    n_comp = len(comps)
    na, nz = grid.log10ages.size, grid.metallicity.size
    k = na * nz
    # every metallicity's columns start 32-byte aligned: TMA needs a 16-byte aligned box origin, and measured on
    # B200 a 32-byte aligned origin streams ~5% faster than a 16-byte aligned one (no further gain at 128 B)
    na_pad = (na + 7) // 8 * 8
    k_pad = (na_pad * nz + 31) // 32 * 32
    lch = CHUNK_COLS // n_comp
    n_chunk = (n_lam + lch - 1) // lch
    grid_scale = float(max(c.max() for c in comps))
    if not np.isfinite(grid_scale) or grid_scale <= 0:
        raise ValueError("grid spectra must be finite and not all zero")
    # dust emission: the absorbed-energy sum through pseudo-bins (see _energy_pseudo_bins) when the exponent is tau x ONE
    # global curve -- a global dust curve, and for two screens a birth-cloud curve proportional to the ISM one
    x_bin0 = x_bins = 0
    x_nodes = x_rows = None
    x_h = x_ratio = 0.0
    wants_energy = (general is None and kappa is not None and not dust_free
                    and getattr(emission_model, "has_dust_emission", lambda k: False)(emission_key))
    if wants_energy and not getattr(dust, "per_galaxy", False) and not os.environ.get("SB2_NO_PSEUDO_BINS"):
        kb = emission_model.dust_curve_birth.get_tau(lam) if two_screens else None
        big = np.abs(kappa) > 1e-6 * np.abs(kappa).max()
        x_ratio = float(np.median(kb[big] / kappa[big])) if two_screens else 0.0
        if not two_screens or np.allclose(kb, x_ratio * kappa, rtol=1e-9, atol=1e-12 * np.abs(kappa).max()):
            nu = 2.99792458e18 / lam
            w_nu = np.empty(n_lam)
            w_nu[1:-1] = 0.5 * (nu[:-2] - nu[2:])
            w_nu[0], w_nu[-1] = 0.5 * (nu[0] - nu[1]), 0.5 * (nu[-2] - nu[-1])
            # which components the screen acts on: both of the two-screen model, else the attenuated one (the first)
            x_nodes, x_h, x_rows = _energy_pseudo_bins(np.asarray(kappa, dtype=np.float64), w_nu / DUST_W0, comps,
                                                       [True, True] if two_screens else [True] + [False] * (n_comp - 1))
            x_bins = X_BINS
            x_bin0 = (n_lam + 191) // 192 * 192
            n_chunk = (x_bin0 + x_bins + lch - 1) // lch
    gt = np.zeros((n_chunk, n_comp, lch, k_pad), dtype=np.float64)
    for ci, comp in enumerate(comps):
        # (age, Z, lam) -> (lam, Z, age) -> rows lam, columns k = iz*na_pad + ia
        pad = np.zeros((n_chunk * lch, nz, na_pad))
        pad[:n_lam, :, :na] = np.transpose(comp, (2, 1, 0)) / grid_scale
        if x_bins and x_rows[ci] is not None:
            pad[x_bin0:x_bin0 + x_bins, :, :na] = x_rows[ci] / grid_scale
        gt[:, ci, :, :nz * na_pad] = pad.reshape(n_chunk, lch, nz * na_pad)
    gt = gt.reshape(n_chunk * CHUNK_COLS, k_pad)
    gt_hi, gt_lo = tf32_split(gt)
    kap = d0 = l2 = kap_birth = None
    if kappa is not None:
        kap = np.zeros(n_chunk * lch, dtype=np.float32)
        kap[:n_lam] = kappa
        if x_bins:
            kap[x_bin0:x_bin0 + x_bins] = x_nodes
        if two_screens:
            kap_birth = np.zeros_like(kap)
            kap_birth[:n_lam] = emission_model.dust_curve_birth.get_tau(lam)
            if x_bins:
                kap_birth[x_bin0:x_bin0 + x_bins] = x_ratio * x_nodes
        if getattr(dust, "per_galaxy", False) and not dust_free:
            # per-galaxy slope / bump amplitude: kappa holds the curve at slope = 0, ampl = 0 (get_tau already used 0 for
            # the string-named ones; a numeric one is broadcast to every galaxy by SynthEngine._fill)
            k0, dd, ll = dust.components(lam)
            kap[:n_lam] = k0
            d0 = np.zeros_like(kap)
            l2 = np.zeros_like(kap)
            d0[:n_lam], l2[:n_lam] = dd, ll
    # ---- filters on the shared axis -> (U, V) weight pairs (SURVEY A9 on a geometric grid):
    #      sample weight = (1-beta) U[n] + beta V[n], denominator = (1-beta) sum(U) + beta sum(V)
    if variant not in ("nu", "lam"):
        raise ValueError("variant must be 'nu' or 'lam'")
    c_left, c_right = ((q - 1) / 2, (1 - 1 / q) / 2) if variant == "nu" else ((1 - 1 / q) / 2, (q - 1) / 2)
    lo_l, hi_l, off_l, su_l, sdv_l, uv = [], [], [], [], [], []
    if general is not None:
        lo_l, hi_l, off_l, su_l, sdv_l = [1], [2], [0], [1.0], [1.0]      # a two-bin placeholder: spec_out multiplies every chunk
        uv = [np.zeros((2 - 1 + 4, 2))]
    for f in (filters if general is None else []):
        if f.lam.shape != lam.shape or np.max(np.abs(f.lam / lam - 1)) > 1e-12:
            raise ValueError(f"filter {f.filter_code} is not tabulated on the grid's wavelength axis; "
                             "use FilterCollection.resample_filters(new_lam=grid.lam)")
        nzi = np.nonzero(f.t > 0)[0]
        if nzi.size < 2:
            raise ValueError(f"filter {f.filter_code} has fewer than two in-band samples on this grid")
        lo, hi = int(nzi[0]), int(nzi[-1])
        if nzi.size != hi - lo + 1:
            raise ValueError(f"filter {f.filter_code} has interior zeros; not supported by the batched path")
        if lo < 1 or hi > n_lam - 2:
            raise ValueError(f"filter {f.filter_code} is truncated by the wavelength grid")
        n = np.arange(lo - 2, hi + 2)                 # table index k -> n = lo - 2 + k
        tpad = np.concatenate([[0.0, 0.0], f.t, [0.0, 0.0]])
        t_n, t_n1 = tpad[n + 2], tpad[n + 3]
        w = np.where((n >= lo) & (n <= hi - 1), c_left + c_right, 0.0)
        w = np.where(n == lo - 1, c_right, w)
        w = np.where(n == hi, c_left, w)
        u, v = w * t_n, w * t_n1
        lo_l.append(lo); hi_l.append(hi); off_l.append(sum(len(x) for x in uv))
        su_l.append(u.sum()); sdv_l.append(v.sum())
        uv.append(np.stack([u, v], 1))
    uv = np.concatenate(uv, 0).astype(np.float32)
    tables = dict(
        n_age=na, n_z=nz, n_lam=n_lam, n_comp=n_comp, n_filt=len(lo_l), n_age_pad=na_pad, k_pad=k_pad, n_chunk=n_chunk,
        log10ages=np.ascontiguousarray(grid.log10ages, dtype=np.float64),
        metallicities=np.ascontiguousarray(grid.metallicity, dtype=np.float64),
        gt_hi=gt_hi, gt_lo=gt_lo, grid_scale=grid_scale, kappa=kap, dust_d0=d0, dust_l2=l2, kappa_birth=kap_birth,
        dust_global=(float(getattr(dust, "slope", 0.0)), float(getattr(dust, "ampl", 0.0))) if d0 is not None else None,
        lam0=float(lam[0]), q=q,
        interp_variant=0 if variant == "nu" else 1,
        filt_lo=np.array(lo_l, dtype=np.int32), filt_hi=np.array(hi_l, dtype=np.int32),
        filt_off=np.array(off_l, dtype=np.int32), filt_uv=np.ascontiguousarray(uv),
        filt_su=np.array(su_l, dtype=np.float64), filt_sdv=np.array(sdv_l, dtype=np.float64),
        # a lone component is the kernel's component A whichever grid it came from: its per-galaxy coefficient
        # travels as coef_att (SynthEngine._fill)
        single_is_unatt=(n_comp == 1 and not has_att),
        general=general,
        x_bin0=x_bin0, x_bins=x_bins,
        # largest tau_V (+ ratio x tau_V_birth) the pseudo-bins are accurate for; beyond it a batch sums over the whole axis
        x_tau_max=(X_H_TAU_MAX / x_h if x_bins else 0.0), x_birth_ratio=x_ratio,
    )
    if igm:
        laf, dla = (igm if isinstance(igm, tuple) else (_igm.INOUE14_LAF, _igm.INOUE14_DLA))
        tables["igm"] = _igm.device_tables(lam, laf, dla)
    else:
        tables["igm"] = None
    tables.update(dust_wnu=None, dust_g=None, dust_duv=None, dust_m_len=0)
    if general is not None and getattr(emission_model, "has_dust_emission", lambda k: False)(emission_key) and kap is not None:
        raise NotImplementedError("dust emission (spectrum 'total') on a non-constant-R wavelength axis: resample grid and "
                                  "filters with generate_constant_R first")
    if getattr(emission_model, "has_dust_emission", lambda k: False)(emission_key) and kap is not None and not dust_free:
        tables.update(_dust_emission_tables(emission_model.dust_emission, lam, n_chunk * lch, uv, lo_l, hi_l, off_l,
                                            tables["igm"]["n_blue"] if tables["igm"] is not None else 0))
        if x_bins:
            tables["dust_wnu"][x_bin0:x_bin0 + x_bins] = 1.0
    lya = getattr(emission_model, "lya_line", lambda k: None)(emission_key)
    if lya is not None:
        # the line sits in the first grid of every recipe, i.e. in the kernel's component A, except when that grid is
        # all zero and the lone component is the second one -- also component A
        vals, lya_bin = lya
        tables["lya_line"] = np.ascontiguousarray(np.asarray(vals, dtype=np.float64).T / grid_scale)   # [iz][ia]
        tables["lya_bin"] = int(lya_bin)
    else:
        tables["lya_line"], tables["lya_bin"] = None, 0
    tab = cosmo.table(z_max=z_table_max)
    tables["cosmo"] = tab
    return tables


class SynthEngine:
    """Device-resident model + batched synthesis (one instance per GPU / process)."""

    def __init__(self, grid: Grid, emission_model: EmissionModel, emission_key: str,
                 filters: FilterCollection, cosmo=Planck18, igm=True, variant="nu", base_mass=1.0e9,
                 max_batch=1 << 20, device=0, fast_math=True, rest_frame=False):
        self.lib = _capi.load()
        if self.lib.sb2_device_count() < 1:
            raise RuntimeError("synference_b200: no CUDA device visible; the hot path has no CPU fallback")
        if rest_frame and igm:
            raise ValueError("rest_frame=True (luminosities through the filters) takes igm=False")
        self.tables = t = build_tables(grid, emission_model, emission_key, filters, cosmo, igm, variant)
        self.tables_lam = np.asarray(grid.lam, dtype=np.float64)
        self.filter_codes = list(filters.filter_codes)
        self.general = t["general"] is not None
        self.n_filt, self.n_lam, self.n_comp = (t["general"]["n_filt"] if self.general else t["n_filt"]), t["n_lam"], t["n_comp"]
        self.variant = variant
        self.k = t["n_age"] * t["n_z"]
        self.k_pad = t["k_pad"]
        self.max_batch = int(max_batch)
        self.device = int(device)
        self.base_mass = float(base_mass)
        self.cosmo = cosmo
        d = _capi.ModelDesc()
        keep = []

        def ptr(a, ctype):
            if a is None:
                return None
            a = np.ascontiguousarray(a)
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(ctype))

        for name in ("n_age", "n_z", "n_lam", "n_comp", "n_filt", "n_age_pad", "k_pad", "n_chunk", "interp_variant"):
            setattr(d, name, int(t[name]))
        d.log10ages, d.metallicities = ptr(t["log10ages"], C.c_double), ptr(t["metallicities"], C.c_double)
        d.gt_hi, d.gt_lo = ptr(t["gt_hi"], C.c_float), ptr(t["gt_lo"], C.c_float)
        d.grid_scale, d.lam0, d.q = t["grid_scale"], t["lam0"], t["q"]
        d.kappa = ptr(t["kappa"], C.c_float)
        d.dust_d0, d.dust_l2 = ptr(t["dust_d0"], C.c_float), ptr(t["dust_l2"], C.c_float)
        d.lya_line, d.lya_bin = ptr(t["lya_line"], C.c_double), int(t["lya_bin"])
        d.kappa_birth = ptr(t["kappa_birth"], C.c_float)
        d.dust_wnu, d.dust_g = ptr(t["dust_wnu"], C.c_float), ptr(t["dust_g"], C.c_float)
        d.dust_duv, d.dust_m_len = ptr(t["dust_duv"], C.c_float), int(t["dust_m_len"])
        d.filt_lo, d.filt_hi = ptr(t["filt_lo"], C.c_int32), ptr(t["filt_hi"], C.c_int32)
        d.filt_off, d.filt_uv = ptr(t["filt_off"], C.c_int32), ptr(t["filt_uv"], C.c_float)
        d.filt_uv_len = int(t["filt_uv"].shape[0])
        d.filt_su, d.filt_sdv = ptr(t["filt_su"], C.c_double), ptr(t["filt_sdv"], C.c_double)
        ig = t["igm"]
        if ig is not None and ig["n_blue"] > 0:
            d.n_blue, d.n_lines = int(ig["n_blue"]), int(ig["n_lines"])
            d.igm_bin_pow = ptr(ig["bin_pow"], C.c_double)
            d.igm_nline, d.igm_lc_on = ptr(ig["nline"], C.c_int32), ptr(ig["lc_on"], C.c_int32)
            d.igm_thr, d.igm_pre = ptr(ig["thr"], C.c_double), ptr(ig["pre"], C.c_double)
        cz = t["cosmo"]
        d.cosmo_n, d.cosmo_smax = int(cz.n_cells), float(cz.s_max)
        d.cosmo_dc, d.cosmo_ddc = ptr(cz.dc, C.c_double), ptr(cz.ddc, C.c_double)
        d.cosmo_age, d.cosmo_dage = ptr(cz.age, C.c_double), ptr(cz.dage, C.c_double)
        d.base_mass, d.max_batch = self.base_mass, self.max_batch
        d.rest_frame = 1 if rest_frame else 0
        d.x_bin0, d.x_bins = int(t["x_bin0"]), int(t["x_bins"])
        if fast_math:
            fm = _fm.build_tables()
            d.fm_log_tab, d.fm_exp_tab = ptr(fm["log_tab"], C.c_double), ptr(fm["exp_tab"], C.c_double)
            d.fm_tail_tab, d.fm_tail_n, d.fm_tail_w = ptr(fm["tail_tab"], C.c_double), int(fm["tail_n"]), float(fm["tail_w"])
        handle = C.c_void_p()
        _capi.check(self.lib.sb2_model_create(C.byref(d), self.device, C.byref(handle)), "sb2_model_create")
        self._h = handle
        self._fs = None
        if self.general:
            g = t["general"]
            fs = C.c_void_p()
            _capi.check(self.lib.sb2_filterset_create(int(g["n_filt"]), g["off"].ctypes.data, g["lam"].ctypes.data, g["t"].ctypes.data,
                                                      self.tables_lam.ctypes.data, int(self.n_lam), 0 if variant == "nu" else 1,
                                                      self.device, C.byref(fs)), "sb2_filterset_create")
            self._fs = fs
        del keep

    def close(self):
        if getattr(self, "_fs", None):
            self.lib.sb2_filterset_destroy(self._fs)
            self._fs = None
        if getattr(self, "_h", None):
            self.lib.sb2_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameter marshalling ------------------------------------------------------------
    def _fill(self, p: GalaxyParams, get_ptr):
        s = _capi.Params()
        s.n = len(p)
        s.redshift, s.log_mass, s.tau_v = get_ptr(p.redshift), get_ptr(p.log_mass), get_ptr(p.tau_v)
        s.sfh_type, s.sfh_stride, s.sfh_rows = int(p.sfh_type), int(p.sfh_rows.shape[1]), get_ptr(p.sfh_rows)
        s.max_age_from_z, s.norm_mask, s.age_zmax_gyr = int(p.max_age_from_z), int(p.norm_mask), float(p.age_zmax_gyr)
        s.zd_type, s.zd_value, s.zd_sigma = int(p.zd_type), get_ptr(p.zd_value), get_ptr(p.zd_sigma)
        if self.tables["single_is_unatt"]:
            s.coef_att, s.coef_unatt = get_ptr(p.coef_unatt), None
        else:
            s.coef_att, s.coef_unatt = get_ptr(p.coef_att), get_ptr(p.coef_unatt)
        s.dust_slope, s.dust_ampl = self._dust_arrays(p, get_ptr)
        if self.tables["lya_line"] is not None and p.fesc_lya is None:
            raise ValueError("the emission model reads fesc_ly_alpha per galaxy: GalaxyParams.fesc_lya is required")
        if self.tables["lya_line"] is None and p.fesc_lya is not None:
            raise ValueError("per-galaxy fesc_lya given, but the emission model has a numeric fesc_ly_alpha (or the spectrum "
                             "has no nebular part); build it with fesc_ly_alpha='fesc_lya'")
        s.fesc_lya = get_ptr(p.fesc_lya)
        if p.tau_v_birth is not None and self.tables["kappa_birth"] is None:
            raise ValueError("tau_v_birth given, but this spectrum / emission model has a single dust screen")
        if p.tau_v_birth is None and self.tables["kappa_birth"] is not None:
            raise ValueError("the emission model has two dust screens: GalaxyParams.tau_v_birth is required")
        s.tau_v_birth = get_ptr(p.tau_v_birth)
        if self.tables.get("x_bins"):
            # pseudo-bins of the absorbed-energy sum: good to < 1e-6 up to x_tau_max; a batch beyond it sums over the axis
            def top(a):
                if a is None:
                    return 0.0
                return float(a.max()) if hasattr(a, "max") else float(np.max(a))
            key = (id(p.tau_v), id(p.tau_v_birth))
            if getattr(p, "_tau_top", (None, 0.0))[0] != key:      # (a DeviceParams batch is filled on every step)
                p._tau_top = (key, top(p.tau_v) + self.tables["x_birth_ratio"] * top(p.tau_v_birth))
            tau_eff = p._tau_top[1]
            s.energy_full_axis = 1 if (not np.isfinite(tau_eff) or tau_eff > self.tables["x_tau_max"]) else 0
        return s

    def _dust_arrays(self, p: GalaxyParams, get_ptr):
        """Per-galaxy dust slope / amplitude pointers; a parameter the curve holds as a number is broadcast."""
        glob = self.tables.get("dust_global")
        if glob is None:
            if p.dust_slope is not None or p.dust_ampl is not None:
                raise ValueError("per-galaxy dust slope / amplitude given, but the model's dust curve has numeric parameters; "
                                 "build it with Calzetti2000(slope='slope', ampl='dust_bump_amplitude')")
            return None, None
        n = len(p)
        sl = p.dust_slope if p.dust_slope is not None else np.full(n, glob[0])
        am = p.dust_ampl if p.dust_ampl is not None else np.full(n, glob[1])
        return get_ptr(sl), get_ptr(am)

    @staticmethod
    def _host_ptr_factory(keep, dtype=np.float64):
        def get(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dtype)
            keep.append(a)
            return a.ctypes.data
        return get

    # ---- host-buffer entry (what a drop-in caller uses) --------------------------------------
    def photometry(self, params: GalaxyParams, scaled=True, out=None, transport="f64", library_out=None):
        """Fluxes [nJy] as a host array ``(N, n_filt)``: float64 scaled by stellar mass
        (``float32(base) * 10**log_mass / base_mass``, ``library.py:4588-4609``) or float32 at base mass.

        Populations larger than ``max_batch`` run batch by batch through the two staging slots of the C ABI, so
        the copies of one batch overlap the kernels of the next.  ``transport="f32"`` sends the parameters over PCIe as
        float32 (what ``draw_from_hypercube`` produces, ``library.py:1098``) and widens them on the device: half the
        host-to-device bytes, exact for float32 draws (use ``max_age_from_z`` so the SFH rows hold the raw draws).

        ``library_out=(matrix, first_column)`` (with ``scaled=False``): the same pass ALSO writes the mass-scaled float64
        fluxes into columns ``[first_column, first_column + N)`` of ``matrix``, a C-contiguous float64 ``(n_filt, N_total)``
        array -- the layout of a library's ``Grid/Photometry`` (``library.py:4739-4742``) -- transposed on the device
        (``sb2_params.scaled_ld``), so a library build needs no host-side cast, multiply or transpose."""
        n = len(params)
        res = out if out is not None else np.empty((n, self.n_filt), dtype=np.float64 if scaled else np.float32)
        if library_out is not None:
            mat, col0 = library_out
            if scaled or self.general or params.log_mass is None:
                raise ValueError("library_out needs scaled=False, a constant-R wavelength axis and params.log_mass")
            if not (isinstance(mat, np.ndarray) and mat.dtype == np.float64 and mat.flags.c_contiguous and mat.ndim == 2
                    and mat.shape[0] == self.n_filt and 0 <= col0 and col0 + n <= mat.shape[1]):
                raise ValueError("library_out: matrix must be C-contiguous float64 (n_filt, N_total) with room for the batch")
        if self.general:
            return self._photometry_general(params, scaled, res)
        pending = []
        for i, a in enumerate(range(0, n, self.max_batch)):
            b = min(n, a + self.max_batch)
            if len(pending) == 2:
                self.wait(pending.pop(0))
            lo = None if library_out is None else (library_out[0], library_out[1] + a)
            pending.append(self.submit(params.slice(slice(a, b)), res[a:b], scaled=scaled, slot=i & 1, transport=transport,
                                       library_out=lo))
        for t in pending:
            self.wait(t)
        return res

    def submit(self, params: GalaxyParams, out, scaled=True, slot=0, transport="f64", library_out=None):
        """Enqueue one batch (``len(params) <= max_batch``) and return a ticket for :meth:`wait`; ``out`` is the
        host array the results land in (pinned memory gives real copy/compute overlap)."""
        assert out.flags.c_contiguous and out.shape == (len(params), self.n_filt)
        assert out.dtype == (np.float64 if scaled else np.float32)
        if self.general:      # the fallback path has no staging slots: run it now
            self._photometry_general(params, scaled, out)
            return (-1, [out])
        keep = [out]
        if transport not in ("f64", "f32"):
            raise ValueError("transport must be 'f64' or 'f32'")
        s = self._fill(params, self._host_ptr_factory(keep, np.float32 if transport == "f32" else np.float64))
        s.host_f32 = 1 if transport == "f32" else 0
        scaled_ptr = out.ctypes.data if scaled else None
        if library_out is not None:      # float32 base rows into `out` and the transposed float64 scaled block in one pass
            mat, col0 = library_out
            keep.append(mat)
            scaled_ptr = mat.ctypes.data + 8 * int(col0)
            s.scaled_ld = int(mat.shape[1])
        rc = self.lib.sb2_synth_photometry_host_submit(self._h, C.byref(s), None if scaled else out.ctypes.data,
                                                       scaled_ptr, int(slot))
        _capi.check(rc, "sb2_synth_photometry_host_submit")
        return (int(slot), keep)

    def wait(self, ticket):
        if ticket[0] < 0:
            return
        _capi.check(self.lib.sb2_synth_photometry_host_wait(self._h, ticket[0]), "sb2_synth_photometry_host_wait")

    # ---- device entry (torch tensors stay on the GPU) ------------------------------------------
    def to_device(self, params: GalaxyParams):
        import torch
        dev = torch.device("cuda", self.device)
        mv = lambda a: None if a is None else torch.as_tensor(  # noqa: E731
            np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        tensors = {k: mv(getattr(params, k)) for k in ("redshift", "log_mass", "tau_v", "sfh_rows", "zd_value", "zd_sigma",
                                                       "coef_att", "coef_unatt")}
        keep = []
        sl, am = self._dust_arrays(params, lambda a: (keep.append(np.ascontiguousarray(a, dtype=np.float64)), keep[-1])[1])
        tensors["dust_slope"], tensors["dust_ampl"] = mv(sl), mv(am)
        tensors["fesc_lya"] = mv(params.fesc_lya)
        tensors["tau_v_birth"] = mv(params.tau_v_birth)
        return DeviceParams(params, tensors)

    def _set_device_ptrs(self, s, tensors):
        ptr = lambda k: None if tensors.get(k) is None else tensors[k].data_ptr()  # noqa: E731
        for k in ("redshift", "log_mass", "tau_v", "sfh_rows", "zd_value", "zd_sigma"):
            setattr(s, k, ptr(k))
        if self.tables["single_is_unatt"]:
            s.coef_att, s.coef_unatt = ptr("coef_unatt"), None
        else:
            s.coef_att, s.coef_unatt = ptr("coef_att"), ptr("coef_unatt")
        s.dust_slope, s.dust_ampl = ptr("dust_slope"), ptr("dust_ampl")
        s.fesc_lya = ptr("fesc_lya")
        s.tau_v_birth = ptr("tau_v_birth")

    def photometry_device(self, dparams: "DeviceParams", flux_base=None, flux_scaled=None, spectra=None):
        """Run one batch whose parameters are already in HBM; outputs are caller-provided torch tensors."""
        import torch
        p = dparams.host
        t = dparams.tensors
        if self.general:
            return self._device_general(dparams, flux_base, flux_scaled, spectra)
        s = self._fill(p, lambda a: None)
        self._set_device_ptrs(s, t)
        st = torch.cuda.current_stream(self.device).cuda_stream
        dp = lambda x: None if x is None else x.data_ptr()  # noqa: E731
        rc = self.lib.sb2_synth_photometry(self._h, C.byref(s), dp(flux_base), dp(flux_scaled), dp(spectra), st)
        _capi.check(rc, "sb2_synth_photometry")

    # ---- general (non-constant-R) wavelength axis: spectra to HBM, then general_filter_kernel -----------------------
    def _general_rows(self):
        return int(max(256, min(32768, self.max_batch, (1 << 30) // (4 * self.n_lam))))

    def _device_general(self, dparams, flux_base=None, flux_scaled=None, spectra=None):
        import torch
        p, t = dparams.host, dparams.tensors
        n = len(p)
        dev = t["redshift"].device
        rows = n if spectra is not None else self._general_rows()
        buf = spectra if spectra is not None else torch.empty((min(rows, n), self.n_lam), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(self.device).cuda_stream
        dp = lambda x: None if x is None else x.data_ptr()  # noqa: E731
        for a in range(0, n, rows):
            b = min(n, a + rows)
            sub = DeviceParams(p.slice(slice(a, b)), {k: (None if v is None else v[a:b]) for k, v in t.items()})
            s = self._fill(sub.host, lambda x: None)
            self._set_device_ptrs(s, sub.tensors)
            sp = buf[a:b] if spectra is not None else buf[:b - a]
            _capi.check(self.lib.sb2_synth_photometry(self._h, C.byref(s), None, None, sp.data_ptr(), st), "sb2_synth_photometry")
            if flux_base is not None or flux_scaled is not None:
                _capi.check(self.lib.sb2_filter_integrate(
                    self._fs, sp.data_ptr(), sub.tensors["redshift"].data_ptr(), dp(sub.tensors.get("log_mass")), self.base_mass,
                    b - a, dp(None if flux_base is None else flux_base[a:b]), dp(None if flux_scaled is None else flux_scaled[a:b]), st),
                    "sb2_filter_integrate")

    def _photometry_general(self, params, scaled, res):
        import torch
        n = len(params)
        dev = torch.device("cuda", self.device)
        for a in range(0, n, self.max_batch):
            b = min(n, a + self.max_batch)
            dpar = self.to_device(params.slice(slice(a, b)))
            out = torch.empty((b - a, self.n_filt), dtype=torch.float64 if scaled else torch.float32, device=dev)
            self._device_general(dpar, flux_base=None if scaled else out, flux_scaled=out if scaled else None)
            res[a:b] = out.cpu().numpy()
        return res

    def spectra(self, params: GalaxyParams, out=None, photometry_out=None):
        """Observed-frame f_nu [nJy] at base mass on the rest-frame axis, ``(N, n_lam)`` float32 (host), through the host entry
        of the C ABI (``sb2_synth_photometry_host`` with ``spec_out``): the batch is walked in slices whose device-to-host
        copies overlap the kernels of the next slice.  ``out`` may be a pinned array (real overlap; e.g. a
        ``torch.empty(...).pin_memory().numpy()`` view) -- this is cfg 5's write path (``library.py:4887-4919``);
        ``photometry_out`` ``(N, n_filt)`` float32 receives the base-mass photometry of the same pass."""
        n = len(params)
        res = out if out is not None else np.empty((n, self.n_lam), dtype=np.float32)
        assert res.flags.c_contiguous and res.shape == (n, self.n_lam) and res.dtype == np.float32
        if self.general and photometry_out is not None:
            photometry_out[...] = self.photometry(params, scaled=False)
            photometry_out = None
        if photometry_out is not None:
            assert photometry_out.flags.c_contiguous and photometry_out.shape == (n, self.n_filt) and photometry_out.dtype == np.float32
        for a in range(0, n, self.max_batch):
            b = min(n, a + self.max_batch)
            keep = []
            s = self._fill(params.slice(slice(a, b)), self._host_ptr_factory(keep))
            fp = None if photometry_out is None else photometry_out[a:b].ctypes.data
            rc = self.lib.sb2_synth_photometry_host(self._h, C.byref(s), fp, None, res[a:b].ctypes.data)
            _capi.check(rc, "sb2_synth_photometry_host")
        return res

    def weights(self, params: GalaxyParams):
        """SFZH weights ``(N, n_age*n_z)`` float64, k = iz*n_age + ia (parity hook)."""
        import torch
        n = len(params)
        dev = torch.device("cuda", self.device)
        out = np.empty((n, self.k), dtype=np.float64)
        for a in range(0, n, self.max_batch):
            b = min(n, a + self.max_batch)
            dpar = self.to_device(params.slice(slice(a, b)))
            s = self._fill(dpar.host, lambda x: None)
            self._set_device_ptrs(s, dpar.tensors)
            w = torch.empty((b - a, self.k), dtype=torch.float64, device=dev)
            st = torch.cuda.current_stream(self.device).cuda_stream
            _capi.check(self.lib.sb2_build_weights(self._h, C.byref(s), w.data_ptr(), st), "sb2_build_weights")
            out[a:b] = w.cpu().numpy()
        return out


    def rest_band_flux(self, params: GalaxyParams, lam_lo: float, lam_hi: float):
        """Observed-frame f_nu [nJy] at base mass averaged over a REST-frame top-hat ``[lam_lo, lam_hi]`` (Angstrom), ``(N,)``
        float64: the photometry convention of A9 (trapezoid of ``f T / nu`` over the in-band samples divided by that of
        ``T / nu``) applied to the synthesised spectrum, reduced on the device -- only N numbers come back.  This is what the
        reference's ``calculate_muv`` measures with its 1500 +- 50 A top-hat (``library.py:100-104, 172-196``)."""
        import torch
        lam = np.asarray(self.tables_lam, dtype=np.float64)
        inb = np.nonzero((lam >= lam_lo) & (lam <= lam_hi))[0]
        if inb.size < 2:
            raise ValueError(f"fewer than two wavelength samples in [{lam_lo}, {lam_hi}] A")
        nu = 2.99792458e18 / lam[inb]
        w = np.zeros(inb.size)
        d = np.abs(np.diff(nu))
        w[:-1] += 0.5 * d
        w[1:] += 0.5 * d
        w = w / nu                                   # trapezoid weights of  integral( . T / nu d nu )
        dev = torch.device("cuda", self.device)
        wt = torch.as_tensor(w / w.sum(), dtype=torch.float64, device=dev)
        n = len(params)
        out = np.empty(n, dtype=np.float64)
        i0, i1 = int(inb[0]), int(inb[-1]) + 1
        for a in range(0, n, self.max_batch):
            b = min(n, a + self.max_batch)
            dpar = self.to_device(params.slice(slice(a, b)))
            spec = torch.empty((b - a, self.n_lam), dtype=torch.float32, device=dev)
            flux = torch.empty((b - a, self.n_filt), dtype=torch.float32, device=dev)
            self.photometry_device(dpar, flux_base=flux, spectra=spec)
            out[a:b] = (spec[:, i0:i1].to(torch.float64) @ wt).cpu().numpy()
        return out

    def sfzh(self, params: GalaxyParams):
        """SFZH ``(N, n_age, n_z)`` float64, normalised to 1 per galaxy (host array; the by-products of
        ``synference_b200.supplementary`` are evaluated from it)."""
        t = self.tables
        return self.weights(params).reshape(len(params), t["n_z"], t["n_age"]).transpose(0, 2, 1)


@dataclass
class DeviceParams:
    host: GalaxyParams
    tensors: dict


def depth_noise_features(flux, sigma, n_scatter=1, normals=None, seed=0, epoch=0, norm_mag_limit=50.0,
                         min_flux_pc_error=0.0, want_flux=True, want_features=True, device=0, set_index=None):
    """Depth scatter + AB feature rows on the GPU (``sb2_depth_noise_features``).

    flux ``(n_gal, n_filt)`` nJy (numpy or CUDA torch tensor), sigma ``(n_filt,)`` nJy -- or ``(n_sets, n_filt)`` together
    with ``set_index (n_filt, n_scatter)`` int: which depth set each block of ``n_gal`` rows uses (``sbi_runner.py:626-647``).
    Returns ``(noisy_flux (n_filt, n_rows) f64 | None, sigma (n_filt, n_rows) f64 | None,
    features (n_rows, 2 n_filt) f32 | None)`` as torch CUDA tensors.
    """
    import torch
    lib = _capi.load()
    dev = torch.device("cuda", device)
    # float32 fluxes (a library's photometry as stored) stay float32 across PCIe and in HBM when only Philox feature rows are
    # wanted: sb2_depth_noise_features_f32 widens them on the device, as numpy does for float32 flux + float64 noise
    is32 = (isinstance(flux, torch.Tensor) and flux.dtype == torch.float32) or (isinstance(flux, np.ndarray) and flux.dtype == np.float32)
    if is32 and normals is None and set_index is None and want_features and not want_flux:
        fl = torch.as_tensor(flux).to(dev).contiguous()
        n_gal, n_filt = fl.shape
        sg = torch.as_tensor(np.asarray(sigma, dtype=np.float64)).to(dev).contiguous()
        feat = torch.empty((n_gal * int(n_scatter), 2 * n_filt), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.sb2_depth_noise_features_f32(fl.data_ptr(), n_gal, n_filt, int(n_scatter), sg.data_ptr(), float(min_flux_pc_error),
                                                  int(seed), int(epoch), float(norm_mag_limit), feat.data_ptr(),
                                                  torch.cuda.current_stream(device).cuda_stream)
        _capi.check(rc, "sb2_depth_noise_features_f32")
        return None, None, feat
    fl = torch.as_tensor(flux, dtype=torch.float64).to(dev).contiguous()
    n_gal, n_filt = fl.shape
    sg = torch.as_tensor(np.asarray(sigma, dtype=np.float64)).to(dev).contiguous()
    n_rows = n_gal * int(n_scatter)
    nz = None
    if normals is not None:
        nz = torch.as_tensor(normals, dtype=torch.float64).to(dev).contiguous()
        assert tuple(nz.shape) == (n_filt, n_rows), "normals must be (n_filt, n_gal*n_scatter)"
    of = torch.empty((n_filt, n_rows), dtype=torch.float64, device=dev) if want_flux else None
    osig = torch.empty((n_filt, n_rows), dtype=torch.float64, device=dev) if want_flux else None
    feat = torch.empty((n_rows, 2 * n_filt), dtype=torch.float32, device=dev) if want_features else None
    with torch.cuda.device(dev):      # the entry points launch on the current device
        dp = lambda x: None if x is None else x.data_ptr()  # noqa: E731
        st = torch.cuda.current_stream(device).cuda_stream
        if set_index is not None:
            if sg.ndim != 2 or sg.shape[1] != n_filt:
                raise ValueError("with set_index, sigma must be (n_sets, n_filt)")
            si = torch.as_tensor(np.asarray(set_index, dtype=np.int32)).to(dev).contiguous()
            if tuple(si.shape) != (n_filt, int(n_scatter)) or int(si.min()) < 0 or int(si.max()) >= sg.shape[0]:
                raise ValueError("set_index must be (n_filt, n_scatter) with entries in [0, n_sets)")
            rc = lib.sb2_depth_noise_features_sets(fl.data_ptr(), n_gal, n_filt, int(n_scatter), sg.data_ptr(), int(sg.shape[0]),
                                                   si.data_ptr(), float(min_flux_pc_error), dp(nz), int(seed), int(epoch),
                                                   float(norm_mag_limit), dp(of), dp(osig), dp(feat), st)
            _capi.check(rc, "sb2_depth_noise_features_sets")
            return of, osig, feat
        rc = lib.sb2_depth_noise_features(fl.data_ptr(), n_gal, n_filt, int(n_scatter), sg.data_ptr(),
                                          float(min_flux_pc_error), dp(nz), int(seed), int(epoch),
                                          float(norm_mag_limit), dp(of), dp(osig), dp(feat), st)
        _capi.check(rc, "sb2_depth_noise_features")
        return of, osig, feat
