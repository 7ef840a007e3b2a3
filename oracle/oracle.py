"""CPU ORACLE for the synference mock-library hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a float64 numpy/scipy restatement of the reference algorithm for the
path  parameters -> SFZH weights -> grid-weighted spectrum -> emission tree / dust ->
observed frame + IGM -> filter integration -> mass scaling -> magnitudes / noise.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference
arm may import it; the product package (``synference_b200``) never does, and the oracle
imports nothing from the product.  All inputs are plain arrays.

PARITY STATUS
  * synference-side arithmetic (unit conversions, depth noise, feature magnitudes, mass
    scaling, constant-R grid, asinh magnitudes) is restated from the files under
    /root/reference and PINNED against the only known-answer checks the reference holds
    (``tests/test_uncertainty_models.py:47-74``; see ``tests/test_oracle_golden.py``).
  * the arithmetic behind the third-party boundary (``synthesizer``: SFZH weights, weighted
    sum, emission tree, Calzetti/N09, Inoue14, filter integration, Planck18) is NOT under
    /root/reference, not installed and unpinned upstream (``pyproject.toml:53``), and the
    reference's own tests assert only finiteness / shapes on it.  For that part this oracle
    follows the frozen pin list in SURVEY.md Appendix A:  **parity unpinned**.

Every function cites the reference lines (relative to /root/reference) or the SURVEY
appendix item it follows.
"""

from __future__ import annotations

import math

import numpy as np
from scipy import integrate, special

# --------------------------------------------------------------------------------------
# A7: Planck18 (astropy.cosmology.Planck18; used at library.py:1206 and inside Synthesizer)
# --------------------------------------------------------------------------------------
_H0, _OM0, _TCMB, _NEFF, _MNU = 67.66, 0.30966, 2.7255, 3.046, 0.06
_C_KMS, _G, _SB, _KB = 299792.458, 6.6743e-11, 5.670374419e-8, 8.617333262e-5
_MPC_M, _GYR_S = 3.0856775814913673e22, 3.15576e16


def _omegas():
    h0 = _H0 * 1e3 / _MPC_M
    rho_c = 3 * h0 * h0 / (8 * math.pi * _G)
    ogam = 4 * _SB / (_C_KMS * 1e3) ** 3 * _TCMB**4 / rho_c
    return h0, ogam


def _nu_rel(z):
    # astropy FLRW.nu_relative_density (Komatsu+11 eq. 26 fitting form), one massive species
    nu_y = _MNU / (_KB * 0.7137658555036082 * _TCMB)
    rel = (1.0 + (0.3173 * nu_y / (1.0 + z)) ** 1.83) ** 0.54644808743 + 2.0
    return 0.22710731766 * (_NEFF / 3.0) * rel


def efunc(z):
    _, ogam = _omegas()
    ode = 1.0 - _OM0 - ogam * (1.0 + _nu_rel(0.0))
    zp1 = 1.0 + z
    return math.sqrt(zp1**3 * (ogam * (1.0 + _nu_rel(z)) * zp1 + _OM0) + ode)


def age_gyr(z):
    """Cosmic age at z [Gyr]:  t_H * int_z^inf dz / ((1+z) E(z))."""
    h0, _ = _omegas()
    f = lambda s: 1.0 / efunc(math.expm1(s))  # noqa: E731   (s = ln(1+z))
    s0 = math.log1p(float(z))
    val = integrate.quad(f, s0, s0 + 60.0, epsabs=0, epsrel=1e-12, limit=400)[0]
    return val / h0 / _GYR_S


def luminosity_distance_cm(z):
    f = lambda zz: 1.0 / efunc(zz)  # noqa: E731
    dc = integrate.quad(f, 0.0, float(z), epsabs=0, epsrel=1e-12, limit=400)[0] * _C_KMS / _H0
    return (1.0 + z) * dc * _MPC_M * 100.0


def max_age_myr(z, max_redshift=20.0):
    """library.py:1206  max_ages = (cosmo.age(z) - cosmo.age(max_redshift)).to(Myr)."""
    a_max = age_gyr(max_redshift)
    return np.array([(age_gyr(zz) - a_max) * 1e3 for zz in np.atleast_1d(z)])


# --------------------------------------------------------------------------------------
# utils.py:257-289  constant-R wavelength grid
# --------------------------------------------------------------------------------------
def constant_r_grid(start, end, R=300):
    x = [float(start)]
    while x[-1] < end:
        x.append(x[-1] * (1.0 + 0.5 / R))
    return np.array(x)


# --------------------------------------------------------------------------------------
# A2: SFH -> age-bin masses (Stars.__init__ -> _get_sfzh; call site library.py:1372-1379)
# --------------------------------------------------------------------------------------
def age_bin_edges(log10ages):
    """e_0 = 0, e_{i+1} = (t_i + t_{i+1})/2 ; the last age bin receives no mass."""
    t = 10.0 ** np.asarray(log10ages, dtype=float)
    return np.concatenate([[0.0], 0.5 * (t[1:] + t[:-1])])


def _phi_diff(u_lo, u_hi):
    """Phi(u_hi) - Phi(u_lo) without cancellation in either tail."""
    s2 = math.sqrt(2.0)
    up = 0.5 * (special.erfc(u_lo / s2) - special.erfc(u_hi / s2))   # good when both > 0
    dn = 0.5 * (special.erfc(-u_hi / s2) - special.erfc(-u_lo / s2))  # good when both < 0
    return np.where(u_lo + u_hi > 0, up, dn)


def sfr_pointwise(kind, p, age):
    """The SFH functional forms of SURVEY A2 as functions of lookback age [yr]."""
    age = np.asarray(age, dtype=float)
    mn, mx = p["min_age"], p["max_age"]
    with np.errstate(all="ignore"):
        if kind == "Constant":
            return ((age >= mn) & (age <= mx)).astype(float)
        inside = (age >= mn) & (age < mx)
        a = np.where(inside, age, mn)
        t = mx - a
        if kind == "Gaussian":
            v = np.exp(-0.5 * ((a - p["peak_age"]) / p["sigma"]) ** 2)
        elif kind == "Exponential":
            v = np.exp(t / p["tau"])
        elif kind == "DecliningExponential":
            v = np.exp(-t / p["tau"])
        elif kind == "DelayedExponential":
            v = t * np.exp(-t / p["tau"])
        elif kind == "LogNormal":
            t0 = math.log(mx - p["peak_age"]) + p["tau"] ** 2
            v = (1.0 / t) * np.exp(-((np.log(t) - t0) ** 2) / 2.0 / p["tau"] ** 2)
        elif kind == "DoublePowerLaw":
            x = a / p["peak_age"]
            v = 1.0 / (x ** p["alpha"] + x ** p["beta"])
        else:
            raise ValueError(kind)
        return np.where(inside, v, 0.0)


def sfh_bin_masses_quad(kind, p, log10ages):
    """Literal form: scipy.integrate.quad of the SFR over each age bin (default tolerances)."""
    e = age_bin_edges(log10ages)
    sf = np.zeros(len(log10ages))
    for i in range(len(e) - 1):
        pts = [x for x in (p["min_age"], p["max_age"]) if e[i] < x < e[i + 1]]
        sf[i] = integrate.quad(lambda a: float(sfr_pointwise(kind, p, a)), e[i], e[i + 1],
                               points=pts or None, limit=200)[0]
    return sf


def sfh_bin_masses(kind, p, log10ages):
    """Closed-form bin integrals (what quad approximates); returns un-normalised masses."""
    e = age_bin_edges(log10ages)
    mn, mx = float(p["min_age"]), float(p["max_age"])
    lo = np.clip(e[:-1], mn, mx)
    hi = np.clip(e[1:], mn, mx)
    with np.errstate(all="ignore"):
        if kind == "Constant":
            m = hi - lo
        elif kind == "Gaussian":
            s = p["sigma"]
            m = s * math.sqrt(2 * math.pi) * _phi_diff((lo - p["peak_age"]) / s, (hi - p["peak_age"]) / s)
        elif kind in ("Exponential", "DecliningExponential"):
            tau = p["tau"] if kind == "Exponential" else -p["tau"]
            shift = (mx - mn) / tau if tau > 0 else 0.0  # common factor, cancels on normalisation
            m = tau * (np.exp((mx - lo) / tau - shift) - np.exp((mx - hi) / tau - shift))
        elif kind == "DelayedExponential":
            tau = p["tau"]
            anti = lambda t: -tau * (t + tau) * np.exp(-t / tau)  # noqa: E731
            m = anti(mx - lo) - anti(mx - hi)
        elif kind == "LogNormal":
            tau = p["tau"]
            t0 = math.log(mx - p["peak_age"]) + tau * tau
            u = lambda a: (np.log(np.maximum(mx - a, 1e-300)) - t0) / tau  # noqa: E731
            m = tau * math.sqrt(2 * math.pi) * _phi_diff(u(hi), u(lo))
        elif kind == "Continuity":
            edges = np.asarray(p["edges"], dtype=float)
            sfr = 10.0 ** (-np.concatenate([[0.0], np.cumsum(p["logsfr_ratios"])]))
            m = np.zeros(len(e) - 1)
            for j in range(len(edges) - 1):
                m += sfr[j] * np.clip(np.minimum(e[1:], edges[j + 1]) - np.maximum(e[:-1], edges[j]), 0, None)
        else:
            return sfh_bin_masses_quad(kind, p, log10ages)
    m = np.where(hi > lo, m, 0.0) if kind != "Continuity" else m
    return np.concatenate([m, [0.0]])


# --------------------------------------------------------------------------------------
# A3: ZDist -> metallicity weights
# --------------------------------------------------------------------------------------
def zdist_weights(kind, value, sigma, metallicities):
    zg = np.asarray(metallicities, dtype=float)
    if kind in ("delta_linear", "delta_log10"):
        x = zg if kind == "delta_linear" else np.log10(zg)
        w = np.zeros(len(zg))
        if value <= x[0]:
            w[0] = 1.0
        elif value >= x[-1]:
            w[-1] = 1.0
        else:
            j = int(np.searchsorted(x, value, side="right") - 1)
            f = (value - x[j]) / (x[j + 1] - x[j])
            w[j], w[j + 1] = 1.0 - f, f
        return w
    x = zg if kind == "normal_linear" else np.log10(zg)
    w = np.exp(-0.5 * ((x - value) / sigma) ** 2)
    return w / w.sum()


def sfzh_weights(sf, zd):
    """w[a, Z] = sf[a] zd[Z], normalised to unit mass (initial_mass is applied as a scale)."""
    w = np.outer(sf, zd)
    return w / w.sum()


# --------------------------------------------------------------------------------------
# A5: emission-model tree;  A6: dust
# --------------------------------------------------------------------------------------
def emission_parts(components, lam, key, fesc=0.0, fesc_ly_alpha=1.0):
    """(attenuated-part grid, unattenuated-part grid) of spectrum ``key``."""
    inc = components["incident"]
    zero = np.zeros_like(inc)
    if key == "incident":
        return zero, inc
    line = np.array(components.get("linecont", zero), dtype=float, copy=True)
    line[..., int(np.argmin(np.abs(np.asarray(lam) - 1215.67)))] *= fesc_ly_alpha
    trans = (1.0 - fesc) * components.get("transmitted", inc)
    neb = (1.0 - fesc) * (line + components.get("nebular_continuum", zero))
    repro, esc = trans + neb, fesc * inc
    return {"transmitted": (zero, trans), "nebular": (zero, neb), "reprocessed": (zero, repro),
            "escaped": (zero, esc), "intrinsic": (zero, repro + esc), "attenuated": (repro, zero),
            "emergent": (repro, esc), "total": (repro, esc)}[key]


def calzetti_k(x):
    x = np.asarray(x, dtype=float)
    return 4.05 + 2.659 * np.where(x < 0.63, -2.156 + 1.509 / x - 0.198 / x**2 + 0.011 / x**3,
                                   -1.857 + 1.040 / x)


def dust_kappa(lam_angstrom, curve="Calzetti2000", slope=0.0, cent_lam=0.2175, ampl=0.0, gamma=0.035):
    """tau(lambda)/tau_V.  Calzetti2000 in the Noll+09 form on a 0.12-2.2 um helper grid, linear
    interpolation with linear extrapolation, then the slope power law."""
    lam_um = np.asarray(lam_angstrom, dtype=float) * 1e-4
    if curve == "PowerLaw":
        return (lam_um / 0.55) ** slope
    x = np.arange(0.12, 2.2, 0.001)
    helper = (calzetti_k(x) + ampl * (x * gamma) ** 2 / ((x**2 - cent_lam**2) ** 2 + (x * gamma) ** 2)) \
        / calzetti_k(0.55)
    f = integrate  # noqa: F841
    from scipy.interpolate import interp1d
    y = interp1d(x, helper, kind="linear", fill_value="extrapolate", bounds_error=False)(lam_um)
    return y * (lam_um / 0.55) ** slope


# --------------------------------------------------------------------------------------
# A8: Inoue+14 IGM (structure of the upstream implementation; coefficient arrays are inputs)
# --------------------------------------------------------------------------------------
def inoue14_transmission(z, lam_obs, laf, dla):
    lobs = np.asarray(lam_obs, dtype=float)
    z1l, z2l, z1d, lam_l = 1.2, 4.7, 2.0, 911.8
    tau = np.zeros_like(lobs)
    for row in laf:  # Lyman-series, Lyman-alpha forest component
        lj, a1, a2, a3 = row[1], row[2], row[3], row[4]
        on = lobs < lj * (1 + z)
        r1 = on & (lobs < lj * (1 + z1l))
        r2 = on & (lobs >= lj * (1 + z1l)) & (lobs < lj * (1 + z2l))
        r3 = on & (lobs >= lj * (1 + z2l))
        tau[r1] += a1 * (lobs[r1] / lj) ** 1.2
        tau[r2] += a2 * (lobs[r2] / lj) ** 3.7
        tau[r3] += a3 * (lobs[r3] / lj) ** 5.5
    for row in dla:  # Lyman-series, DLA component
        lj, a1, a2 = row[1], row[2], row[3]
        on = lobs < lj * (1 + z)
        r1 = on & (lobs < lj * (1 + z1d))
        r2 = on & (lobs >= lj * (1 + z1d))
        tau[r1] += a1 * (lobs[r1] / lj) ** 2.0
        tau[r2] += a2 * (lobs[r2] / lj) ** 3.0
    x0 = lobs < lam_l * (1 + z)
    xl = lobs / lam_l
    zp = 1.0 + z
    lc = np.zeros_like(lobs)
    # Lyman continuum, DLA
    if z < z1d:
        lc[x0] += 0.2113 * zp**2 - 0.07661 * zp**2.3 * xl[x0] ** -0.3 - 0.1347 * xl[x0] ** 2
    else:
        x1 = lobs >= lam_l * (1 + z1d)
        m = x0 & x1
        lc[m] += 0.04696 * zp**3 - 0.01779 * zp**3.3 * xl[m] ** -0.3 - 0.02916 * xl[m] ** 3
        m = x0 & ~x1
        lc[m] += (0.6340 + 0.04696 * zp**3 - 0.01779 * zp**3.3 * xl[m] ** -0.3
                  - 0.1347 * xl[m] ** 2 - 0.2905 * xl[m] ** -0.3)
    # Lyman continuum, LAF
    if z < z1l:
        lc[x0] += 0.3248 * (xl[x0] ** 1.2 - zp**-0.9 * xl[x0] ** 2.1)
    elif z < z2l:
        x1 = lobs >= lam_l * (1 + z1l)
        m = x0 & x1
        lc[m] += 2.545e-2 * (zp**1.6 * xl[m] ** 2.1 - xl[m] ** 3.7)
        m = x0 & ~x1
        lc[m] += 2.545e-2 * zp**1.6 * xl[m] ** 2.1 + 0.3248 * xl[m] ** 1.2 - 0.2496 * xl[m] ** 2.1
    else:
        x1 = lobs > lam_l * (1 + z2l)
        x2 = (lobs >= lam_l * (1 + z1l)) & (lobs < lam_l * (1 + z2l))
        x3 = lobs < lam_l * (1 + z1l)
        m = x0 & x1
        lc[m] += 5.221e-4 * (zp**3.4 * xl[m] ** 2.1 - xl[m] ** 5.5)
        m = x0 & x2
        lc[m] += 5.221e-4 * zp**3.4 * xl[m] ** 2.1 + 0.2182 * xl[m] ** 2.1 - 2.545e-2 * xl[m] ** 3.7
        m = x0 & x3
        lc[m] += 5.221e-4 * zp**3.4 * xl[m] ** 2.1 + 0.3248 * xl[m] ** 1.2 - 3.140e-2 * xl[m] ** 2.1
    return np.exp(-(tau + lc))


# --------------------------------------------------------------------------------------
# A9: filter integration (Sed.get_photo_fnu -> Filter.apply_filter)
# --------------------------------------------------------------------------------------
_C_ANG = 2.99792458e18  # Angstrom / s


def apply_filter(fnu, lam_obs, filt_lam, filt_t, variant="nu"):
    """trapz(f T / x, x) / trapz(T / x, x) over samples with T > 0; x = nu_obs (or lam_obs).

    The filter's own table is re-interpolated linearly *in the integration variable* onto
    the spectrum's observed abscissa with 0 outside its range; ValueError when no sample is in band.
    """
    if variant == "nu":
        x = _C_ANG / lam_obs
        xp = _C_ANG / np.asarray(filt_lam)[::-1]
        t = np.interp(x, xp, np.asarray(filt_t)[::-1], left=0.0, right=0.0)
    else:
        x = lam_obs
        t = np.interp(x, filt_lam, filt_t, left=0.0, right=0.0)
    keep = t > 0
    if not keep.any():
        raise ValueError("filter lies entirely outside the spectrum")
    xx, tt, ff = x[keep], t[keep], fnu[keep]
    return np.trapezoid(ff * tt / xx, xx) / np.trapezoid(tt / xx, xx)


# --------------------------------------------------------------------------------------
# The whole path for a batch of galaxies
# --------------------------------------------------------------------------------------
def weights_for(gal, log10ages, metallicities):
    sf = sfh_bin_masses(gal["sfh_kind"], gal["sfh"], log10ages)
    zd = zdist_weights(gal["zd_kind"], gal["zd_value"], gal.get("zd_sigma", 0.0), metallicities)
    return sfzh_weights(sf, zd)


def dust_emission_shape(lam, kind="Greybody", temperature=40.0, emissivity=1.5):
    """A5 pin: the generator's L_nu, nu^beta B_nu(T) (beta = 0 for a Blackbody), normalised to unit integral over all
    frequencies -- here by numerical quadrature in x = h nu / k T (the product uses the closed form Gamma * zeta)."""
    from scipy.integrate import quad
    beta = 0.0 if kind == "Blackbody" else float(emissivity)
    h_over_k = 6.62607015e-34 / 1.380649e-23
    nu = 2.99792458e18 / np.asarray(lam, dtype=float)
    x = h_over_k * nu / float(temperature)
    with np.errstate(over="ignore", under="ignore"):
        f = x ** (3.0 + beta) / np.expm1(x)
    norm = quad(lambda t: t ** (3.0 + beta) / math.expm1(t) if t < 700 else 0.0, 0.0, 700.0, epsabs=0, epsrel=1e-12, limit=400)[0]
    return f / norm * (h_over_k / float(temperature))


def synthesize(galaxies, log10ages, metallicities, lam, components, filters, *, key="intrinsic",
               fesc=0.0, fesc_ly_alpha=1.0, dust=None, igm=None, variant="nu", base_mass=1e9,
               return_spectra=False, dl_cm=None, two_screens=None, dust_emission=None):
    """Fluxes [nJy] of every galaxy through every filter at ``base_mass`` Msun.

    galaxies : list of dicts with redshift, tau_v, sfh_kind, sfh (dict, ages in yr),
               zd_kind, zd_value, zd_sigma and optionally fesc (per-galaxy escape fraction: the emission tree
               of SURVEY A5 is then evaluated with that galaxy's value instead of the global ``fesc``; likewise
               fesc_ly_alpha) and
               dust_slope / dust_ampl (per-galaxy shape of the attenuation curve)
    filters  : list of (lam_table [A], transmission) on each filter's own axis
    dust     : None or dict(curve=..., slope=..., ampl=...)  ;  igm : None or (laf, dla)
    two_screens : None or dict(age_pivot=log10 yr, dust_birth=dict(...)): stars with log10age < age_pivot are attenuated by
               exp(-tau_v_birth kappa_birth - tau_v kappa), the others by exp(-tau_v kappa) (galaxy key tau_v_birth)
    dust_emission : None or dict(kind=, temperature=, emissivity=): adds E_abs * shape(nu), E_abs = trapezoid over nu of the
               light the screen(s) removed (energy balance, A5 'total')
    """
    lam = np.asarray(lam, dtype=float)
    g_att, g_un = emission_parts(components, lam, key, fesc, fesc_ly_alpha)
    na, nz = len(log10ages), len(metallicities)
    g_att2, g_un2 = g_att.reshape(na * nz, -1), g_un.reshape(na * nz, -1)
    kappa = dust_kappa(lam, **dust) if dust is not None else None
    if two_screens is not None:
        kappa_birth = dust_kappa(lam, **two_screens["dust_birth"])
        young = (np.repeat(np.asarray(log10ages)[:, None], nz, 1) < two_screens["age_pivot"]).reshape(-1)
    dust_shape = dust_emission_shape(lam, **dust_emission) if dust_emission is not None else None
    nu_rest = 2.99792458e18 / lam
    out = np.zeros((len(galaxies), len(filters)))
    spectra = np.zeros((len(galaxies), len(lam))) if return_spectra else None
    for g, gal in enumerate(galaxies):
        w = weights_for(gal, log10ages, metallicities).reshape(-1)
        if "fesc" in gal or "fesc_ly_alpha" in gal:
            ga, gu = emission_parts(components, lam, key, float(gal.get("fesc", fesc)),
                                    float(gal.get("fesc_ly_alpha", fesc_ly_alpha)))
            g_att2, g_un2 = ga.reshape(na * nz, -1), gu.reshape(na * nz, -1)
        lnu = w @ g_un2  # A4: grid-weighted sum, erg/s/Hz per Msun
        att = w @ g_att2
        before_dust = att
        if kappa is not None:
            kap = kappa
            if "dust_slope" in gal or "dust_ampl" in gal:
                dd = dict(dust)
                dd["slope"] = float(gal.get("dust_slope", dd.get("slope", 0.0)))
                dd["ampl"] = float(gal.get("dust_ampl", dd.get("ampl", 0.0)))
                kap = dust_kappa(lam, **dd)
            if two_screens is not None:
                att = (w * young) @ g_att2 * np.exp(-gal.get("tau_v_birth", 0.0) * kappa_birth - gal.get("tau_v", 0.0) * kap) \
                    + (w * ~young) @ g_att2 * np.exp(-gal.get("tau_v", 0.0) * kap)
            else:
                att = att * np.exp(-gal.get("tau_v", 0.0) * kap)
        if dust_shape is not None and kappa is not None:
            e_abs = -np.trapezoid(before_dust - att, nu_rest)          # nu decreases along the axis
            att = att + e_abs * dust_shape
        lnu = (lnu + att) * base_mass
        z = float(gal["redshift"])
        dl = luminosity_distance_cm(z) if dl_cm is None else dl_cm[g]
        with np.errstate(all="ignore"):
            fnu = lnu * (1.0 + z) / (4.0 * math.pi * dl * dl) * 1e23 * 1e9  # A7, nJy
        lam_obs = lam * (1.0 + z)
        if igm is not None:
            fnu = fnu * inoue14_transmission(z, lam_obs, igm[0], igm[1])
        if return_spectra:
            spectra[g] = fnu
        for f, (fl, ft) in enumerate(filters):
            out[g, f] = apply_filter(fnu, lam_obs, fl, ft, variant)
    return (out, spectra) if return_spectra else out


# --------------------------------------------------------------------------------------
# Spectroscopic path: utils.py:129-182 variable-width Gaussian, utils.py:185-254 transform_spectrum.
# The flux-conserving rebin is the third-party `spectres` package (Carnall 2017, arXiv:1705.05165; not installed
# here, unpinned): bin edges at the midpoints of the wavelengths, end bins symmetric about their centres; a new bin is
# the old bins' fluxes weighted by the overlapping widths; new bins not fully covered get `fill`.
# --------------------------------------------------------------------------------------
def spectres_bins(wavs):
    wavs = np.asarray(wavs, dtype=float)
    edges = np.empty(wavs.size + 1)
    edges[0] = wavs[0] - (wavs[1] - wavs[0]) / 2
    edges[-1] = wavs[-1] + (wavs[-1] - wavs[-2]) / 2
    edges[1:-1] = (wavs[1:] + wavs[:-1]) / 2
    return edges, np.diff(edges)


def spectres_resample(new_wavs, spec_wavs, spec_fluxes, fill=0.0):
    old_edges, old_widths = spectres_bins(spec_wavs)
    new_edges, _ = spectres_bins(new_wavs)
    flux = np.asarray(spec_fluxes, dtype=float)
    out = np.full(len(new_wavs), float(fill))
    start = stop = 0
    for j in range(len(new_wavs)):
        if new_edges[j] < old_edges[0] or new_edges[j + 1] > old_edges[-1]:
            continue
        while old_edges[start + 1] <= new_edges[j]:
            start += 1
        while old_edges[stop + 1] < new_edges[j + 1]:
            stop += 1
        if stop == start:
            out[j] = flux[start]
            continue
        w = old_widths[start:stop + 1].copy()
        w[0] *= (old_edges[start + 1] - new_edges[j]) / (old_edges[start + 1] - old_edges[start])
        w[-1] *= (new_edges[j + 1] - old_edges[stop]) / (old_edges[stop + 1] - old_edges[stop])
        out[j] = np.sum(w * flux[start:stop + 1]) / np.sum(w)
    return out


def convolve_variable_width_gaussian(flux, sigma_pixels, trunc=4.0):
    """utils.py:129-182: per output pixel its own normalised Gaussian (sigma in pixels, truncated at ceil(trunc sigma)),
    nearest-edge padding, pixels with sigma <= 0.01 copied."""
    flux = np.asarray(flux, dtype=float)
    n = flux.size
    out = np.empty(n)
    for i in range(n):
        sg = float(sigma_pixels[i])
        if sg <= 0.01:
            out[i] = flux[i]
            continue
        hw = int(np.ceil(sg * trunc))
        x = np.arange(-hw, hw + 1)
        k = np.exp(-0.5 * (x / sg) ** 2)
        k /= k.sum()
        out[i] = np.dot(flux[np.clip(i + x, 0, n - 1)], k)
    return out


def transform_spectrum(theory_wave, theory_flux, z, observed_wave, resolution_curve_wave, resolution_curve_r,
                       theory_r=np.inf, trunc_constant=4.0):
    """utils.py:185-254: redshift the axis, smooth to the instrument's R(lambda) (in quadrature with the model's own
    resolution; sigma in pixels uses the MEDIAN pixel width of the redshifted axis), rebin to the observed pixels."""
    wz = np.asarray(theory_wave, dtype=float) * (1 + z)
    c = 2 * np.sqrt(2 * np.log(2))
    s_inst = wz / np.interp(wz, resolution_curve_wave, resolution_curve_r) / c
    r_th = np.full_like(wz, theory_r) if np.ndim(theory_r) == 0 else np.asarray(theory_r, dtype=float)
    s_th = wz / r_th / c
    s_pix = np.sqrt(np.maximum(s_inst**2 - s_th**2, 0.0)) / np.median(np.diff(wz))
    conv = convolve_variable_width_gaussian(theory_flux, s_pix, trunc=trunc_constant)
    return np.asarray(observed_wave), spectres_resample(observed_wave, wz, conv, fill=0.0)


def scale_to_mass(flux_base, log_mass, log_base_mass=9.0):
    """library.py:4588-4609  photometry cast to float32, then multiplied by the float64 mass ratio."""
    scale = 10.0 ** np.asarray(log_mass, dtype=float) / 10.0**log_base_mass
    return np.asarray(flux_base, dtype=np.float32) * scale[:, None]


# --------------------------------------------------------------------------------------
# noise_models.py:55-73 static converters ; noise_models.py:76-208 depth model
# --------------------------------------------------------------------------------------
def ab_to_jy(m):
    return 10 ** (-0.4 * (np.asarray(m, dtype=float) - 8.90))


def jy_to_ab(f):
    with np.errstate(all="ignore"):
        return -2.5 * np.log10(np.asarray(f, dtype=float)) + 8.90


def ab_err_to_jy(merr, fjy):
    return (np.asarray(fjy, dtype=float) * merr * np.log(10)) / 2.5


def jy_err_to_ab(ferr, fjy):
    with np.errstate(all="ignore"):
        return np.abs((2.5 / np.log(10)) * (np.asarray(ferr, dtype=float) / np.asarray(fjy, dtype=float)))


def depth_model_sigma_jy(depth_ab, sigma_level=5.0):
    return ab_to_jy(depth_ab) / sigma_level  # noise_models.py:104-105


def empirical_apply_noise(flux, model, draws, true_flux_units=None, out_units=None):
    """noise_models.py:818-880 (GeneralEmpiricalUncertaintyModel.apply_noise) with the random numbers injected per element.

    ``model``: dict(centers, median, std, extrapolate, flux_unit, interpolation_flux_unit, sigma_clip, error_type,
    upper_limits, snr_threshold, upper_limit_value, ul_flux_behaviour, ul_err_value, min_err, max_err); units are "AB" or
    the size of a linear unit in Jy.  ``draws`` (4, n): uniform for the sigma draw, normal for the scatter (uniform when
    sigma-clipped), uniform for the "observed" re-draw, uniform for the scatter about the upper limit.  The reference draws
    the last three only for the elements that need them (compacted arrays); a caller reproducing its stream scatters them.
    Returns (noisy flux, sigma) in ``out_units``."""
    from scipy import stats
    c, med, sd = (np.asarray(model[k], dtype=float) for k in ("centers", "median", "std"))

    def interp(y, v):
        v = np.asarray(v, dtype=float)
        out = np.interp(v, c, y)                                  # end-value fill
        if model.get("extrapolate", False):
            lo, hi = v < c[0], v > c[-1]
            out = np.where(lo, y[0] + (v - c[0]) * (y[1] - y[0]) / (c[1] - c[0]), out)
            out = np.where(hi, y[-2] + (v - c[-2]) * (y[-1] - y[-2]) / (c[-1] - c[-2]), out)
        return out

    def to_jy(f, e, unit):
        if unit == "AB":
            fj = ab_to_jy(f)
            return fj, ab_err_to_jy(e, fj)
        return f * unit, e * unit

    def convert(f, e, src, dst):
        if src == dst:
            return f, e
        if src == "AB":
            fj, ej = to_jy(f, e, "AB")
            return fj / dst, ej / dst
        if dst == "AB":
            fj, ej = f * src, e * src
            return jy_to_ab(fj), jy_err_to_ab(ej, fj)
        return f * (src / dst), e * (src / dst)

    def sample(f, u):
        mu, ss = interp(med, f), np.maximum(0, interp(sd, f))
        a = (0 - mu) / np.where(ss > 1e-9, ss, 1)
        return mu + ss * stats.truncnorm.ppf(u, a, np.inf)

    def below(f, e):
        fj, ej = to_jy(f, e, iu)
        with np.errstate(all="ignore"):
            snr = fj / ej
        return ~np.isfinite(snr) | (snr < model["snr_threshold"])

    iu = model["interpolation_flux_unit"]
    flux = np.asarray(flux, dtype=float)
    src = model["flux_unit"] if true_flux_units is None else true_flux_units
    dst = model["flux_unit"] if out_units is None else out_units
    f_int, _ = convert(flux, np.zeros_like(flux), src, iu)
    sig = sample(f_int, draws[0])
    init = below(f_int, sig) if model["upper_limits"] else np.zeros(flux.shape, dtype=bool)
    zz = stats.truncnorm.ppf(draws[1], -model["sigma_clip"], model["sigma_clip"]) if model.get("sigma_clip") is not None else draws[1]
    noisy = np.where(init, f_int, f_int + (0.0 + sig * zz))
    final = sample(noisy, draws[2]) if model["error_type"] == "observed" else sig
    if model["upper_limits"] and model.get("upper_limit_value") is not None:
        mask = init | below(noisy, final)
        beh, ul = model["ul_flux_behaviour"], model["upper_limit_value"]
        if beh == "scatter_limit":
            repl = ul + float(np.maximum(0, interp(sd, ul))) * stats.truncnorm.ppf(draws[3], -3, 3)
        else:
            repl = np.full(flux.shape, ul if beh == "upper_limit" else float(beh))
        noisy = np.where(mask, repl, noisy)
        final = np.where(mask, model["ul_err_value"], final)
    with np.errstate(all="ignore"):
        of, os_ = convert(noisy, final, iu, dst)
    return of, np.clip(os_, model["min_err"], model["max_err"])


def asinh_mag(f_jy, b_jy):
    """utils.py:647-675"""
    return -2.5 * np.log10(np.e) * (np.arcsinh(np.asarray(f_jy, dtype=float) / (2 * b_jy)) + np.log(b_jy / 3631.0))


def asinh_mag_err(f_jy, e_jy, b_jy):
    """utils.py:678-704"""
    return 2.5 * np.log10(np.e) * np.asarray(e_jy, dtype=float) / np.sqrt(np.asarray(f_jy, dtype=float) ** 2 + (2 * b_jy) ** 2)


def empirical_asinh_apply_noise(flux_jy, model, draws):
    """noise_models.py:507-557 (AsinhEmpiricalUncertaintyModel.apply_noise) with injected draws (3, n): sigma uniform,
    scatter normal, second sigma uniform.  ``model``: dict(centers, median, std, extrapolate, b [Jy], interpolation unit
    ("asinh" or the size of a linear unit in Jy), error_type, min_err, max_err).  Returns (asinh magnitudes, errors)."""
    from scipy import stats
    c, med, sd = (np.asarray(model[k], dtype=float) for k in ("centers", "median", "std"))

    def interp(y, v):
        out = np.interp(v, c, y)
        if model.get("extrapolate", False):
            out = np.where(v < c[0], y[0] + (v - c[0]) * (y[1] - y[0]) / (c[1] - c[0]), out)
            out = np.where(v > c[-1], y[-2] + (v - c[-2]) * (y[-1] - y[-2]) / (c[-1] - c[-2]), out)
        return out

    def sample(v, u):
        mu, ss = interp(med, v), np.maximum(0, interp(sd, v))
        return mu + ss * stats.truncnorm.ppf(u, (0 - mu) / np.where(ss > 1e-9, ss, 1), np.inf)

    f = np.asarray(flux_jy, dtype=float)
    b, iu = model["b"], model["interpolation_flux_unit"]
    if iu == "asinh":
        m_true = asinh_mag(f, b)
        e0 = sample(m_true, draws[0])
        m_noisy = m_true + (0.0 + e0 * draws[1])
        err = e0 if model["error_type"] == "empirical" else sample(m_noisy, draws[2])
    else:
        e0 = sample(f / iu, draws[0]) * iu
        noisy = f + (0.0 + e0 * draws[1])
        m_noisy = asinh_mag(noisy, b)
        e1 = sample(noisy / iu, draws[2]) * iu if model["error_type"] == "empirical" else e0
        err = asinh_mag_err(noisy, e1, b)
    return m_noisy, np.clip(err, model["min_err"], model["max_err"])


def depth_model_apply_noise(flux_jy, depth_ab, z, sigma_level=5.0, out_units=None,
                            min_err=0.0, max_err=np.inf):
    """DepthUncertaintyModel.apply_noise with the normal draws ``z`` injected
    (np.random.normal(loc, scale) == loc + scale*z; noise_models.py:146-159)."""
    sigma = depth_model_sigma_jy(depth_ab, sigma_level)
    noisy = np.asarray(flux_jy, dtype=float) + (0.0 + sigma * np.asarray(z, dtype=float))
    unc = np.ones_like(noisy) * sigma
    if out_units == "AB":
        unc = jy_err_to_ab(unc, noisy)
        noisy = jy_to_ab(noisy)
    return noisy, np.clip(unc, min_err, max_err)


# --------------------------------------------------------------------------------------
# sbi_runner.py:580-691 _apply_depths ; :1698-1716 AB features ; :1927-1932 clip
# --------------------------------------------------------------------------------------
def apply_depths(phot, depths_std, z, n_scatters, min_flux_pc_error=0.0, depth_indices=None):
    """phot (m, n) -> (m, n_scatters*n) noisy copy and the sigma used; ``z`` are injected normals.
    2-D ``depths_std`` (k, m) with ``depth_indices`` (m, n_scatters): sbi_runner.py:626-647 -- the picked set is expanded with
    ``np.repeat(..., n, axis=1)``, i.e. columns [j n, (j+1) n) of the repeated array use the set drawn for (row, j)."""
    m, n = phot.shape
    rep = np.repeat(phot, n_scatters, axis=1)
    depths_std = np.asarray(depths_std, dtype=float)
    if depths_std.ndim == 2:
        sel = depths_std[np.asarray(depth_indices), np.arange(m)[:, None]]       # (m, n_scatters)
        std = np.repeat(sel, n, axis=1)
    else:
        std = np.repeat(depths_std[:, None], n_scatters * n, axis=1)
    if min_flux_pc_error > 0.0:
        std = np.maximum(std, rep * min_flux_pc_error / 100.0)
    return rep + (0 + std * z), std


def ab_features(phot_njy, err_njy, norm_mag_limit=50.0):
    """Flux -> AB feature rows: uJy conversion, magnitude error, negative flux -> limit, clip."""
    f_ujy = phot_njy * 1e-3
    e_ujy = err_njy * 1e-3
    with np.errstate(all="ignore"):
        merr = 2.5 * e_ujy / (np.log(10) * f_ujy)
        mag = -2.5 * np.log10(f_ujy) + 23.9
    mag[f_ujy < 0] = norm_mag_limit
    mag[mag > norm_mag_limit] = norm_mag_limit
    return mag, merr


# utils.py:647-704
def f_jy_to_asinh(f_jy, f_b):
    return -2.5 * np.log10(np.e) * (np.arcsinh(f_jy / (2 * f_b)) + np.log(f_b / 3631.0))


def f_jy_err_to_asinh(f_jy, f_err, f_b):
    return 2.5 * np.log10(np.e) * f_err / np.sqrt(f_jy**2 + (2 * f_b) ** 2)
