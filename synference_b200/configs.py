"""The BASELINE.json workload configurations, built through the public API.

cfg1  README quickstart: LogNormal SFH, single metallicity, intrinsic (stellar+nebular), no dust,
      8 NIRCam wide filters, no dust (``README.md:86-134``)
cfg2  LogNormal SFH + Calzetti dust screen, z 0-10 with IGM, 20 NIRCam+MIRI filters
cfg3  continuity (non-parametric) SFH, Normal metallicity distribution, z 0-15
Parameter draws follow ``tests/conftest.py:132-148`` / SURVEY 8d (Latin hypercube, rng=42).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .cosmology import Planck18
from .engine import GalaxyParams
from .parametric import (Calzetti2000, Grid, Instrument, IntrinsicEmission, PacmanEmission, SFH, ZDist,
                         ZDistArray)
from .sampling import continuity_sfh_array, draw_from_hypercube, generate_sfh_basis
from .synthetic import NIRCAM_MIRI20, NIRCAM_WIDE8, synthetic_filters, synthetic_grid
from .utils import generate_constant_R


@dataclass
class Workload:
    name: str
    grid: Grid
    instrument: Instrument
    emission_model: object
    emission_key: str
    params: GalaxyParams
    samples: dict
    igm: bool = True
    max_redshift: float = 20.0

    @property
    def filters(self):
        return self.instrument.filters


def _model(filter_codes, z_grid_max):
    raw = synthetic_filters(filter_codes)
    lam = generate_constant_R(R=300, auto_start_stop=True, filterset=raw, max_redshift=z_grid_max)
    filters = synthetic_filters(filter_codes, new_lam=lam)
    grid = synthetic_grid(lam)
    return grid, Instrument("JWST", filters=filters)


def make_workload(name: str, n_gal: int, seed: int = 42) -> Workload:
    if name == "cfg1":
        grid, inst = _model(NIRCAM_WIDE8, 15.0)
        pr = {"log_stellar_mass": (8.0, 12.0), "redshift": (0.01, 10.0), "log_zmet": (-4.0, -1.4),
              "peak_age_norm": (0.0, 0.99), "tau": (0.2, 2.0)}
        s = draw_from_hypercube(pr, N=n_gal, rng=seed)
        em, key, tau_v = IntrinsicEmission(grid=grid), "intrinsic", None
    elif name in ("cfg2", "cfg4", "cfg5"):
        grid, inst = _model(NIRCAM_MIRI20, 10.0)
        pr = {"log_stellar_mass": (8.0, 12.0), "redshift": (0.01, 10.0), "log_zmet": (-4.0, -1.4),
              "peak_age_norm": (0.0, 0.99), "tau": (0.2, 2.0), "tau_v": (0.0, 3.0)}
        s = draw_from_hypercube(pr, N=n_gal, rng=seed)
        em = PacmanEmission(grid=grid, fesc=0.0, fesc_ly_alpha=1.0, dust_curve=Calzetti2000(), tau_v="tau_v")
        key, tau_v = "emergent", s["tau_v"]
    elif name == "cfg3":
        grid, inst = _model(NIRCAM_MIRI20, 15.0)
        pr = {"log_stellar_mass": (8.0, 12.0), "redshift": (0.01, 15.0), "log_zmet": (-4.0, -1.4),
              "zmet_sigma": (0.05, 0.5), "tau_v": (0.0, 3.0)}
        s = draw_from_hypercube(pr, N=n_gal, rng=seed)
        em = PacmanEmission(grid=grid, fesc=0.0, fesc_ly_alpha=1.0, dust_curve=Calzetti2000(), tau_v="tau_v")
        key, tau_v = "emergent", s["tau_v"]
    else:
        raise ValueError(f"unknown workload {name}")
    z = np.asarray(s["redshift"], dtype=np.float64)
    if name == "cfg3":
        rng = np.random.default_rng(seed)
        nb = 6
        ratios = np.clip(rng.standard_t(2, size=(n_gal, nb - 1)) * 1.0, -30, 30)
        tuniv = np.asarray((Planck18.age(z) - Planck18.age(20.0)).to("yr").value)
        # Prospector-style bins: 0-30 Myr, 30-100 Myr, then log-spaced to 0.85 t_univ, then to t_univ
        edges = np.empty((n_gal, nb + 1))
        edges[:, 0], edges[:, 1], edges[:, 2] = 1.0, 3.0e7, 1.0e8
        top = np.maximum(tuniv, 2.0e8)
        inner = np.exp(np.linspace(np.log(1.0e8), np.log(0.85 * top), nb - 2).T) if nb > 3 else None
        edges[:, 2:nb] = inner
        edges[:, nb] = top
        agebins = np.stack([np.log10(edges[:, :-1]), np.log10(edges[:, 1:])], axis=2)
        sfhs = continuity_sfh_array(ratios, agebins, z)
        zd = ZDistArray.normal(s["log_zmet"], s["zmet_sigma"], log10=True)
    else:
        sfhs, _ = generate_sfh_basis(SFH.LogNormal, ["tau", "peak_age_norm"],
                                     [np.asarray(s["tau"]), np.asarray(s["peak_age_norm"])], z,
                                     max_redshift=20, cosmo=Planck18)
        zd = ZDistArray.delta(log10metallicity=s["log_zmet"])
    params = GalaxyParams.from_objects(z, sfhs, zd, log_mass=s["log_stellar_mass"], tau_v=tau_v)
    return Workload(name, grid, inst, em, key, params, s, igm=True)
