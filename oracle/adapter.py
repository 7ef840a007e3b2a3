"""TEST INFRASTRUCTURE: turn a struct-of-arrays parameter block into the oracle's per-galaxy dicts.

Takes plain arrays / duck-typed objects only (nothing is imported from the product package).
"""

from __future__ import annotations

import numpy as np

from . import oracle as O

SFH_NAMES = {0: "Constant", 1: "Gaussian", 2: "Exponential", 3: "DecliningExponential",
             4: "DelayedExponential", 5: "LogNormal", 6: "DoublePowerLaw", 7: "Continuity"}
SFH_PARAMS = {"Constant": (), "Gaussian": ("peak_age", "sigma"), "Exponential": ("tau",),
              "DecliningExponential": ("tau",), "DelayedExponential": ("tau",),
              "LogNormal": ("tau", "peak_age"), "DoublePowerLaw": ("peak_age", "alpha", "beta")}
ZD_NAMES = {0: "delta_linear", 1: "delta_log10", 2: "normal_linear", 3: "normal_log10"}


def galaxies_from_params(p):
    """``p`` has redshift, sfh_type, sfh_rows, zd_type, zd_value, zd_sigma, tau_v (GalaxyParams-like)."""
    kind = SFH_NAMES[int(p.sfh_type)]
    out = []
    for i in range(len(p.redshift)):
        row = np.zeros(24)
        row[:p.sfh_rows.shape[1]] = p.sfh_rows[i]
        if kind == "Continuity":
            nb = int(row[2])
            sfh = dict(min_age=row[0], max_age=row[1], edges=row[3:3 + nb + 1].copy(),
                       logsfr_ratios=row[3 + nb + 1:3 + 2 * nb].copy())
        else:
            sfh = dict(min_age=row[0], max_age=row[1])
            for j, name in enumerate(SFH_PARAMS[kind]):
                sfh[name] = row[2 + j]
        out.append(dict(redshift=float(p.redshift[i]),
                        tau_v=0.0 if p.tau_v is None else float(p.tau_v[i]),
                        sfh_kind=kind, sfh=sfh, zd_kind=ZD_NAMES[int(p.zd_type)],
                        zd_value=float(p.zd_value[i]),
                        zd_sigma=0.0 if p.zd_sigma is None else float(p.zd_sigma[i])))
    return out


def weights_matrix(p, log10ages, metallicities):
    """(N, n_z*n_age) SFZH weights in the product's k = iz*n_age + ia order."""
    gals = galaxies_from_params(p)
    return np.stack([O.weights_for(g, log10ages, metallicities).T.reshape(-1) for g in gals])
