"""Planck18 background cosmology without astropy.

The reference takes ``astropy.cosmology.Planck18`` (``library.py:1146``,
``library.py:1206`` for ``age(z)``; the luminosity distance is used inside the
observed-frame step).  astropy is not available here, so the
same flat LCDM model with photons and one massive neutrino species (astropy's
Komatsu-style fitting function) is restated.  Host code evaluates it through a
dense cubic-Hermite table in ``s = ln(1+z)`` (exact derivatives are known
analytically); the same table is uploaded to the GPU for the per-galaxy flux
scale.  The oracle integrates with ``scipy.integrate.quad`` independently.
"""

from __future__ import annotations

import numpy as np

from .units import Gyr, Mpc, Quantity

__all__ = ["FlatLambdaCDM", "Planck18", "CosmologyTable"]

_C_KMS = 299792.458
_G_SI = 6.6743e-11
_SIGMA_SB = 5.670374419e-8
_KB_EVK = 8.617333262e-5
_MPC_M = 3.0856775814913673e22
_GYR_S = 3.15576e16
_MPC_CM = 3.0856775814913673e24


class CosmologyTable:
    """Cubic-Hermite tables of comoving distance [Mpc] and age [Gyr] over s=ln(1+z)."""

    def __init__(self, s_max, n_cells, dc, ddc, age, dage):
        self.s_max = float(s_max)
        self.n_cells = int(n_cells)
        self.ds = self.s_max / self.n_cells
        self.dc, self.ddc, self.age, self.dage = dc, ddc, age, dage

    @staticmethod
    def _hermite(y, dy, ds, s):
        x = np.clip(np.asarray(s, dtype=float) / ds, 0.0, len(y) - 1 - 1e-12)
        k = np.floor(x).astype(np.int64)
        t = x - k
        h00 = (1 + 2 * t) * (1 - t) ** 2
        h10 = t * (1 - t) ** 2
        h01 = t * t * (3 - 2 * t)
        h11 = t * t * (t - 1)
        return h00 * y[k] + h10 * ds * dy[k] + h01 * y[k + 1] + h11 * ds * dy[k + 1]

    def comoving_distance_mpc(self, z):
        return self._hermite(self.dc, self.ddc, self.ds, np.log1p(np.asarray(z, dtype=float)))

    def age_gyr(self, z):
        return self._hermite(self.age, self.dage, self.ds, np.log1p(np.asarray(z, dtype=float)))


class FlatLambdaCDM:
    """Flat LCDM with radiation and massive neutrinos (astropy conventions)."""

    def __init__(self, H0, Om0, Tcmb0=0.0, Neff=3.04, m_nu=(0.0,), Ob0=None, name=None):
        self.H0, self.Om0, self.Tcmb0, self.Neff, self.Ob0 = H0, Om0, Tcmb0, Neff, Ob0
        self.m_nu = tuple(float(m) for m in m_nu)
        self.name = name or "FlatLambdaCDM"
        h0_si = H0 * 1.0e3 / _MPC_M
        rho_crit = 3.0 * h0_si**2 / (8.0 * np.pi * _G_SI)
        c_si = _C_KMS * 1.0e3
        self.Ogamma0 = 4.0 * _SIGMA_SB / c_si**3 * Tcmb0**4 / rho_crit
        nnu = int(np.floor(Neff))
        self._neff_per_nu = Neff / nnu if nnu > 0 else 0.0
        massive = [m for m in self.m_nu if m > 0]
        self._nmasslessnu = nnu - len(massive)
        tnu0 = 0.7137658555036082 * Tcmb0
        self._nu_y = np.array([m / (_KB_EVK * tnu0) for m in massive]) if Tcmb0 > 0 else np.zeros(0)
        self.Onu0 = self.Ogamma0 * self.nu_relative_density(0.0)
        self.Ode0 = 1.0 - Om0 - self.Ogamma0 - self.Onu0
        self.hubble_time_gyr = 1.0 / h0_si / _GYR_S
        self.hubble_distance_mpc = _C_KMS / H0
        self._table = None

    def nu_relative_density(self, z):
        prefac = 0.22710731766
        z = np.asarray(z, dtype=float)
        if self._nu_y.size == 0:
            return prefac * self.Neff * np.ones_like(z)
        p, invp, k = 1.83, 0.54644808743, 0.3173
        y = self._nu_y[(None,) * z.ndim] / (1.0 + z[..., None])
        rel = ((1.0 + (k * y) ** p) ** invp).sum(-1) + self._nmasslessnu
        return prefac * self._neff_per_nu * rel

    def efunc(self, z):
        z = np.asarray(z, dtype=float)
        zp1 = 1.0 + z
        o_r = self.Ogamma0 * (1.0 + self.nu_relative_density(z))
        return np.sqrt(zp1**3 * (o_r * zp1 + self.Om0) + self.Ode0)

    # ---- table -----------------------------------------------------------
    def table(self, z_max=100.0, n_cells=4096) -> CosmologyTable:
        if self._table is not None and self._table.s_max >= np.log1p(z_max) - 1e-12 \
                and self._table.n_cells == n_cells:
            return self._table
        s_max = float(np.log1p(z_max))
        s = np.linspace(0.0, s_max, n_cells + 1)
        xg, wg = np.polynomial.legendre.leggauss(12)

        def cell_integrals(f, a, b):
            mid, half = 0.5 * (a + b), 0.5 * (b - a)
            pts = mid[:, None] + half[:, None] * xg[None, :]
            return (f(pts) * wg[None, :]).sum(1) * half

        inv_e = lambda ss: 1.0 / self.efunc(np.expm1(ss))  # noqa: E731
        # comoving distance: D_H * int_0^s e^s / E ds
        f_dc = lambda ss: np.exp(ss) * inv_e(ss)  # noqa: E731
        dc = np.concatenate([[0.0], np.cumsum(cell_integrals(f_dc, s[:-1], s[1:]))])
        dc *= self.hubble_distance_mpc
        ddc = self.hubble_distance_mpc * f_dc(s)
        # age: t_H * int_s^inf ds / E ; tail beyond s_max integrated numerically
        cells = cell_integrals(inv_e, s[:-1], s[1:])
        st = np.linspace(s_max, s_max + 40.0, 4001)
        tail = cell_integrals(inv_e, st[:-1], st[1:]).sum()
        age = np.concatenate([np.cumsum(cells[::-1])[::-1], [0.0]]) + tail
        age *= self.hubble_time_gyr
        dage = -self.hubble_time_gyr * inv_e(s)
        self._table = CosmologyTable(s_max, n_cells, dc, ddc, age, dage)
        return self._table

    # ---- astropy-like accessors -----------------------------------------
    def age(self, z):
        return Quantity(self.table().age_gyr(z), Gyr)

    def comoving_distance(self, z):
        return Quantity(self.table().comoving_distance_mpc(z), Mpc)

    def luminosity_distance(self, z):
        z = np.asarray(z, dtype=float)
        return Quantity((1.0 + z) * self.table().comoving_distance_mpc(z), Mpc)

    def luminosity_distance_cm(self, z):
        return np.asarray(self.luminosity_distance(z)) * _MPC_CM

    def __repr__(self):
        return f"{self.name}(H0={self.H0}, Om0={self.Om0}, Tcmb0={self.Tcmb0}, Neff={self.Neff}, m_nu={self.m_nu})"


Planck18 = FlatLambdaCDM(H0=67.66, Om0=0.30966, Tcmb0=2.7255, Neff=3.046,
                         m_nu=(0.0, 0.0, 0.06), Ob0=0.04897, name="Planck18")
