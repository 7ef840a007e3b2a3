"""GPU parity tests for the empirical uncertainty models on the device: sb2_empirical_noise through
`apply_empirical_noise_models`, against (1) golden outputs of the reference's own GeneralEmpiricalUncertaintyModel.apply_noise
(numpy stream seeded, the consumed draws replayed per element) and (2) the float64 oracle for unit / rule combinations.

Tolerance: float64 on both sides; the inverse normal CDF and 10**x come from different math libraries -> rtol 1e-9."""

import os

import numpy as np
import pytest

import synference_b200 as S
from oracle import oracle as O
from tests.test_oracle_golden import EMP_CASES, empirical_case

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_noise_golden.npz"))
RTOL = 1e-9


def host_model(case, **over):
    """The product's model object with the golden case's binned statistics (already_binned=True path)."""
    m = empirical_case(case)
    kw = dict(flux_unit="AB", sigma_clip=m["sigma_clip"], error_type=m["error_type"], upper_limits=m["upper_limits"],
              treat_as_upper_limits_below=m["snr_threshold"] if m["upper_limits"] else None,
              upper_limit_flux_behaviour=m["ul_flux_behaviour"], min_flux_error=m["min_err"], max_flux_error=m["max_err"],
              upper_limit_flux_err_behaviour={"limits": "flux", "limits_const": "sig_1"}.get(case, "flux"), return_noise=True)
    kw.update(over)
    mod = S.GeneralEmpiricalUncertaintyModel(m["centers"], None, already_binned=True, bin_median_errors=m["median"],
                                             bin_std_errors=m["std"], **kw)
    mod.upper_limit_value = m["upper_limit_value"]
    return mod


@pytest.mark.parametrize("case", list(EMP_CASES))
def test_golden_outputs_of_the_reference_class(case):
    mod = host_model(case)
    f, s = S.apply_empirical_noise_models(G["en_true_ab"][None, :], ["F"], {"F": mod}, N_scatters=1, flux_units="AB",
                                          normed_flux_units="AB", return_errors=True, draws=G[f"en_{case}_draws"][:, None, :])
    np.testing.assert_allclose(f[0], G[f"en_{case}_out_flux"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(s[0], G[f"en_{case}_out_sigma"], rtol=RTOL, atol=1e-12)


def test_units_rules_and_several_filters_against_the_oracle():
    """Interpolation in nJy / uJy / AB, inputs in nJy, outputs in AB / uJy / nJy, extrapolation, every upper-limit rule."""
    rng = np.random.default_rng(3)
    n = 5000
    flux_njy = np.abs(rng.lognormal(3.0, 1.5, n)) + 0.01
    centers = np.geomspace(0.5, 5000.0, 24)
    med = 2.0 + 0.02 * centers + rng.uniform(0, 0.2, 24)
    sd = 0.3 + 0.004 * centers
    cfgs = [
        dict(iu="nJy", out="AB", extrapolate=False, kw=dict()),
        dict(iu="uJy", out="uJy", extrapolate=True, kw=dict(sigma_clip=2.5, error_type="observed")),
        dict(iu="nJy", out="nJy", extrapolate=False, kw=dict(upper_limits=True, treat_as_upper_limits_below=3.0,
             upper_limit_flux_behaviour=7.5, upper_limit_flux_err_behaviour="max", max_flux_error=40.0)),
        dict(iu="nJy", out="AB", extrapolate=True, kw=dict(upper_limits=True, treat_as_upper_limits_below=2.0,
             upper_limit_flux_behaviour="scatter_limit", upper_limit_flux_err_behaviour="upper_limit", error_type="observed")),
    ]
    names, models, omodels = [], {}, []
    size = {"nJy": 1e-9, "uJy": 1e-6, "AB": "AB"}
    for i, c in enumerate(cfgs):
        scale = 1e-3 if c["iu"] == "uJy" else 1.0
        mod = S.GeneralEmpiricalUncertaintyModel(centers * scale, None, flux_unit=c["iu"], already_binned=True,
                                                 bin_median_errors=med * scale, bin_std_errors=sd * scale,
                                                 extrapolate=c["extrapolate"], return_noise=True, **c["kw"])
        if c["kw"].get("upper_limits"):
            mod.upper_limit_value = 12.0 * scale                       # as _setup_upper_limit_interpolator would leave it
        names.append(f"F{i}")
        models[f"F{i}"] = mod
        kw = c["kw"]
        om = dict(centers=centers * scale, median=med * scale, std=sd * scale, extrapolate=c["extrapolate"],
                  flux_unit=size[c["iu"]], interpolation_flux_unit=size[c["iu"]], sigma_clip=kw.get("sigma_clip"),
                  error_type=kw.get("error_type", "empirical"), upper_limits=kw.get("upper_limits", False),
                  snr_threshold=kw.get("treat_as_upper_limits_below", 0.0),
                  upper_limit_value=mod.upper_limit_value, ul_flux_behaviour=kw.get("upper_limit_flux_behaviour", "scatter_limit"),
                  ul_err_value=0.0, min_err=0.0, max_err=kw.get("max_flux_error", np.inf))
        if om["upper_limits"]:
            om["ul_err_value"] = float(mod._apply_error_behaviour(np.zeros(1), np.ones(1, dtype=bool))[0])
        omodels.append((om, size[c["out"]]))
    phot = np.repeat(flux_njy[None, :], len(cfgs), 0)
    draws = np.stack([rng.uniform(1e-6, 1 - 1e-6, (len(cfgs), n)), rng.standard_normal((len(cfgs), n)),
                      rng.uniform(1e-6, 1 - 1e-6, (len(cfgs), n)), rng.uniform(1e-6, 1 - 1e-6, (len(cfgs), n))])
    draws[1, 1] = rng.uniform(1e-6, 1 - 1e-6, n)                        # the sigma-clipped filter takes a uniform
    for out_unit in ("AB", "uJy"):
        f, s = S.apply_empirical_noise_models(phot, names, models, N_scatters=1, flux_units="nJy", normed_flux_units=out_unit,
                                              return_errors=True, draws=draws)
        for i, (om, _) in enumerate(omodels):
            wf, ws = O.empirical_apply_noise(flux_njy, om, draws[:, i], true_flux_units=1e-9, out_units=size[out_unit])
            ok = np.isfinite(wf)
            assert ok.mean() > 0.9 and np.array_equal(np.isfinite(f[i]), ok)
            np.testing.assert_allclose(f[i][ok], wf[ok], rtol=RTOL, atol=1e-12)
            oks = np.isfinite(ws)
            np.testing.assert_allclose(s[i][oks], ws[oks], rtol=RTOL, atol=1e-12)
    # the upper-limit rules did fire
    f, s = S.apply_empirical_noise_models(phot, names, models, N_scatters=1, flux_units="nJy", normed_flux_units="nJy",
                                          return_errors=True, draws=draws)
    assert (f[2] == 7.5).sum() > 10 and (s[2][f[2] == 7.5] == 40.0).all()


def test_philox_mode_statistics_layout_and_errors():
    import torch
    rng = np.random.default_rng(8)
    centers = np.linspace(20.0, 31.0, 16)
    mod = S.GeneralEmpiricalUncertaintyModel(centers, None, flux_unit="AB", already_binned=True,
                                             bin_median_errors=0.02 + np.exp((centers - 28) / 1.5), bin_std_errors=np.full(16, 0.01),
                                             return_noise=True)
    n_gal, n_sc = 40000, 3
    phot = rng.uniform(22, 27, (2, n_gal))
    f, s = S.apply_empirical_noise_models(phot, ["a", "b"], {"a": mod, "b": mod}, N_scatters=n_sc, flux_units="AB",
                                          normed_flux_units="AB", return_errors=True, seed=5, epoch=0)
    assert f.shape == (2, n_gal * n_sc)
    rep = np.repeat(phot, n_sc, axis=1)                                  # np.repeat layout: scatters of a galaxy are adjacent
    zs = (f - rep) / s
    assert abs(zs.mean()) < 0.01 and abs(zs.std() - 1) < 0.02            # sigma is drawn per element, mean sigma ~ mu(f)
    mu = np.interp(rep, centers, mod.median_error_in_bin)
    assert np.all(s > 0) and abs(np.mean(s / mu) - 1) < 0.05 and abs(np.std(s - mu) - 0.01) < 1e-3
    assert abs(np.corrcoef(zs[0], zs[1])[0, 1]) < 0.01 and abs(np.corrcoef(zs[0, :-1], zs[0, 1:])[0, 1]) < 0.01
    f2 = S.apply_empirical_noise_models(phot, ["a", "b"], {"a": mod, "b": mod}, N_scatters=n_sc, flux_units="AB", seed=5, epoch=0)
    f3 = S.apply_empirical_noise_models(phot, ["a", "b"], {"a": mod, "b": mod}, N_scatters=n_sc, flux_units="AB", seed=5, epoch=1)
    assert np.array_equal(f, f2) and not np.array_equal(f, f3)
    t = S.apply_empirical_noise_models(torch.as_tensor(phot, device="cuda"), ["a", "b"], {"a": mod, "b": mod}, N_scatters=n_sc,
                                       flux_units="AB", seed=5, epoch=0)
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), f)
    with pytest.raises(ValueError):
        S.apply_empirical_noise_models(phot, ["a", "b"], {"a": mod}, N_scatters=1)
    with pytest.raises(ValueError):
        S.apply_empirical_noise_models(phot, ["a"], {"a": mod, "b": mod}, N_scatters=1)
    with pytest.raises(ValueError):
        S.apply_empirical_noise_models(phot, ["a", "b"], [mod, mod])


@pytest.mark.parametrize("case", ["asinh", "asinh_observed", "asinh_flux_interp"])
def test_asinh_model_golden_outputs_of_the_reference_class(case):
    """AsinhEmpiricalUncertaintyModel on the device (noise_models.py:507-557; asinh magnitudes, utils.py:647-704)."""
    from synference_b200.units import Quantity
    from tests.test_oracle_golden import ASINH_CASES
    kw = ASINH_CASES[case]
    mod = S.AsinhEmpiricalUncertaintyModel(error_type=kw["error_type"], max_flux_error=kw["max_err"], return_noise=True,
                                           interpolation_flux_unit="asinh" if kw["interpolation_flux_unit"] == "asinh" else "nJy")
    mod.bin_centers, mod.median_error_in_bin, mod.std_error_in_bin = (G[f"ea_{case}_centers"], G[f"ea_{case}_median"],
                                                                      G[f"ea_{case}_std"])
    mod.b = Quantity(float(G[f"ea_{case}_b"]), "Jy")
    mod._create_interpolators()
    n = G["ea_true_jy"].size
    draws = np.full((4, 1, n), 0.5)
    draws[:3, 0] = G[f"ea_{case}_draws"]
    m, e = S.apply_empirical_noise_models(G["ea_true_jy"][None, :], ["F"], {"F": mod}, N_scatters=1, flux_units="Jy",
                                          return_errors=True, draws=draws)
    np.testing.assert_allclose(m[0], G[f"ea_{case}_out_mag"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(e[0], G[f"ea_{case}_out_err"], rtol=RTOL, atol=1e-12)
    # inputs in nJy give the same magnitudes; the Philox mode runs and keeps the noise level
    m2 = S.apply_empirical_noise_models(G["ea_true_jy"][None, :] * 1e9, ["F"], {"F": mod}, N_scatters=1, flux_units="nJy", draws=draws)
    np.testing.assert_allclose(m2[0], m[0], rtol=1e-9)
    mp, ep = S.apply_empirical_noise_models(np.repeat(G["ea_true_jy"][None, :], 1, 0), ["F"], {"F": mod}, N_scatters=200,
                                            flux_units="Jy", return_errors=True, seed=3)
    assert np.isfinite(mp).all() and abs(np.median(ep) / np.median(e[0]) - 1) < 0.05


def test_feature_builder_with_empirical_models():
    """create_feature_array_from_raw_photometry(..., empirical_noise_models=...) (sbi_runner.py:1678-1692): rows are the
    models' noisy AB magnitudes (clipped at norm_mag_limit) followed by their errors."""
    rng = np.random.default_rng(6)
    centers = np.linspace(20.0, 31.0, 16)
    mod = S.GeneralEmpiricalUncertaintyModel(centers, None, flux_unit="AB", already_binned=True,
                                             bin_median_errors=0.02 + np.exp((centers - 28) / 1.5), bin_std_errors=np.full(16, 0.01),
                                             return_noise=True)
    names = ["a", "b", "c"]
    grid = np.abs(rng.lognormal(4.0, 1.0, (3, 2000))) + 1.0                    # nJy
    feat, fnames, par = S.create_feature_array_from_raw_photometry(
        grid, names, scatter_fluxes=2, empirical_noise_models={k: mod for k in names}, include_errors_in_feature_array=True,
        parameter_array=rng.uniform(0, 1, (2000, 2)), seed=11)
    assert feat.shape == (4000, 6) and feat.dtype == np.float32 and fnames == names + [f"unc_{k}" for k in names]
    m0 = -2.5 * np.log10(np.repeat(grid, 2, axis=1).T * 1e-9) + 8.90
    zs = (feat[:, :3] - m0) / feat[:, 3:]
    assert abs(zs.mean()) < 0.05 and abs(zs.std() - 1) < 0.05 and par.shape == (4000, 2)
    mu = np.interp(m0, centers, mod.median_error_in_bin)
    assert abs(np.mean(feat[:, 3:] / mu) - 1) < 0.05


def _same_law(want, got, what):
    """Two samples of one distribution: equal share of non-finite values, and empirical CDFs that agree to 0.008 (5 standard
    errors at 100 000 draws) at the first sample's 2 / 16 / 50 / 84 / 98 % points -- evaluated a relative 1e-4 to either side
    of each point, which steps over float32 rounding of a point mass (a constant a rule puts in place of many sources)."""
    fin_w, fin_g = np.isfinite(want), np.isfinite(got)
    assert abs(fin_w.mean() - fin_g.mean()) < 0.01, what          # log of a negative noisy flux: NaN on both sides
    w, g = want[fin_w], got[fin_g]
    for t in np.quantile(w, [0.02, 0.16, 0.5, 0.84, 0.98]):
        for eps in (-1e-4, 1e-4):
            tt = t + eps * abs(t)
            assert abs(np.mean(w <= tt) - np.mean(g <= tt)) < 0.008, (what, t, eps, np.mean(w <= tt), np.mean(g <= tt))


def test_production_kernel_draws_from_the_same_law_as_the_parity_kernel():
    """`empirical_noise_fast_kernel` (Philox, float32 arithmetic) against `empirical_noise_kernel` (float64, injected numpy
    draws) for unit changes, sigma clipping, re-drawn errors, both upper-limit rules and the asinh models: at a handful of
    fixed true fluxes the two outputs are samples of one distribution, so their quantiles and the share of replaced sources
    agree within sampling error (`_same_law`)."""
    rng = np.random.default_rng(17)
    centers = np.geomspace(0.5, 5000.0, 24)
    med = 2.0 + 0.02 * centers
    sd = 0.3 + 0.004 * centers
    m_per = 100_000
    levels = np.array([4.0, 30.0, 400.0, 3000.0, 9000.0])                   # nJy, the last one beyond the table
    flux = np.repeat(levels, m_per)[None, :]
    cfgs = [
        dict(iu="nJy", out="AB", extrapolate=False, kw=dict()),
        dict(iu="uJy", out="uJy", extrapolate=True, kw=dict(sigma_clip=2.5, error_type="observed")),
        dict(iu="nJy", out="nJy", extrapolate=False, kw=dict(upper_limits=True, treat_as_upper_limits_below=3.0,
             upper_limit_flux_behaviour=7.5, upper_limit_flux_err_behaviour="max", max_flux_error=40.0)),
        dict(iu="nJy", out="AB", extrapolate=True, kw=dict(upper_limits=True, treat_as_upper_limits_below=2.0,
             upper_limit_flux_behaviour="scatter_limit", upper_limit_flux_err_behaviour="upper_limit", error_type="observed")),
    ]
    for c in cfgs:
        scale = {"nJy": 1.0, "uJy": 1e-3}[c["iu"]]
        mod = S.GeneralEmpiricalUncertaintyModel(centers * scale, None, flux_unit=c["iu"], already_binned=True,
                                                 bin_median_errors=med * scale, bin_std_errors=sd * scale, return_noise=True,
                                                 extrapolate=c["extrapolate"], **c["kw"])
        if c["kw"].get("upper_limits"):
            mod.upper_limit_value = 6.0 * scale
        n = flux.shape[1]
        draws = np.stack([rng.uniform(0, 1, (1, n)), rng.uniform(0, 1, (1, n)) if "sigma_clip" in c["kw"] else rng.standard_normal((1, n)),
                          rng.uniform(0, 1, (1, n)), rng.uniform(0, 1, (1, n))])
        want_f, want_s = S.apply_empirical_noise_models(flux, ["F"], {"F": mod}, N_scatters=1, flux_units="nJy",
                                                        normed_flux_units=c["out"], return_errors=True, draws=draws)
        got_f, got_s = S.apply_empirical_noise_models(flux, ["F"], {"F": mod}, N_scatters=1, flux_units="nJy",
                                                      normed_flux_units=c["out"], return_errors=True, seed=9, epoch=2)
        if c["out"] == "AB":       # compare on a linear scale: a magnitude's far tail (noisy flux -> 0) has no stable quantiles
            want_s, got_s = want_s * 10 ** (-0.4 * want_f), got_s * 10 ** (-0.4 * got_f)
            want_f, got_f = 10 ** (-0.4 * want_f), 10 ** (-0.4 * got_f)
        for k in range(len(levels)):
            sl = slice(k * m_per, (k + 1) * m_per)
            _same_law(want_f[0, sl], got_f[0, sl], (c, k, "flux"))
            _same_law(want_s[0, sl], got_s[0, sl], (c, k, "sigma"))
    # asinh magnitudes: tables in asinh mags and in a linear unit
    from synference_b200.units import Quantity
    for iu in ("asinh", "nJy"):
        mod = S.AsinhEmpiricalUncertaintyModel(error_type="empirical", return_noise=True, interpolation_flux_unit=iu)
        b = 5.0
        mod.b = Quantity(b * 1e-9, "Jy")
        if iu == "asinh":
            cm = np.linspace(22.0, 32.0, 20)
            mod.bin_centers, mod.median_error_in_bin, mod.std_error_in_bin = cm, 0.02 + 0.3 * np.exp((cm - 29) / 1.2), np.full(20, 0.01)
        else:
            mod.bin_centers, mod.median_error_in_bin, mod.std_error_in_bin = centers, med, sd
        mod._create_interpolators()
        n = flux.shape[1]
        draws = np.stack([rng.uniform(0, 1, (1, n)), rng.standard_normal((1, n)), rng.uniform(0, 1, (1, n)), np.full((1, n), 0.5)])
        want_f, want_s = S.apply_empirical_noise_models(flux, ["F"], {"F": mod}, N_scatters=1, flux_units="nJy", return_errors=True, draws=draws)
        got_f, got_s = S.apply_empirical_noise_models(flux, ["F"], {"F": mod}, N_scatters=1, flux_units="nJy", return_errors=True, seed=4)
        for k in range(len(levels)):
            sl = slice(k * m_per, (k + 1) * m_per)
            _same_law(want_f[0, sl], got_f[0, sl], (iu, k, "mag"))
            _same_law(want_s[0, sl], got_s[0, sl], (iu, k, "err"))
    # an odd row count takes the scalar tail path and every element is still written
    f_odd = S.apply_empirical_noise_models(flux[:, :10001], ["F"], {"F": mod}, N_scatters=1, flux_units="nJy", seed=4)
    assert f_odd.shape == (1, 10001) and np.isfinite(f_odd).all()
