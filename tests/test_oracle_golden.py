"""Pin the oracle (and the host-side product code) against the reference.

Sources of truth, strongest first:
  1. ``tests/golden/reference_noise_golden.npz`` - outputs of the reference's OWN code
     (noise_models.py, utils.py, SBI_Fitter._apply_depths, draw_from_hypercube) run offline by
     ``tests/golden/make_golden_from_reference.py``;
  2. the known-answer checks in the reference's tests (``tests/test_uncertainty_models.py:47-74``);
  3. internal consistency of the unpinned Synthesizer-side restatement (closed forms vs quad, C vs numpy).
"""

import os

import numpy as np
import pytest

from oracle import adapter as A, oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_noise_golden.npz"))


# ---- 1. golden vectors from the reference's own code -------------------------------------------
def test_converters_match_reference_code():
    np.testing.assert_array_equal(O.ab_to_jy(G["conv_mags"]), G["conv_ab_to_jy"])
    np.testing.assert_array_equal(O.jy_to_ab(G["conv_ab_to_jy"]), G["conv_jy_to_ab"])
    np.testing.assert_array_equal(O.ab_err_to_jy(G["conv_merr"], G["conv_ab_to_jy"]), G["conv_ab_err_to_jy"])
    np.testing.assert_array_equal(O.jy_err_to_ab(G["conv_ab_err_to_jy"], G["conv_ab_to_jy"]), G["conv_jy_err_to_ab"])


def test_depth_model_matches_reference_code_bit_exact():
    assert O.depth_model_sigma_jy(29.0, 5) == float(G["depth_sigma_jy"])
    noisy, unc = O.depth_model_apply_noise(G["depth_flux_ujy"] * 1e-6, 29.0, G["depth_z"])
    np.testing.assert_array_equal(noisy, G["depth_noisy_jy"])
    np.testing.assert_array_equal(unc, G["depth_unc_jy"])
    noisy_ab, unc_ab = O.depth_model_apply_noise(G["depth_flux_ujy"] * 1e3 * 1e-9, 29.0, G["depth_z"], out_units="AB")
    np.testing.assert_allclose(noisy_ab, G["depth_noisy_ab"], rtol=0, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(unc_ab, G["depth_unc_ab"], rtol=1e-12, equal_nan=True)


def test_apply_depths_matches_reference_code_bit_exact():
    std = G["ad_depths"] / 5
    out, err = O.apply_depths(G["ad_phot"], std, G["ad_z"], 3)
    np.testing.assert_array_equal(out, G["ad_out"])
    np.testing.assert_array_equal(err, G["ad_err"])
    out, err = O.apply_depths(G["ad_phot"], std, G["ad_z"], 3, min_flux_pc_error=10.0)
    np.testing.assert_array_equal(out, G["ad_out_pc"])
    np.testing.assert_array_equal(err, G["ad_err_pc"])


def test_apply_depths_with_depth_sets_matches_reference_code():
    """2-D depths (sbi_runner.py:626-647): randint pick per (filter, scatter), expanded with np.repeat(..., n, axis=1)."""
    noisy, std = O.apply_depths(G["ad_phot"], G["ad2_depths"] / 5.0, G["ad2_z"], 3, depth_indices=G["ad2_idx"])
    np.testing.assert_array_equal(std, G["ad2_err"])
    np.testing.assert_array_equal(noisy, G["ad2_out"])


def test_asinh_and_constant_r_match_reference_code():
    np.testing.assert_allclose(O.f_jy_to_asinh(G["asinh_f"], float(G["asinh_b"])), G["asinh_mag"], rtol=1e-14)
    np.testing.assert_allclose(O.f_jy_err_to_asinh(G["asinh_f"], G["asinh_e"], float(G["asinh_b"])), G["asinh_err"], rtol=1e-14)
    np.testing.assert_array_equal(O.constant_r_grid(1000.0, 1100.0, 300), G["const_r_300"])


def test_spectroscopic_path_matches_reference_code():
    """utils.py:129-254 run offline: the variable-width Gaussian is the reference's own loop; transform_spectrum is the
    reference's own function around the oracle's restatement of `spectres` (third party, unpinned)."""
    conv = O.convolve_variable_width_gaussian(G["sp_flux"], G["sp_sigma_pix"], 4.0)
    np.testing.assert_allclose(conv, G["sp_conv"], rtol=1e-13)
    for z, want in zip(G["sp_z"], G["sp_out"]):
        _, got = O.transform_spectrum(G["sp_wave"], G["sp_flux"], float(z), G["sp_obs_wave"], G["sp_res_wave"], G["sp_res_r"])
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
    _, got = O.transform_spectrum(G["sp_wave"], G["sp_flux"], 2.0, G["sp_obs_wave"], G["sp_res_wave"], G["sp_res_r"], theory_r=1000.0)
    np.testing.assert_allclose(got, G["sp_out_r1000"], rtol=1e-12)
    assert (G["sp_out"][-1] == 0).sum() > 0 and (G["sp_out"][0] > 0).all()     # z = 14: the bluest pixels are not covered -> fill
    # the rebin conserves flux: a constant spectrum stays constant, integral over covered pixels is preserved
    w = G["sp_wave"]
    flat = O.spectres_resample(G["sp_obs_wave"], w, np.full(w.size, 3.0))
    np.testing.assert_allclose(flat, 3.0, rtol=1e-13)


EMP_CASES = {"plain": dict(), "clip": dict(sigma_clip=2.0),
             "limits": dict(upper_limits=True, snr_threshold=3.0, ul_flux_behaviour="scatter_limit", error_type="observed",
                            max_err=5.0),
             "limits_const": dict(upper_limits=True, snr_threshold=2.0, ul_flux_behaviour="upper_limit", min_err=0.02)}


def empirical_case(name):
    """The oracle's model dict for one of the golden cases (tests/golden/make_golden_from_reference.py)."""
    m = dict(centers=G[f"en_{name}_centers"], median=G[f"en_{name}_median"], std=G[f"en_{name}_std"], extrapolate=False,
             flux_unit="AB", interpolation_flux_unit="AB", sigma_clip=None, error_type="empirical", upper_limits=False,
             snr_threshold=0.0, upper_limit_value=None, ul_flux_behaviour="scatter_limit", ul_err_value=0.0, min_err=0.0,
             max_err=np.inf)
    m.update(EMP_CASES[name])
    if m["upper_limits"]:
        m["upper_limit_value"] = float(G[f"en_{name}_ul_value"])
        m["ul_err_value"] = float(G[f"en_{name}_ul_err"])
    return m


@pytest.mark.parametrize("name", list(EMP_CASES))
def test_empirical_noise_oracle_matches_reference_class(name):
    """GeneralEmpiricalUncertaintyModel.apply_noise run by the reference's own code with numpy's global stream seeded; the
    oracle gets the same random numbers per element and must reproduce fluxes and errors."""
    f, s = O.empirical_apply_noise(G["en_true_ab"], empirical_case(name), G[f"en_{name}_draws"])
    np.testing.assert_allclose(f, G[f"en_{name}_out_flux"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(s, G[f"en_{name}_out_sigma"], rtol=1e-12, atol=1e-12)
    if name.startswith("limits"):
        assert (f == float(G[f"en_{name}_ul_value"])).sum() > 5 or name == "limits"      # some sources became upper limits


ASINH_CASES = {"asinh": dict(error_type="empirical", interpolation_flux_unit="asinh", max_err=np.inf),
               "asinh_observed": dict(error_type="observed", interpolation_flux_unit="asinh", max_err=0.8),
               "asinh_flux_interp": dict(error_type="empirical", interpolation_flux_unit=1e-9, max_err=np.inf)}


def asinh_case(name):
    m = dict(centers=G[f"ea_{name}_centers"], median=G[f"ea_{name}_median"], std=G[f"ea_{name}_std"], extrapolate=False,
             b=float(G[f"ea_{name}_b"]), min_err=0.0)
    m.update(ASINH_CASES[name])
    return m


@pytest.mark.parametrize("name", [k for k in ASINH_CASES if f"ea_{k}_out_mag" in G.files])
def test_asinh_noise_oracle_matches_reference_class(name):
    """AsinhEmpiricalUncertaintyModel.apply_noise by the reference's own code (global numpy stream seeded, draws replayed)."""
    m, e = O.empirical_asinh_apply_noise(G["ea_true_jy"], asinh_case(name), G[f"ea_{name}_draws"])
    np.testing.assert_allclose(m, G[f"ea_{name}_out_mag"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(e, G[f"ea_{name}_out_err"], rtol=1e-12, atol=1e-12)


def test_product_host_code_matches_reference_code():
    """The API-level host functions of the product against the same vectors."""
    import synference_b200 as S
    np.testing.assert_array_equal(np.asarray(S.UncertaintyModel.ab_to_jy(G["conv_mags"])), G["conv_ab_to_jy"])
    np.testing.assert_array_equal(S.UncertaintyModel.jy_to_ab(S.Quantity(G["conv_ab_to_jy"], "Jy")), G["conv_jy_to_ab"])
    np.testing.assert_array_equal(np.asarray(S.generate_constant_R(R=300, start=1000 * S.Angstrom, end=1100 * S.Angstrom)),
                                  G["const_r_300"])
    b = S.Quantity(float(G["asinh_b"]), "Jy")
    np.testing.assert_allclose(S.f_jy_to_asinh(S.Quantity(G["asinh_f"], "Jy"), b), G["asinh_mag"], rtol=1e-14)
    np.testing.assert_allclose(np.asarray(S.asinh_to_f_jy(G["asinh_mag"], b)), G["asinh_back"], rtol=1e-12, atol=1e-24)
    # draw_from_hypercube: same scipy engine, same seed -> identical float32 draws (tests/conftest.py:132-148)
    pr = {"redshift": (0.01, 10), "masses": (5, 11), "tau_v": (0, 2), "peak_age": (0, 0.99), "tau": (0.1, 1.5),
          "log_zmet": (-3, -1.39)}
    d = S.draw_from_hypercube(pr, N=100, rng=42, unlog_keys=["masses"])
    for k, v in d.items():
        assert np.asarray(v).dtype == np.float32
        np.testing.assert_array_equal(np.asarray(v), G[f"lhc_{k}"])
    # depth model through the product class with numpy's global stream
    m = S.DepthUncertaintyModel(depth_ab=29.0, depth_sigma_level=5, return_noise=True)
    assert float(m.sigma.value) == float(G["depth_sigma_jy"])
    np.random.seed(1234)
    noisy, unc = m.apply_noise(S.Quantity(G["depth_flux_ujy"], "uJy"))
    np.testing.assert_array_equal(np.asarray(noisy), G["depth_noisy_jy"])
    np.random.seed(1234)
    noisy_ab, unc_ab = m.apply_noise(G["depth_flux_ujy"] * 1e3, true_flux_units="nJy", out_units="AB")
    np.testing.assert_allclose(noisy_ab, G["depth_noisy_ab"], atol=1e-12, equal_nan=True)
    # empirical model construction (binning + interpolators)
    gm = S.GeneralEmpiricalUncertaintyModel(G["emp_obs"], G["emp_err"], flux_unit="AB", log_bins=False, num_bins=16)
    np.testing.assert_allclose(gm.bin_centers, G["emp_centers"], rtol=1e-14)
    np.testing.assert_allclose(gm.median_error_in_bin, G["emp_median"], rtol=1e-14)
    np.testing.assert_allclose(gm._mu_sigma_interpolator(G["emp_probe"]), G["emp_mu"], rtol=1e-13)
    np.testing.assert_allclose(gm._sigma_sigma_interpolator(G["emp_probe"]), G["emp_sig"], rtol=1e-13)


# ---- 2. the reference's own known-answer tests ---------------------------------------------------
def test_reference_known_answers():
    assert O.ab_to_jy(23.9) == pytest.approx(1e-6)                     # tests/test_uncertainty_models.py:49-51
    assert O.jy_to_ab(1e-6) == pytest.approx(23.9, abs=1e-2)           # :53-55
    fe = O.ab_err_to_jy(0.1, 1e-6)
    assert O.jy_err_to_ab(fe, 1e-6) == pytest.approx(0.1)              # :57-61
    assert O.depth_model_sigma_jy(25.0, 5) == pytest.approx(O.ab_to_jy(25.0) / 5.0)  # :67-74


# ---- 3. internal consistency of the Synthesizer-side restatement ---------------------------------
LOG10AGES = np.round(np.arange(6.0, 11.0 + 1e-9, 0.1), 1)


@pytest.mark.parametrize("kind,p", [
    ("LogNormal", dict(min_age=0.0, max_age=3e8, tau=0.5, peak_age=1e8)),        # tests/conftest.py:102-105
    ("LogNormal", dict(min_age=0.0, max_age=5e9, tau=1.7, peak_age=4.9e9)),
    ("DelayedExponential", dict(min_age=0.0, max_age=2e9, tau=5e8)),
    ("Gaussian", dict(min_age=0.0, max_age=2e9, peak_age=5e8, sigma=1e8)),
    ("Exponential", dict(min_age=1e7, max_age=2e9, tau=5e8)),
    ("DecliningExponential", dict(min_age=0.0, max_age=2e9, tau=5e8)),
    ("Constant", dict(min_age=1e7, max_age=1e8)),
])
def test_closed_form_bin_masses_equal_quad(kind, p):
    a = O.sfh_bin_masses(kind, p, LOG10AGES)
    b = O.sfh_bin_masses_quad(kind, p, LOG10AGES)
    assert a[-1] == 0.0
    np.testing.assert_allclose(a / a.sum(), b / b.sum(), atol=1e-9 * (a / a.sum()).max())


def test_zdist_weights():
    z = np.array([1e-5, 1e-4, 1e-3, 2e-3, 3e-3, 4e-3, 6e-3, 8e-3, 1e-2, 1.4e-2, 2e-2, 3e-2, 4e-2])
    w = O.zdist_weights("delta_log10", -2.5, 0.0, z)
    assert w.sum() == pytest.approx(1.0) and (w > 0).sum() == 2
    j = np.nonzero(w)[0]
    assert np.sum(w[j] * np.log10(z[j])) == pytest.approx(-2.5)          # linear interpolation in log10 Z
    assert O.zdist_weights("delta_log10", -9.0, 0.0, z)[0] == 1.0         # clamped below
    assert O.zdist_weights("delta_linear", 1.0, 0.0, z)[-1] == 1.0        # clamped above
    w = O.zdist_weights("normal_log10", -2.0, 0.3, z)
    assert w.sum() == pytest.approx(1.0) and np.argmax(w) == np.argmin(np.abs(np.log10(z) + 2.0))


def test_planck18_known_values():
    assert O.age_gyr(0.0) == pytest.approx(13.7869, abs=2e-4)             # astropy Planck18.age(0)
    assert O.luminosity_distance_cm(1.0) / 3.0856775814913673e24 == pytest.approx(6791.27, abs=0.05)
    from synference_b200.cosmology import Planck18
    for z in (0.0, 0.5, 3.0, 7.0, 15.0, 20.0):                             # product table vs oracle quad
        assert float(Planck18.age(z).value) == pytest.approx(O.age_gyr(z), rel=1e-9)
        if z > 0:
            assert float(Planck18.luminosity_distance(z).value) * 3.0856775814913673e24 == \
                pytest.approx(O.luminosity_distance_cm(z), rel=1e-9)


def test_igm_product_table_matches_oracle_loops():
    from synference_b200 import igm as I
    lam = O.constant_r_grid(500.0, 1400.0, 300)
    for z in (0.3, 1.19, 1.21, 2.5, 4.69, 4.71, 6.0, 9.9):
        t_or = O.inoue14_transmission(z, lam * (1 + z), I.INOUE14_LAF, I.INOUE14_DLA)
        t_pr = I.transmission(z, lam * (1 + z))
        np.testing.assert_allclose(t_pr, t_or, rtol=1e-12, atol=1e-300)
        assert np.all(t_or[lam > 1216] == 1.0) and np.all(t_or <= 1.0)


def test_c_oracle_equals_numpy_oracle():
    from oracle import c_oracle as CO
    from synference_b200 import igm as I
    from synference_b200.configs import make_workload
    CO.build()
    for name in ("cfg1", "cfg2", "cfg3"):
        w = make_workload(name, 24)
        lam = np.asarray(w.grid.lam)
        filt = [(f.lam, f.t) for f in w.filters]
        dust = dict(curve="Calzetti2000") if w.emission_model.dust_curve is not None else None
        fo = O.synthesize(A.galaxies_from_params(w.params), w.grid.log10ages, w.grid.metallicity, lam,
                          w.grid.spectra, filt, key=w.emission_key, dust=dust, igm=(I.INOUE14_LAF, I.INOUE14_DLA))
        ga, gu = O.emission_parts(w.grid.spectra, lam, w.emission_key)
        fc = CO.synthesize(w.params, w.grid.log10ages, w.grid.metallicity, lam, ga, gu, filt,
                           kappa=O.dust_kappa(lam) if dust else None, igm=(I.INOUE14_LAF, I.INOUE14_DLA))
        ok = np.abs(fo) > 1e-200
        np.testing.assert_allclose(fc[ok], fo[ok], rtol=1e-9)


def test_filter_variants_differ_at_the_expected_level():
    """nu- and lambda-integration differ at O((dlam/lam)^2) ~ 3e-6 on an R=300 grid."""
    lam = O.constant_r_grid(3000.0, 60000.0, 300)
    fnu = (lam / 1e4) ** -1.3
    fl = np.linspace(10000.0, 13000.0, 400)
    ft = np.exp(-0.5 * ((fl - 11500.0) / 700.0) ** 2)
    a = O.apply_filter(fnu, lam * 1.7, fl, ft, "nu")
    b = O.apply_filter(fnu, lam * 1.7, fl, ft, "lam")
    assert 0 < abs(a / b - 1) < 5e-5
    with pytest.raises(ValueError):
        O.apply_filter(fnu, lam * 100.0, fl, ft, "nu")


def test_c_oracle_extensions_match_numpy_oracle():
    """The C restatement's two-screen attenuation and dust-emission energy balance (used to check 20 000-galaxy batches on
    the GPU) against the numpy oracle, galaxy by galaxy."""
    from oracle import adapter as A, c_oracle as CO, oracle as O
    from synference_b200 import igm as I
    from synference_b200.configs import make_workload
    w = make_workload("cfg2", 24)
    lam = np.asarray(w.grid.lam)
    filt = [(f.lam, f.t) for f in w.filters]
    p = w.params
    tau_b = np.random.default_rng(0).uniform(0, 3, len(p))
    gals = A.galaxies_from_params(p)
    for g, tb in zip(gals, tau_b):
        g["tau_v_birth"] = float(tb)
    ts = dict(age_pivot=7.0, dust_birth=dict(curve="Calzetti2000", slope=-0.7))
    de = dict(kind="Greybody", temperature=40.0, emissivity=1.5)
    want = O.synthesize(gals, w.grid.log10ages, w.grid.metallicity, lam, w.grid.spectra, filt, key="total", fesc_ly_alpha=0.4,
                        dust=dict(curve="Calzetti2000"), igm=(I.INOUE14_LAF, I.INOUE14_DLA), two_screens=ts, dust_emission=de)
    ga, gu = O.emission_parts(w.grid.spectra, lam, "total", 0.0, 0.4)
    got = CO.synthesize(p, w.grid.log10ages, w.grid.metallicity, lam, ga, gu, filt, kappa=O.dust_kappa(lam),
                        igm=(I.INOUE14_LAF, I.INOUE14_DLA),
                        two_screens=dict(age_pivot=7.0, kappa_birth=O.dust_kappa(lam, slope=-0.7), tau_v_birth=tau_b),
                        dust_shape=O.dust_emission_shape(lam, **de))
    np.testing.assert_allclose(got, want, rtol=1e-10)
    # single screen + emission
    want1 = O.synthesize(gals, w.grid.log10ages, w.grid.metallicity, lam, w.grid.spectra, filt, key="total", fesc=0.1, fesc_ly_alpha=0.5,
                         dust=dict(curve="Calzetti2000"), igm=(I.INOUE14_LAF, I.INOUE14_DLA), dust_emission=de)
    ga, gu = O.emission_parts(w.grid.spectra, lam, "total", 0.1, 0.5)
    got1 = CO.synthesize(p, w.grid.log10ages, w.grid.metallicity, lam, ga, gu, filt, kappa=O.dust_kappa(lam),
                         igm=(I.INOUE14_LAF, I.INOUE14_DLA), dust_shape=O.dust_emission_shape(lam, **de))
    np.testing.assert_allclose(got1, want1, rtol=1e-10)
