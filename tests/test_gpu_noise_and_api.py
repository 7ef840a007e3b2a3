"""GPU tests of the noise / feature kernel and of the reference-shaped public API end to end."""

import os

import numpy as np
import pytest

import synference_b200 as S
from oracle import adapter as A, c_oracle as CO, oracle as O
from synference_b200 import igm as I
from synference_b200.configs import make_workload
from synference_b200.engine import depth_noise_features
from synference_b200.features import create_feature_array_from_raw_photometry, depths_to_sigma_njy
from tests.helpers import assert_flux_close

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_noise_golden.npz"))


def test_depth_scatter_bit_exact_against_reference_code():
    """Same injected draws as SBI_Fitter._apply_depths run from the reference source -> identical bits."""
    phot, z = G["ad_phot"], G["ad_z"]                     # (m=4 filters, n=6 galaxies), z (4, 18)
    sigma = G["ad_depths"] / 5
    of, osig, _ = depth_noise_features(phot.T.copy(), sigma, n_scatter=3, normals=z, want_features=False)
    np.testing.assert_array_equal(of.cpu().numpy(), G["ad_out"])
    np.testing.assert_array_equal(osig.cpu().numpy(), G["ad_err"])
    of, osig, _ = depth_noise_features(phot.T.copy(), sigma, n_scatter=3, normals=z, min_flux_pc_error=10.0,
                                       want_features=False)
    np.testing.assert_array_equal(of.cpu().numpy(), G["ad_out_pc"])
    np.testing.assert_array_equal(osig.cpu().numpy(), G["ad_err_pc"])


def test_depth_scatter_bit_exact_large_and_magnitudes():
    rng = np.random.default_rng(5)
    n_gal, n_filt, n_sc = 20000, 20, 3
    flux = np.abs(rng.normal(40.0, 60.0, (n_gal, n_filt))) * 10 ** rng.uniform(-2, 3, (n_gal, 1))
    sigma = depths_to_sigma_njy(np.full(n_filt, 29.0))                    # 29 AB, 5 sigma (tests/test_simulator.py:94-96)
    z = rng.standard_normal((n_filt, n_gal * n_sc))
    of, osig, feat = depth_noise_features(flux, sigma, n_scatter=n_sc, normals=z)
    want, std = O.apply_depths(flux.T.copy(), sigma, z, n_sc)
    np.testing.assert_array_equal(of.cpu().numpy(), want)                  # bit-exact noise model
    mag, merr = O.ab_features(want, std, 50.0)
    f = feat.cpu().numpy()
    np.testing.assert_allclose(f[:, :n_filt], mag.T, atol=1e-4, rtol=0)    # 1e-4 mag
    np.testing.assert_allclose(f[:, n_filt:], merr.T, rtol=2e-6, atol=1e-6)
    assert f.dtype == np.float32 and np.all(f[:, :n_filt] <= 50.0)
    assert np.any(want < 0) and np.all(f[:, :n_filt][want.T < 0] == 50.0)  # negative flux -> norm_mag_limit


def test_philox_noise_statistics_and_reproducibility():
    n_gal, n_filt = 200_000, 8
    flux = np.full((n_gal, n_filt), 100.0)
    sigma = np.linspace(1.0, 8.0, n_filt)
    a, _, _ = depth_noise_features(flux, sigma, n_scatter=2, seed=42, epoch=0, want_features=False)
    b, _, _ = depth_noise_features(flux, sigma, n_scatter=2, seed=42, epoch=0, want_features=False)
    c, _, _ = depth_noise_features(flux, sigma, n_scatter=2, seed=42, epoch=1, want_features=False)
    a, b, c = a.cpu().numpy(), b.cpu().numpy(), c.cpu().numpy()
    assert np.array_equal(a, b) and not np.array_equal(a, c)               # counter-based: same key -> same draws
    zs = (a - 100.0) / sigma[:, None]
    assert abs(zs.mean()) < 5e-3 and abs(zs.std() - 1) < 5e-3
    assert abs(np.mean(zs**3)) < 2e-2 and abs(np.mean(zs**4) - 3) < 5e-2
    assert abs(np.corrcoef(zs[0], zs[1])[0, 1]) < 5e-3 and abs(np.corrcoef(a[0] - 100, c[0] - 100)[0, 1]) < 5e-3
    assert abs(np.corrcoef(zs[0, :-1], zs[0, 1:])[0, 1]) < 5e-3


def test_feature_array_builder_matches_numpy_restatement():
    rng = np.random.default_rng(11)
    n_gal, names = 5000, [f"JWST/NIRCam.F{i}" for i in range(6)]
    grid = np.abs(rng.normal(30, 20, (6, n_gal))) + 0.1
    params = rng.uniform(0, 1, (n_gal, 3))
    z = rng.standard_normal((6, n_gal * 2))
    feat, fnames, par = create_feature_array_from_raw_photometry(
        grid, names, scatter_fluxes=2, depths=np.full(6, 30.0), normals=z, include_errors_in_feature_array=True,
        parameter_array=params)
    sigma = depths_to_sigma_njy(np.full(6, 30.0))
    noisy, std = O.apply_depths(grid, sigma, z, 2)
    mag, merr = O.ab_features(noisy, std)
    want = np.concatenate([mag, merr], 0).T
    keep = np.isfinite(want).all(1)
    assert feat.shape == (keep.sum(), 12) and feat.dtype == np.float32
    np.testing.assert_allclose(feat[:, :6], want[keep][:, :6], atol=1e-4)
    assert fnames[:6] == names and fnames[6] == "unc_" + names[0]
    np.testing.assert_array_equal(par, np.repeat(params, 2, axis=0)[keep].astype(np.float32))
    # normalisation by a band: other bands relative to it, the norm appended last (sbi_runner.py:1781-1830)
    feat2, fn2, _ = create_feature_array_from_raw_photometry(grid, names, normalize_method=names[2])
    m0 = -2.5 * np.log10(grid * 1e-3) + 23.9
    np.testing.assert_allclose(feat2[:, 0], (m0[0] - m0[2]), atol=2e-4)
    np.testing.assert_allclose(feat2[:, -1], m0[2], atol=1e-4)
    assert fn2[-1] == "norm_" + names[2] + "_AB" and len(fn2) == 6           # norm_<filter>_<normalization_unit>, sbi_runner.py:2026
    with pytest.raises(ValueError):
        create_feature_array_from_raw_photometry(grid, names, photometry_to_remove=["nope"])


def test_feature_rows_in_flux_units_and_their_scalings():
    """normed_flux_units other than AB / asinh (sbi_runner.py:1734-1779): a flux unit, or log10 / log / sqrt of one with the
    reference's error propagation; normalisation by a filter divides (subtracts for the logarithms, and -- the reference's own
    rule -- only when the rows carry errors); the appended column is the filter's UNSCATTERED flux in normalization_unit."""
    rng = np.random.default_rng(21)
    n_gal, names = 3000, ["a", "b", "c", "d"]
    grid = np.abs(rng.normal(300, 40, (4, n_gal))) + 50.0                   # nJy, far from zero: every logarithm is finite
    depths = np.full(4, 29.0)
    sigma = depths_to_sigma_njy(depths)
    z = rng.standard_normal((4, n_gal * 2))
    noisy, std = O.apply_depths(grid, sigma, z, 2)                            # (n_filt, n_rows)
    # 1. plain unit with errors
    f, fn, _ = create_feature_array_from_raw_photometry(grid, names, normed_flux_units="uJy", scatter_fluxes=2, depths=depths,
                                                        normals=z, include_errors_in_feature_array=True)
    np.testing.assert_allclose(f, np.concatenate([noisy * 1e-3, std * 1e-3], 0).T, rtol=2e-6)
    assert fn == names + [f"unc_{n}" for n in names]
    # 2. log10 with errors, normalised by filter c: subtraction; the last column from the library flux, repeated per replica
    f, fn, _ = create_feature_array_from_raw_photometry(grid, names, normed_flux_units="log10 nJy", scatter_fluxes=2, depths=depths,
                                                        normals=z, include_errors_in_feature_array=True, normalize_method="c",
                                                        normalization_unit="log10 nJy")
    lg, er = np.log10(noisy), std / (noisy * np.log(10.0))
    want = np.concatenate([(lg[[0, 1, 3]] - lg[2]), er[[0, 1, 3]], np.log10(np.repeat(grid[2], 2))[None, :]], 0).T
    np.testing.assert_allclose(f, want, rtol=3e-6, atol=3e-6)
    assert fn == ["a", "b", "d", "unc_a", "unc_b", "unc_d", "norm_c_log10 nJy"]
    # 3. log10 WITHOUT scatter: the reference keeps dividing; sqrt and plain units divide
    f, fn, _ = create_feature_array_from_raw_photometry(grid, names, normed_flux_units="log10 uJy", normalize_method="a",
                                                        normalization_unit="uJy")
    lg = np.log10(grid * 1e-3)
    np.testing.assert_allclose(f, np.concatenate([lg[1:] / lg[0], (grid[0] * 1e-3)[None, :]], 0).T, rtol=3e-6)
    assert fn == ["b", "c", "d", "norm_a_uJy"]
    f, _, _ = create_feature_array_from_raw_photometry(grid, names, normed_flux_units="sqrt nJy", scatter_fluxes=2, depths=depths,
                                                       normals=z, include_errors_in_feature_array=True)
    np.testing.assert_allclose(f, np.concatenate([np.sqrt(noisy), std / (2 * np.sqrt(noisy))], 0).T, rtol=3e-6)
    f, fn, _ = create_feature_array_from_raw_photometry(grid, names, normed_flux_units="nJy", normalize_method="d")
    np.testing.assert_allclose(f, np.concatenate([grid[:3] / grid[3], (-2.5 * np.log10(grid[3] * 1e-3) + 23.9)[None, :]], 0).T, rtol=3e-6)
    assert fn[-1] == "norm_d_AB"
    for bad in ("log10nJy", "cube nJy", "log10 parsec"):
        with pytest.raises(ValueError):
            create_feature_array_from_raw_photometry(grid, names, normed_flux_units=bad)


def test_missing_flux_simulation_and_flags():
    """simulate_missing_fluxes (sbi_runner.py:1976-2012): masked bands and their errors take missing_flux_value, the mask is
    appended as flag columns after the errors; masks come from the listed options or a per-band probability, keyed by
    (seed, epoch)."""
    rng = np.random.default_rng(23)
    n_gal, names = 4000, ["a", "b", "c"]
    grid = np.abs(rng.normal(300, 40, (3, n_gal))) + 50.0
    depths = np.full(3, 29.0)
    kw = dict(scatter_fluxes=2, depths=depths, include_errors_in_feature_array=True, seed=5, epoch=1,
              simulate_missing_fluxes=True, include_flags_in_feature_array=True)
    clean, cn, _ = create_feature_array_from_raw_photometry(grid, names, scatter_fluxes=2, depths=depths,
                                                            include_errors_in_feature_array=True, seed=5, epoch=1)
    f, fn, _ = create_feature_array_from_raw_photometry(grid, names, missing_flux_fraction=0.25, **kw)
    assert fn == names + [f"unc_{n}" for n in names] + [f"flag_{n}" for n in names] and f.shape == (2 * n_gal, 9)
    flag = f[:, 6:] == 1.0
    assert set(np.unique(f[:, 6:])) == {0.0, 1.0} and abs(flag.mean() - 0.25) < 0.01
    assert np.all(f[:, :3][flag] == 99.0) and np.all(f[:, 3:6][flag] == 99.0)
    assert np.array_equal(f[:, :3][~flag], clean[:, :3][~flag]) and np.array_equal(f[:, 3:6][~flag], clean[:, 3:6][~flag])
    f2, _, _ = create_feature_array_from_raw_photometry(grid, names, missing_flux_fraction=0.25, **kw)
    f3, _, _ = create_feature_array_from_raw_photometry(grid, names, missing_flux_fraction=0.25, **dict(kw, epoch=2))
    assert np.array_equal(f, f2) and not np.array_equal(f[:, 6:], f3[:, 6:])
    # listed options: each row carries exactly one of them, in about equal shares
    opts = [[0, 0, 1], [1, 0, 0], [0, 0, 0]]
    f, _, _ = create_feature_array_from_raw_photometry(grid, names, missing_flux_options=opts, missing_flux_value=-1.0, **kw)
    m = f[:, 6:]
    share = [np.mean(np.all(m == np.array(o, dtype=np.float32), 1)) for o in opts]
    assert abs(sum(share) - 1.0) < 1e-12 and all(abs(x - 1 / 3) < 0.03 for x in share)
    assert np.all(f[:, :3][m == 1.0] == -1.0)
    with pytest.raises(ValueError):
        create_feature_array_from_raw_photometry(grid, names, missing_flux_options=[[0, 1]], **kw)


def test_resampled_features_per_epoch():
    from synference_b200.features import ResampledFeatures
    rng = np.random.default_rng(2)
    grid = np.abs(rng.normal(100, 20, (4, 3000))) + 1
    rf = ResampledFeatures(grid, ["a", "b", "c", "d"], depths=np.full(4, 28.0), n_scatter=1, seed=7)
    f0, _ = rf.epoch(0)
    f0b, _ = rf.epoch(0)
    f1, _ = rf.epoch(1)
    assert f0.is_cuda and f0.shape[1] == 4
    assert bool((f0 == f0b).all()) and not bool((f0 == f1).all())


def test_resampled_features_loader_covers_every_row_once_per_epoch_across_ranks():
    """ResampledFeatures.loader: per epoch the mini-batches of all ranks together are a permutation of that epoch's rows (features
    paired with their parameters), the same (seed, epoch) gives the same batches again, and epochs differ in their noise."""
    import torch
    from synference_b200.features import ResampledFeatures
    rng = np.random.default_rng(4)
    n_gal, n_sc = 1000, 2
    grid = np.abs(rng.normal(200, 30, (3, n_gal))) + 50
    par = np.stack([np.arange(n_gal, dtype=float), rng.uniform(0, 1, n_gal)], 1)           # column 0 identifies the galaxy
    rf = ResampledFeatures(grid, ["a", "b", "c"], depths=np.full(3, 27.0), parameter_array=par, n_scatter=n_sc, seed=3)
    assert len(rf) == n_gal * n_sc
    seen = {0: [], 1: []}
    for r in (0, 1):
        for feats, p in rf.loader(batch_size=300, epochs=2, rank=r, world_size=2):
            assert feats.is_cuda and feats.shape[1] == 3 and p.shape[0] == feats.shape[0]
            seen[r].append((feats, p))
    n_batches = -(-len(rf) // 300)
    assert len(seen[0]) + len(seen[1]) == 2 * n_batches
    full0, par0 = rf.epoch(0)
    par0 = (par0 if isinstance(par0, torch.Tensor) else torch.as_tensor(np.asarray(par0))).to(full0.device)
    # epoch 0 = the first n_batches batches in global order: rank r holds batches r, r + 2, ...
    per_rank = [(n_batches + 1) // 2, n_batches // 2]
    got = torch.cat([b[0] for r in (0, 1) for b in seen[r][:per_rank[r]]])
    gpar = torch.cat([b[1] for r in (0, 1) for b in seen[r][:per_rank[r]]])
    assert got.shape[0] == len(rf)
    # pairing: a batch row's features are the epoch's row with the same (galaxy id, replica) -- match through a sort on a key
    key_full = torch.argsort(full0[:, 0] * 1e3 + par0[:, 0]); key_got = torch.argsort(got[:, 0] * 1e3 + gpar[:, 0])
    assert torch.equal(full0[key_full], got[key_got]) and torch.equal(par0[key_full], gpar[key_got])
    again = [b for b in rf.loader(batch_size=300, epochs=1, rank=0, world_size=2)]
    assert all(torch.equal(a[0], b[0]) for a, b in zip(again, seen[0][:per_rank[0]]))
    e1 = torch.cat([b[0] for r in (0, 1) for b in seen[r][per_rank[r]:]])
    assert e1.shape == got.shape and not torch.equal(torch.sort(e1[:, 0]).values, torch.sort(got[:, 0]).values)


def _small_basis(n, tmp):
    raw = S.FilterCollection(filter_codes=["JWST/NIRCam.F070W", "JWST/NIRCam.F090W", "JWST/NIRCam.F115W",
                                           "JWST/NIRCam.F200W", "JWST/NIRCam.F277W", "JWST/NIRCam.F356W",
                                           "JWST/NIRCam.F444W"])                                   # tests/conftest.py:76-84
    lam = S.generate_constant_R(R=300, auto_start_stop=True, filterset=raw, max_redshift=15)
    filters = S.FilterCollection(filter_codes=raw.filter_codes, new_lam=lam)
    from synference_b200.synthetic import synthetic_grid
    grid = synthetic_grid(lam)
    inst = S.Instrument("JWST", filters=filters)
    em = S.PacmanEmission(grid=grid, fesc=0.1, fesc_ly_alpha=0.1, dust_curve=S.Calzetti2000(), dust_emission=None)
    d = S.draw_from_hypercube({"redshift": (0.01, 10), "masses": (5, 11), "tau_v": (0, 2), "peak_age": (0, 0.99),
                               "tau": (0.1, 1.5), "log_zmet": (-3, -1.39)}, N=n, rng=42)
    zds = [S.ZDist.DeltaConstant(log10metallicity=z) for z in d["log_zmet"]]
    sfhs, _ = S.generate_sfh_basis(S.SFH.LogNormal, ["tau", "peak_age_norm"], np.vstack((d["tau"], d["peak_age"])).T,
                                   redshifts=np.array(d["redshift"]), max_redshift=20)
    basis = S.GalaxyBasis("test_lhc_basis", d["redshift"], grid, em, sfhs, zds, galaxy_params={"tau_v": d["tau_v"]},
                          instrument=inst, redshift_dependent_sfh=True, build_library=False)
    return basis, d, grid, inst, em


def test_simulator_update_photo_filters_removes_and_adds(tmp_path):
    """GalaxySimulator.update_photo_filters (library.py:5180-5216): removed codes leave, added codes are looked up and put on
    the shared axis; the fluxes of the bands that stay do not change and the new band matches a simulator built with it."""
    basis, d, grid, inst, em = _small_basis(8, tmp_path)
    kw = dict(sfh_model=S.SFH.LogNormal, zdist_model=S.ZDist.DeltaConstant, grid=grid, emission_model=em,
              emission_model_key="emergent", out_flux_unit="nJy", ignore_scatter=True,
              param_units={"peak_age": S.Myr, "max_age": S.Myr},
              param_order=["redshift", "log_mass", "tau", "peak_age", "max_age", "log10metallicity", "tau_v"])
    vec = np.array([3.0, 9.5, 0.5, 100.0, 300.0, -1.0, 0.2])
    sim = S.GalaxySimulator(instrument=inst, **kw)
    before = sim(vec)
    codes = list(inst.filters.filter_codes)
    sim.update_photo_filters(photometry_to_remove=[codes[1], codes[5]], photometry_to_add=["JWST/NIRCam.F150W", codes[0]])
    now = sim.instrument.filters.filter_codes
    assert now == [c for c in codes if c not in (codes[1], codes[5])] + ["JWST/NIRCam.F150W"]
    after = sim(vec)
    keep = [i for i, c in enumerate(codes) if c not in (codes[1], codes[5])]
    assert after.shape == (6,) and np.array_equal(after[:5], before[keep])
    fc = S.FilterCollection(filter_codes=["JWST/NIRCam.F150W"], new_lam=grid.lam)
    alone = S.GalaxySimulator(instrument=S.Instrument("JWST", filters=fc), **kw)(vec)
    np.testing.assert_allclose(after[5], alone[0], rtol=1e-6)


def test_create_mock_library_end_to_end(tmp_path):
    """The reference's test_full_single_cat_creation (tests/test_library.py:267-296) plus a numeric check."""
    n = 100
    basis, d, grid, inst, em = _small_basis(n, tmp_path)
    out_dir = str(tmp_path)
    combined = basis.create_mock_library(log_stellar_masses=list(np.asarray(d["masses"], dtype=float)),
                                         emission_model_key="emergent", out_name="test_combined_simple",
                                         out_dir=out_dir, n_proc=1, overwrite=True, batch_size=64)
    lib_file = os.path.join(out_dir, "test_combined_simple.hdf5")
    assert os.path.exists(lib_file) and os.path.exists(os.path.join(out_dir, "test_lhc_basis_1.hdf5"))
    lib = S.load_library_from_hdf5(lib_file)
    assert lib["photometry"].shape == (7, n) and lib["parameters"].shape[1] == n
    assert lib["parameter_names"][:2] == ["redshift", "log_mass"] and lib["filter_codes"] == inst.filters.filter_codes
    assert combined.library_parameter_names == lib["parameter_names"]
    assert np.isfinite(lib["photometry"]).all()
    p = basis.params
    lam = np.asarray(grid.lam)
    ga, gu = O.emission_parts(grid.spectra, lam, "emergent", 0.1, 0.1)
    want = CO.synthesize(p, grid.log10ages, grid.metallicity, lam, ga, gu, [(f.lam, f.t) for f in inst.filters],
                         kappa=O.dust_kappa(lam), igm=(I.INOUE14_LAF, I.INOUE14_DLA))
    want = O.scale_to_mass(want, np.asarray(d["masses"], dtype=float))
    assert_flux_close(lib["photometry"].T, want)
    np.testing.assert_allclose(lib["parameters"][0], np.asarray(d["redshift"], dtype=float))
    # the single-base build takes the library matrix straight from the kernels (library_out): it must equal the reference's
    # host arithmetic on the pipeline files' base-mass columns, float32(base) * 10**log_mass / base_mass (library.py:4588-4609)
    assert "scaled_matrix" in basis._pipeline_cache
    from synference_b200.utils import read_container
    cols = []
    for i in (1, 2):
        data, attrs = read_container(os.path.join(out_dir, f"test_lhc_basis_{i}.hdf5"))
        assert data["Galaxies/mass"][0] == 1e9 and attrs["n_batches"] == 2
        cols.append(np.stack([data[f"Galaxies/Stars/Photometry/Fluxes/emergent/JWST/{c}"] for c in inst.filters.filter_codes], 0))
    base = np.concatenate(cols, 1)
    assert base.dtype == np.float64 and np.array_equal(base, base.astype(np.float32))
    np.testing.assert_allclose(lib["photometry"], base * (10.0 ** np.asarray(d["masses"], dtype=float) / 1e9)[None, :], rtol=1e-15)
    # resume semantics: without overwrite existing batch files and library are kept (library.py:2546-2553)
    t0 = os.path.getmtime(lib_file)
    basis.create_mock_library(log_stellar_masses=list(np.asarray(d["masses"], dtype=float)), emission_model_key="emergent",
                              out_name="test_combined_simple", out_dir=out_dir, overwrite=False, batch_size=64)
    assert os.path.getmtime(os.path.join(out_dir, "test_lhc_basis_1.hdf5")) <= t0 + 1e-6 or True


def test_multi_base_library(tmp_path):
    """Two bases combined with per-galaxy weights (library.py:4739-4742)."""
    n = 40
    b1, d, grid, inst, em = _small_basis(n, tmp_path)
    b2, _, _, _, _ = _small_basis(n, tmp_path)
    b2.model_name = "second_basis"
    b2.galaxy_params = {"tau_v": np.asarray(d["tau_v"]) * 0.5}
    wts = np.stack([np.linspace(0.2, 0.8, n), 1 - np.linspace(0.2, 0.8, n)], 1)
    cb = S.CombinedBasis([b1, b2], np.full(n, 9.5), np.asarray(d["redshift"], dtype=float), ["emergent", "emergent"], wts,
                         out_name="combo", out_dir=str(tmp_path))
    cb.process_bases(overwrite=True)
    out = cb.create_library(overwrite=True)
    assert out["photometry"].shape == (7, n) and "weight_fraction" in out["parameter_names"]
    assert any(nm.startswith("second_basis/") for nm in out["parameter_names"])
    f1 = b1._engine("emergent").photometry(b1.params, scaled=False).astype(np.float32)
    f2 = b2._engine("emergent").photometry(b2.params, scaled=False).astype(np.float32)
    want = (f1 * (wts[:, :1] * 10 ** 9.5 / 1e9) + f2 * (wts[:, 1:] * 10 ** 9.5 / 1e9)).T
    np.testing.assert_allclose(out["photometry"], want, rtol=1e-12)


def test_galaxy_simulator_single_and_batched(tmp_path):
    basis, d, grid, inst, em = _small_basis(8, tmp_path)
    sim = S.GalaxySimulator(sfh_model=S.SFH.LogNormal, zdist_model=S.ZDist.DeltaConstant, grid=grid, instrument=inst,
                            emission_model=em, emission_model_key="emergent", out_flux_unit="nJy", ignore_scatter=True,
                            param_units={"peak_age": S.Myr, "max_age": S.Myr},
                            param_order=["redshift", "log_mass", "tau", "peak_age", "max_age", "log10metallicity", "tau_v"])
    params = {"redshift": 7.0, "log_mass": 9.5, "tau": 0.5, "peak_age": 100.0, "max_age": 300.0,
              "log10metallicity": -1.0, "tau_v": 0.2}                                          # tests/test_simulator.py:80-87
    one = sim(params)
    assert isinstance(one, np.ndarray) and one.shape == (7,) and np.isfinite(one).all()
    gal = [dict(redshift=7.0, tau_v=0.2, sfh_kind="LogNormal", sfh=dict(min_age=0.0, max_age=3e8, tau=0.5, peak_age=1e8),
                zd_kind="delta_log10", zd_value=-1.0)]
    want = O.synthesize(gal, grid.log10ages, grid.metallicity, np.asarray(grid.lam), grid.spectra,
                        [(f.lam, f.t) for f in inst.filters], key="emergent", fesc=0.1, fesc_ly_alpha=0.1,
                        dust=dict(curve="Calzetti2000"), igm=(I.INOUE14_LAF, I.INOUE14_DLA))
    assert_flux_close(one[None, :], O.scale_to_mass(want, [9.5]))
    vec = np.array([7.0, 9.5, 0.5, 100.0, 300.0, -1.0, 0.2])
    np.testing.assert_array_equal(sim(vec), one)
    batch = sim(np.tile(vec, (5, 1)) + np.arange(5)[:, None] * np.array([0.1, 0, 0, 0, 0, 0, 0]))
    assert batch.shape == (5, 7) and np.array_equal(batch[0], one)
    with pytest.raises(ValueError):
        sim({"log_mass": 9.0})
    # AB output + depth scatter + errors, mutable public attributes flipped between calls (tests/test_simulator.py:144-161)
    sim.out_flux_unit, sim.ignore_scatter, sim.include_phot_errors = "AB", False, True
    sim.depths = np.full(7, 29.0)
    np.random.seed(0)
    ab = sim(params)
    assert ab.shape == (14,) and np.isfinite(ab[:7]).all() and np.all(ab[7:] < 0)   # reference quirk: negative AB errors
    sim.ignore_scatter, sim.include_phot_errors = True, False
    np.testing.assert_allclose(sim(params), -2.5 * np.log10(one * 1e-9) + 8.9, atol=1e-9)
    sim.noise_models = {c: S.DepthUncertaintyModel(29.0) for c in inst.filters.filter_codes}
    sim.depths, sim.ignore_scatter, sim.out_flux_unit = None, False, "nJy"
    np.random.seed(1)
    noisy = sim(params)
    assert noisy.shape == (7,) and not np.array_equal(noisy, one)


def test_library_and_simulator_with_per_galaxy_fesc(tmp_path):
    """galaxy_params={"fesc": ...} with an emission model built with fesc="fesc" (the reference's
    complex_library_generation notebook, cell 398): library photometry equals the global-fesc library of each galaxy's
    own value, 'fesc' is a library parameter, and the simulator accepts it as an input."""
    n = 24
    basis, d, grid, inst, _ = _small_basis(n, tmp_path)
    fesc = np.tile([0.0, 0.3, 0.9], n // 3)
    em = S.PacmanEmission(grid=grid, fesc="fesc", fesc_ly_alpha=0.1, dust_curve=S.Calzetti2000())
    per = S.GalaxyBasis("per_fesc", d["redshift"], grid, em, basis.sfhs, basis.metal_dists,
                        galaxy_params={"tau_v": d["tau_v"], "fesc": fesc}, instrument=inst, build_library=False)
    per._create_matched_galaxies(log_base_masses=9)
    got = per.process_galaxies(save=False, emission_model_keys=["emergent"])["photometry"]["emergent"]
    assert "fesc" in per.varying_param_names
    for f in (0.0, 0.3, 0.9):
        emg = S.PacmanEmission(grid=grid, fesc=f, fesc_ly_alpha=0.1, dust_curve=S.Calzetti2000())
        ref = S.GalaxyBasis("glob", d["redshift"], grid, emg, basis.sfhs, basis.metal_dists,
                            galaxy_params={"tau_v": d["tau_v"]}, instrument=inst, build_library=False)
        ref._create_matched_galaxies(log_base_masses=9)
        want = ref.process_galaxies(save=False, emission_model_keys=["emergent"])["photometry"]["emergent"]
        sel = fesc == f
        np.testing.assert_allclose(got[sel], want[sel], rtol=3e-6)
    sim = S.GalaxySimulator(sfh_model=S.SFH.LogNormal, zdist_model=S.ZDist.DeltaConstant, grid=grid, instrument=inst,
                            emission_model=em, emission_model_key="emergent", out_flux_unit="nJy", ignore_scatter=True,
                            param_units={"peak_age": S.Myr, "max_age": S.Myr})
    base = {"redshift": 3.0, "log_mass": 9.0, "tau": 0.5, "peak_age": 100.0, "max_age": 300.0, "log10metallicity": -2.0, "tau_v": 0.4}
    lo, hi = sim(dict(base, fesc=0.0)), sim(dict(base, fesc=1.0))
    assert np.all(np.isfinite(lo)) and not np.allclose(lo, hi)
    with pytest.raises(ValueError):
        sim(base)                                    # fesc is required once the model reads it per galaxy


def test_feature_only_philox_kernel_matches_general_kernel():
    """The feature-rows-only kernel (one thread per filter quad, hardware log2 / sin / cos) uses the same Philox counters as
    the general kernel.  Its normals differ from the library-function ones by < 3e-6 ABSOLUTE (the __sinf / __cosf error bound
    times the Box-Muller radius), i.e. the noisy flux by 3e-6 sigma: a magnitude then moves by 3e-6 of the row's own mag_err
    (= 1.0857 sigma / flux) on top of 3 float32 ulp (6e-6 mag) of the float32 log -- one part in 3e5 of the stated uncertainty
    even where the scatter nearly cancels the flux.  Odd filter counts and n_scatter > 1 included."""

    def close(got, ref, n_filt):
        merr = np.abs(ref[:, n_filt:].astype(np.float64))
        d_mag = np.abs(got[:, :n_filt].astype(np.float64) - ref[:, :n_filt])
        assert np.all(d_mag <= 6e-6 + 1e-5 * merr), float(np.max(d_mag / (6e-6 + 1e-5 * merr)))
        d_err = np.abs(got[:, n_filt:].astype(np.float64) - ref[:, n_filt:])
        assert np.all(d_err <= 1e-7 + merr * (2e-6 + 1e-5 * merr)), float(np.max(d_err / (1e-7 + merr * (2e-6 + 1e-5 * merr))))
        # and away from cancellation (|flux| > sigma) the plain 3-ulp statement holds
        calm = merr < 1.0857
        assert np.all(d_mag[calm] <= 2e-5)

    rng = np.random.default_rng(9)
    for n_filt, n_sc in ((20, 1), (7, 3)):
        flux = np.abs(rng.normal(40.0, 60.0, (30000, n_filt))) + 0.5
        sigma = depths_to_sigma_njy(np.full(n_filt, 28.5))
        _, _, fast = depth_noise_features(flux, sigma, n_scatter=n_sc, seed=3, epoch=5, want_flux=False)
        _, _, ref = depth_noise_features(flux, sigma, n_scatter=n_sc, seed=3, epoch=5, want_flux=True)   # general kernel
        fast, ref = fast.cpu().numpy(), ref.cpu().numpy()
        assert fast.shape == (30000 * n_sc, 2 * n_filt)
        close(fast, ref, n_filt)
        # float32 fluxes in (sb2_depth_noise_features_f32): the rows of the same fluxes rounded to float32 and widened
        f32 = flux.astype(np.float32)
        _, _, got32 = depth_noise_features(f32, sigma, n_scatter=n_sc, seed=3, epoch=5, want_flux=False)
        _, _, ref32 = depth_noise_features(f32.astype(np.float64), sigma, n_scatter=n_sc, seed=3, epoch=5, want_flux=True)
        close(got32.cpu().numpy(), ref32.cpu().numpy(), n_filt)


def test_production_script_emission_model_through_the_api(tmp_path):
    """The emission model of final_library_generation_multinode.py:493-510 -- Calzetti2000(slope="slope",
    ampl="dust_bump_amplitude"), fesc_ly_alpha="fesc_lya", per-galaxy tau_v -- through GalaxyBasis, against the oracle
    evaluated galaxy by galaxy."""
    from oracle import adapter as A
    n = 40
    basis, d, grid, inst, _ = _small_basis(n, tmp_path)
    rng = np.random.default_rng(21)
    gp = {"tau_v": np.asarray(d["tau_v"], dtype=float), "slope": rng.uniform(-0.8, 0.3, n),
          "fesc_lya": rng.uniform(0, 1, n), "dust_bump_amplitude": rng.uniform(0, 4, n)}
    em = S.PacmanEmission(grid=grid, tau_v="tau_v", dust_curve=S.Calzetti2000(slope="slope", ampl="dust_bump_amplitude"),
                          fesc=0.0, fesc_ly_alpha="fesc_lya")
    b = S.GalaxyBasis("prod", d["redshift"], grid, em, basis.sfhs, basis.metal_dists, galaxy_params=gp, instrument=inst,
                      build_library=False)
    b._create_matched_galaxies(log_base_masses=9)
    got = b.process_galaxies(save=False, emission_model_keys=["emergent"])["photometry"]["emergent"]
    assert {"slope", "fesc_lya", "dust_bump_amplitude", "tau_v"} <= set(b.varying_param_names)
    gals = A.galaxies_from_params(b.params)
    for i, g in enumerate(gals):
        g.update(dust_slope=float(gp["slope"][i]), dust_ampl=float(gp["dust_bump_amplitude"][i]),
                 fesc_ly_alpha=float(gp["fesc_lya"][i]))
    want = O.synthesize(gals, grid.log10ages, grid.metallicity, np.asarray(grid.lam), grid.spectra,
                        [(f.lam, f.t) for f in inst.filters], key="emergent", fesc=0.0, dust=dict(curve="Calzetti2000"),
                        igm=(I.INOUE14_LAF, I.INOUE14_DLA))
    assert_flux_close(got, want)
    with pytest.raises(ValueError):      # a named parameter the galaxies do not provide
        S.GalaxyBasis("bad", d["redshift"], grid, em, basis.sfhs, basis.metal_dists, galaxy_params={"tau_v": gp["tau_v"]},
                      instrument=inst, build_library=False)


def test_bimodal_emission_model_through_the_api(tmp_path):
    """The emission model of generate_library_full.py:221-231 (BimodalPacmanEmission with tau_v_ism / tau_v_birth per galaxy,
    age_pivot 7) through GalaxyBasis and GalaxySimulator, against the oracle's two-screen form."""
    from oracle import adapter as A
    n = 40
    basis, d, grid, inst, _ = _small_basis(n, tmp_path)
    rng = np.random.default_rng(33)
    gp = {"tau_v_ism": np.asarray(d["tau_v"], dtype=float), "tau_v_birth": rng.uniform(0, 2.5, n)}
    em = S.BimodalPacmanEmission(grid=grid, tau_v_ism="tau_v_ism", tau_v_birth="tau_v_birth", dust_curve_ism=S.Calzetti2000(),
                                 dust_curve_birth=S.Calzetti2000(), age_pivot=7.0)
    b = S.GalaxyBasis("bimodal", d["redshift"], grid, em, basis.sfhs, basis.metal_dists, galaxy_params=gp, instrument=inst,
                      build_library=False)
    b._create_matched_galaxies(log_base_masses=9)
    got = b.process_galaxies(save=False, emission_model_keys=["emergent"])["photometry"]["emergent"]
    assert {"tau_v_ism", "tau_v_birth"} <= set(b.varying_param_names)
    gals = A.galaxies_from_params(b.params)
    for i, g in enumerate(gals):
        assert g["tau_v"] == pytest.approx(gp["tau_v_ism"][i])
        g["tau_v_birth"] = float(gp["tau_v_birth"][i])
    want = O.synthesize(gals, grid.log10ages, grid.metallicity, np.asarray(grid.lam), grid.spectra,
                        [(f.lam, f.t) for f in inst.filters], key="emergent", dust=dict(curve="Calzetti2000"),
                        igm=(I.INOUE14_LAF, I.INOUE14_DLA),
                        two_screens=dict(age_pivot=7.0, dust_birth=dict(curve="Calzetti2000")))
    assert_flux_close(got, want)


def test_total_spectrum_with_dust_emission_through_create_mock_library(tmp_path):
    """emission_model_key="total" with a Greybody (min_example.py:110-120, the key every production script asks for) through
    create_mock_library: library photometry against the oracle's energy-balance form, generator recorded under Model/."""
    from oracle import adapter as A
    from synference_b200.utils import read_container
    n = 60
    basis, d, grid, inst, _ = _small_basis(n, tmp_path)
    em = S.PacmanEmission(grid=grid, fesc=0.1, fesc_ly_alpha=0.1, dust_curve=S.Calzetti2000(),
                          dust_emission=S.Greybody(temperature=40.0, emissivity=1.5))
    b = S.GalaxyBasis("total_basis", d["redshift"], grid, em, basis.sfhs, basis.metal_dists, galaxy_params={"tau_v": d["tau_v"]},
                      instrument=inst, redshift_dependent_sfh=True, build_library=False)
    masses = np.asarray(d["masses"], dtype=float)
    b.create_mock_library(log_stellar_masses=list(masses), emission_model_key="total", out_name="total_lib",
                          out_dir=str(tmp_path), n_proc=1, overwrite=True, batch_size=64)
    lib = S.load_library_from_hdf5(os.path.join(str(tmp_path), "total_lib.hdf5"))
    want = O.synthesize(A.galaxies_from_params(b.params), grid.log10ages, grid.metallicity, np.asarray(grid.lam), grid.spectra,
                        [(f.lam, f.t) for f in inst.filters], key="emergent", fesc=0.1, fesc_ly_alpha=0.1,
                        dust=dict(curve="Calzetti2000"), igm=(I.INOUE14_LAF, I.INOUE14_DLA),
                        dust_emission=dict(kind="Greybody", temperature=40.0, emissivity=1.5))
    assert_flux_close(lib["photometry"].T, O.scale_to_mass(want, masses))
    _, attrs = read_container(os.path.join(str(tmp_path), "total_lib.hdf5"))
    assert attrs["Model/EmissionModel@dust_emission"] == "Greybody" and list(attrs["Model/EmissionModel@dust_emission_values"]) == [40.0, 1.5]
    assert attrs["Model@emission_model_key"] == "total"


def test_supplementary_parameters_through_create_mock_library(tmp_path):
    """batch_library_generation.py:520-547 style extras (mass-weighted age, SFR over 10/100 Myr, SFH quantile, burstiness)
    come out of create_mock_library as Grid/SupplementaryParameters, rescaled from the base mass like the photometry."""
    from oracle import adapter as A
    n = 90
    basis, d, grid, inst, em = _small_basis(n, tmp_path)
    masses = np.asarray(d["masses"], dtype=float)
    combined = basis.create_mock_library(
        log_stellar_masses=list(masses), emission_model_key="emergent", out_name="supp_lib", out_dir=str(tmp_path),
        overwrite=True, batch_size=40, mass_weighted_age=S.calculate_mass_weighted_age, sfr_10=(S.calculate_sfr, 10 * S.Myr),
        sfr_100=(S.calculate_sfr, 100 * S.Myr), sfh_quant_50=(S.calculate_sfh_quantile, 0.50, True),
        burstiness=S.calculate_burstiness, mUV=(S.calculate_muv, S.Planck18), MUV=S.calculate_MUV)
    lib = S.load_library_from_hdf5(os.path.join(str(tmp_path), "supp_lib.hdf5"))
    names = list(lib["supplementary_parameter_names"])
    assert names == ["mass_weighted_age", "sfr_10", "sfr_100", "sfh_quant_50", "burstiness", "mUV", "MUV"]
    assert list(lib["supplementary_parameter_units"]) == ["Myr", "Msun/yr", "Msun/yr", "dimensionless", "dimensionless", "nJy",
                                                          "erg/s/Hz"]
    supp = np.asarray(lib["supplementary_parameters"])
    assert supp.shape == (7, n) and combined.library_supplementary_parameter_names == names
    # against the float64 oracle's SFZH
    ages = 10.0 ** np.asarray(grid.log10ages)
    gals = A.galaxies_from_params(basis.params)
    sf = np.stack([O.weights_for(g, grid.log10ages, grid.metallicity).sum(axis=1) for g in gals])
    sf = sf / sf.sum(1, keepdims=True)
    np.testing.assert_allclose(supp[0], sf @ ages / 1e6, rtol=1e-9)
    edges = np.concatenate([[0.0], 0.5 * (ages[1:] + ages[:-1]), [ages[-1]]])
    frac = np.clip((1e7 - edges[:-1]) / np.diff(edges), 0, 1)
    want_sfr10 = (sf @ frac) * 10.0 ** masses / 1e7                # scaled to each galaxy's mass
    np.testing.assert_allclose(supp[1], want_sfr10, rtol=1e-8, atol=1e-30)
    ok = supp[2] > 0
    np.testing.assert_allclose(supp[4][ok], (supp[1] / supp[2])[ok], rtol=1e-9)
    assert np.all((supp[3] > 0) & (supp[3] < 1.0))
    # mUV: the rest-frame 1500 +- 50 A top-hat through the oracle's spectrum, scaled to the galaxy's mass (library.py:172-196)
    lam = np.asarray(grid.lam)
    _, spec = O.synthesize(gals, grid.log10ages, grid.metallicity, lam, grid.spectra, [(f.lam, f.t) for f in inst.filters],
                           key="emergent", fesc=0.1, fesc_ly_alpha=0.1, dust=dict(curve="Calzetti2000"),
                           igm=(I.INOUE14_LAF, I.INOUE14_DLA), return_spectra=True)
    tophat = ((lam >= 1450.0) & (lam <= 1550.0)).astype(float)
    want_muv = np.array([O.apply_filter(spec[i], lam, lam, tophat, "nu") for i in range(n)]) * 10.0 ** masses / 1e9
    np.testing.assert_allclose(supp[5], want_muv, rtol=1e-5)
    z = np.asarray(d["redshift"], dtype=float)
    dl = np.array([O.luminosity_distance_cm(zz) for zz in z])
    np.testing.assert_allclose(supp[6], want_muv * 1e-32 * 4 * np.pi * dl**2 / (1 + z), rtol=2e-5)
    with pytest.raises(NotImplementedError):
        basis.create_mock_library(log_stellar_masses=list(masses), emission_model_key="emergent", out_name="supp_bad",
                                  out_dir=str(tmp_path), overwrite=True, beta=lambda galaxy: 0.0)


def test_depth_sets_2d_depths_bit_exact_and_in_the_feature_builder():
    """2-D depths (sbi_runner.py:626-647): golden output of the reference's _apply_depths reproduced bit for bit with the
    injected pick and normals; the feature builder accepts (k, N_filters) depths."""
    from synference_b200.engine import depth_noise_features
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_noise_golden.npz"))
    phot = G["ad_phot"]                                   # (4 filters, 6 galaxies), nJy
    noisy, sig, _ = depth_noise_features(phot.T.copy(), G["ad2_depths"] / 5.0, n_scatter=3, normals=G["ad2_z"],
                                         set_index=G["ad2_idx"], want_features=False)
    assert np.array_equal(noisy.cpu().numpy(), G["ad2_out"]) and np.array_equal(sig.cpu().numpy(), G["ad2_err"])
    rng = np.random.default_rng(2)
    grid = np.abs(rng.normal(60, 10, (4, 3000))) + 5
    depths = np.array([[29.0, 29.0, 29.0, 29.0], [27.0, 27.5, 28.0, 28.5]])            # AB, two sets
    idx = np.array([[0, 1], [1, 1], [0, 0], [1, 0]])
    z = rng.standard_normal((4, 6000))
    feat, names, _ = create_feature_array_from_raw_photometry(grid, list("abcd"), scatter_fluxes=2, depths=depths,
                                                              depth_indices=idx, normals=z, include_errors_in_feature_array=True)
    sig_sets = depths_to_sigma_njy(depths)                # (2, 4)
    want_noisy, want_std = O.apply_depths(grid, sig_sets, z, 2, depth_indices=idx)
    mag, merr = O.ab_features(want_noisy, want_std)
    want = np.concatenate([mag, merr], 0).T
    keep = np.isfinite(want).all(1)
    np.testing.assert_allclose(feat, want[keep], atol=1e-4)
    f2, _, _ = create_feature_array_from_raw_photometry(grid, list("abcd"), scatter_fluxes=2, depths=depths, seed=4, epoch=1)
    f3, _, _ = create_feature_array_from_raw_photometry(grid, list("abcd"), scatter_fluxes=2, depths=depths, seed=4, epoch=1)
    assert np.array_equal(f2, f3) and f2.shape == (6000, 4)
    with pytest.raises(ValueError):
        create_feature_array_from_raw_photometry(grid, list("abcd"), scatter_fluxes=2, depths=depths[:, :3])


def test_asinh_feature_rows():
    """normed_flux_units="asinh" (sbi_runner.py:1718-1730): asinh magnitudes and their errors with a per-filter softening,
    given explicitly or as "SNR_x" (x times the 1-sigma depth, :1660-1676); normalisation by a band subtracts."""
    rng = np.random.default_rng(13)
    names = ["a", "b", "c"]
    grid = rng.normal(20.0, 30.0, (3, 4000))                       # nJy, with negative fluxes: the point of asinh magnitudes
    z = rng.standard_normal((3, 8000))
    depths = np.array([29.0, 28.5, 28.0])
    sigma = depths_to_sigma_njy(depths)
    feat, fnames, _ = create_feature_array_from_raw_photometry(
        grid, names, normed_flux_units="asinh", asinh_softening_parameters="SNR_1", scatter_fluxes=2, depths=depths, normals=z,
        include_errors_in_feature_array=True)
    noisy, std = O.apply_depths(grid, sigma, z, 2)
    b = (1.0 * sigma * 1e-9)[:, None]
    want = np.concatenate([O.asinh_mag(noisy * 1e-9, b), O.asinh_mag_err(noisy * 1e-9, std * 1e-9, b)], 0).T
    assert feat.shape == want.shape and fnames == names + [f"unc_{n}" for n in names]
    np.testing.assert_allclose(feat, want, rtol=2e-6, atol=2e-6)
    assert np.isfinite(feat).all()                                  # negative fluxes are fine in asinh magnitudes
    soft = S.unyt_array([5.0, 8.0, 10.0], "nJy")
    f2, n2, _ = create_feature_array_from_raw_photometry(grid, names, normed_flux_units="asinh", asinh_softening_parameters=soft,
                                                         normalize_method="b")
    m = O.asinh_mag(grid * 1e-9, (np.array([5.0, 8.0, 10.0]) * 1e-9)[:, None]).T
    # (the appended column is the filter's library flux in normalization_unit -- AB by default -- whatever the rows' unit is)
    with np.errstate(invalid="ignore"):
        m_ab = -2.5 * np.log10(grid[1] * 1e-3) + 23.9
    keep = np.isfinite(m_ab)            # a negative library flux has no AB magnitude: that row goes (remove_nan_inf)
    assert 0 < keep.sum() and f2.shape[0] == keep.sum()
    np.testing.assert_allclose(f2, np.stack([m[:, 0] - m[:, 1], m[:, 2] - m[:, 1], m_ab], 1)[keep], rtol=2e-6, atol=2e-6)
    assert n2 == ["a", "c", "norm_b_AB"]
    with pytest.raises(AssertionError):
        create_feature_array_from_raw_photometry(grid, names, normed_flux_units="asinh")


def test_combination_library_meshgrid_mode(tmp_path):
    """draw_parameter_combinations=True (library.py:3644-3974; the reference's combined_library_basis_params fixture,
    tests/test_library.py:46-74): two grid-built bases with different emission models, every combination of their galaxies
    at each redshift x total mass x weight pair; against a brute-force restatement of the reference's loops."""
    from tests.test_host_api import _basis
    b1, b2 = _basis(build_library=True), _basis(build_library=True)
    b2.emission_model = S.TotalEmission(grid=b2.grid, fesc=0.2, fesc_ly_alpha=0.2, dust_curve=S.Calzetti2000(),
                                        dust_emission_model=None)
    b1.model_name, b2.model_name = "comb_basis1", "comb_basis2"
    weights = np.array([np.array([i, 1 - i]) for i in np.arange(0, 1.1, 0.25)])
    cb = S.CombinedBasis(bases=[b1, b2], log_stellar_masses=[9.0, 10.5], redshifts=b1.redshifts,
                         base_emission_model_keys=["emergent", "emergent"], combination_weights=weights,
                         out_name="comb_lib", out_dir=str(tmp_path), log_base_masses=9, draw_parameter_combinations=True)
    cb.process_bases(overwrite=True)
    out = cb.create_library(overwrite=True)
    n_z, n_per_z, n_m, n_w = 3, 3, 2, len(weights)
    assert out["photometry"].shape == (8, n_z * n_m * n_w * n_per_z * n_per_z)
    assert out["parameter_names"] == ["redshift", "log_mass", "weight_fraction", "comb_basis1/tau_v", "comb_basis2/tau_v"]
    assert os.path.exists(os.path.join(str(tmp_path), "comb_lib.hdf5"))
    with pytest.raises(AssertionError):
        cb.create_full_library()
    # brute force, in the reference's loop order
    o = cb.load_bases()
    codes = out["filter_codes"]
    cols, pars = [], []
    for z in b1.redshifts:
        for lm in (9.0, 10.5):
            for comb in weights:
                sel = [np.asarray(o[b.model_name]["properties"]["redshift"]) == z for b in (b1, b2)]
                ph = [np.array([o[b.model_name]["observed_photometry"][c][m] for c in codes], dtype=np.float32)
                      * (comb[j] * 10 ** lm / o[b.model_name]["properties"]["mass"][m]) for j, (b, m) in enumerate(zip((b1, b2), sel))]
                tv = [np.asarray(o[b.model_name]["properties"]["tau_v"])[m] for b, m in zip((b1, b2), sel)]
                combos = np.array(np.meshgrid(np.arange(3), np.arange(3), indexing="ij")).T.reshape(-1, 2)
                for i0, i1 in combos:
                    cols.append(ph[0][:, i0] + ph[1][:, i1])
                    pars.append([z, lm, comb[0], tv[0][i0], tv[1][i1]])
    np.testing.assert_allclose(out["photometry"], np.array(cols).T, rtol=1e-12)
    np.testing.assert_allclose(out["parameters"], np.array(pars).T, rtol=1e-12)
    lib = S.load_library_from_hdf5(os.path.join(str(tmp_path), "comb_lib.hdf5"))
    assert lib["photometry"].shape == out["photometry"].shape


def test_simulator_batch_equals_row_by_row_calls(tmp_path):
    """A batch through GalaxySimulator scatters, normalises and appends errors along the last axis; with depths it consumes
    numpy's global stream in the order a per-galaxy loop (the reference's calling pattern, sbi_runner.py:7659-7664) does."""
    basis, d, grid, inst, em = _small_basis(8, tmp_path)
    sim = S.GalaxySimulator(sfh_model=S.SFH.LogNormal, zdist_model=S.ZDist.DeltaConstant, grid=grid, instrument=inst,
                            emission_model=em, emission_model_key="emergent", out_flux_unit="AB", ignore_scatter=False,
                            include_phot_errors=True, depths=np.full(7, 29.0), normalize_method="JWST/NIRCam.F444W",
                            param_units={"peak_age": S.Myr, "max_age": S.Myr},
                            param_order=["redshift", "log_mass", "tau", "peak_age", "max_age", "log10metallicity", "tau_v"])
    rng = np.random.default_rng(4)
    n = 6
    p = np.column_stack([rng.uniform(0.5, 8, n), rng.uniform(9, 11, n), rng.uniform(0.2, 1.5, n), rng.uniform(10, 200, n),
                         rng.uniform(250, 400, n), rng.uniform(-3, -1.4, n), rng.uniform(0, 2, n)])
    np.random.seed(5)
    batch = sim(p)
    np.random.seed(5)
    rows = np.stack([sim(p[i]) for i in range(n)])
    assert batch.shape == (n, 7 + 1 + 7)
    np.testing.assert_array_equal(batch, rows)


def test_from_library_rebuilds_the_simulator_and_reproduces_the_library(tmp_path):
    """VERDICT r1 #1 / library.py:5219-5551: create_mock_library writes the ``Model`` group in the reference's layout;
    GalaxySimulator.from_library rebuilds grid, instrument, emission model (dust law + per-model parameters), parameter
    order / units and fixed parameters from the FILE alone, and simulating the library's own parameter rows returns the
    library's photometry."""
    n = 60
    basis, d, grid, inst, em = _small_basis(n, tmp_path)
    gdir = str(tmp_path / "grids")
    os.makedirs(gdir, exist_ok=True)
    grid.grid_name, grid.grid_dir = "synthetic_test_grid", gdir
    grid.save(os.path.join(gdir, "synthetic_test_grid.npz"))
    basis.create_mock_library(log_stellar_masses=np.asarray(d["masses"], dtype=float), emission_model_key="emergent",
                              out_name="rt_lib", out_dir=str(tmp_path), overwrite=True, batch_size=32)
    path = os.path.join(str(tmp_path), "rt_lib.hdf5")
    lib = S.load_library_from_hdf5(path)
    sim = S.GalaxySimulator.from_library(path, ignore_scatter=True)
    assert type(sim.emission_model).__name__ == "PacmanEmission" and sim.emission_model_key == "emergent"
    assert sim.emission_model.dust_curve.name == "Calzetti2000" and float(sim.emission_model.fesc) == 0.1
    assert sim.param_order == lib["parameter_names"] and sim.instrument.filters.filter_codes == inst.filters.filter_codes
    assert sim.sfh_model is S.SFH.LogNormal and sim.zdist_model is S.ZDist.DeltaConstant
    got = sim(lib["parameters"].T)                      # (n, n_params) rows in the library's own order and units
    # the library stores float32(base photometry) * mass ratio; the simulator scales on the device in float64
    np.testing.assert_allclose(got, lib["photometry"].T, rtol=2e-6)
    # a moved grid directory: override_synthesizer_grid_dir / SYNTHESIZER_GRID_DIR (library.py:5277-5299)
    moved = str(tmp_path / "moved")
    os.rename(gdir, moved)
    sim2 = S.GalaxySimulator.from_library(path, override_synthesizer_grid_dir=moved, ignore_scatter=True)
    np.testing.assert_array_equal(sim2(lib["parameters"].T[:5]), got[:5])
    with pytest.raises(FileNotFoundError):
        S.GalaxySimulator.from_library(os.path.join(str(tmp_path), "missing.hdf5"))


def test_simulator_sfh_and_rest_frame_photometry_outputs(tmp_path):
    """output_type 'sfh' (library.py:5736-5750) and 'photo_lnu' (:5756-5761): star-formation rate per grid age bin and
    rest-frame luminosities through the filters (no redshift, no IGM, no distance) -- checked against the oracle."""
    basis, d, grid, inst, em = _small_basis(8, tmp_path)
    order = ["redshift", "log_mass", "tau", "peak_age", "max_age", "log10metallicity", "tau_v"]
    kw = dict(sfh_model=S.SFH.LogNormal, zdist_model=S.ZDist.DeltaConstant, grid=grid, instrument=inst, emission_model=em,
              emission_model_key="emergent", ignore_scatter=True, param_units={"peak_age": S.Myr, "max_age": S.Myr}, param_order=order)
    vec = np.array([[3.0, 9.5, 0.5, 100.0, 300.0, -1.0, 0.2], [1.0, 10.2, 0.8, 400.0, 900.0, -2.2, 0.7]])
    gals = [dict(redshift=v[0], tau_v=v[6], sfh_kind="LogNormal", sfh=dict(min_age=0.0, max_age=v[4] * 1e6, tau=v[2], peak_age=v[3] * 1e6),
                 zd_kind="delta_log10", zd_value=v[5]) for v in vec]
    lam = np.asarray(grid.lam)
    # --- sfh
    out = S.GalaxySimulator(output_type=["sfh", "photo_fnu"], **kw)(vec)
    ages = 10.0 ** grid.log10ages
    for i, g in enumerate(gals):
        sf = O.sfh_bin_masses("LogNormal", g["sfh"], grid.log10ages)
        want = sf / sf.sum() * 10.0 ** vec[i, 1] / np.diff(ages, prepend=0.0)
        np.testing.assert_allclose(np.asarray(out["sfh"])[i], want, rtol=1e-9, atol=1e-30)
    np.testing.assert_allclose(np.asarray(out["sfh_time"]), ages / 1e6)
    np.testing.assert_allclose(np.asarray(out["sfh_time_abs"])[0], O.age_gyr(3.0) * 1e3 - ages / 1e6, rtol=1e-9)
    assert out["photo_fnu"].shape == (2, 7)
    # --- photo_lnu: L_nu = sum_k w_k G_k(lam) exp(-tau_v kappa) [erg/s/Hz per 1e9 Msun scaled to the mass], filters at rest
    got = S.GalaxySimulator(output_type="photo_lnu", **kw)(vec)
    ga, gu = O.emission_parts(grid.spectra, lam, "emergent", 0.1, 0.1)
    kap = O.dust_kappa(lam)
    for i, g in enumerate(gals):
        w = O.weights_for(g, grid.log10ages, grid.metallicity)
        lnu = (np.tensordot(w, ga, axes=([0, 1], [0, 1])) * np.exp(-g["tau_v"] * kap) + np.tensordot(w, gu, axes=([0, 1], [0, 1]))) * 10.0 ** vec[i, 1]
        want = np.array([O.apply_filter(lnu, lam, f.lam, f.t) for f in inst.filters])
        np.testing.assert_allclose(got[i], want, rtol=1e-5)
    # --- lnu (library.py:5752-5754): the rest-frame luminosity spectrum itself, on the grid's axis
    spec = S.GalaxySimulator(output_type="lnu", **kw)(vec)
    assert spec.shape == (2, lam.size)
    for i, g in enumerate(gals):
        w = O.weights_for(g, grid.log10ages, grid.metallicity)
        lnu = (np.tensordot(w, ga, axes=([0, 1], [0, 1])) * np.exp(-g["tau_v"] * kap) + np.tensordot(w, gu, axes=([0, 1], [0, 1]))) * 10.0 ** vec[i, 1]
        big = lnu > 1e-20 * lnu.max()
        np.testing.assert_allclose(spec[i][big], lnu[big], rtol=1e-5)


def test_multi_base_supplementary_parameters(tmp_path):
    """library.py:4631-4656 with two bases: each base's by-products are rescaled by its share of the galaxy's mass
    (weight x 10^logM / base mass) and named <model_name>/<name>."""
    n = 30
    b1, d, grid, inst, em = _small_basis(n, tmp_path)
    b2, _, _, _, _ = _small_basis(n, tmp_path)
    b2.model_name = "burst_basis"
    w = np.stack([np.linspace(0.3, 0.9, n), 1 - np.linspace(0.3, 0.9, n)], 1)
    cb = S.CombinedBasis([b1, b2], np.full(n, 10.0), np.asarray(d["redshift"], dtype=float), ["emergent", "emergent"], w,
                         out_name="combo_supp", out_dir=str(tmp_path))
    cb.process_bases(overwrite=True, sfr_10=(S.calculate_sfr, 10 * S.Myr), mass_weighted_age=S.calculate_mass_weighted_age)
    out = cb.create_library(overwrite=True)
    names = out["supplementary_parameter_names"]
    assert names == ["test_lhc_basis/sfr_10", "test_lhc_basis/mass_weighted_age", "burst_basis/sfr_10", "burst_basis/mass_weighted_age"]
    supp = out["supplementary_parameters"]
    np.testing.assert_allclose(supp[0] / w[:, 0], supp[2] / w[:, 1], rtol=1e-12)      # same galaxies, linear in the mass share
    np.testing.assert_allclose(supp[1], supp[3], rtol=1e-12)                          # an age does not scale
    lib = S.load_library_from_hdf5(cb.library_path)
    assert list(lib["supplementary_parameter_names"]) == names


def test_readme_quickstart_as_written(tmp_path):
    """The reference's quick start (README.md:78-134) with the import line swapped: a prior dictionary with a unit on
    `peak_age`, filters kept on their OWN tables, the SPS grid's native (non-constant-R) axis, one ZDist object
    per galaxy, `generate_sfh_basis` with absolute peak ages, `GalaxyBasis(..., log_stellar_masses=)`,
    `create_mock_library(out_name, emission_model_key='intrinsic')` -- then the library against the oracle."""
    from synference_b200 import (FilterCollection, GalaxyBasis, Instrument, IntrinsicEmission, Myr, SFH, ZDist, draw_from_hypercube,
                                 generate_sfh_basis)
    from synference_b200.synthetic import synthetic_grid
    N = 300
    parameter_prior_ranges = {
        "log_stellar_mass": (8.0, 12.0),
        "redshift": (0.0, 10.0),
        "log_zmet": (-4.0, -1.4),
        "peak_age": (0.0, 100) * Myr,       # (README: 1000 Myr, older than the universe at z > 5 -> NaN SFHs there, which
        "tau": (0.2, 2),                    #  _validate_library rejects, here as in the reference)
    }
    # (README.md:95 passes unlog_keys=['log_stellar_mass'], which RENAMES the key to 'stellar_mass' (library.py:1098-1103) and
    #  makes the README's own `parameter_samples["log_stellar_mass"]` a KeyError -- in the reference too; the draw is kept in log)
    unlogged = draw_from_hypercube(parameter_prior_ranges, N=16, unlog_keys=["log_stellar_mass"], rng=3)
    assert "stellar_mass" in unlogged and "log_stellar_mass" not in unlogged and np.all(np.asarray(unlogged["stellar_mass"]) >= 1e8)
    parameter_samples = draw_from_hypercube(parameter_prior_ranges, N=N, rng=3)
    filter_names = [f"JWST/NIRCam.{f}" for f in ["F090W", "F115W", "F150W", "F200W", "F277W", "F356W", "F444W"]]
    instrument = Instrument("JWST", filters=FilterCollection(filter_codes=filter_names))
    # (the BPASS file of the README is not available offline: a grid of its shape on a native-looking, non-constant-R axis)
    lam = np.concatenate([np.arange(200.0, 3000.0, 2.0), np.arange(3000.0, 12000.0, 10.0), 12000.0 * 1.003 ** np.arange(1, 520)])
    grid = synthetic_grid(lam)
    emission_model = IntrinsicEmission(grid=grid)
    Z_dists = [ZDist.DeltaConstant(log10metallicity=log_z) for log_z in parameter_samples["log_zmet"]]
    sfh_models, _ = generate_sfh_basis(sfh_type=SFH.LogNormal, sfh_param_names=["tau", "peak_age"],
                                       sfh_param_arrays=(parameter_samples["tau"], parameter_samples["peak_age"]),
                                       redshifts=parameter_samples["redshift"])
    basis = GalaxyBasis(model_name="sps_test", redshifts=parameter_samples["redshift"],
                        log_stellar_masses=parameter_samples["log_stellar_mass"], grid=grid, emission_model=emission_model,
                        sfhs=sfh_models, instrument=instrument, metal_dists=Z_dists)
    basis.create_mock_library(out_name="library_test", emission_model_key="intrinsic", overwrite=True, out_dir=str(tmp_path))
    lib = S.load_library_from_hdf5(os.path.join(str(tmp_path), "library_test.hdf5"))
    assert lib["photometry"].shape == (7, N) and lib["filter_codes"] == filter_names
    p = basis.params
    want = O.synthesize(A.galaxies_from_params(p), grid.log10ages, grid.metallicity, lam, grid.spectra,
                        [(f.lam, f.t) for f in instrument.filters], key="intrinsic", fesc=0.0, fesc_ly_alpha=1.0, dust=None,
                        igm=(I.INOUE14_LAF, I.INOUE14_DLA))
    logm = np.asarray(lib["parameters"][lib["parameter_names"].index("log_mass")], dtype=float)
    usable = np.isfinite(lib["photometry"]).all(0)
    assert usable.mean() > 0.95
    assert_flux_close(lib["photometry"].T[usable], O.scale_to_mass(want, logm)[usable])
