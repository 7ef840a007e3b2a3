"""Host-side API behaviour mirrored from the reference (no GPU needed)."""

import ctypes
import os
import re

import numpy as np
import pytest

import synference_b200 as S
from synference_b200 import _capi
from synference_b200.configs import make_workload
from synference_b200.synthetic import NIRCAM_WIDE8, synthetic_filters, synthetic_grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- draw_from_hypercube (library.py:1021-1115) ---------------------------------------------------
def test_draw_from_hypercube_contract():
    pr = {"log_stellar_mass": (8.0, 12.0), "redshift": (0.0, 10.0), "peak_age": (0.0, 1000) * S.Myr, "tau": (0.2, 2)}
    d = S.draw_from_hypercube(pr, N=500, rng=42, unlog_keys=["log_stellar_mass"])
    assert set(d) == {"stellar_mass", "redshift", "peak_age", "tau"}          # log_ prefix dropped (:1103)
    assert np.asarray(d["redshift"]).dtype == np.float32                       # float32 cast (:1098)
    assert str(d["peak_age"].units) == "Myr"
    assert 1e8 <= d["stellar_mass"].min() and d["stellar_mass"].max() <= 1e12
    # Latin hypercube: one sample per stratum along each axis
    strata = np.floor(np.asarray(d["redshift"], dtype=np.float64) / 10.0 * 500).astype(int)
    assert len(np.unique(strata)) >= 498
    with pytest.raises(AssertionError):
        S.draw_from_hypercube({"a": (1.0, 0.0)}, N=4)
    with pytest.raises(ValueError):
        S.draw_from_hypercube({"a": (0.0, 1.0)}, N=4, model=lambda d: None)


# ---- generate_sfh_basis (library.py:1137-1334) ----------------------------------------------------
def test_generate_sfh_basis_max_age_and_norm():
    z = np.array([0.5, 3.0, 7.0])
    sfhs, zz = S.generate_sfh_basis(S.SFH.LogNormal, ["tau", "peak_age_norm"],
                                    [np.array([0.5, 1.0, 1.5]), np.array([0.1, 0.5, 0.9])], redshifts=z, max_redshift=20)
    assert len(sfhs) == 3 and np.array_equal(zz, z)
    max_age_myr = (S.Planck18.age(z) - S.Planck18.age(20)).to("Myr").value
    np.testing.assert_allclose(sfhs.max_age, max_age_myr * 1e6, rtol=1e-12)
    s1 = sfhs[1]
    assert s1.redshift == 3.0 and s1.tau == 1.0
    assert s1.peak_age == pytest.approx(0.5 * max_age_myr[1] * 1e6)            # *_norm scaled by max_age (:1287-1289)
    assert "peak_age" in s1.parameters and "max_age" in s1.parameters
    # units on ordinary parameters, scalar redshift broadcast, explicit max_age capped
    sf2, _ = S.generate_sfh_basis(S.SFH.DelayedExponential, ["tau", "max_age"],
                                  [np.array([100.0, 200.0]) * S.Myr, np.array([50.0, 1e5]) * S.Myr], redshifts=2.0)
    assert sf2.rows[0, 1] == pytest.approx(50e6) and sf2.rows[1, 1] == pytest.approx(float(max_age_myr[0]) * 0 + sf2.rows[1, 1])
    assert sf2.rows[1, 1] < 1e11 and sf2.rows[0, 2] == pytest.approx(100e6)
    with pytest.raises(ValueError):
        S.generate_sfh_basis(S.SFH.LogNormal, ["tau"], [np.ones(2)], redshifts="bad")
    # iterate_redshifts: every redshift x every row
    sf3, _ = S.generate_sfh_basis(S.SFH.LogNormal, ["tau", "peak_age"], [np.array([0.5, 1.0]), np.array([10.0, 20.0]) * S.Myr],
                                  redshifts=np.array([1.0, 2.0, 3.0]), iterate_redshifts=True)
    assert len(sf3) == 6 and list(sf3.redshifts) == [1.0, 1.0, 2.0, 2.0, 3.0, 3.0]


def test_sfh_objects_and_packing():
    s = S.SFH.LogNormal(tau=0.5, peak_age=100 * S.Myr, max_age=300 * S.Myr)      # tests/conftest.py:102-105
    assert s.max_age == 3e8 and s.peak_age == 1e8
    assert s.get_sfr(np.array([-1.0, 1e7, 4e8])).tolist()[0] == 0 and s.get_sfr(np.array([4e8]))[0] == 0
    zd = S.ZDist.DeltaConstant(log10metallicity=-1.0)
    with pytest.raises(ValueError):
        S.ZDist.DeltaConstant()
    p = S.GalaxyParams.from_objects(np.array([6.0, 7.0]), [s, s], [zd, zd], log_mass=[9, 10], tau_v=[0.2, 0.3])
    assert p.sfh_type == 5 and p.sfh_rows.shape == (2, 4) and p.zd_type == 1
    with pytest.raises(ValueError):
        S.GalaxyParams.from_objects(np.array([1.0, 2.0]), [s, S.SFH.Constant(max_age=1e8 * S.yr)], [zd, zd])


def _basis(n=12, build_library=False):
    raw = synthetic_filters(NIRCAM_WIDE8)
    lam = S.generate_constant_R(R=300, auto_start_stop=True, filterset=raw, max_redshift=15)
    inst = S.Instrument("JWST", filters=synthetic_filters(NIRCAM_WIDE8, new_lam=lam))
    grid = synthetic_grid(lam)
    em = S.PacmanEmission(grid=grid, fesc=0.1, fesc_ly_alpha=0.1, dust_curve=S.Calzetti2000())
    if build_library:
        return S.GalaxyBasis("test_basis", np.array([6.0, 7.0, 8.0]), grid, em,
                             [S.SFH.LogNormal(tau=0.5, peak_age=100 * S.Myr, max_age=300 * S.Myr)],
                             [S.ZDist.DeltaConstant(log10metallicity=-1.0)], galaxy_params={"tau_v": [0.2, 0.3, 0.4]},
                             instrument=inst, build_library=True)
    d = S.draw_from_hypercube({"redshift": (0.01, 10), "masses": (5, 11), "tau_v": (0, 2), "peak_age": (0, 0.99),
                               "tau": (0.1, 1.5), "log_zmet": (-3, -1.39)}, N=n, rng=42)
    sfhs, _ = S.generate_sfh_basis(S.SFH.LogNormal, ["tau", "peak_age_norm"], np.vstack((d["tau"], d["peak_age"])).T,
                                   redshifts=np.array(d["redshift"]))
    zds = [S.ZDist.DeltaConstant(log10metallicity=z) for z in d["log_zmet"]]
    return S.GalaxyBasis("test_lhc_basis", d["redshift"], grid, em, sfhs, zds, galaxy_params={"tau_v": d["tau_v"]},
                         instrument=inst, log_stellar_masses=d["masses"])


def test_galaxy_basis_matched_parameters():
    b = _basis(12)
    gals = b._create_matched_galaxies()
    assert len(gals) == 12 and len(b.params) == 12
    assert set(b.varying_param_names) >= {"redshift", "tau_v", "tau", "peak_age", "log10metallicity", "max_age"}
    assert gals[3]["all_params"]["tau_v"] == pytest.approx(float(b.galaxy_params["tau_v"][3]))
    mask = np.zeros(12, bool)
    mask[4:9] = True
    b._create_matched_galaxies(galaxies_mask=mask)
    assert len(b.params) == 5 and b.params.redshift[0] == pytest.approx(float(b.redshifts[4]))


def test_galaxy_basis_combinatorial_order():
    b = _basis(build_library=True)
    b._create_galaxies()
    assert len(b.params) == 9                                                   # 3 z x 1 sfh x 1 Z x 3 tau_v
    assert list(b.params.redshift) == [6.0] * 3 + [7.0] * 3 + [8.0] * 3
    assert list(b.params.tau_v[:3]) == [0.2, 0.3, 0.4]
    assert b.varying_param_names == ["tau_v", "redshift"] and "tau" in b.fixed_param_names
    with pytest.raises(ValueError):
        _basis(4)._create_galaxies()


def test_unsupported_features_fail_loudly():
    b = _basis(4)
    with pytest.raises(NotImplementedError):
        S.GalaxyBasis("x", b.redshifts, b.grid, b.emission_model, b.sfhs, b.metal_dists,
                      galaxy_params={"slope": np.zeros(4)}, instrument=b.instrument)
    lya = S.PacmanEmission(grid=b.grid, fesc=0.1, fesc_ly_alpha="fesc_lya", dust_curve=S.Calzetti2000())
    assert lya.lya_per_galaxy and lya.lya_line("emergent") is not None and lya.lya_line("incident") is None
    per = S.PacmanEmission(grid=b.grid, fesc="fesc", dust_curve=S.Calzetti2000())     # per-galaxy fesc IS supported ...
    with pytest.raises(ValueError):                                                  # ... but must then be provided
        S.GalaxyBasis("x", b.redshifts, b.grid, per, b.sfhs, b.metal_dists, galaxy_params={"tau_v": np.ones(4)},
                      instrument=b.instrument)
    with pytest.raises(ValueError):                                                  # and not with a global-fesc model
        S.GalaxyBasis("x", b.redshifts, b.grid, b.emission_model, b.sfhs, b.metal_dists,
                      galaxy_params={"fesc": np.full(4, 0.2)}, instrument=b.instrument)
    with pytest.raises(ValueError):
        b.emission_model.recipe("nonsense")
    with pytest.raises(ValueError):
        S.IntrinsicEmission(grid=b.grid).recipe("emergent")


def test_validate_and_save_library_roundtrip(tmp_path):
    b = _basis(4)
    cb = S.CombinedBasis([b], [9.0] * 4, b.redshifts, ["emergent"], None, out_name="lib", out_dir=str(tmp_path))
    good = {"photometry": np.ones((8, 4)), "parameters": np.ones((3, 4)), "parameter_names": ["redshift", "log_mass", "tau_v"],
            "filter_codes": list(b.instrument.filters.filter_codes), "parameter_units": ["dimensionless", "log10_Msun", "mag"],
            "supplementary_parameters": np.zeros((0, 4)), "supplementary_parameter_names": [], "supplementary_parameter_units": []}
    cb.save_library(good, overwrite=True)
    lib = S.load_library_from_hdf5(os.path.join(str(tmp_path), "lib.hdf5"))
    assert lib["photometry"].shape == (8, 4) and lib["parameter_names"][:2] == ["redshift", "log_mass"]
    assert lib["photometry_units"] == "nJy" and lib["filter_codes"] == good["filter_codes"]
    bad = dict(good, photometry=np.full((8, 4), np.nan))
    with pytest.raises(ValueError, match="NaN"):
        cb._validate_library(bad)
    with pytest.raises(ValueError, match="infinite"):
        cb._validate_library(dict(good, photometry=np.full((8, 4), np.inf)))


def test_noise_model_serialisation_roundtrip(tmp_path):
    path = os.path.join(str(tmp_path), "models.hdf5")
    m = S.DepthUncertaintyModel(depth_ab=26.5, depth_sigma_level=10)
    S.save_unc_model_to_hdf5(m, path, "depth_test", overwrite=True)
    m2 = S.load_unc_model_from_hdf5(path, "depth_test")
    assert isinstance(m2, S.DepthUncertaintyModel) and m2.depth_ab == 26.5
    assert float(m2.sigma.value) == pytest.approx(float(m.sigma.value))
    rng = np.random.default_rng(0)
    mag = np.linspace(20, 28, 3000)
    err = 0.05 + np.exp((mag - 26) / 1.5) + rng.normal(0, 0.02, mag.size)
    g = S.GeneralEmpiricalUncertaintyModel(mag + rng.normal(0, 0.01, mag.size), err, flux_unit="AB", log_bins=False,
                                           return_noise=True)
    S.save_unc_model_to_hdf5(g, path, "emp", overwrite=True)
    g2 = S.load_unc_model_from_hdf5(path, "emp")
    np.testing.assert_allclose(g2.bin_centers, g.bin_centers)
    np.random.seed(3)
    f, s = g2.apply_noise(np.full(2000, 25.0), true_flux_units="AB", out_units="AB")
    assert np.isfinite(f).all() and np.std(f) == pytest.approx(float(g._mu_sigma_interpolator(25.0)), rel=0.25)
    fj = np.asarray(S.UncertaintyModel.ab_to_jy(mag))
    a = S.AsinhEmpiricalUncertaintyModel(fj, np.asarray(S.UncertaintyModel.ab_err_to_jy(err, fj)), return_noise=True)
    S.save_unc_model_to_hdf5(a, path, "asinh", overwrite=True)
    a2 = S.load_unc_model_from_hdf5(path, "asinh")
    assert float(a2.b.value) == pytest.approx(float(a.b.value))
    m_as, e_as = a2.apply_noise(S.Quantity(fj[:100], "Jy"))
    assert np.isfinite(m_as).all() and np.all(e_as >= 0)
    with pytest.raises(ValueError):
        S.save_unc_model_to_hdf5(m, path, "depth_test")


# ---- C ABI: the library builds here, loads and exports every declared symbol ------------------------
def test_capi_exports_every_declared_symbol(native_lib):
    header = open(os.path.join(ROOT, "include", "synference_b200.h")).read()
    declared = set(re.findall(r"\b(sb2_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_capi.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(native_lib, sym), sym
    assert ctypes.sizeof(_capi.Params) >= 13 * 8


def test_no_cpu_fallback(native_lib):
    """Without a GPU the product must fail loudly (never route through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    w = make_workload("cfg1", 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        S.SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters)
    src = "".join(open(os.path.join(ROOT, "synference_b200", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "synference_b200")) if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src


def test_shard_bounds_follow_reference_rule():
    from synference_b200.distributed import shard_bounds, shard_counts
    assert [shard_bounds(10, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]   # library.py:3130-3137
    assert sum(shard_counts(1_000_003, 8)) == 1_000_003


def test_per_galaxy_fesc_recipe_and_coefficients():
    """fesc="fesc": grids are built for fesc = 0 and (1 - fesc, fesc) become per-galaxy coefficients; summing
    coefficient * grid reproduces the global-fesc recipe for every spectrum of the tree."""
    import numpy as np
    from synference_b200.configs import make_workload
    from synference_b200.parametric import Calzetti2000, PacmanEmission
    w = make_workload("cfg1", 4)
    per = PacmanEmission(grid=w.grid, fesc="fesc", fesc_ly_alpha=0.4, dust_curve=Calzetti2000())
    assert per.fesc_per_galaxy and per.fesc_name == "fesc"
    for f in (0.0, 0.25, 1.0):
        glob = PacmanEmission(grid=w.grid, fesc=f, fesc_ly_alpha=0.4, dust_curve=Calzetti2000())
        for key in ("incident", "transmitted", "nebular", "reprocessed", "escaped", "intrinsic", "attenuated", "emergent", "total"):
            a, u = per.recipe(key)
            ca, cu = per.coefficients(key, np.array([f]))
            ca = 1.0 if ca is None else ca[0]
            cu = 1.0 if cu is None else cu[0]
            ga, gu = glob.recipe(key)
            if per.dust_free(key):          # both grids bypass the screen: compare the sum
                np.testing.assert_allclose(ca * a + cu * u, ga + gu, rtol=1e-14, atol=0)
            else:
                np.testing.assert_allclose(ca * a, ga, rtol=1e-14, atol=0)
                np.testing.assert_allclose(cu * u, gu, rtol=1e-14, atol=0)
    import pytest
    with pytest.raises(ValueError):
        per.coefficients("emergent", np.array([1.2]))
    # per-galaxy Lyman-alpha escape: grids are lowered without the line bin, lya_line() carries it
    lya = PacmanEmission(grid=w.grid, fesc=0.25, fesc_ly_alpha="fesc_lya", dust_curve=Calzetti2000())
    full = PacmanEmission(grid=w.grid, fesc=0.25, fesc_ly_alpha=1.0, dust_curve=Calzetti2000())
    vals, i = lya.lya_line("emergent")
    a0, _ = lya.recipe("emergent")
    a1, _ = full.recipe("emergent")
    rebuilt = a0.copy()
    rebuilt[..., i] += vals
    np.testing.assert_allclose(rebuilt, a1, rtol=1e-14, atol=0)


def test_supplementary_by_products_of_the_sfzh():
    """synference_b200.supplementary: history-based callbacks of library.py:223-241, 427-442, 468-526 evaluated per batch.
    calculate_sfh_quantile is pinned on the reference's own function (golden vectors); the others on their definitions."""
    import os
    from synference_b200 import supplementary as SP
    from synference_b200.cosmology import Planck18
    from synference_b200.units import Myr
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_noise_golden.npz"))
    ages, sfh, z = G["q_ages"], G["q_sfh"], G["q_z"]
    n_z = 3
    zfrac = np.array([0.2, 0.5, 0.3])
    sfzh = sfh[:, :, None] * zfrac[None, None, :]
    ctx = SP._Context(sfzh, np.log10(ages), z, Planck18)
    for q in (25, 50, 90):
        np.testing.assert_allclose(SP.calculate_sfh_quantile(ctx, q / 100), G[f"q_{q}"], rtol=1e-12)
    np.testing.assert_allclose(SP.calculate_sfh_quantile(ctx, 0.5, True, Planck18), G["q_50_norm"], rtol=1e-9)
    np.testing.assert_allclose(SP.calculate_mass_weighted_age(ctx), (sfh @ ages) / sfh.sum(1) / 1e6, rtol=1e-13)
    # SFR over a timescale: brute-force integral of the piecewise-uniform history
    edges = np.concatenate([[0.0], 0.5 * (ages[1:] + ages[:-1]), [ages[-1]]])
    dens = sfh / np.diff(edges)
    t = np.linspace(0, 3e7, 300001)
    mid = 0.5 * (t[1:] + t[:-1])
    mass = (dens[:, np.searchsorted(edges, mid, side="right") - 1] * np.diff(t)).sum(1)
    np.testing.assert_allclose(SP.calculate_sfr(ctx, 30 * Myr), mass / 3e7, rtol=1e-4)     # (the brute-force sum has 100 yr steps)
    np.testing.assert_allclose(SP.calculate_burstiness(ctx), SP.calculate_sfr(ctx, 1e7) / SP.calculate_sfr(ctx, 1e8), rtol=1e-13)
    grid = type("G", (), {"stellar_fraction": np.linspace(1.0, 0.5, ages.size)[:, None] * np.ones((1, n_z))})()
    np.testing.assert_allclose(SP.calculate_surviving_mass(ctx, grid), np.log10(sfh @ np.linspace(1.0, 0.5, ages.size)), rtol=1e-13)
    out = SP.evaluate({"mwa": SP.calculate_mass_weighted_age, "sfr_10": (SP.calculate_sfr, 10 * Myr),
                       "q50n": (SP.calculate_sfh_quantile, 0.5, True)}, sfzh, np.log10(ages), z, Planck18)
    assert [out[k][1] for k in ("mwa", "sfr_10", "q50n")] == ["Myr", "Msun/yr", "dimensionless"]
    assert [SP.scales_with_mass(u) for u in ("Myr", "Msun/yr", "log10_Msun")] == ["none", "linear", "log"]
    with pytest.raises(NotImplementedError):
        SP.check_supported({"beta": lambda galaxy: 0.0})


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's keys,
    runnable without a GPU; under torchrun ranks other than 0 print nothing."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-sample", "400"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "galaxies/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["gpu_launches"] == 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                         "--warmup", "0", "--ref-sample", "200"], capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_new_device_paths_fail_loudly_without_a_gpu():
    """No CPU fallback anywhere: on a box without a CUDA device the spectral and empirical-noise entry points raise."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a GPU-less box")
    lam = 0.05 * (1 + 0.5 / 300) ** np.arange(500)
    with pytest.raises(RuntimeError):
        S.SpectrumResampler(lam, np.linspace(0.6, 1.0, 50), np.array([0.5, 1.1]), np.array([30.0, 100.0]))
    with pytest.raises(RuntimeError):
        S.transform_spectrum(lam, np.ones(500), 1.0, np.linspace(0.6, 1.0, 50), np.array([0.5, 1.1]), np.array([30.0, 100.0]))
    c = np.linspace(20.0, 30.0, 8)
    mod = S.GeneralEmpiricalUncertaintyModel(c, None, flux_unit="AB", already_binned=True, bin_median_errors=np.full(8, 0.1),
                                             bin_std_errors=np.full(8, 0.01))
    with pytest.raises(RuntimeError):
        S.apply_empirical_noise_models(np.full((1, 4), 25.0), ["a"], {"a": mod}, N_scatters=1)
    # the model still lowers to the C struct on the host
    m = mod.device_model("nJy", "AB")
    assert m.n_bins == 8 and m.internal_is_ab == 1 and m.in_is_ab == 0 and m.in_to_jy == 1e-9 and m.out_is_ab == 1


def test_generate_sfh_grid_and_emission_models():
    """library.py:742-873 / :931-1018: meshgrid of drawn redshifts x SFH parameters (redshift-dependent upper bounds) and one
    emission model per drawn parameter combination."""
    from scipy.stats import uniform
    import synference_b200 as S
    from synference_b200.synthetic import synthetic_grid
    np.random.seed(3)
    pri = {"tau": {"prior": uniform, "min": 0.1, "max": 1.5, "size": 4},
           "peak_age": {"prior": uniform, "min": 0.0, "max": 13000.0, "size": 6, "units": S.Myr, "depends_on": "max_redshift"}}
    sfhs, combos = S.generate_sfh_grid(S.SFH.LogNormal, pri, {"prior": uniform, "min": 0.5, "max": 8.0, "size": 3}, max_redshift=15)
    assert combos.shape == (3 * 4 * 6, 3) and len(sfhs) == 72
    z = combos[:, 0]
    want_max = (S.Planck18.age(z) - S.Planck18.age(15)).to("Myr").value
    np.testing.assert_allclose(sfhs.rows[:, 1], want_max * 1e6, rtol=1e-12)
    np.testing.assert_allclose(sfhs.rows[:, 2], combos[:, 1])                  # tau
    np.testing.assert_allclose(sfhs.rows[:, 3], combos[:, 2] * 1e6)            # peak_age Myr -> yr
    assert np.all(combos[:, 2].reshape(3, 4, 6).max(axis=(0, 1)) <= want_max.max())   # capped by the available age
    one = sfhs[5]
    assert type(one).__name__ == "_LogNormal" and abs(one.redshift - z[5]) < 1e-12
    lam = S.generate_constant_R(R=100, start=900 * S.Angstrom, end=5e4 * S.Angstrom)
    grid = synthetic_grid(lam)
    models, out = S.generate_emission_models(S.PacmanEmission, {"fesc": {"prior": uniform, "min": 0.0, "max": 0.5, "size": 3}},
                                             grid, fixed_params={"dust_curve": S.Calzetti2000()})
    assert len(models) == 3 and len(out["fesc"]) == 3 and all(0 <= m.fesc <= 0.5 for m in models)
    assert all(m.dust_curve.name == "Calzetti2000" for m in models)


def test_uncertainty_models_from_an_epochs_style_table():
    """noise_models.py:1159-1330 on a synthetic catalogue passed as a dict of columns (no astropy needed)."""
    import synference_b200 as S
    rng = np.random.default_rng(0)
    n = 4000
    depth = rng.normal(29.0, 0.2, n)
    mag = rng.uniform(23, 30, n)
    flux = 10 ** (-0.4 * (mag - 8.9))
    tab = {"MAG_APER_F444W_aper_corr": mag, "FLUX_APER_F444W_aper_corr_Jy": flux, "loc_depth_F444W": depth}
    tab["MAG_APER_F444W_aper_corr"][:5] = -99
    ms = S.create_uncertainty_models_from_EPOCHS_cat(tab, "F444W", new_band_names=["JWST/NIRCam.F444W"])
    m = ms["JWST/NIRCam.F444W"]
    assert isinstance(m, S.GeneralEmpiricalUncertaintyModel) and m.return_noise
    np.random.seed(1)
    noisy, sig = m.apply_noise(np.full(100, 26.0), true_flux_units="AB", out_units="AB")
    want_sigma = 2.5 / np.log(10) * (10 ** (-0.4 * (29.0 - 8.9)) / 5) / 10 ** (-0.4 * (26.0 - 8.9))
    assert abs(np.median(sig) / want_sigma - 1) < 0.3
    d = S.create_uncertainty_models_from_EPOCHS_cat(tab, ["F444W"], model_class="depth")["F444W"]
    assert isinstance(d, S.DepthUncertaintyModel) and abs(d.depth_ab - np.median(depth)) < 1e-12
    a = S.create_uncertainty_models_from_EPOCHS_cat(tab, ["F444W"], model_class="asinh")["F444W"]
    assert isinstance(a, S.AsinhEmpiricalUncertaintyModel)
    with pytest.raises(ValueError):
        S.create_uncertainty_models_from_EPOCHS_cat(tab, "F200W")
    with pytest.raises(ValueError):
        S.create_uncertainty_models_from_EPOCHS_cat(tab, "F444W", model_class="nope")


def test_energy_full_axis_flag_follows_the_largest_optical_depth():
    """SynthEngine._fill (host logic, no GPU): a model with absorbed-energy pseudo-bins asks for the sum over the whole axis
    (sb2_params.energy_full_axis) when a batch's largest tau_V (+ ratio x tau_V_birth) is beyond the range the pseudo-bins are
    accurate for, or is not finite."""
    from synference_b200.engine import GalaxyParams, SynthEngine
    eng = SynthEngine.__new__(SynthEngine)
    eng.tables = dict(single_is_unatt=False, lya_line=None, kappa_birth=np.zeros(4, np.float32), dust_global=None,
                      x_bins=192, x_tau_max=10.0, x_birth_ratio=1.0)
    n = 5
    base = dict(redshift=np.ones(n), sfh_type=5, sfh_rows=np.zeros((n, 4)), zd_type=1, zd_value=np.full(n, -2.0), zd_sigma=None,
                log_mass=None)
    for tau, birth, want in ((3.0, 2.0, 0), (6.0, 3.9, 0), (6.0, 4.1, 1), (11.0, 0.0, 1), (np.nan, 0.0, 1)):
        p = GalaxyParams(tau_v=np.full(n, tau), tau_v_birth=np.full(n, birth), **base)
        assert eng._fill(p, lambda a: None).energy_full_axis == want, (tau, birth)
    eng.tables["x_bins"] = 0
    p = GalaxyParams(tau_v=np.full(n, 50.0), tau_v_birth=np.full(n, 0.0), **base)
    assert eng._fill(p, lambda a: None).energy_full_axis == 0


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """The C ABI is plain structs: every field of the ctypes mirrors in synference_b200/_capi.py must sit at the offset a C
    compiler gives it in include/synference_b200.h (a drifted mirror would scramble every argument behind it).  gcc compiles a
    probe that prints sizeof and offsetof for every field; the field NAMES come from the mirrors, so a field missing on
    either side fails to compile or to compare."""
    import ctypes
    import shutil
    import subprocess
    from synference_b200 import _capi
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = [("sb2_model_desc", _capi.ModelDesc), ("sb2_params", _capi.Params), ("sb2_resample_desc", _capi.ResampleDesc),
             ("sb2_empirical_model", _capi.EmpiricalModel)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "synference_b200.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, *_ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in pairs:
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, *_ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
    # and the header has no field the mirror lacks: same number of members (counted by the compiler through the sizes above
    # only if every one is mirrored; a trailing unmirrored field shows up as a size mismatch)
