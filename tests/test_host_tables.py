"""CPU checks of the host-side derivations the CUDA kernels consume.

``emulate_*`` replay, in float64 numpy, exactly the arithmetic the kernels do with the device tables
(geometric-axis filter shift (m, beta) with packed (U, DV) weights; separable Inoue+14 power tables and
prefix sums; TF32 hi/lo split).  Comparing them with the oracle's general semantics (np.interp of each
filter onto the observed abscissa, T>0 compress, trapezoid; line-by-line IGM loops) validates the
table construction without a GPU.
"""

import numpy as np
import pytest

from oracle import adapter as A, oracle as O
from synference_b200 import igm as I
from synference_b200.configs import make_workload
from synference_b200.engine import build_tables, geometric_ratio, tf32_split


def emulate_igm(tab, z):
    """exp(-tau) for the n_blue bins, from the separable tables (what prep_kernel does)."""
    nb, bp, pre, thr = tab["n_blue"], tab["bin_pow"], tab["pre"], tab["thr"]
    zp = 1.0 + z
    Z = {p: zp**p for p in I.Z_POWERS}
    b12, b21, b37, b55 = bp[0] * Z[1.2], bp[1] * Z[2.1], bp[2] * Z[3.7], bp[3] * Z[5.5]
    bm3, b2, b3, xl = bp[4] * Z[-0.3], bp[5] * Z[2.0], bp[6] * Z[3.0], bp[7] * zp
    n1 = (thr[0][None, :] > xl[:, None]).sum(1)
    n2 = (thr[1][None, :] > xl[:, None]).sum(1)
    nd = (thr[2][None, :] > xl[:, None]).sum(1)
    J = tab["nline"]
    a, b, c = np.minimum(J, n1), np.minimum(J, n2), np.minimum(J, nd)
    tau = (b12 * pre[0][a] + b37 * (pre[1][b] - pre[1][a]) + b55 * (pre[2][J] - pre[2][b])
           + b2 * pre[3][c] + b3 * (pre[4][J] - pre[4][c]))
    lc = np.zeros(nb)
    if z < 2.0:
        lc += 0.2113 * Z[2.0] - 0.07661 * Z[2.3] * bm3 - 0.1347 * b2
    else:
        lc += np.where(xl >= 3.0, 0.04696 * Z[3.0] - 0.01779 * Z[3.3] * bm3 - 0.02916 * b3,
                       0.6340 + 0.04696 * Z[3.0] - 0.01779 * Z[3.3] * bm3 - 0.1347 * b2 - 0.2905 * bm3)
    if z < 1.2:
        lc += 0.3248 * (b12 - Z[-0.9] * b21)
    elif z < 4.7:
        lc += np.where(xl >= 2.2, 2.545e-2 * (Z[1.6] * b21 - b37),
                       2.545e-2 * Z[1.6] * b21 + 0.3248 * b12 - 0.2496 * b21)
    else:
        lc += np.where(xl > 5.7, 5.221e-4 * (Z[3.4] * b21 - b55),
                       np.where(xl >= 2.2, 5.221e-4 * Z[3.4] * b21 + 0.2182 * b21 - 2.545e-2 * b37,
                                5.221e-4 * Z[3.4] * b21 + 0.3248 * b12 - 3.140e-2 * b21))
    return np.exp(-(tau + np.where(tab["lc_on"] > 0, lc, 0.0)))


def emulate_filters(t, spec, z):
    """num/den of every filter from the packed (U, DV) tables with the (m, beta) shift."""
    q = t["q"]
    s = np.log1p(z)
    m = int(np.floor(s / np.log(q)))
    r = np.exp(s - m * np.log(q))
    beta = (1 - 1 / r) / (1 - 1 / q) if t["interp_variant"] == 0 else (r - 1) / (q - 1)
    out = np.zeros(t["n_filt"])
    i = np.arange(t["n_lam"])
    for f in range(t["n_filt"]):
        lo, hi, off = int(t["filt_lo"][f]), int(t["filt_hi"][f]), int(t["filt_off"][f])
        k = np.clip(i + m - (lo - 2), 0, hi - lo + 3)
        uv = t["filt_uv"][off + k].astype(np.float64)
        num = np.sum(spec * ((1 - beta) * uv[:, 0] + beta * uv[:, 1]))
        out[f] = num / ((1 - beta) * t["filt_su"][f] + beta * t["filt_sdv"][f])
    return out


@pytest.mark.parametrize("variant", ["nu", "lam"])
def test_filter_tables_reproduce_general_filter_integration(variant):
    w = make_workload("cfg2", 4)
    t = build_tables(w.grid, w.emission_model, w.emission_key, w.filters, variant=variant)
    lam = np.asarray(w.grid.lam)
    rng = np.random.default_rng(0)
    for z in (0.013, 0.5, 2.7181, 6.02, 9.97):
        spec = np.exp(rng.normal(0, 1, lam.size)) * (lam / 1e4) ** -1.0
        want = np.array([O.apply_filter(spec, lam * (1 + z), f.lam, f.t, variant) for f in w.filters])
        got = emulate_filters(t, spec, z)
        # the table uses float32 weights -> 1e-7 level; the semantics (compress end-points, shift,
        # blend) would show up at 1e-5..1e-3 if wrong
        np.testing.assert_allclose(got, want, rtol=2e-6)


def test_igm_tables_reproduce_line_by_line_transmission():
    w = make_workload("cfg1", 4)
    lam = np.asarray(w.grid.lam)
    tab = I.device_tables(lam)
    assert np.all(lam[:tab["n_blue"]] < 1215.67) and lam[tab["n_blue"]] >= 1215.67
    for z in (0.05, 1.19, 1.2, 1.9999, 2.0, 3.3, 4.6999, 4.7, 7.5, 14.9):
        want = O.inoue14_transmission(z, lam * (1 + z), I.INOUE14_LAF, I.INOUE14_DLA)
        got = emulate_igm(tab, z)
        tau_w = -np.log(np.maximum(want[:tab["n_blue"]], 1e-300))
        tau_g = -np.log(np.maximum(got, 1e-300))
        ok = tau_w < 600
        np.testing.assert_allclose(tau_g[ok], tau_w[ok], rtol=1e-10, atol=1e-12)
        assert np.all(want[tab["n_blue"]:] == 1.0)


def test_tf32_split_properties():
    rng = np.random.default_rng(1)
    x = np.exp(rng.uniform(-30, 5, 10000)) * rng.choice([1.0, 1.0, 0.0], 10000)
    hi, lo = tf32_split(x)
    assert np.all((hi.view(np.uint32) & 0x1FFF) == 0) and np.all((lo.view(np.uint32) & 0x1FFF) == 0)
    nz = x > 0
    assert np.max(np.abs(hi[nz] / x[nz] - 1)) <= 2.0**-11 * 1.001
    assert np.max(np.abs((hi[nz].astype(np.float64) + lo[nz]) / x[nz] - 1)) <= 2.0**-21
    assert np.all(hi[~nz] == 0) and np.all(lo[~nz] == 0)


def test_grid_layout_and_scaling():
    w = make_workload("cfg2", 4)
    t = build_tables(w.grid, w.emission_model, w.emission_key, w.filters)
    na, nz, nl, nap = t["n_age"], t["n_z"], t["n_lam"], t["n_age_pad"]
    # every metallicity's block of columns starts 32-byte aligned (TMA box origin of a bracket-grouped tile)
    assert nap % 8 == 0 and nap >= na
    assert t["k_pad"] % 32 == 0 and t["k_pad"] >= nap * nz and t["n_chunk"] * 256 // t["n_comp"] >= nl
    att, un = w.emission_model.recipe(w.emission_key)
    comp = att if att.any() else un
    g = (t["gt_hi"].astype(np.float64) + t["gt_lo"]) * t["grid_scale"]
    for (ia, iz, il) in ((0, 0, 0), (50, 12, nl - 1), (17, 5, 1234), (3, 9, 255), (3, 9, 256)):
        assert g[il, iz * nap + ia] == pytest.approx(comp[ia, iz, il], rel=1e-6)
    assert np.all(g[nl:] == 0) and np.all(g[:, nap * nz:] == 0)
    pad_cols = np.concatenate([np.arange(iz * nap + na, (iz + 1) * nap) for iz in range(nz)])
    assert np.all(g[:, pad_cols] == 0)


def test_two_component_layout():
    from synference_b200.parametric import Calzetti2000, PacmanEmission
    w = make_workload("cfg2", 4)
    em = PacmanEmission(grid=w.grid, fesc=0.2, fesc_ly_alpha=0.5, dust_curve=Calzetti2000())
    t = build_tables(w.grid, em, "emergent", w.filters)
    assert t["n_comp"] == 2
    att, un = em.recipe("emergent")
    nap = t["n_age_pad"]
    g = (t["gt_hi"].astype(np.float64) + t["gt_lo"]) * t["grid_scale"]
    il, ia, iz = 700, 20, 4
    chunk, j = divmod(il, 128)
    assert g[chunk * 256 + j, iz * nap + ia] == pytest.approx(att[ia, iz, il], rel=1e-6)
    assert g[chunk * 256 + 128 + j, iz * nap + ia] == pytest.approx(un[ia, iz, il], rel=1e-6)
    # Ly-alpha escape applies to the single bin nearest 1215.67 A only (A5)
    lam = np.asarray(w.grid.lam)
    jl = int(np.argmin(np.abs(lam - 1215.67)))
    ga, gu = O.emission_parts(w.grid.spectra, lam, "emergent", 0.2, 0.5)
    np.testing.assert_allclose(att, ga, rtol=1e-14)
    np.testing.assert_allclose(un, gu, rtol=1e-14)
    full = PacmanEmission(grid=w.grid, fesc=0.2, fesc_ly_alpha=1.0, dust_curve=Calzetti2000()).recipe("emergent")[0]
    diff = np.nonzero(np.any(full != att, axis=(0, 1)))[0]
    assert list(diff) == [jl]


def test_two_screen_layout():
    """BimodalPacmanEmission lowers to (young, old) reprocessed components split at age_pivot, plus a second kappa table."""
    from synference_b200.parametric import BimodalPacmanEmission, Calzetti2000, PacmanEmission
    w = make_workload("cfg2", 4)
    em = BimodalPacmanEmission(grid=w.grid, dust_curve_ism=Calzetti2000(), dust_curve_birth=Calzetti2000(slope=-0.7),
                               age_pivot=7.0, fesc_ly_alpha=0.4)
    t = build_tables(w.grid, em, "emergent", w.filters)
    assert t["n_comp"] == 2 and t["kappa_birth"] is not None and t["kappa"] is not None
    lam = np.asarray(w.grid.lam)
    nl = len(lam)
    np.testing.assert_allclose(t["kappa"][:nl], O.dust_kappa(lam, curve="Calzetti2000"), rtol=1e-6)
    np.testing.assert_allclose(t["kappa_birth"][:nl], O.dust_kappa(lam, curve="Calzetti2000", slope=-0.7), rtol=1e-6)
    assert np.all(t["kappa_birth"][nl:] == 0)
    young, old = em.recipe("emergent")
    whole = PacmanEmission(grid=w.grid, fesc=0.0, fesc_ly_alpha=0.4, dust_curve=Calzetti2000()).recipe("emergent")[0]
    np.testing.assert_array_equal(young + old, whole)
    is_young = np.asarray(w.grid.log10ages) < 7.0
    assert is_young.any() and (~is_young).any()
    assert np.all(young[~is_young] == 0) and np.all(old[is_young] == 0)
    # keys without dust keep the single-sum form
    assert build_tables(w.grid, em, "reprocessed", w.filters)["kappa_birth"] is None
    with pytest.raises(NotImplementedError):
        BimodalPacmanEmission(grid=w.grid, dust_curve_ism=Calzetti2000(), dust_curve_birth=Calzetti2000(), fesc=0.1)


def test_dust_emission_tables():
    """Greybody shape (closed-form normalisation vs the oracle's quadrature), trapezoid weights in frequency, and the per-shift
    filter table: dust_duv[m][f] must equal the filter numerators of the emission's spectrum shifted by m bins."""
    from synference_b200.engine import DUST_W0
    from synference_b200.parametric import Calzetti2000, Greybody, PacmanEmission
    w = make_workload("cfg2", 4)
    em = PacmanEmission(grid=w.grid, fesc=0.1, dust_curve=Calzetti2000(), dust_emission=Greybody(40.0, 1.5))
    t = build_tables(w.grid, em, "total", w.filters)
    lam = np.asarray(w.grid.lam)
    nl = len(lam)
    shape = O.dust_emission_shape(lam, kind="Greybody", temperature=40.0, emissivity=1.5)
    g = t["dust_g"].astype(np.float64) / DUST_W0
    big = shape > 1e-20 * shape.max()
    np.testing.assert_allclose(g[big], shape[big], rtol=1e-6)
    flat = np.ones(nl)
    nu = 2.99792458e18 / lam
    assert float(t["dust_wnu"][:nl].astype(np.float64) @ flat) * DUST_W0 == pytest.approx(nu[0] - nu[-1], rel=1e-6)
    x0, xn = t["x_bin0"], t["x_bins"]
    assert xn == 192 and x0 % 192 == 0 and x0 >= nl and t["n_chunk"] * (256 // t["n_comp"]) >= x0 + xn
    assert np.all(t["dust_wnu"][nl:x0] == 0) and np.all(t["dust_wnu"][x0:x0 + xn] == 1) and np.all(t["dust_wnu"][x0 + xn:] == 0)
    assert np.all(t["dust_g"][:t["igm"]["n_blue"]] == 0)
    duv = t["dust_duv"].reshape(t["dust_m_len"], t["n_filt"], 2).astype(np.float64)
    uv = t["filt_uv"].astype(np.float64)
    for f in (0, 7, t["n_filt"] - 1):
        lo, off = int(t["filt_lo"][f]), int(t["filt_off"][f])
        end = int(t["filt_off"][f + 1]) if f + 1 < t["n_filt"] else uv.shape[0]
        for m in (0, 1, 250, 1400, t["dust_m_len"] - 1):
            num = np.zeros(2)
            for k in range(end - off):
                i = lo - 2 + k - m                   # table entry k is the sample n = lo - 2 + k = i + m
                if 0 <= i < nl:
                    num += t["dust_g"][i].astype(np.float64) * uv[off + k]
            np.testing.assert_allclose(duv[m, f], num, rtol=1e-6, atol=1e-30)
    assert build_tables(w.grid, em, "emergent", w.filters)["dust_wnu"] is None
    with pytest.raises(NotImplementedError):
        PacmanEmission(grid=w.grid, dust_curve=Calzetti2000(), dust_emission=object())


@pytest.mark.parametrize("model", ["one_component", "two_components", "two_screens"])
def test_absorbed_energy_pseudo_bins_reproduce_the_sum_over_the_axis(model):
    """The pseudo-bin rows appended to the grid (engine._energy_pseudo_bins): for random (age, Z) weights and optical depths up
    to the stated range, sum_j [W . rows_j] (1 - exp(-tau node_j)) equals the absorbed energy summed over the whole axis,
    sum_i wnu_i [W . grid_i] (1 - exp(-tau kappa_i)), to < 1e-6 -- evaluated from the very tables the kernel reads (float32
    kappa / nodes, TF32 hi + lo grid rows)."""
    from synference_b200.parametric import BimodalPacmanEmission, Calzetti2000, Greybody, PacmanEmission
    w = make_workload("cfg2", 4)
    if model == "two_screens":
        em = BimodalPacmanEmission(grid=w.grid, dust_curve_ism=Calzetti2000(), dust_curve_birth=Calzetti2000(), age_pivot=7.0,
                                   dust_emission_ism=Greybody(40.0, 1.5), dust_emission_birth=Greybody(40.0, 1.5))
    else:
        em = PacmanEmission(grid=w.grid, fesc=0.1 if model == "two_components" else 0.0, dust_curve=Calzetti2000(),
                            dust_emission=Greybody(40.0, 1.5))
    t = build_tables(w.grid, em, "total", w.filters)
    nl, nc, x0, xn = t["n_lam"], t["n_comp"], t["x_bin0"], t["x_bins"]
    assert xn == 192 and t["x_tau_max"] > 9.0
    lch = 256 // nc
    g = (t["gt_hi"].astype(np.float64) + t["gt_lo"].astype(np.float64)).reshape(t["n_chunk"], nc, lch, t["k_pad"])
    kap, wnu = t["kappa"].astype(np.float64), t["dust_wnu"].astype(np.float64)
    rng = np.random.default_rng(2)
    absorbing = [0, 1] if model == "two_screens" else [0]
    for ci in range(nc):
        rows = g[:, ci].reshape(t["n_chunk"] * lch, t["k_pad"])
        if ci not in absorbing:
            assert np.all(rows[x0:x0 + xn] == 0)
            continue
        for trial in range(6):
            wgt = np.zeros(t["k_pad"])
            pick = rng.choice(t["n_age_pad"] * t["n_z"], 5 if trial else 1, replace=False)     # a few bins, as a real SFH gives
            wgt[pick] = rng.uniform(0.1, 1.0, pick.size)
            light = rows @ wgt
            for tau in (0.003, 0.3, 1.0, 3.0, 9.0):
                exact = np.sum(wnu[:nl] * light[:nl] * -np.expm1(-tau * kap[:nl]))
                pseudo = np.sum(light[x0:x0 + xn] * -np.expm1(-tau * kap[x0:x0 + xn]))
                scale = np.sum(np.abs(wnu[:nl] * light[:nl] * np.expm1(-tau * kap[:nl])))
                assert abs(pseudo - exact) <= 1e-6 * scale, (model, ci, trial, tau, pseudo, exact)
    if model == "two_screens":
        np.testing.assert_allclose(t["kappa_birth"][x0:x0 + xn], t["x_birth_ratio"] * t["kappa"][x0:x0 + xn], rtol=1e-6)
    # a per-galaxy dust curve has no single kappa axis: no pseudo-bins
    em_pg = PacmanEmission(grid=w.grid, dust_curve=Calzetti2000(slope="slope", ampl="ampl"), dust_emission=Greybody(40.0, 1.5))
    assert build_tables(w.grid, em_pg, "total", w.filters)["x_bins"] == 0


def test_non_geometric_axis_is_rejected():
    with pytest.raises(ValueError, match="constant-R"):
        geometric_ratio(np.linspace(1000.0, 2000.0, 100))


def test_dust_curve_matches_oracle():
    from synference_b200.parametric import Calzetti2000, PowerLaw
    lam = O.constant_r_grid(600.0, 3e5, 300)
    np.testing.assert_allclose(Calzetti2000().get_tau(lam), O.dust_kappa(lam), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(Calzetti2000(slope=-0.3, ampl=2.0).get_tau(lam),
                               O.dust_kappa(lam, slope=-0.3, ampl=2.0), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(PowerLaw(-0.7).get_tau(lam), O.dust_kappa(lam, curve="PowerLaw", slope=-0.7), rtol=1e-12)
    assert Calzetti2000().get_tau(np.array([5500.0]))[0] == pytest.approx(1.0, abs=1e-12)


def test_weights_adapter_order():
    w = make_workload("cfg3", 8)
    W = A.weights_matrix(w.params, w.grid.log10ages, w.grid.metallicity)
    assert W.shape == (8, 663) and np.allclose(W.sum(1), 1.0)
    na = 51
    assert np.all(W.reshape(8, 13, na)[:, :, -1] == 0)   # last age bin receives no mass
