"""cProfile of create_mock_library for 1 M galaxies (host side around the kernels)."""
import cProfile, os, pstats, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synference_b200 as S
from synference_b200.synthetic import synthetic_grid
n = 1000000
raw = S.FilterCollection(filter_codes=["JWST/NIRCam.F070W", "JWST/NIRCam.F090W", "JWST/NIRCam.F115W", "JWST/NIRCam.F150W",
                                       "JWST/NIRCam.F200W", "JWST/NIRCam.F277W", "JWST/NIRCam.F356W", "JWST/NIRCam.F444W"])
lam = S.generate_constant_R(R=300, auto_start_stop=True, filterset=raw, max_redshift=15)
inst = S.Instrument("JWST", filters=S.FilterCollection(filter_codes=raw.filter_codes, new_lam=lam))
grid = synthetic_grid(lam)
em = S.PacmanEmission(grid=grid, fesc=0.1, fesc_ly_alpha=0.1, dust_curve=S.Calzetti2000())
d = S.draw_from_hypercube({"redshift": (0.01, 10), "masses": (5, 11), "tau_v": (0, 2), "peak_age": (0, 0.99),
                           "tau": (0.1, 1.5), "log_zmet": (-3, -1.39)}, N=n, rng=42)
sfhs, _ = S.generate_sfh_basis(S.SFH.LogNormal, ["tau", "peak_age_norm"], np.vstack((d["tau"], d["peak_age"])).T,
                               redshifts=np.array(d["redshift"]), max_redshift=20)
zds = [S.ZDist.DeltaConstant(log10metallicity=z) for z in d["log_zmet"]]
basis = S.GalaxyBasis("api_basis", d["redshift"], grid, em, sfhs, zds, galaxy_params={"tau_v": d["tau_v"]}, instrument=inst,
                      redshift_dependent_sfh=True, build_library=False)
pr = cProfile.Profile(); pr.enable()
basis.create_mock_library(log_stellar_masses=np.asarray(d["masses"], dtype=float), emission_model_key="emergent",
                          out_name="api_lib", out_dir=tempfile.mkdtemp(), overwrite=True, batch_size=250000)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
