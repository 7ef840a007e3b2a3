"""cProfile of create_mock_library for 1 M galaxies of the bench workload (host side around the kernels)."""
import cProfile, os, pstats, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synference_b200 as S
from synference_b200.configs import make_workload
n = int(os.environ.get("LIB_N", "1000000"))
w0 = make_workload("cfg2", 64)
d = S.draw_from_hypercube({"log_stellar_mass": (8.0, 12.0), "redshift": (0.01, 10.0), "log_zmet": (-4.0, -1.4),
                           "peak_age_norm": (0.0, 0.99), "tau": (0.2, 2.0), "tau_v": (0.0, 3.0)}, N=n, rng=42)
sfhs, _ = S.generate_sfh_basis(S.SFH.LogNormal, ["tau", "peak_age_norm"], np.vstack((d["tau"], d["peak_age_norm"])).T,
                               redshifts=np.array(d["redshift"]), max_redshift=20)
zds = S.generate_metallicity_distribution(S.ZDist.DeltaConstant, log10metallicity=np.asarray(d["log_zmet"], dtype=float))
basis = S.GalaxyBasis("api_basis", d["redshift"], w0.grid, w0.emission_model, sfhs, zds, galaxy_params={"tau_v": d["tau_v"]},
                      instrument=w0.instrument, redshift_dependent_sfh=True, build_library=False)
basis._engine(w0.emission_key, max_batch=250_000)
pr = cProfile.Profile(); pr.enable()
basis.create_mock_library(log_stellar_masses=np.asarray(d["log_stellar_mass"], dtype=float), emission_model_key=w0.emission_key,
                          out_name="api_lib", out_dir=tempfile.mkdtemp(), overwrite=True, batch_size=250000)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(40)
