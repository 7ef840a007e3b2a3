"""cProfile of one-galaxy GalaxySimulator calls (host-side overheads around the 15 kernel launches)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synference_b200 as S
from synference_b200.synthetic import synthetic_grid
raw = S.FilterCollection(filter_codes=["JWST/NIRCam.F070W", "JWST/NIRCam.F090W", "JWST/NIRCam.F115W", "JWST/NIRCam.F200W",
                                       "JWST/NIRCam.F277W", "JWST/NIRCam.F356W", "JWST/NIRCam.F444W"])
lam = S.generate_constant_R(R=300, auto_start_stop=True, filterset=raw, max_redshift=15)
inst = S.Instrument("JWST", filters=S.FilterCollection(filter_codes=raw.filter_codes, new_lam=lam))
grid = synthetic_grid(lam)
em = S.PacmanEmission(grid=grid, fesc=0.1, fesc_ly_alpha=0.1, dust_curve=S.Calzetti2000())
sim = S.GalaxySimulator(sfh_model=S.SFH.LogNormal, zdist_model=S.ZDist.DeltaConstant, grid=grid, instrument=inst,
                        emission_model=em, emission_model_key="emergent", out_flux_unit="nJy", ignore_scatter=True,
                        param_units={"peak_age": S.Myr, "max_age": S.Myr},
                        param_order=["redshift", "log_mass", "tau", "peak_age", "max_age", "log10metallicity", "tau_v"])
v = np.array([7.0, 9.5, 0.5, 100.0, 300.0, -1.0, 0.2])
for _ in range(20):
    sim(v)
pr = cProfile.Profile(); pr.enable()
for _ in range(500):
    sim(v)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
