"""GalaxySimulator.simulate latency and throughput (the reference builds one Synthesizer galaxy per call,
ms to 100 ms each; callers loop per sample, sbi_runner.py:7659-7664).  Wall clock per call for batches of 1 ... 100 k
parameter vectors through the public API (host arrays in, host arrays out)."""
import json, os, sys, time
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
rng = np.random.default_rng(0)
res = {}
for n in (1, 100, 10_000, 100_000):
    p = np.column_stack([rng.uniform(0.5, 10, n), rng.uniform(8, 11, n), rng.uniform(0.2, 1.5, n), rng.uniform(10, 200, n),
                         rng.uniform(250, 400, n), rng.uniform(-3, -1.4, n), rng.uniform(0, 2, n)])
    arg = p[0] if n == 1 else p
    for _ in range(3):
        sim(arg)
    reps = 200 if n == 1 else 20
    t0 = time.perf_counter()
    for _ in range(reps):
        out = sim(arg)
    dt = (time.perf_counter() - t0) / reps
    res[str(n)] = {"ms_per_call": dt * 1e3, "galaxies_per_s": n / dt}
print(json.dumps({"api": "GalaxySimulator.__call__ (host arrays in and out)", "n_filt": 7, "batches": res}))
