"""Device-resident step time of the emission-model variants on cfg 2's population (1 M galaxies, 20 filters): the default
single-screen model, and the 'extras' instantiations (per-galaxy dust shape, two screens, dust emission / key 'total')."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from synference_b200.configs import make_workload
from synference_b200.engine import SynthEngine
from synference_b200.parametric import BimodalPacmanEmission, Calzetti2000, Greybody, PacmanEmission

n = int(os.environ.get("VAR_N", "1000000"))
w = make_workload("cfg2", n)
rng = np.random.default_rng(1)
cases = {
    "cfg2 default (fesc=0, emergent)": (w.emission_model, w.emission_key, {}),
    "fesc=0.1 two components (emergent)": (PacmanEmission(grid=w.grid, fesc=0.1, fesc_ly_alpha=0.5, dust_curve=Calzetti2000()), "emergent", {}),
    "per-galaxy slope + bump (emergent)": (PacmanEmission(grid=w.grid, fesc=0.0, dust_curve=Calzetti2000(slope="slope", ampl="ampl")),
                                           "emergent", dict(dust_slope=rng.uniform(-1, 0.4, n), dust_ampl=rng.uniform(0, 5, n))),
    "two screens (emergent)": (BimodalPacmanEmission(grid=w.grid, dust_curve_ism=Calzetti2000(), dust_curve_birth=Calzetti2000(),
                                                     age_pivot=7.0), "emergent", dict(tau_v_birth=rng.uniform(0, 2, n))),
    "dust emission (total, fesc=0.1)": (PacmanEmission(grid=w.grid, fesc=0.1, fesc_ly_alpha=0.5, dust_curve=Calzetti2000(),
                                                       dust_emission=Greybody(40.0, 1.5)), "total", {}),
    "two screens + dust emission (total)": (BimodalPacmanEmission(grid=w.grid, dust_curve_ism=Calzetti2000(), dust_curve_birth=Calzetti2000(),
                                                                  age_pivot=7.0, dust_emission_ism=Greybody(40.0, 1.5),
                                                                  dust_emission_birth=Greybody(40.0, 1.5)), "total",
                                            dict(tau_v_birth=rng.uniform(0, 2, n))),
}
only = os.environ.get("VAR_ONLY")          # substring of the case names to run (an ncu capture wants one)
out = {}
for name, (em, key, extra) in cases.items():
    if only and only not in name:
        continue
    eng = SynthEngine(w.grid, em, key, w.filters, max_batch=n)
    p = w.params.slice(slice(0, n))
    for k, v in extra.items():
        setattr(p, k, v)
    dpar = eng.to_device(p)
    flux = torch.empty((n, eng.n_filt), dtype=torch.float32, device="cuda")
    for _ in range(3):
        eng.photometry_device(dpar, flux_base=flux)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        eng.photometry_device(dpar, flux_base=flux)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    out[name] = {"ms_per_step": ms, "galaxies_per_s": n / ms * 1e3, "n_comp": eng.n_comp}
    eng.close(); del eng, dpar, flux
    torch.cuda.empty_cache()
print(json.dumps(out))
