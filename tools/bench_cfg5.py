"""BASELINE cfg 5 on one GPU: cfg 2 physics with photometry AND ~1000-pixel PRISM-like spectra out.  One step = one batch
through contraction kernel (full-wavelength output kept on the device) + resample kernel; device-resident, CUDA events."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from synference_b200.configs import make_workload
from synference_b200.engine import SynthEngine
from synference_b200.spectral import SpectrumResampler

n = int(os.environ.get("CFG5_N", "262144"))
w = make_workload("cfg2", n)
eng = SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters, max_batch=n)
dpar = eng.to_device(w.params)
spec = torch.empty((n, eng.n_lam), dtype=torch.float32, device="cuda")
flux = torch.empty((n, eng.n_filt), dtype=torch.float32, device="cuda")
ow = np.linspace(0.6, 5.3, 1000)
rw = np.linspace(0.55, 5.4, 80)
rr = 30.0 + 270.0 * ((rw - 0.55) / 4.85) ** 1.3
plan = SpectrumResampler(np.asarray(w.grid.lam) * 1e-4, ow, rw, rr)
zd = torch.as_tensor(np.asarray(w.params.redshift, dtype=np.float64), device="cuda")


def step():
    eng.photometry_device(dpar, flux_base=flux, spectra=spec)
    return plan.transform(spec, zd)


for _ in range(3):
    px = step()
torch.cuda.synchronize()
reps = 10
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(reps):
    px = step()
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / reps
print(json.dumps({"workload": "cfg5 on one GPU: cfg2 physics, photometry + 1000-pixel spectra out", "galaxies": n,
                  "ms_per_step": ms, "galaxies_per_s": n / ms * 1e3, "resample_ms": plan.last_ms(),
                  "output_bytes_per_galaxy": 4 * (1000 + eng.n_filt)}))
