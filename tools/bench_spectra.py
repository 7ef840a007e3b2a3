"""Full-wavelength write path (cfg5's device side): spectra + photometry for a device-resident batch."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from synference_b200.configs import make_workload
from synference_b200.engine import SynthEngine
n = int(os.environ.get("SPEC_N", "262144"))
w = make_workload("cfg2", n)
eng = SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters, max_batch=n)
dp = eng.to_device(w.params)
flux = torch.empty((n, eng.n_filt), dtype=torch.float32, device="cuda")
spec = torch.empty((n, eng.n_lam), dtype=torch.float32, device="cuda")
for _ in range(3):
    eng.photometry_device(dp, flux_base=flux, spectra=spec)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5):
    eng.photometry_device(dp, flux_base=flux, spectra=spec)
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / 5
print(json.dumps({"galaxies": n, "n_lam": eng.n_lam, "ms_per_step": ms, "galaxies_per_s": n / ms * 1e3,
                  "spectra_write_gbs": n * eng.n_lam * 4 / ms / 1e6}))
