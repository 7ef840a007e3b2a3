"""Small driver for ncu captures: a few steps of the device-resident hot path (no CPU baseline, no e2e)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from synference_b200.configs import make_workload
from synference_b200.engine import SynthEngine
n = int(os.environ.get("PROF_N", "262144")); steps = int(os.environ.get("PROF_STEPS", "3"))
w = make_workload(os.environ.get("PROF_CFG", "cfg2"), n)
eng = SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters, max_batch=n)
dp = eng.to_device(w.params)
flux = torch.empty((n, eng.n_filt), dtype=torch.float32, device="cuda")
for _ in range(steps):
    eng.photometry_device(dp, flux_base=flux)
torch.cuda.synchronize()
print("ok", float(flux.sum()))
