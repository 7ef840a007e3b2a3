"""Depth-noise + AB-feature kernel against the HBM roofline (4*N_f read + 8*N_f written per row for the feature
rows; this build reads float64 fluxes, so 8*N_f + 8*N_f = 320 B/row at 20 filters).  CUDA events, L2 flushed by size."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from synference_b200.engine import depth_noise_features
from synference_b200.features import depths_to_sigma_njy

n, nf = int(os.environ.get("NOISE_ROWS", "10000000")), 20
flux = torch.rand((n, nf), dtype=torch.float64, device="cuda") * 100 + 1
sigma = depths_to_sigma_njy(np.full(nf, 29.0), 5.0, nf)
for e in range(3):
    depth_noise_features(flux, sigma, seed=42, epoch=e, want_flux=False, want_features=True)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
t0.record()
for e in range(reps):
    depth_noise_features(flux, sigma, seed=42, epoch=3 + e, want_flux=False, want_features=True)
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / reps
bytes_alg = n * nf * (8 + 8)          # float64 flux in, (mag, mag_err) float32 pairs out
peak = 6539.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(json.dumps({"kernel": "depth_noise_kernel (Philox + AB features)", "rows": n, "n_filt": nf, "ms": ms,
                  "rows_per_s": n / ms * 1e3, "achieved_gbs": bytes_alg / ms / 1e6, "peak_gbs": peak,
                  "frac": bytes_alg / ms / 1e6 / peak, "bytes_per_row": bytes_alg / n}))
