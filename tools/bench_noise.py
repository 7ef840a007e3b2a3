"""Depth-noise + AB-feature kernel against the HBM roofline (4*N_f read + 8*N_f written per row for the feature
rows with float32 fluxes; 8*N_f + 8*N_f with float64 fluxes).  CUDA events, inputs >> L2."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from synference_b200.engine import depth_noise_features
from synference_b200.features import depths_to_sigma_njy

n, nf = int(os.environ.get("NOISE_ROWS", "10000000")), 20
sigma = depths_to_sigma_njy(np.full(nf, 29.0), 5.0, nf)
peak = 6539.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
for dt, in_bytes in ((torch.float64, 8), (torch.float32, 4)):
    flux = torch.rand((n, nf), dtype=dt, device="cuda") * 100 + 1
    for e in range(3):
        depth_noise_features(flux, sigma, seed=42, epoch=e, want_flux=False, want_features=True)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    t0.record()
    for e in range(reps):
        depth_noise_features(flux, sigma, seed=42, epoch=3 + e, want_flux=False, want_features=True)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / reps          # (includes the allocation of the output tensor by torch's caching allocator)
    bytes_alg = n * nf * (in_bytes + 8)      # flux in, (mag, mag_err) float32 pairs out
    print(json.dumps({"kernel": "depth_noise_feat_kernel (Philox + AB features)", "flux_dtype": str(dt), "rows": n, "n_filt": nf,
                      "ms": ms, "rows_per_s": n / ms * 1e3, "achieved_gbs": bytes_alg / ms / 1e6, "peak_gbs": peak,
                      "frac": bytes_alg / ms / 1e6 / peak, "bytes_per_row": bytes_alg / n}))
    del flux
