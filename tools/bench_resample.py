"""Spectroscopic resample kernel (cfg 5 shapes: 3790-bin constant-R model axis, 1000 PRISM-like pixels 0.6-5.3 um, z 0-10)
against the HBM roofline.  Algorithmic bytes per galaxy = 4 * (model bins under the observed window) read + 4 * n_px
written (4000 B written per galaxy).  CUDA events on the launching stream, inputs (4 GB) larger than L2.
A bounded CPU sample of the float64 oracle (the reference's per-galaxy Python loop restated) is timed beside it."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from synference_b200.spectral import SpectrumResampler

n = int(os.environ.get("RESAMPLE_N", "262144"))
tw = 0.04 * (1 + 0.5 / 300) ** np.arange(3790)
ow = np.linspace(0.6, 5.3, 1000)
rw = np.linspace(0.55, 5.4, 80)
rr = 30.0 + 270.0 * ((rw - 0.55) / 4.85) ** 1.3
rng = np.random.default_rng(42)
z = rng.uniform(0.0, 10.0, n)
spec = torch.rand((n, tw.size), dtype=torch.float32, device="cuda") + 0.5
zd = torch.as_tensor(z, device="cuda")
plan = SpectrumResampler(tw, ow, rw, rr)
for _ in range(3):
    out = plan.transform(spec, zd)
torch.cuda.synchronize()
reps = 10
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(reps):
    out = plan.transform(spec, zd)
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / reps
edges = np.concatenate([[tw[0] - (tw[1] - tw[0]) / 2], (tw[1:] + tw[:-1]) / 2, [tw[-1] + (tw[-1] - tw[-2]) / 2]])
lo = np.searchsorted(edges, (ow[0] - (ow[1] - ow[0]) / 2) / (1 + z), side="right") - 1
hi = np.searchsorted(edges, (ow[-1] + (ow[-1] - ow[-2]) / 2) / (1 + z), side="left") - 1
bins = np.clip(hi, 0, tw.size - 1) - np.clip(lo, 0, tw.size - 1) + 1
bytes_alg = float(4 * bins.sum() + 4 * ow.size * n)
peak = 6539.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
res = {"kernel": "resample_kernel (redshift + variable-width Gaussian + flux-conserving rebin)", "galaxies": n,
       "n_lam": int(tw.size), "n_px": int(ow.size), "ms": ms, "galaxies_per_s": n / ms * 1e3,
       "achieved_gbs": bytes_alg / ms / 1e6, "peak_gbs": peak, "frac": bytes_alg / ms / 1e6 / peak,
       "bytes_per_galaxy": bytes_alg / n, "full_row_bytes_per_galaxy": 4 * (tw.size + ow.size)}
if not os.environ.get("RESAMPLE_NO_CPU"):
    from oracle import oracle as O
    m = 24
    sp = spec[:m].cpu().numpy().astype(np.float64)
    t = time.perf_counter()
    for i in range(m):
        O.transform_spectrum(tw, sp[i], z[i], ow, rw, rr)
    dt = time.perf_counter() - t
    res["cpu_baseline"] = {"value": m / dt, "unit": "galaxies/s", "cores": 1, "kind": "port",
                           "sample": f"{m} galaxies, float64 numpy restatement of utils.py:129-254 (oracle/oracle.py), {dt:.1f} s"}
print(json.dumps(res))
