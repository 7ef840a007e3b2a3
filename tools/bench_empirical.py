"""Empirical-noise kernel (GeneralEmpiricalUncertaintyModel.apply_noise for all filters and rows) against the HBM roofline:
8 B read + 16 B written per (filter, row) in the Philox mode.  CUDA events, working set (4.8 GB) >> L2.  The scipy-based host
class is timed beside it on a bounded sample (one filter row)."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synference_b200 as S
from synference_b200 import _capi

n, nf = int(os.environ.get("EMP_ROWS", "10000000")), 20
centers = np.linspace(20.0, 31.0, 20)
kw = dict(flux_unit="AB", already_binned=True, bin_median_errors=0.02 + np.exp((centers - 28) / 1.5),
          bin_std_errors=0.005 + 0.2 * np.exp((centers - 28) / 1.5), return_noise=True)
variants = {"plain": {}, "upper_limits+observed": dict(upper_limits=True, treat_as_upper_limits_below=3.0, error_type="observed")}
peak = 6539.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
out = {"kernel": "empirical_noise_kernel", "rows": n, "n_filt": nf, "peak_gbs": peak, "bytes_per_element": 24}
lib = _capi.load()
flux = torch.rand((nf, n), dtype=torch.float64, device="cuda") * 7 + 22
of, os_ = torch.empty_like(flux), torch.empty_like(flux)
for name, extra in variants.items():
    mod = S.GeneralEmpiricalUncertaintyModel(centers, None, **kw, **extra)
    if extra:
        mod.upper_limit_value = 29.0
    models = (_capi.EmpiricalModel * nf)(*[mod.device_model("AB", "AB") for _ in range(nf)])
    st = torch.cuda.current_stream().cuda_stream
    run = lambda e: _capi.check(lib.sb2_empirical_noise(C.c_void_p(flux.data_ptr()), n, nf, models, None, 42, e,
                                                        C.c_void_p(of.data_ptr()), C.c_void_p(os_.data_ptr()), C.c_void_p(st)), "emp")
    for e in range(10):          # (the first launches of a fresh process run at ramping clocks: 1.9 ms instead of 1.34)
        run(e)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps, blocks = 20, []
    for b in range(3):
        t0.record()
        for e in range(reps):
            run(10 + b * reps + e)
        t1.record(); torch.cuda.synchronize()
        blocks.append(t0.elapsed_time(t1) / reps)
    ms = sorted(blocks)[1]       # median of three blocks of 20 launches
    out[name] = {"ms": ms, "elements_per_s": n * nf / ms * 1e3, "rows_per_s": n / ms * 1e3,
                 "achieved_gbs": 24.0 * n * nf / ms / 1e6, "frac": 24.0 * n * nf / ms / 1e6 / peak}
    m = 200000
    x = flux[0, :m].cpu().numpy()
    t = time.perf_counter(); mod.apply_noise(x); dt = time.perf_counter() - t
    out[name]["cpu_baseline"] = {"value": m / dt, "unit": "elements/s", "cores": 1, "kind": "port",
                                 "sample": f"{m} elements of one filter row, numpy/scipy host class, {dt:.2f} s"}
print(json.dumps(out))
