"""Where the first call's second goes: CUDA context, table lowering, sb2_model_create (uploads, workspace, TMA maps), first step."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t = {}
t0 = time.perf_counter()
import numpy as np, torch
t["import numpy+torch"] = time.perf_counter() - t0
t0 = time.perf_counter(); torch.zeros(1, device="cuda"); torch.cuda.synchronize(); t["cuda context (torch)"] = time.perf_counter() - t0
t0 = time.perf_counter()
from synference_b200.configs import make_workload
from synference_b200.engine import SynthEngine, build_tables
w = make_workload("cfg2", 100_000)
t["import package + workload objects"] = time.perf_counter() - t0
t0 = time.perf_counter(); build_tables(w.grid, w.emission_model, w.emission_key, w.filters); t["build_tables (host numpy)"] = time.perf_counter() - t0
for mb in (40_000, 250_000, 1 << 20):
    t0 = time.perf_counter()
    eng = SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters, max_batch=mb)
    torch.cuda.synchronize()
    t[f"SynthEngine(max_batch={mb}) incl. build_tables"] = time.perf_counter() - t0
    p = w.params.slice(slice(0, min(mb, 100_000)))
    t0 = time.perf_counter(); eng.photometry(p, scaled=False); t[f"first photometry call ({len(p)} galaxies, max_batch={mb})"] = time.perf_counter() - t0
    t0 = time.perf_counter(); eng.photometry(p, scaled=False); t[f"second photometry call (max_batch={mb})"] = time.perf_counter() - t0
    eng.close()
print(json.dumps({k: round(v, 4) for k, v in t.items()}, indent=1))
