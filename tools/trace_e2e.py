"""Per-batch timeline of the host entry (SB2_TRACE=1: every submit is synchronised and prints when its H2D copy, kernels and D2H
copy finished): 1M galaxies of cfg2, float32 parameter transport.  Serial phases on one B200: H2D 1.6 ms (32 MB from pageable
numpy arrays), kernels 3.1 ms, D2H 1.4 ms (80 MB into pinned memory)."""
import os, sys, numpy as np, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SB2_TRACE"] = "1"
from synference_b200.configs import make_workload
from synference_b200.engine import SynthEngine
import bench
n = 1_000_000
w = make_workload("cfg2", n)
eng = SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters, max_batch=n)
p32 = bench.raw_draw_params(w)
outs = [torch.empty((n, eng.n_filt), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
tk = []
t0 = time.perf_counter()
for i in range(8):
    if len(tk) == 2:
        eng.wait(tk.pop(0))
    tk.append(eng.submit(p32, outs[i & 1], scaled=False, slot=i & 1, transport="f32"))
for t in tk:
    eng.wait(t)
print("total s", time.perf_counter() - t0)
