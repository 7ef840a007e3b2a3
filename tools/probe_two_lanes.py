"""Probe: do the prep kernels of one batch overlap the contraction kernel of another when two models run on two streams?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from synference_b200.configs import make_workload
from synference_b200.engine import SynthEngine

n = int(os.environ.get("PROBE_N", "1000000"))
w = make_workload("cfg2", n)
lanes = int(os.environ.get("PROBE_LANES", "2"))
engs = [SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters, max_batch=n) for _ in range(lanes)]
dpars = [e.to_device(w.params) for e in engs]
outs = [torch.empty((n, e.n_filt), dtype=torch.float32, device="cuda") for e in engs]
streams = [torch.cuda.Stream() for _ in engs]


def run(k_steps, use_lanes):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(k_steps):
        i = k % use_lanes
        with torch.cuda.stream(streams[i]):
            engs[i].photometry_device(dpars[i], flux_base=outs[i])
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k_steps * 1e3


for use in (1, lanes):
    run(4, use)
    ms = run(20, use)
    print(f"lanes={use}: {ms:.3f} ms/step  {n / ms * 1e3 / 1e6:.1f} M gal/s", flush=True)
assert torch.equal(outs[0], outs[-1])
