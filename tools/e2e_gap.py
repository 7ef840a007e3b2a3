"""Why the streaming host entry is ~10 % slower than the device-resident step: stage times of a batch whose kernels ran while
the neighbouring batches' copies were in flight, against a lone device-resident batch, and the wall clock of 20 streamed steps
with 2 and with 3 batches submitted ahead."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from synference_b200.configs import make_workload
from synference_b200.engine import SynthEngine
n = 1_000_000
w = make_workload("cfg2", n)
eng = SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters, max_batch=n)
pinned = bench.raw_draw_params(w)
for k in ("redshift", "log_mass", "tau_v", "zd_value", "zd_sigma", "sfh_rows"):
    a = getattr(pinned, k)
    if a is not None:
        setattr(pinned, k, torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).pin_memory().numpy())
outs = [torch.empty((n, eng.n_filt), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
st = np.zeros(3, dtype=np.float32)
def stages():
    eng.lib.sb2_last_stage_ms(eng._h, st.ctypes.data_as(C.POINTER(C.c_float)))
    return [round(float(x), 3) for x in st]
dp = eng.to_device(w.params)
flux = torch.empty((n, eng.n_filt), dtype=torch.float32, device="cuda")
for _ in range(5):
    eng.photometry_device(dp, flux_base=flux)
torch.cuda.synchronize()
print("device-resident stages [sort, weights+igm, contraction]:", stages())
for rep in range(2):
    tk = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(20):
        if len(tk) == 2:
            eng.wait(tk.pop(0))
        tk.append(eng.submit(pinned, outs[i & 1], scaled=False, slot=i & 1, transport="f32"))
    for t in tk:
        eng.wait(t)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"streamed, 20 steps: {dt * 50:.3f} ms per step; stages of the last batch:", stages())
# host time of one submit call (everything it enqueues)
torch.cuda.synchronize(); t0 = time.perf_counter()
t = eng.submit(pinned, outs[0], scaled=False, slot=0, transport="f32")
print(f"one submit call returns after {1e3 * (time.perf_counter() - t0):.3f} ms of host time")
eng.wait(t)
