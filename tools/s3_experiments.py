"""Bottleneck experiments for synth3_kernel (library built with -DSB2_EXPERIMENTS): contraction-stage time per SB2_DBG mask."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "one":
    import numpy as np, ctypes as C, torch
    from synference_b200.configs import make_workload
    from synference_b200.engine import SynthEngine
    n = int(os.environ.get("PROF_N", "1000000"))
    w = make_workload("cfg2", n)
    eng = SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters, max_batch=n)
    dp = eng.to_device(w.params)
    flux = torch.empty((n, eng.n_filt), dtype=torch.float32, device="cuda")
    st = np.zeros(3, dtype=np.float32)
    ts = []
    for i in range(6):
        eng.photometry_device(dp, flux_base=flux)
        eng.lib.sb2_last_stage_ms(eng._h, st.ctypes.data_as(C.POINTER(C.c_float)))
        if i >= 2:
            ts.append(st.copy())
    print(json.dumps({"dbg": os.environ.get("SB2_DBG", "0"), "stage_ms": [float(x) for x in np.mean(ts, 0)]}))
else:
    for mask in [int(x) for x in os.environ.get("MASKS", "0,2048,2560,3584").split(",")]:
        env = dict(os.environ, SB2_DBG=str(mask))
        r = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True, timeout=300)
        print(mask, r.stdout.strip()[-200:] or r.stderr.strip()[-300:], flush=True)
