"""Dense-K contraction (cfg3) against the number of CTAs launched: each CTA re-reads its tile's weights for every wavelength
chunk, so the CTAs' live weight tiles compete for the L2.  Prints the stage times per SB2_DENSE_GRID value."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "one":
    import numpy as np, ctypes as C, torch
    from synference_b200.configs import make_workload
    from synference_b200.engine import SynthEngine
    n = int(os.environ.get("PROF_N", "500000"))
    w = make_workload("cfg3", n)
    eng = SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters, max_batch=n)
    dp = eng.to_device(w.params)
    flux = torch.empty((n, eng.n_filt), dtype=torch.float32, device="cuda")
    st = np.zeros(3, dtype=np.float32)
    ts = []
    for i in range(5):
        eng.photometry_device(dp, flux_base=flux)
        eng.lib.sb2_last_stage_ms(eng._h, st.ctypes.data_as(C.POINTER(C.c_float)))
        if i >= 2:
            ts.append(st.copy())
    print(json.dumps({"grid": os.environ.get("SB2_DENSE_GRID", "all"), "galaxies": n, "stage_ms": [float(x) for x in np.mean(ts, 0)]}))
else:
    for g in os.environ.get("GRIDS", "0,132,120,112,104,96,74").split(","):
        env = dict(os.environ, SB2_DENSE_GRID=g)
        r = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True, timeout=300)
        print(g, r.stdout.strip()[-200:] or r.stderr.strip()[-300:], flush=True)
