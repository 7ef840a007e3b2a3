"""GPU diagnostic: weights / spectra / flux parity vs the oracle with error statistics (not a test)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from synference_b200.configs import make_workload
from synference_b200.engine import SynthEngine
from synference_b200 import igm as I
from oracle import adapter as A, oracle as O

def relerr(a, b, floor=1e-30):
    scale = np.abs(b).max(axis=-1, keepdims=True)
    mask = np.abs(b) > floor * scale
    r = np.zeros_like(b); r[mask] = np.abs(a[mask] - b[mask]) / np.abs(b[mask])
    return r, mask

names = sys.argv[1:] or ["cfg1", "cfg2", "cfg3"]
N = int(os.environ.get("DIAG_N", "300"))
for name in names:
    w = make_workload(name, N)
    t0 = time.time()
    eng = SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters, max_batch=4096)
    print(f"[{name}] engine up in {time.time()-t0:.2f}s n_lam={eng.n_lam} n_comp={eng.n_comp} n_filt={eng.n_filt}", flush=True)
    W = eng.weights(w.params)
    Wo = A.weights_matrix(w.params, w.grid.log10ages, w.grid.metallicity)
    print(f"[{name}] weights max abs err {np.abs(W-Wo).max():.3e}  rowsum dev {np.abs(W.sum(1)-1).max():.2e}", flush=True)
    gals = A.galaxies_from_params(w.params)
    filt = [(f.lam, f.t) for f in w.filters]
    dust = dict(curve="Calzetti2000") if w.emission_model.dust_curve is not None else None
    t0 = time.time()
    fo, so = O.synthesize(gals, w.grid.log10ages, w.grid.metallicity, np.asarray(w.grid.lam), w.grid.spectra, filt,
                          key=w.emission_key, fesc=w.emission_model.fesc, fesc_ly_alpha=w.emission_model.fesc_ly_alpha,
                          dust=dust, igm=(I.INOUE14_LAF, I.INOUE14_DLA), return_spectra=True)
    print(f"[{name}] oracle {N} galaxies in {time.time()-t0:.2f}s", flush=True)
    spec = eng.spectra(w.params).astype(np.float64)
    r, mask = relerr(spec, so, 1e-25)
    print(f"[{name}] spectra rel err: max {r.max():.3e}  p99.9 {np.quantile(r[mask], 0.999):.3e} median {np.median(r[mask]):.3e}", flush=True)
    j = np.unravel_index(np.argmax(r), r.shape); print("   worst at", j, spec[j], so[j], 'z=', w.params.redshift[j[0]], 'lam=', np.asarray(w.grid.lam)[j[1]])
    fb = eng.photometry(w.params, scaled=False).astype(np.float64)
    r, mask = relerr(fb, fo, 1e-30)
    print(f"[{name}] flux rel err: max {r.max():.3e} p99 {np.quantile(r[mask],0.99):.3e} median {np.median(r[mask]):.3e} nan {np.isnan(fb).sum()}", flush=True)
    j = np.unravel_index(np.argmax(r), r.shape); print("   worst at", j, fb[j], fo[j], 'z=', w.params.redshift[j[0]])
    fs = eng.photometry(w.params, scaled=True)
    fso = O.scale_to_mass(fo, w.params.log_mass)
    r, mask = relerr(fs, fso, 1e-30)
    print(f"[{name}] scaled flux rel err max {r.max():.3e}", flush=True)
    eng.close()
