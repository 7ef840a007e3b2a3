import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synference_b200.configs import make_workload
from synference_b200.engine import SynthEngine
from synference_b200 import igm as I
from oracle import c_oracle as CO, oracle as O
n = 1_000_000
w = make_workload("cfg2", n)
eng = SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters, max_batch=n)
sub = slice(0, 20000)
p = w.params.slice(sub)
a = eng.photometry(w.params, scaled=False)[sub].astype(np.float64)
lam = np.asarray(w.grid.lam)
ga, gu = O.emission_parts(w.grid.spectra, lam, w.emission_key)
want, spec = CO.synthesize(p, w.grid.log10ages, w.grid.metallicity, lam, ga, gu, [(f.lam, f.t) for f in w.filters], kappa=O.dust_kappa(lam), igm=(I.INOUE14_LAF, I.INOUE14_DLA), return_spectra=True)
ref = np.abs(want).max(1, keepdims=True)
big = np.abs(want) > 1e-30 * ref
err = np.where(big, np.abs(a - want) / np.abs(want), 0)
order = np.argsort(err.ravel())[::-1][:15]
for o in order:
    g, f = divmod(o, want.shape[1])
    print(f"g={g} f={f} err={err[g,f]:.3e} got={a[g,f]:.6e} want={want[g,f]:.6e} ratio_to_max={want[g,f]/ref[g,0]:.2e} z={p.redshift[g]:.4f} tau_v={p.tau_v[g]:.3f} tau={p.sfh_rows[g,2]:.3f} pk/mx={p.sfh_rows[g,3]/p.sfh_rows[g,1]:.3f} logZ={p.zd_value[g]:.3f}")
print("signed mean rel err", np.mean(((a - want) / want)[big]))
b = eng.photometry(p, scaled=False).astype(np.float64)
errb = np.where(big, np.abs(b - want) / np.abs(want), 0)
print("same galaxies as their own batch: max", errb.max(), " in 1M batch: max", err.max())
W = eng.weights(p.slice(slice(0, 2000)))
from oracle import adapter as A
Wo = A.weights_matrix(p.slice(slice(0,2000)), w.grid.log10ages, w.grid.metallicity)
print("weights max abs", np.abs(W - Wo).max())
