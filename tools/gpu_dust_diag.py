"""Diagnostic: where does the 'total' (dust emission) photometry differ from the oracle?"""
import numpy as np
from oracle import adapter as A, oracle as O
from synference_b200 import igm as I
from synference_b200.configs import make_workload
from synference_b200.engine import SynthEngine
from synference_b200.parametric import Calzetti2000, Greybody, PacmanEmission

n = 200
w = make_workload("cfg2", n)
em = PacmanEmission(grid=w.grid, fesc=0.0, fesc_ly_alpha=0.5, dust_curve=Calzetti2000(), dust_emission=Greybody(40.0, 1.5))
eng = SynthEngine(w.grid, em, "total", w.filters, max_batch=4096)
p = w.params.slice(slice(0, n))
p.redshift = p.redshift.copy()
p.redshift[:40] = np.linspace(0.02, 1.5, 40)
got = eng.photometry(p, scaled=False)
gals = A.galaxies_from_params(p)
lam = np.asarray(w.grid.lam)
filt = [(f.lam, f.t) for f in w.filters]
kw = dict(key="emergent", dust=dict(curve="Calzetti2000"), igm=(I.INOUE14_LAF, I.INOUE14_DLA), fesc_ly_alpha=0.5)
want = O.synthesize(gals, w.grid.log10ages, w.grid.metallicity, lam, w.grid.spectra, filt,
                    dust_emission=dict(kind="Greybody", temperature=40.0, emissivity=1.5), **kw)
bare = O.synthesize(gals, w.grid.log10ages, w.grid.metallicity, lam, w.grid.spectra, filt, **kw)
err = np.abs(got - want) / np.abs(want)
bad = np.argwhere(err > 1e-5)
print("n bad", len(bad), "max", err.max())
for g, f in bad[:30]:
    print(g, f, "z", p.redshift[g], "tau", p.tau_v[g], "got", got[g, f], "want", want[g, f], "bare", bare[g, f],
          "dust share", (want[g, f] - bare[g, f]) / want[g, f], "ratio of dust parts", (got[g, f] - bare[g, f]) / (want[g, f] - bare[g, f]))
