import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from synference_b200.configs import make_workload
from synference_b200.engine import SynthEngine
os.environ["SB2_CTA_PAIR"] = "1"
w = make_workload(os.environ.get("DIAG_CFG", "cfg2"), int(os.environ.get("DIAG_N", "300")))
eng = SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters, max_batch=1 << 15)
try:
    s = eng.spectra(w.params)
    print("ok", float(np.nansum(s)))
except Exception as e:
    print("FAILED:", str(e)[:300])
    print("watchdog:", eng.lib.sb2_wait_debug(eng._h).decode())
