"""TEST INFRASTRUCTURE: ctypes wrapper of the C/OpenMP oracle (``oracle_c.c``).

``build()`` compiles it with gcc into ``oracle/_build/liboracle.so`` (git-ignored, travels to the
GPU box).  Imports nothing from the product package; inputs are plain arrays.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "oracle_c.c")
LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    if os.path.isfile(LIB) and not force and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"],
                   check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.oracle_luminosity_distance_cm.restype = C.c_double
        _lib.oracle_luminosity_distance_cm.argtypes = [C.c_double]
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def synthesize(p, log10ages, metallicities, lam, g_att, g_un, filters, *, kappa=None, igm=None,
               variant="nu", base_mass=1e9, nthreads=0, return_spectra=False, two_screens=None, dust_shape=None):
    """Same contract as ``oracle.synthesize`` but threaded C.  ``p`` is GalaxyParams-like;
    g_att / g_un are (n_age, n_z, n_lam) float64 or None; filters = [(lam_table, t_table), ...].
    two_screens: dict(age_pivot=log10 yr, kappa_birth=(n_lam,), tau_v_birth=(n,)) -- SURVEY A5 birth cloud + ISM;
    dust_shape: (n_lam,) unit-integral emission spectrum for the energy balance ('total')."""
    lib = load()
    f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    z, tv = f64(p.redshift), f64(p.tau_v)
    rows = f64(p.sfh_rows)
    zv, zs = f64(p.zd_value), f64(p.zd_sigma)
    n = z.shape[0]
    la, zm, lm = f64(log10ages), f64(metallicities), f64(lam)
    ga = None if g_att is None or not np.any(g_att) else f64(g_att)
    gu = None if g_un is None or not np.any(g_un) else f64(g_un)
    kap = f64(kappa)
    off = np.zeros(len(filters) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(f[0]) for f in filters])
    fl = f64(np.concatenate([f[0] for f in filters]))
    ft = f64(np.concatenate([f[1] for f in filters]))
    laf = f64(igm[0]) if igm is not None else None
    dla = f64(igm[1]) if igm is not None else None
    out = np.empty((n, len(filters)))
    spec = np.empty((n, lm.size)) if return_spectra else None
    kb = tvb = young = None
    if two_screens is not None:
        kb, tvb = f64(two_screens["kappa_birth"]), f64(two_screens["tau_v_birth"])
        young = np.ascontiguousarray(np.asarray(log10ages) < two_screens["age_pivot"], dtype=np.int32)
    ds = f64(dust_shape)
    rc = lib.oracle_synthesize_ex(
        C.c_int64(n), _p(z), _p(tv), C.c_int(int(p.sfh_type)), C.c_int(rows.shape[1]), _p(rows),
        C.c_int(int(p.zd_type)), _p(zv), _p(zs), C.c_int(la.size), C.c_int(zm.size), C.c_int(lm.size),
        _p(la), _p(zm), _p(lm), _p(ga), _p(gu), _p(kap), C.c_int(1 if igm is not None else 0), _p(laf), _p(dla),
        C.c_int(0 if laf is None else laf.shape[0]), C.c_int(len(filters)), _p(off), _p(fl), _p(ft),
        C.c_int(0 if variant == "nu" else 1), C.c_double(base_mass), C.c_int(int(nthreads)), _p(out), _p(spec),
        _p(kb), _p(tvb), _p(young), _p(ds))
    if rc != 0:
        raise ValueError(f"filter lies entirely outside the spectrum for galaxy {rc - 1}")
    return (out, spec) if return_spectra else out


def num_threads():
    return int(load().oracle_num_threads())
