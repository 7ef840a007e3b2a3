"""Shared assertions for the GPU parity tests."""
import numpy as np

FLUX_RTOL = 1e-5


def assert_flux_close(got, want, rtol=FLUX_RTOL):
    """Relative parity on every band within 30 decades of the galaxy's brightest band; bands below that
    (Lyman-continuum dropouts whose exp(-tau) underflows float32) only have to be equally negligible."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    ref = np.abs(want).max(axis=-1, keepdims=True)
    big = np.abs(want) > 1e-30 * ref
    assert np.isfinite(got).all()
    err = np.abs(got[big] - want[big]) / np.abs(want[big])
    assert err.max() < rtol, f"max rel err {err.max():.3e}"
    assert np.all(np.abs(got) <= np.where(big, np.inf, 1e-25 * ref))
    return err.max()
