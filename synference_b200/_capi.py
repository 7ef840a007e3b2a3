"""ctypes binding of ``include/synference_b200.h`` (the C ABI of the CUDA hot path).

The shared library is built in-tree by ``__graft_entry__.build()`` (plain ``nvcc``) as
``synference_b200/csrc/libsynference_b200.so``.  There is no CPU fallback: if the library
is missing, or no B200 is present, the compute entry points raise.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SB2_LIB") or os.path.join(_HERE, "csrc", "libsynference_b200.so")  # SB2_LIB: A/B builds

SFH_ROW = 24

EXPORTED_SYMBOLS = (
    "sb2_last_error", "sb2_device_count", "sb2_model_create", "sb2_model_destroy", "sb2_build_weights",
    "sb2_synth_photometry", "sb2_synth_photometry_host", "sb2_depth_noise_features", "sb2_last_stage_ms", "sb2_wait_debug",
    "sb2_synth_photometry_host_submit", "sb2_synth_photometry_host_wait",
    "sb2_resampler_create", "sb2_resampler_destroy", "sb2_resample_spectra", "sb2_resample_spectra_host", "sb2_resample_last_ms",
    "sb2_empirical_noise", "sb2_depth_noise_features_sets", "sb2_depth_noise_features_f32", "sb2_kernel_launches",
    "sb2_filterset_create", "sb2_filterset_destroy", "sb2_filter_integrate",
)

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)


class ModelDesc(C.Structure):
    _fields_ = [
        ("n_age", C.c_int32), ("n_z", C.c_int32), ("n_lam", C.c_int32), ("n_comp", C.c_int32),
        ("n_filt", C.c_int32), ("n_age_pad", C.c_int32), ("k_pad", C.c_int32), ("n_chunk", C.c_int32),
        ("log10ages", _dp), ("metallicities", _dp),
        ("gt_hi", _fp), ("gt_lo", _fp), ("grid_scale", C.c_double),
        ("kappa", _fp), ("lam0", C.c_double), ("q", C.c_double), ("interp_variant", C.c_int32),
        ("filt_lo", _ip), ("filt_hi", _ip), ("filt_off", _ip), ("filt_uv", _fp), ("filt_uv_len", C.c_int32),
        ("filt_su", _dp), ("filt_sdv", _dp),
        ("n_blue", C.c_int32), ("n_lines", C.c_int32),
        ("igm_bin_pow", _dp), ("igm_nline", _ip), ("igm_lc_on", _ip), ("igm_thr", _dp), ("igm_pre", _dp),
        ("cosmo_n", C.c_int32), ("cosmo_smax", C.c_double),
        ("cosmo_dc", _dp), ("cosmo_ddc", _dp), ("cosmo_age", _dp), ("cosmo_dage", _dp),
        ("base_mass", C.c_double), ("max_batch", C.c_int64),
        ("dust_d0", _fp), ("dust_l2", _fp),
        ("lya_line", _dp), ("lya_bin", C.c_int32), ("kappa_birth", _fp),
        ("dust_wnu", _fp), ("dust_g", _fp), ("dust_duv", _fp), ("dust_m_len", C.c_int32),
        ("fm_log_tab", _dp), ("fm_exp_tab", _dp), ("fm_tail_tab", _dp), ("fm_tail_n", C.c_int32), ("fm_tail_w", C.c_double),
        ("rest_frame", C.c_int32), ("x_bin0", C.c_int32), ("x_bins", C.c_int32),
    ]


class Params(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("redshift", C.c_void_p), ("log_mass", C.c_void_p), ("tau_v", C.c_void_p),
        ("sfh_type", C.c_int32), ("sfh_stride", C.c_int32), ("sfh_rows", C.c_void_p),
        ("max_age_from_z", C.c_int32), ("norm_mask", C.c_uint32), ("age_zmax_gyr", C.c_double),
        ("zd_type", C.c_int32), ("zd_value", C.c_void_p), ("zd_sigma", C.c_void_p),
        ("coef_att", C.c_void_p), ("coef_unatt", C.c_void_p),
        ("dust_slope", C.c_void_p), ("dust_ampl", C.c_void_p), ("fesc_lya", C.c_void_p),
        ("tau_v_birth", C.c_void_p),
        ("host_f32", C.c_int32),
        ("scaled_ld", C.c_int64),
        ("energy_full_axis", C.c_int32),
    ]


class ResampleDesc(C.Structure):
    _fields_ = [
        ("n_lam", C.c_int32), ("n_px", C.c_int32), ("n_res", C.c_int32),
        ("theory_wave", _dp), ("observed_wave", _dp), ("res_wave", _dp), ("res_r", _dp), ("theory_r", _dp),
        ("theory_r_scalar", C.c_double), ("trunc", C.c_double), ("fill", C.c_double),
    ]


EMP_MAX_BINS = 64


class EmpiricalModel(C.Structure):
    _fields_ = [
        ("n_bins", C.c_int32), ("extrapolate", C.c_int32),
        ("internal_is_ab", C.c_int32), ("in_is_ab", C.c_int32), ("out_is_ab", C.c_int32),
        ("observed_error", C.c_int32), ("upper_limits", C.c_int32), ("ul_active", C.c_int32),
        ("internal_to_jy", C.c_double), ("in_to_jy", C.c_double), ("out_to_jy", C.c_double),
        ("sigma_clip", C.c_double), ("snr_threshold", C.c_double), ("ul_flux", C.c_double),
        ("ul_scatter_std", C.c_double), ("ul_err", C.c_double), ("min_err", C.c_double), ("max_err", C.c_double),
        ("asinh_mode", C.c_int32), ("reserved_", C.c_int32), ("asinh_b", C.c_double),
        ("centers", C.c_double * EMP_MAX_BINS), ("median", C.c_double * EMP_MAX_BINS), ("stdev", C.c_double * EMP_MAX_BINS),
    ]


_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load():
    """Load the CUDA library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  synference_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.sb2_last_error.restype = C.c_char_p
    lib.sb2_device_count.restype = C.c_int
    lib.sb2_kernel_launches.restype = C.c_longlong
    lib.sb2_filterset_create.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int,
                                         C.POINTER(C.c_void_p)]
    lib.sb2_filterset_destroy.argtypes = [C.c_void_p]
    lib.sb2_filter_integrate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int64, C.c_void_p, C.c_void_p,
                                         C.c_void_p]
    lib.sb2_model_create.argtypes = [C.POINTER(ModelDesc), C.c_int, C.POINTER(C.c_void_p)]
    lib.sb2_model_destroy.argtypes = [C.c_void_p]
    lib.sb2_wait_debug.argtypes = [C.c_void_p]
    lib.sb2_wait_debug.restype = C.c_char_p
    lib.sb2_build_weights.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p]
    lib.sb2_synth_photometry.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]
    lib.sb2_synth_photometry_host.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sb2_synth_photometry_host_submit.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_int]
    lib.sb2_synth_photometry_host_wait.argtypes = [C.c_void_p, C.c_int]
    lib.sb2_last_stage_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.sb2_depth_noise_features.argtypes = [
        C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_double, C.c_void_p, C.c_uint64,
        C.c_uint64, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sb2_depth_noise_features_f32.argtypes = [
        C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_double, C.c_uint64, C.c_uint64, C.c_double, C.c_void_p,
        C.c_void_p]
    lib.sb2_depth_noise_features_sets.argtypes = [
        C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_double, C.c_void_p, C.c_uint64,
        C.c_uint64, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sb2_resampler_create.argtypes = [C.POINTER(ResampleDesc), C.c_int, C.POINTER(C.c_void_p)]
    lib.sb2_resampler_destroy.argtypes = [C.c_void_p]
    lib.sb2_resample_spectra.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.sb2_resample_spectra_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.sb2_resample_last_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.sb2_empirical_noise.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(EmpiricalModel), C.c_void_p, C.c_uint64,
                                        C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().sb2_last_error().decode(errors="replace")
        exc = ValueError if rc == -1 else RuntimeError
        raise exc(f"{what} failed ({rc}): {msg}")
