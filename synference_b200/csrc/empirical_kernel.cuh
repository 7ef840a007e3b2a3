// Empirical uncertainty models on the device.
//
// Restates GeneralEmpiricalUncertaintyModel.apply_noise (noise_models.py:818-880) for every (filter, row) of a library at
// once -- the reference walks the filters in Python and calls scipy per filter (sbi_runner.py:813-903):
//   1. flux -> the model's interpolation unit (linear unit change, or AB <-> Jy: noise_models.py:55-73);
//   2. sigma ~ TruncNorm(mu_sigma(f), sigma_sigma(f); >= 0), both linear interpolants of the binned catalogue statistics
//      with end-value fill or linear extrapolation (noise_models.py:347-390);
//   3. sources already below the SNR threshold are not scattered; the others get N(0, sigma) (or a sigma-clipped normal);
//   4. error_type "observed": sigma is drawn again at the noisy flux;
//   5. sources below the SNR threshold before or after the scatter are replaced by the upper-limit rule (a constant, or the
//      limit plus a +-3 sigma truncated normal) and their error by the error rule (a constant resolved on the host);
//   6. back to the output unit, error clipped to [min, max].
// Truncated normals are drawn by inversion, x = Phi^-1(Phi(a) + u (Phi(b) - Phi(a))) evaluated on the tail that keeps
// precision, so that a test can inject the uniforms and compare with scipy's ppf.  Draws: injected ([4][n_filt][n]: sigma
// uniform, scatter normal (or uniform when sigma-clipped), re-draw uniform, limit-scatter uniform) or Philox4x32-10 keyed by
// (seed, epoch) with counter (row, filter).  The injected (parity) mode is float64 throughout; the Philox (production) mode
// evaluates the normal CDF / inverse CDF and Box-Muller in float32.  One thread per (filter, row); 8 B read + 16 B written per
// element (+ 32 B of injected draws in the parity mode).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "noise_kernel.cuh"

namespace sb2 {

constexpr int kEmpMaxBins = 64;

struct EmpiricalModelDev {
  int n_bins, extrapolate;
  int internal_is_ab, in_is_ab, out_is_ab, observed_error, upper_limits, ul_active;
  double internal_to_jy, in_to_jy, out_to_jy;     // size of the linear unit in Jy (unused for AB)
  double sigma_clip;                               // < 0: plain normal
  double snr_threshold, ul_flux, ul_scatter_std, ul_err;   // ul_scatter_std < 0: ul_flux is the replacement value itself
  double min_err, max_err;
  int asinh_mode, pad_;                            // 0: general model; 1: asinh magnitudes, tables in asinh mags;
  double asinh_b;                                  // 2: asinh magnitudes out, tables in a linear flux unit.  b in Jy
  double centers[kEmpMaxBins], median[kEmpMaxBins], stdev[kEmpMaxBins];
};

struct EmpiricalArgs {
  const double* flux;      // [n_filt][n]
  long long n;
  int n_filt;
  const EmpiricalModelDev* models;   // [n_filt] device
  const double* draws;     // [4][n_filt][n] or nullptr
  unsigned long long seed, epoch;
  double* out_flux;        // [n_filt][n]
  double* out_sigma;       // [n_filt][n]
};

__device__ __forceinline__ double ab_to_jy(double m) { return pow(10.0, -0.4 * (m - 8.90)); }
__device__ __forceinline__ double jy_to_ab(double f) { return -2.5 * log10(f) + 8.90; }
__device__ __forceinline__ double ab_err_to_jy(double e, double fjy) { return (fjy * e * 2.302585092994046) / 2.5; }
__device__ __forceinline__ double jy_err_to_ab(double e, double fjy) { return fabs((2.5 / 2.302585092994046) * (e / fjy)); }

// inverse CDF of the standard normal truncated to [a, b] at probability u
__device__ __forceinline__ double truncnorm_ppf(double u, double a, double b) {
  if (a > 0.0) {   // work on the mirrored (left) tail, where the CDF values are small and precise
    const double pa = normcdf(-a), pb = isinf(b) ? 0.0 : normcdf(-b);
    return -normcdfinv(pb + (1.0 - u) * (pa - pb));
  }
  const double pa = normcdf(a), pb = isinf(b) ? 1.0 : normcdf(b);
  return normcdfinv(pa + u * (pb - pa));
}

// utils.py:647-704: asinh magnitudes with softening b [Jy]
__device__ __forceinline__ double jy_to_asinh(double f, double b) {
  return -1.0857362047581296 * (asinh(f / (2.0 * b)) + log(b / 3631.0));
}
__device__ __forceinline__ double jy_err_to_asinh(double f, double e, double b) {
  return 1.0857362047581296 * e / sqrt(f * f + (2.0 * b) * (2.0 * b));
}

__device__ __forceinline__ void to_jy(const EmpiricalModelDev& M, double f, double e, double& fj, double& ej) {
  if (M.internal_is_ab) { fj = ab_to_jy(f); ej = ab_err_to_jy(e, fj); }
  else { fj = f * M.internal_to_jy; ej = e * M.internal_to_jy; }
}
__device__ __forceinline__ bool below_snr(const EmpiricalModelDev& M, double f, double e) {
  double fj, ej;
  to_jy(M, f, e, fj, ej);
  const double snr = fj / ej;
  return !isfinite(snr) || snr < M.snr_threshold;
}
// mu_sigma(v) and sigma_sigma(v): scipy interp1d(kind="linear", bounds_error=False, fill_value=(y0, y_last) | "extrapolate");
// the two tables share the abscissa, hence the bin search
__device__ __forceinline__ void interp_both(const EmpiricalModelDev& M, double v, double& mu, double& ss) {
  const double* x = M.centers;
  const int n = M.n_bins;
  if (!(v == v)) { mu = ss = v; return; }
  int lo, hi;
  if (v < x[0]) {
    if (!M.extrapolate) { mu = M.median[0]; ss = M.stdev[0]; return; }
    const double d = v - x[0], w = x[1] - x[0];
    mu = M.median[0] + d * ((M.median[1] - M.median[0]) / w);
    ss = M.stdev[0] + d * ((M.stdev[1] - M.stdev[0]) / w);
    return;
  }
  if (v > x[n - 1]) {
    if (!M.extrapolate) { mu = M.median[n - 1]; ss = M.stdev[n - 1]; return; }
    const double d = v - x[n - 2], w = x[n - 1] - x[n - 2];
    mu = M.median[n - 2] + d * ((M.median[n - 1] - M.median[n - 2]) / w);
    ss = M.stdev[n - 2] + d * ((M.stdev[n - 1] - M.stdev[n - 2]) / w);
    return;
  }
  lo = 0; hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (x[mid] <= v) lo = mid; else hi = mid;
  }
  const double d = v - x[lo], w = x[hi] - x[lo];
  mu = ((M.median[hi] - M.median[lo]) / w) * d + M.median[lo];
  ss = ((M.stdev[hi] - M.stdev[lo]) / w) * d + M.stdev[lo];
}

// float32 special functions for the production (Philox) mode: the draw's relative accuracy (1e-6) is far below its spread
__device__ __forceinline__ double truncnorm_ppf_fast(double u, double a, double b) {
  const float af = (float)a, bf = (float)b, uf = (float)u;
  if (af > 0.f) {
    const float pa = normcdff(-af), pb = isinf(bf) ? 0.f : normcdff(-bf);
    return (double)(-normcdfinvf(fmaf(1.f - uf, pa - pb, pb)));
  }
  const float pa = normcdff(af), pb = isinf(bf) ? 1.f : normcdff(bf);
  return (double)normcdfinvf(fminf(fmaf(uf, pb - pa, pa), 0.99999994f));
}

template <bool kFast>
__device__ __forceinline__ double sample_sigma(const EmpiricalModelDev& M, double f, double u) {
  double mu, ss;
  interp_both(M, f, mu, ss);
  ss = fmax(0.0, ss);
  const double a = (0.0 - mu) / (ss > 1e-9 ? ss : 1.0);
  return mu + ss * (kFast ? truncnorm_ppf_fast(u, a, INFINITY) : truncnorm_ppf(u, a, INFINITY));
}

// kFast: Philox draws + float32 special functions (production); otherwise injected draws, float64 throughout (parity)
template <bool kFast>
__global__ void __launch_bounds__(256) empirical_noise_kernel(EmpiricalArgs A) {
  __shared__ EmpiricalModelDev M;
  const int f = blockIdx.y;
  {   // this block's filter model -> shared memory
    const int* src = reinterpret_cast<const int*>(A.models + f);
    int* dst = reinterpret_cast<int*>(&M);
    for (int i = threadIdx.x; i < (int)(sizeof(EmpiricalModelDev) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const long long stride = (long long)A.n_filt * A.n;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < A.n; r += (long long)gridDim.x * blockDim.x) {
    const long long idx = (long long)f * A.n + r;
    double u_sig, z_noise, u_obs, u_lim;
    if (!kFast) {
      u_sig = A.draws[idx]; z_noise = A.draws[stride + idx]; u_obs = A.draws[2 * stride + idx]; u_lim = A.draws[3 * stride + idx];
    } else {
      uint32_t a4[4], b4[4] = {0u, 0u, 0u, 0u};
      philox4x32_10((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)f, (uint32_t)A.epoch, (uint32_t)A.seed,
                    (uint32_t)(A.seed >> 32) ^ (uint32_t)(A.epoch >> 32), a4);
      if (M.sigma_clip < 0.0)
        philox4x32_10((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)f | 0x80000000u, (uint32_t)A.epoch, (uint32_t)A.seed,
                      (uint32_t)(A.seed >> 32) ^ (uint32_t)(A.epoch >> 32), b4);
      // 24-bit uniforms in (0, 1); the scatter normal by Box-Muller in float32 (as depth_noise_kernel)
      auto uni = [](uint32_t w) { return ((float)(w >> 8) + 0.5f) * (1.0f / 16777216.0f); };
      u_sig = uni(a4[0]); u_obs = uni(a4[1]); u_lim = uni(a4[2]);
      if (M.sigma_clip >= 0.0) {
        z_noise = uni(a4[3]);
      } else {
        float sn, cs;
        sincospif(2.0f * uni(b4[1]), &sn, &cs);
        z_noise = (double)(sqrtf(-2.0f * logf(uni(b4[0]))) * cs);
      }
    }
    const double fin = A.flux[idx];
    if (M.asinh_mode) {   // AsinhEmpiricalUncertaintyModel.apply_noise (noise_models.py:507-557); the scatter is always normal
      const double fjy = M.in_is_ab ? ab_to_jy(fin) : fin * M.in_to_jy;
      double m_noisy, err;
      if (M.asinh_mode == 1) {
        const double m_true = jy_to_asinh(fjy, M.asinh_b);
        const double e0 = sample_sigma<kFast>(M, m_true, u_sig);
        m_noisy = m_true + (0.0 + e0 * z_noise);
        err = M.observed_error ? sample_sigma<kFast>(M, m_noisy, u_obs) : e0;
      } else {
        const double e0 = sample_sigma<kFast>(M, fjy / M.internal_to_jy, u_sig) * M.internal_to_jy;
        const double njy = fjy + (0.0 + e0 * z_noise);
        m_noisy = jy_to_asinh(njy, M.asinh_b);
        // (in this branch the reference re-draws for error_type "empirical" and keeps the first draw otherwise)
        const double e1 = M.observed_error ? e0 : sample_sigma<kFast>(M, njy / M.internal_to_jy, u_obs) * M.internal_to_jy;
        err = jy_err_to_asinh(njy, e1, M.asinh_b);
      }
      A.out_flux[idx] = m_noisy;
      if (A.out_sigma) A.out_sigma[idx] = fmin(fmax(err, M.min_err), M.max_err);
      continue;
    }
    // 1. to the interpolation unit
    double fi;
    if (M.in_is_ab == M.internal_is_ab) fi = M.in_is_ab ? fin : fin * (M.in_to_jy / M.internal_to_jy);
    else if (M.in_is_ab) fi = ab_to_jy(fin) / M.internal_to_jy;
    else fi = jy_to_ab(fin * M.in_to_jy);
    // 2. sampled uncertainty at the true flux
    const double sig0 = sample_sigma<kFast>(M, fi, u_sig);
    // 3. scatter, unless the source is already below the SNR threshold
    const bool init_lim = M.upper_limits && below_snr(M, fi, sig0);
    double noisy = fi;
    if (!init_lim) {
      const double zz = M.sigma_clip >= 0.0 ? (kFast ? truncnorm_ppf_fast(z_noise, -M.sigma_clip, M.sigma_clip) : truncnorm_ppf(z_noise, -M.sigma_clip, M.sigma_clip)) : z_noise;
      noisy = fi + (0.0 + sig0 * zz);
    }
    // 4. "observed" errors: drawn again at the noisy flux
    double sig = M.observed_error ? sample_sigma<kFast>(M, noisy, u_obs) : sig0;
    // 5. upper limits
    if (M.upper_limits && M.ul_active && (init_lim || below_snr(M, noisy, sig))) {
      noisy = M.ul_scatter_std >= 0.0 ? M.ul_flux + (0.0 + M.ul_scatter_std * (kFast ? truncnorm_ppf_fast(u_lim, -3.0, 3.0) : truncnorm_ppf(u_lim, -3.0, 3.0))) : M.ul_flux;
      sig = M.ul_err;
    }
    // 6. to the output unit, error clip
    double of, os;
    if (M.out_is_ab == M.internal_is_ab) {
      const double c = M.out_is_ab ? 1.0 : M.internal_to_jy / M.out_to_jy;
      of = M.out_is_ab ? noisy : noisy * c;
      os = M.out_is_ab ? sig : sig * c;
    } else if (M.internal_is_ab) {
      const double fj = ab_to_jy(noisy);
      of = fj / M.out_to_jy;
      os = ab_err_to_jy(sig, fj) / M.out_to_jy;
    } else {
      const double fj = noisy * M.internal_to_jy, ej = sig * M.internal_to_jy;
      of = jy_to_ab(fj);
      os = jy_err_to_ab(ej, fj);
    }
    os = fmin(fmax(os, M.min_err), M.max_err);
    A.out_flux[idx] = of;
    if (A.out_sigma) A.out_sigma[idx] = os;
  }
}

}  // namespace sb2
