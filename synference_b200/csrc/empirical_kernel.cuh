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
// (seed, epoch) with counter (row pair, filter).  Two kernels: `empirical_noise_kernel` (parity mode: injected draws, float64
// throughout, one thread per (filter, row), + 32 B of draws per element) and `empirical_noise_fast_kernel` (production mode:
// Philox draws, float32 arithmetic on float64 inputs / outputs, tables with precomputed slopes in shared memory, two rows per
// thread so that one Philox block and one Box-Muller pair serve both and the accesses are 16-byte).  8 B read + 16 B written
// per element: the production kernel is meant to run at the HBM roofline, not at the special-function rate.
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
  // Philox round keys of (seed, epoch), filled by the host: as kernel parameters they are constant-bank operands of the
  // rounds' XORs (bumping the key in registers cost 20 integer adds per Philox block)
  uint32_t rk0[10], rk1[10];
};

// Philox4x32-10 with the round keys given (same function of (counter, key) as philox4x32_10)
__device__ __forceinline__ void philox4x32_10_keyed(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&rk0)[10],
                                                    const uint32_t (&rk1)[10], uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ rk0[r]; c1 = lo1; c2 = hi0 ^ c3 ^ rk1[r]; c3 = lo0;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

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

__device__ __forceinline__ double sample_sigma(const EmpiricalModelDev& M, double f, double u) {
  double mu, ss;
  interp_both(M, f, mu, ss);
  ss = fmax(0.0, ss);
  const double a = (0.0 - mu) / (ss > 1e-9 ? ss : 1.0);
  return mu + ss * truncnorm_ppf(u, a, INFINITY);
}

// parity mode: injected draws, float64 throughout
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
    const double u_sig = A.draws[idx], z_noise = A.draws[stride + idx], u_obs = A.draws[2 * stride + idx], u_lim = A.draws[3 * stride + idx];
    const double fin = A.flux[idx];
    if (M.asinh_mode) {   // AsinhEmpiricalUncertaintyModel.apply_noise (noise_models.py:507-557); the scatter is always normal
      const double fjy = M.in_is_ab ? ab_to_jy(fin) : fin * M.in_to_jy;
      double m_noisy, err;
      if (M.asinh_mode == 1) {
        const double m_true = jy_to_asinh(fjy, M.asinh_b);
        const double e0 = sample_sigma(M, m_true, u_sig);
        m_noisy = m_true + (0.0 + e0 * z_noise);
        err = M.observed_error ? sample_sigma(M, m_noisy, u_obs) : e0;
      } else {
        const double e0 = sample_sigma(M, fjy / M.internal_to_jy, u_sig) * M.internal_to_jy;
        const double njy = fjy + (0.0 + e0 * z_noise);
        m_noisy = jy_to_asinh(njy, M.asinh_b);
        // (in this branch the reference re-draws for error_type "empirical" and keeps the first draw otherwise)
        const double e1 = M.observed_error ? e0 : sample_sigma(M, njy / M.internal_to_jy, u_obs) * M.internal_to_jy;
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
    const double sig0 = sample_sigma(M, fi, u_sig);
    // 3. scatter, unless the source is already below the SNR threshold
    const bool init_lim = M.upper_limits && below_snr(M, fi, sig0);
    double noisy = fi;
    if (!init_lim) {
      const double zz = M.sigma_clip >= 0.0 ? truncnorm_ppf(z_noise, -M.sigma_clip, M.sigma_clip) : z_noise;
      noisy = fi + (0.0 + sig0 * zz);
    }
    // 4. "observed" errors: drawn again at the noisy flux
    double sig = M.observed_error ? sample_sigma(M, noisy, u_obs) : sig0;
    // 5. upper limits
    if (M.upper_limits && M.ul_active && (init_lim || below_snr(M, noisy, sig))) {
      noisy = M.ul_scatter_std >= 0.0 ? M.ul_flux + (0.0 + M.ul_scatter_std * truncnorm_ppf(u_lim, -3.0, 3.0)) : M.ul_flux;
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

// ---- production mode -------------------------------------------------------------------------------------------------------
// float32 helpers.  Accuracy is set against the quantities' own spread (errors are a few per cent of a flux, drawn at random):
// the normal quantile is good to 1e-6 absolute, the CDF to 3e-7 absolute, magnitudes to a float32 ulp (2e-6 mag).
struct EmpFastTab {
  float x[kEmpMaxBins];        // centres, padded with +inf (the search reads x alone)
  float4 seg[kEmpMaxBins];     // per segment: x, median, its slope, stdev -- one 16-byte read
  float dsd[kEmpMaxBins];      // slope of the stdev
  float x0, xlast;
  float lo_clamp, hi_clamp;    // x0, xlast when values beyond the table take the end values; -inf, +inf when it extrapolates
  float clip_p0, clip_dp;      // Phi(-clip), Phi(clip) - Phi(-clip): the sigma-clipped scatter by inversion
  float lim_p0, lim_dp;        // the same for the +-3 sigma scatter about an upper limit
  float log_b;                 // asinh modes: log(b / 3631 Jy)
  float inv_dx;                // > 0: the centres are equally spaced, 1 / spacing
  float c_out_f;
  float min_err, max_err;
  double c_in, c_out;          // linear unit changes: input -> interpolation unit, interpolation unit -> output (1 for AB)
  int n_bins, extrapolate, simple;
};

// standard normal quantile, p in [3e-8, 1 - 6e-8]: sqrt(2) erfinv(2p - 1) with Giles' single-precision polynomials in
// w = -log(1 - (2p - 1)^2) = -log(4 p (1 - p)), which keeps its precision in both tails
__device__ __forceinline__ float ndtri_fast(float p) {
  const float x = fmaf(2.f, p, -1.f);
  float w = -__logf(4.f * p * (1.f - p));
  float q;
  if (w < 5.f) {
    w -= 2.5f;
    q = 2.81022636e-08f;
    q = fmaf(q, w, 3.43273939e-07f); q = fmaf(q, w, -3.5233877e-06f); q = fmaf(q, w, -4.39150654e-06f);
    q = fmaf(q, w, 0.00021858087f); q = fmaf(q, w, -0.00125372503f); q = fmaf(q, w, -0.00417768164f);
    q = fmaf(q, w, 0.246640727f); q = fmaf(q, w, 1.50140941f);
  } else {
    w = sqrtf(w) - 3.f;
    q = -0.000200214257f;
    q = fmaf(q, w, 0.000100950558f); q = fmaf(q, w, 0.00134934322f); q = fmaf(q, w, -0.00367342844f);
    q = fmaf(q, w, 0.00573950773f); q = fmaf(q, w, -0.0076224613f); q = fmaf(q, w, 0.00943887047f);
    q = fmaf(q, w, 1.00167406f); q = fmaf(q, w, 2.83297682f);
  }
  return q * x * 1.41421356f;
}
// Phi(t) for t <= 0: erfc by Abramowitz & Stegun 7.1.26 (|error| < 1.5e-7 on erfc)
__device__ __forceinline__ float ncdf_neg_fast(float t) {
  const float x = -0.70710678f * t;
  const float k = __fdividef(1.f, fmaf(0.3275911f, x, 1.f));
  float s = 1.061405429f;
  s = fmaf(s, k, -1.453152027f); s = fmaf(s, k, 1.421413741f); s = fmaf(s, k, -0.284496736f); s = fmaf(s, k, 0.254829592f);
  return 0.5f * s * k * __expf(-x * x);
}
__device__ __forceinline__ float ab_to_jy_f(float m) { return exp2f(-1.3287712379549449f * (m - 8.90f)); }
__device__ __forceinline__ float jy_to_ab_f(float f) { return fmaf(-0.7525749891599529f, log2f(f), 8.90f); }

// mu_sigma(v), sigma_sigma(v) from the slope tables, then sigma ~ TruncNorm(mu, ss; >= 0) at probability u
__device__ __forceinline__ float emp_sigma_fast(const EmpFastTab& T, int n_bins, int extrapolate, float v, float u) {
  const float vc = fminf(fmaxf(v, T.lo_clamp), T.hi_clamp);
  int lo = 0;
  if (T.inv_dx > 0.f) {   // equally spaced centres (linear bins, none dropped): the segment by arithmetic.  A value within
                          // rounding of a centre may land in the neighbouring segment, where the interpolant is the same.
    lo = (int)((vc - T.x0) * T.inv_dx);       // (NaN and out-of-range values convert to a clamped integer)
  } else {
#pragma unroll
    for (int s = kEmpMaxBins / 2; s; s >>= 1) lo = (T.x[lo + s] <= vc) ? lo + s : lo;
  }
  lo = max(min(lo, n_bins - 2), 0);
  const float4 sg = T.seg[lo];
  const float d = vc - sg.x;
  const float mu = fmaf(sg.z, d, sg.y);
  const float ss = fmaxf(0.f, fmaf(T.dsd[lo], d, sg.w));
  const float a = __fdividef(-mu, ss > 1e-9f ? ss : 1.f);
  float x;
  if (a > 0.f) {   // negative mean error (an extrapolated table): the mirrored tail, library functions (rare)
    x = -normcdfinvf((1.f - u) * normcdff(-a));
  } else {
    const float pa = ncdf_neg_fast(a);
    x = ndtri_fast(fminf(fmaf(u, 1.f - pa, pa), 0.99999994f));
  }
  return (v == v) ? fmaf(ss, x, mu) : v;      // NaN in, NaN out
}
__device__ __forceinline__ bool emp_below_snr_fast(const EmpiricalModelDev& M, float f, float e) {
  // AB: flux / (flux e ln10 / 2.5) does not depend on the flux; linear units: the unit cancels
  const float snr = M.internal_is_ab ? 1.0857362f / e : f / e;
  return !isfinite(snr) || !isfinite(f) || snr < (float)M.snr_threshold;
}

// The common case as its own straight-line body (no unit change other than a factor, no clipping, no re-draw, no upper
// limits): the general body below spends as many instructions on its (block-uniform) switches as on arithmetic.
__device__ __forceinline__ void emp_element_simple(const EmpFastTab& T, double fin, float u_sig, float zz, double& of, double& os) {
  const double fi_d = fin * T.c_in;
  const float sig0 = emp_sigma_fast(T, T.n_bins, T.extrapolate, (float)fi_d, u_sig);
  of = (fi_d + (double)(sig0 * zz)) * T.c_out;
  os = (double)fminf(fmaxf(sig0 * T.c_out_f, T.min_err), T.max_err);
}

// one element: uniforms u_sig, u_obs, u_lim in (0, 1); zz the scatter's standard normal, or a uniform when the model clips
__device__ __forceinline__ void emp_element_fast(const EmpiricalModelDev& M, const EmpFastTab& T, double fin, float u_sig, float zz,
                                                 float u_obs, float u_lim, double& of, double& os) {
  const int nb = M.n_bins, ex = M.extrapolate;
  if (M.asinh_mode) {   // AsinhEmpiricalUncertaintyModel.apply_noise (noise_models.py:507-557)
    const float b = (float)M.asinh_b;
    const float fjy = M.in_is_ab ? ab_to_jy_f((float)fin) : (float)(fin * M.in_to_jy);
    float m_noisy, err;
    if (M.asinh_mode == 1) {
      const float m_true = -1.0857362f * (asinhf(__fdividef(fjy, 2.f * b)) + T.log_b);
      const float e0 = emp_sigma_fast(T, nb, ex, m_true, u_sig);
      m_noisy = fmaf(e0, zz, m_true);
      err = M.observed_error ? emp_sigma_fast(T, nb, ex, m_noisy, u_obs) : e0;
    } else {
      const float itj = (float)M.internal_to_jy;
      const float e0 = emp_sigma_fast(T, nb, ex, __fdividef(fjy, itj), u_sig) * itj;
      const float njy = fmaf(e0, zz, fjy);
      m_noisy = -1.0857362f * (asinhf(__fdividef(njy, 2.f * b)) + T.log_b);
      const float e1 = M.observed_error ? e0 : emp_sigma_fast(T, nb, ex, __fdividef(njy, itj), u_obs) * itj;
      err = 1.0857362f * e1 * rsqrtf(fmaf(njy, njy, 4.f * b * b));
    }
    of = (double)m_noisy;
    os = fmin(fmax((double)err, M.min_err), M.max_err);
    return;
  }
  // 1. to the interpolation unit.  Where the unit change is linear the value stays in float64 and only the noise term is
  //    float32, so an unscattered element (an upper limit kept as it is) comes back exactly
  double fi_d;
  if (M.in_is_ab == M.internal_is_ab) fi_d = M.in_is_ab ? fin : fin * (M.in_to_jy / M.internal_to_jy);
  else if (M.in_is_ab) fi_d = (double)(ab_to_jy_f((float)fin) / (float)M.internal_to_jy);
  else fi_d = (double)jy_to_ab_f((float)(fin * M.in_to_jy));
  const float fi = (float)fi_d;
  // 2. sampled uncertainty at the true flux
  const float sig0 = emp_sigma_fast(T, nb, ex, fi, u_sig);
  // 3. scatter, unless the source is already below the SNR threshold
  const bool init_lim = M.upper_limits && emp_below_snr_fast(M, fi, sig0);
  if (M.sigma_clip >= 0.0) zz = ndtri_fast(fminf(fmaf(zz, T.clip_dp, T.clip_p0), 0.99999994f));
  double noisy_d = init_lim ? fi_d : fi_d + (double)(sig0 * zz);
  const float noisy = (float)noisy_d;
  // 4. "observed" errors: drawn again at the noisy flux
  float sig = M.observed_error ? emp_sigma_fast(T, nb, ex, noisy, u_obs) : sig0;
  // 5. upper limits
  if (M.upper_limits && M.ul_active && (init_lim || emp_below_snr_fast(M, noisy, sig))) {
    noisy_d = M.ul_scatter_std >= 0.0 ? M.ul_flux + M.ul_scatter_std * (double)ndtri_fast(fmaf(u_lim, T.lim_dp, T.lim_p0)) : M.ul_flux;
    sig = (float)M.ul_err;
  }
  // 6. to the output unit, error clip
  double s_d;
  if (M.out_is_ab == M.internal_is_ab) {
    const double c = M.out_is_ab ? 1.0 : M.internal_to_jy / M.out_to_jy;
    of = noisy_d * c;
    s_d = (double)sig * c;
  } else if (M.internal_is_ab) {
    const float fj = ab_to_jy_f((float)noisy_d), inv = 1.f / (float)M.out_to_jy;
    of = (double)(fj * inv);
    s_d = (double)(fj * sig * 0.92103404f * inv);
  } else {
    const float fj = (float)noisy_d * (float)M.internal_to_jy;
    of = (double)jy_to_ab_f(fj);
    s_d = (double)fabsf(1.0857362f * __fdividef(sig * (float)M.internal_to_jy, fj));
  }
  os = fmin(fmax(s_d, M.min_err), M.max_err);
}

__global__ void __launch_bounds__(256, 4) empirical_noise_fast_kernel(EmpiricalArgs A) {
  __shared__ EmpiricalModelDev M;
  __shared__ EmpFastTab T;
  const int f = blockIdx.y;
  {
    const int* src = reinterpret_cast<const int*>(A.models + f);
    int* dst = reinterpret_cast<int*>(&M);
    for (int i = threadIdx.x; i < (int)(sizeof(EmpiricalModelDev) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kEmpMaxBins; i += blockDim.x) {
    const int n = M.n_bins;
    const bool in = i < n, seg = i < n - 1;
    const double w = seg ? M.centers[i + 1] - M.centers[i] : 1.0;
    T.x[i] = in ? (float)M.centers[i] : INFINITY;
    T.seg[i] = make_float4(in ? (float)M.centers[i] : 0.f, in ? (float)M.median[i] : 0.f,
                           seg ? (float)((M.median[i + 1] - M.median[i]) / w) : 0.f, in ? (float)M.stdev[i] : 0.f);
    T.dsd[i] = seg ? (float)((M.stdev[i + 1] - M.stdev[i]) / w) : 0.f;
  }
  if (threadIdx.x == 0) {
    T.x0 = (float)M.centers[0]; T.xlast = (float)M.centers[M.n_bins - 1];
    T.lo_clamp = M.extrapolate ? -INFINITY : T.x0; T.hi_clamp = M.extrapolate ? INFINITY : T.xlast;
    const double c = M.sigma_clip >= 0.0 ? M.sigma_clip : 0.0;
    T.clip_p0 = (float)normcdf(-c); T.clip_dp = (float)(normcdf(c) - normcdf(-c));
    T.lim_p0 = (float)normcdf(-3.0); T.lim_dp = (float)(normcdf(3.0) - normcdf(-3.0));
    T.log_b = M.asinh_mode ? (float)log(M.asinh_b / 3631.0) : 0.f;
    T.min_err = (float)M.min_err; T.max_err = (float)M.max_err;
    T.n_bins = M.n_bins; T.extrapolate = M.extrapolate;
    const bool same = M.in_is_ab == M.internal_is_ab && M.out_is_ab == M.internal_is_ab;
    T.c_in = (same && !M.internal_is_ab) ? M.in_to_jy / M.internal_to_jy : 1.0;
    T.c_out = (same && !M.internal_is_ab) ? M.internal_to_jy / M.out_to_jy : 1.0;
    T.simple = same && !M.asinh_mode && !M.upper_limits && !M.observed_error && M.sigma_clip < 0.0;
    T.c_out_f = (float)T.c_out;
    const double dx = (M.centers[M.n_bins - 1] - M.centers[0]) / (M.n_bins - 1);
    bool even = dx > 0.0;
    for (int i = 1; i < M.n_bins && even; ++i) even = fabs(M.centers[i] - (M.centers[0] + i * dx)) <= 1e-6 * dx;
    T.inv_dx = even ? (float)(1.0 / dx) : 0.f;
  }
  __syncthreads();
  // the second Philox block (re-draw and upper-limit uniforms) only for the models that consume it
  const bool need2 = M.observed_error || M.asinh_mode == 2 || (M.upper_limits && M.ul_active && M.ul_scatter_std >= 0.0);
  const bool clipped = !M.asinh_mode && M.sigma_clip >= 0.0;
  const bool simple = T.simple != 0;
  const bool vec = (A.n & 1) == 0 && ((reinterpret_cast<uintptr_t>(A.flux) | reinterpret_cast<uintptr_t>(A.out_flux) |
                                       reinterpret_cast<uintptr_t>(A.out_sigma)) & 15) == 0;
  const long long n_pairs = (A.n + 1) >> 1;
  const double* in = A.flux + (long long)f * A.n;
  double* o_f = A.out_flux + (long long)f * A.n;
  double* o_s = A.out_sigma ? A.out_sigma + (long long)f * A.n : nullptr;
  // 23-bit uniforms (k + 1/2) / 2^23: exact in float32, inside (0, 1), so no clamp
  auto uni = [](uint32_t w) { return fmaf((float)(w >> 9), 1.0f / 8388608.0f, 0.5f / 8388608.0f); };
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n_pairs; q += (long long)gridDim.x * blockDim.x) {
    const long long r = 2 * q;
    const bool two = r + 1 < A.n;
    double fin[2];
    if (vec) { const double2 v = *reinterpret_cast<const double2*>(in + r); fin[0] = v.x; fin[1] = v.y; }
    else { fin[0] = in[r]; fin[1] = two ? in[r + 1] : in[r]; }
    uint32_t a4[4];
    philox4x32_10_keyed((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)f | 0x40000000u, (uint32_t)A.epoch, A.rk0, A.rk1, a4);
    float zz[2];
    if (clipped) { zz[0] = uni(a4[2]); zz[1] = uni(a4[3]); }
    else {   // one Box-Muller pair: both normals are used
      float sn, cs;
      __sincosf(6.2831853f * (uni(a4[3]) - 0.5f), &sn, &cs);
      float rad;
      asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(fmaxf(-2.f * __logf(uni(a4[2])), 0.f)));
      zz[0] = rad * cs; zz[1] = rad * sn;
    }
    double of[2], os[2];
    if (simple) {
      emp_element_simple(T, fin[0], uni(a4[0]), zz[0], of[0], os[0]);
      emp_element_simple(T, fin[1], uni(a4[1]), zz[1], of[1], os[1]);
    } else {
      uint32_t b4[4] = {0x80000000u, 0x80000000u, 0x80000000u, 0x80000000u};
      if (need2) philox4x32_10_keyed((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)f | 0xC0000000u, (uint32_t)A.epoch, A.rk0, A.rk1, b4);
      emp_element_fast(M, T, fin[0], uni(a4[0]), zz[0], uni(b4[0]), uni(b4[2]), of[0], os[0]);
      emp_element_fast(M, T, fin[1], uni(a4[1]), zz[1], uni(b4[1]), uni(b4[3]), of[1], os[1]);
    }
    if (vec) {
      *reinterpret_cast<double2*>(o_f + r) = make_double2(of[0], of[1]);
      if (o_s) *reinterpret_cast<double2*>(o_s + r) = make_double2(os[0], os[1]);
    } else {
      o_f[r] = of[0];
      if (o_s) o_s[r] = os[0];
      if (two) { o_f[r + 1] = of[1]; if (o_s) o_s[r + 1] = os[1]; }
    }
  }
}

}  // namespace sb2
