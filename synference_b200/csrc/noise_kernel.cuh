// Depth-based scatter + flux -> AB-magnitude feature rows (elementwise, HBM-bound).
//
// Restates SBI_Fitter._apply_depths (sbi_runner.py:580-691: np.repeat columns, sigma = depth/level,
// optional percentage floor, flux + N(0, sigma)) and the AB branch of
// create_feature_array_from_raw_photometry (sbi_runner.py:1698-1716 magnitudes and errors,
// :1927-1932 clip at norm_mag_limit, :2150 float32 rows).
//
// Two RNG modes: injected float64 normal draws (bit-exact against numpy: the noisy flux is
// flux + (0 + sigma*z) with separately rounded multiply and add, no FMA contraction), or
// counter-based Philox4x32-10 keyed by (seed, epoch) with counter (row, filter) + Box-Muller.
// One thread per output row; loads/stores over rows are coalesced in the [filter][row] arrays.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sb2 {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct NoiseArgs {
  const double* flux;     // [n_gal][n_filt]
  long long n_gal;
  int n_filt, n_scatter;
  const double* sigma;    // [n_filt], or [n_sets][n_filt] with set_index
  const int* set_index;   // [n_filt][n_scatter] which depth set each (filter, block of n_gal rows) uses, or nullptr
  double min_pc;
  const double* normals;  // [n_filt][n_rows] or nullptr
  unsigned long long seed, epoch;
  double mag_limit;
  double* out_flux;       // [n_filt][n_rows] or nullptr
  double* out_sigma;      // [n_filt][n_rows] or nullptr
  float* out_feat;        // [n_rows][2*n_filt] or nullptr
};

__global__ void __launch_bounds__(256) depth_noise_kernel(NoiseArgs A) {
  const long long n_rows = A.n_gal * A.n_scatter;
  const double ln10 = 2.302585092994046;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
       r += (long long)gridDim.x * blockDim.x) {
    const long long g = r / A.n_scatter;
    for (int f0 = 0; f0 < A.n_filt; f0 += 2) {
      float zf[2] = {0.f, 0.f};
      if (A.normals == nullptr) {  // one Philox block -> two Box-Muller normals (filters f0, f0+1)
        uint32_t rnd[4];
        philox4x32_10((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)(f0 >> 1), (uint32_t)A.epoch,
                      (uint32_t)A.seed, (uint32_t)(A.seed >> 32) ^ (uint32_t)(A.epoch >> 32), rnd);
        const float u1 = ((float)(rnd[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float u2 = ((float)(rnd[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float rad = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        zf[0] = rad * cs; zf[1] = rad * sn;
      }
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const int f = f0 + d;
        if (f >= A.n_filt) break;
        const double rep = A.flux[g * A.n_filt + f];
        // 2-D depths (sbi_runner.py:626-647): rows [j n_gal, (j+1) n_gal) of the repeated array share the depth set drawn
        // for (filter, j) -- the reference expands its (m, N_scatters) pick with np.repeat(..., n, axis=1)
        double sd = A.set_index ? A.sigma[(long long)A.set_index[f * A.n_scatter + (int)(r / A.n_gal)] * A.n_filt + f] : A.sigma[f];
        if (A.min_pc > 0.0) sd = fmax(sd, __ddiv_rn(__dmul_rn(rep, A.min_pc), 100.0));
        const double z = A.normals ? A.normals[(long long)f * n_rows + r] : (double)zf[d];
        const double noisy = __dadd_rn(rep, __dadd_rn(0.0, __dmul_rn(sd, z)));
        if (A.out_flux) A.out_flux[(long long)f * n_rows + r] = noisy;
        if (A.out_sigma) A.out_sigma[(long long)f * n_rows + r] = sd;
        if (A.out_feat) {
          const double f_ujy = noisy * 1e-3, e_ujy = sd * 1e-3;
          double merr = 2.5 * e_ujy / (ln10 * f_ujy);
          double mag = -2.5 * log10(f_ujy) + 23.9;
          if (f_ujy < 0.0) mag = A.mag_limit;
          if (mag > A.mag_limit) mag = A.mag_limit;
          A.out_feat[r * (2LL * A.n_filt) + f] = (float)mag;
          A.out_feat[r * (2LL * A.n_filt) + A.n_filt + f] = (float)merr;
        }
      }
    }
  }
}

// Feature rows only, Philox draws (the per-epoch resampling of a training set): one thread per (row, filter PAIR),
// so the float64 fluxes are read as coalesced 16-byte pairs, one Philox block yields the pair's two normals, and the
// (mag, mag_err) rows are written as coalesced float2.  Same counters, hence the same draws, as depth_noise_kernel; the
// noisy flux is formed in float64 exactly as there, only log10 and the error ratio run in float32 (|d mag| < 6e-6,
// tolerance 1e-4 mag).  HBM-bound: 16 B read + 16 B written per pair.
__global__ void __launch_bounds__(256) depth_noise_feat_kernel(NoiseArgs A) {
  const long long n_rows = A.n_gal * A.n_scatter;
  const int npair = (A.n_filt + 1) >> 1;
  const long long total = n_rows * npair;
  const bool even = (A.n_filt & 1) == 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / npair;
    const int pr = (int)(i - r * npair), f0 = 2 * pr;
    const long long g = r / A.n_scatter;
    uint32_t rnd[4];
    philox4x32_10((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)pr, (uint32_t)A.epoch, (uint32_t)A.seed,
                  (uint32_t)(A.seed >> 32) ^ (uint32_t)(A.epoch >> 32), rnd);
    const float u1 = ((float)(rnd[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((float)(rnd[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    const float zf[2] = {rad * cs, rad * sn};
    double rep[2];
    if (even) {
      const double2 v = *reinterpret_cast<const double2*>(A.flux + g * A.n_filt + f0);
      rep[0] = v.x; rep[1] = v.y;
    } else {
      rep[0] = A.flux[g * A.n_filt + f0];
      rep[1] = (f0 + 1 < A.n_filt) ? A.flux[g * A.n_filt + f0 + 1] : 1.0;
    }
    float mag[2], merr[2];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const int f = min(f0 + d, A.n_filt - 1);
      double sd = A.sigma[f];
      if (A.min_pc > 0.0) sd = fmax(sd, __ddiv_rn(__dmul_rn(rep[d], A.min_pc), 100.0));
      const double noisy = __dadd_rn(rep[d], __dadd_rn(0.0, __dmul_rn(sd, (double)zf[d])));
      const float f_ujy = (float)(noisy * 1e-3), e_ujy = (float)(sd * 1e-3);
      float m = -2.5f * log10f(f_ujy) + 23.9f;
      const float lim = (float)A.mag_limit;
      if (f_ujy < 0.f) m = lim;
      if (m > lim) m = lim;
      mag[d] = m;
      merr[d] = 2.5f * e_ujy / (2.302585092994046f * f_ujy);
    }
    float* row = A.out_feat + r * (2LL * A.n_filt);
    if (even) {
      *reinterpret_cast<float2*>(row + f0) = make_float2(mag[0], mag[1]);
      *reinterpret_cast<float2*>(row + A.n_filt + f0) = make_float2(merr[0], merr[1]);
    } else {
      row[f0] = mag[0]; row[A.n_filt + f0] = merr[0];
      if (f0 + 1 < A.n_filt) { row[f0 + 1] = mag[1]; row[A.n_filt + f0 + 1] = merr[1]; }
    }
  }
}

}  // namespace sb2
