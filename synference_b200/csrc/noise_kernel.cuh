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
  const float* flux32;    // the same as float32 (feature-only kernel; widened on the device, as numpy would), or nullptr
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

// Philox draws of one (row, filter QUAD): counter (row, quad, epoch), key (seed, epoch); the block's four words make two
// Box-Muller pairs = the normals of filters 4q .. 4q+3.  kFast: hardware log2 / sin / cos (feature-only kernel; the
// normals then differ from the library functions' by ~1e-6 relative, far inside the 1e-4 mag tolerance of the rows).
template <bool kFast>
__device__ __forceinline__ void philox_normals4(long long r, int quad, unsigned long long seed, unsigned long long epoch, float (&z)[4]) {
  uint32_t rnd[4];
  philox4x32_10((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)quad, (uint32_t)epoch, (uint32_t)seed,
                (uint32_t)(seed >> 32) ^ (uint32_t)(epoch >> 32), rnd);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = ((float)(rnd[2 * h] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((float)(rnd[2 * h + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float rad, sn, cs;
    if constexpr (kFast) {
      rad = sqrtf(-1.3862943611198906f * __log2f(u1));       // -2 ln u = -2 ln2 log2 u
      __sincosf(6.283185307179586f * u2, &sn, &cs);
    } else {
      rad = sqrtf(-2.0f * logf(u1));
      sincospif(2.0f * u2, &sn, &cs);
    }
    z[2 * h] = rad * cs; z[2 * h + 1] = rad * sn;
  }
}

__global__ void __launch_bounds__(256) depth_noise_kernel(NoiseArgs A) {
  const long long n_rows = A.n_gal * A.n_scatter;
  const double ln10 = 2.302585092994046;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
       r += (long long)gridDim.x * blockDim.x) {
    const long long g = r / A.n_scatter;
    for (int f0 = 0; f0 < A.n_filt; f0 += 4) {
      float zf[4] = {0.f, 0.f, 0.f, 0.f};
      if (A.normals == nullptr) philox_normals4<false>(r, f0 >> 2, A.seed, A.epoch, zf);
#pragma unroll
      for (int d = 0; d < 4; ++d) {
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

// Feature rows only, Philox draws (the per-epoch resampling of a training set): one thread per (row, filter QUAD).  The
// fluxes of the quad are read as one or two 16-byte loads (float32 | float64 input), one Philox block yields its four
// normals, and (mag, mag_err) are written as float4.  Same counters as depth_noise_kernel; the noisy flux is formed in
// float64 exactly as there, log2 / sin / cos / the error ratio run on the special-function unit (|d mag| < 6e-6,
// tolerance 1e-4 mag).  HBM-bound: 16 | 32 B read + 32 B written per quad.  The (row, quad) of a thread come from a
// per-block decomposition (a block walks whole rows), so the loop has no 64-bit division.
template <typename T>
__global__ void __launch_bounds__(256) depth_noise_feat_kernel(NoiseArgs A) {
  const long long n_rows = A.n_gal * A.n_scatter;
  const int nf = A.n_filt, nquad = (nf + 3) >> 2;
  const int rows_pb = 256 / nquad;                       // whole rows per block and pass (nquad <= 8 for 32 filters)
  const int rl = (int)threadIdx.x / nquad, q = (int)threadIdx.x - rl * nquad, f0 = 4 * q;
  if (rl >= rows_pb) return;
  const bool vec = (nf & 3) == 0;
  const T* flux = sizeof(T) == 8 ? reinterpret_cast<const T*>(A.flux) : reinterpret_cast<const T*>(A.flux32);
  const float lim = (float)A.mag_limit;
  double sdq[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) sdq[d] = A.sigma[min(f0 + d, nf - 1)];
  const bool small = n_rows < 0x7fffffffLL;
  for (long long r = (long long)blockIdx.x * rows_pb + rl; r < n_rows; r += (long long)gridDim.x * rows_pb) {
    const long long g = A.n_scatter == 1 ? r : (small ? (long long)((unsigned)r / (unsigned)A.n_scatter) : r / A.n_scatter);
    float zf[4];
    philox_normals4<true>(r, q, A.seed, A.epoch, zf);
    double rep[4];
    const T* src = flux + g * nf + f0;
    if (vec) {
      if constexpr (sizeof(T) == 8) {
        const double2 a = *reinterpret_cast<const double2*>(src), b = *reinterpret_cast<const double2*>(src + 2);
        rep[0] = a.x; rep[1] = a.y; rep[2] = b.x; rep[3] = b.y;
      } else {
        const float4 a = *reinterpret_cast<const float4*>(src);
        rep[0] = a.x; rep[1] = a.y; rep[2] = a.z; rep[3] = a.w;
      }
    } else {
#pragma unroll
      for (int d = 0; d < 4; ++d) rep[d] = (f0 + d < nf) ? (double)src[d] : 1.0;
    }
    float mag[4], merr[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      double sd = sdq[d];
      if (A.min_pc > 0.0) sd = fmax(sd, __ddiv_rn(__dmul_rn(rep[d], A.min_pc), 100.0));
      const double noisy = __dadd_rn(rep[d], __dadd_rn(0.0, __dmul_rn(sd, (double)zf[d])));
      const float f_ujy = (float)(noisy * 1e-3), e_ujy = (float)(sd * 1e-3);
      float m = fmaf(-0.7525749891599529f, __log2f(f_ujy), 23.9f);     // -2.5 log10 f = -2.5 log10(2) log2 f
      if (f_ujy < 0.f) m = lim;
      if (m > lim) m = lim;
      mag[d] = m;
      merr[d] = __fdividef(1.0857362047581296f * e_ujy, f_ujy);        // 2.5 / ln 10
    }
    float* row = A.out_feat + r * (2LL * nf);
    if (vec) {
      *reinterpret_cast<float4*>(row + f0) = make_float4(mag[0], mag[1], mag[2], mag[3]);
      *reinterpret_cast<float4*>(row + nf + f0) = make_float4(merr[0], merr[1], merr[2], merr[3]);
    } else {
#pragma unroll
      for (int d = 0; d < 4; ++d)
        if (f0 + d < nf) { row[f0 + d] = mag[d]; row[nf + f0 + d] = merr[d]; }
    }
  }
}

}  // namespace sb2
