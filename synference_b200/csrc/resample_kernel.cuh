// Spectroscopic training path: one library spectrum -> one instrument-frame feature row.
//   redshift the wavelength axis, smooth with a Gaussian whose width varies per pixel (instrument R(lambda) in quadrature
//   with the model's own resolution; utils.py:129-182), rebin flux-conservingly to the observed pixels (spectres, as called
//   by transform_spectrum, utils.py:185-254).  The reference does this galaxy by galaxy in Python
//   (sbi_runner.py:1322-1334); here one CTA takes one galaxy:
//     1. the observed pixel range picks the rest-frame bins [k_lo, k_hi] that can contribute (binary search on bin edges),
//     2. each thread evaluates sigma_pix of its bins; the block agrees on the widest kernel half-width H,
//     3. the slice [k_lo - H, k_hi + H] of the spectrum is staged in shared memory (indices clamped to the axis: the
//        reference pads with the nearest edge value), i.e. only the part of the row that matters is read from HBM,
//     4. smoothing: thread per bin, symmetric taps, weights exp(-x^2 / 2 sigma^2) from ex2 (a multiplicative recurrence was
//        rejected: the rounding of its seed acts like a relative error sigma^2 * 6e-8 on sigma),
//     5. rebin: thread per observed pixel, two binary searches on the edges (read-only cache), overlap-weighted mean.
// HBM-bound by design: algorithmic bytes per galaxy = 4 * (bins under the observed window + 2H) read + 4 * n_px written.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb2 {

struct ResampleArgs {
  const float* spectra;      // [n][n_lam]
  const double* redshift;    // [n]
  float* out;                // [n][n_px]
  long long n;
  int n_lam, n_px, n_res;
  const double* old_edges;   // [n_lam + 1] rest-frame bin edges (spectres convention)
  const double* new_edges;   // [n_px + 1]  observed-frame pixel edges
  const float* lam;          // [n_lam] rest-frame wavelengths
  const float* lam_k;        // [n_lam] lam / (median(diff(lam)) * 2 sqrt(2 ln 2)): sigma_pix = lam_k * sqrt(1/R_inst^2 - 1/R_th^2)
  const float* qth2;         // [n_lam] 1 / R_theory^2 (0: no intrinsic broadening)
  const float* res_wave;     // [n_res] resolution curve abscissa (observed frame, increasing)
  const float* res_r;        // [n_res]
  const float* res_slope;    // [n_res - 1] (R[j+1] - R[j]) / (wave[j+1] - wave[j])
  float trunc;
  float fill;
  int h_cap;                 // widest kernel half-width the shared-memory staging buffer was sized for
  int nb_max;                // most bins the launch reserved shared memory for (any redshift's window fits: see capi)
  // Search accelerators (nullptr: plain binary search).  A table over uniform steps of log2(wavelength) holds, per bucket, the
  // last edge / curve node at or below the bucket's left end: a lower bound from which the exact answer is a few steps away.
  const int* edge_lut;       // [lut_n]
  const int* res_lut;        // [res_lut_n]
  int lut_n, res_lut_n;
  float lut_u0, lut_inv_du, res_u0, res_inv_du;
};

constexpr int kResampleThreads = 256;

// A lower bound for "last table entry <= x" from a log2-uniform bucket table: the bucket is found from an approximate
// logarithm, so one bucket of margin is taken; entries are then walked forward with the EXACT comparison.
__device__ __forceinline__ int lut_guess(const int* __restrict__ lut, int lut_n, float u0, float inv_du, float x) {
  const int b = (int)floorf((__log2f(x) - u0) * inv_du) - 1;
  return __ldg(lut + min(max(b, 0), lut_n - 1));
}

// largest k in [0, n] with edges[k] * s <= x  (0 if none: callers test the ends of the axis themselves)
__device__ __forceinline__ int last_edge_le(const ResampleArgs& A, double s, float inv_sf, double x) {
  const double* __restrict__ e = A.old_edges;
  const int n = A.n_lam;
  if (A.edge_lut) {
    int k = lut_guess(A.edge_lut, A.lut_n, A.lut_u0, A.lut_inv_du, (float)x * inv_sf);
    while (k < n && __ldg(e + k + 1) * s <= x) ++k;
    return k;
  }
  int lo = -1, hi = n + 1;   // invariant: e[lo]*s <= x < e[hi]*s (with sentinels)
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(e + mid) * s <= x) lo = mid; else hi = mid;
  }
  return max(lo, 0);
}
// largest k in [0, n] with edges[k] * s < x  (0 if none)
__device__ __forceinline__ int last_edge_lt(const ResampleArgs& A, double s, float inv_sf, double x) {
  const double* __restrict__ e = A.old_edges;
  const int n = A.n_lam;
  if (A.edge_lut) {
    int k = lut_guess(A.edge_lut, A.lut_n, A.lut_u0, A.lut_inv_du, (float)x * inv_sf);
    while (k > 0 && !(__ldg(e + k) * s < x)) --k;     // the table answers "<=": step back over an edge equal to x
    while (k < n && __ldg(e + k + 1) * s < x) ++k;
    return k;
  }
  int lo = -1, hi = n + 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(e + mid) * s < x) lo = mid; else hi = mid;
  }
  return max(lo, 0);
}

// np.interp(x, xp, fp): linear, clamped to the end values
__device__ __forceinline__ float interp_clamped(const ResampleArgs& A, float x) {
  const float* __restrict__ xp = A.res_wave;
  const float* __restrict__ fp = A.res_r;
  const int n = A.n_res;
  if (x <= __ldg(xp)) return __ldg(fp);
  if (x >= __ldg(xp + n - 1)) return __ldg(fp + n - 1);
  int lo = 0, hi = n - 1;    // xp[lo] <= x < xp[hi]
  if (A.res_lut) {
    lo = lut_guess(A.res_lut, A.res_lut_n, A.res_u0, A.res_inv_du, x);
    while (lo < n - 2 && __ldg(xp + lo + 1) <= x) ++lo;
  } else {
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(xp + mid) <= x) lo = mid; else hi = mid;
    }
  }
  return fmaf(x - __ldg(xp + lo), __ldg(A.res_slope + lo), __ldg(fp + lo));
}

__global__ void __launch_bounds__(kResampleThreads) resample_kernel(const __grid_constant__ ResampleArgs A) {
  extern __shared__ float smem[];       // [n_bins] + [n_bins + 2H]; the launch reserves 2 n_lam + 2 h_cap floats
  __shared__ int s_h;
  const int tid = threadIdx.x;
  for (long long g = blockIdx.x; g < A.n; g += gridDim.x) {
    const double s = 1.0 + A.redshift[g];
    const float sf = (float)s;
    const float inv_sf = 1.f / sf;
    float* out = A.out + (size_t)g * A.n_px;
    const bool bad = !(s > 0.0) || !isfinite(s);
    // ---- 1. rest-frame bins that the observed window can touch
    int k_lo = 0, k_hi = -1;
    if (!bad) {
      k_lo = min(last_edge_le(A, s, inv_sf, __ldg(A.new_edges)), A.n_lam - 1);
      k_hi = min(last_edge_lt(A, s, inv_sf, __ldg(A.new_edges + A.n_px)), A.n_lam - 1);
      // (window entirely off either end of the axis: every pixel gets `fill` in step 5; the range below is then harmless)
    }
    // (a window wider than the launch sized shared memory for cannot happen by construction; if it ever did, the row is
    //  marked NaN rather than written past the buffers)
    const bool too_wide = !bad && (k_hi - k_lo + 1) > A.nb_max;
    if (bad || too_wide || k_hi < k_lo) {   // nothing of the spectrum under the observed window (or an unusable redshift)
      for (int j = tid; j < A.n_px; j += kResampleThreads) out[j] = (bad || too_wide) ? __int_as_float(0x7fc00000) : A.fill;
      continue;
    }
    const int n_bins = k_hi - k_lo + 1;
    float* s_conv = smem;              // [n_bins]: sigma_pix of each bin, then its smoothed flux
    float* s_flux = smem + n_bins;     // [n_bins + 2H]: the staged slice
    // ---- 2. kernel widths.  sigma_pix = sigma_wave / pixel, with sigma_wave = lam (1+z) sqrt(1/R_i^2 - 1/R_t^2) / 2.3548 and
    //         pixel = (1+z) median(diff(lam)): the (1+z) cancels, z only enters through R_inst(lam (1+z)).
    if (tid == 0) s_h = 0;
    __syncthreads();
    int h_max = 0;
    for (int i = k_lo + tid; i <= k_hi; i += kResampleThreads) {
      const float q = __frcp_rn(interp_clamped(A, __ldg(A.lam + i) * sf));
      const float v = fmaf(q, q, -__ldg(A.qth2 + i));
      const float sg = v > 0.f ? __ldg(A.lam_k + i) * sqrtf(v) : 0.f;
      s_conv[i - k_lo] = sg;
      if (sg > 0.01f) h_max = max(h_max, (int)ceilf(sg * A.trunc));
    }
    h_max = __reduce_max_sync(0xffffffffu, h_max);
    if ((tid & 31) == 0) atomicMax(&s_h, h_max);
    __syncthreads();
    // (kernels wider than the staging buffer allows -- possible only far outside any real instrument's R(lambda) -- read
    //  their taps from global memory instead: slower, same result)
    const bool staged = s_h <= A.h_cap;
    const int H = staged ? s_h : 0;
    // ---- 3. stage flux[clamp(k_lo - H + p)] for p in [0, n_bins + 2H)
    const int n_stage = n_bins + 2 * H;
    const float* row = A.spectra + (size_t)g * A.n_lam;
    for (int p = tid; p < n_stage; p += kResampleThreads)
      s_flux[p] = __ldg(row + min(max(k_lo - H + p, 0), A.n_lam - 1));
    __syncthreads();
    // ---- 4. smoothing (each thread overwrites the sigma it stored itself)
    for (int i = k_lo + tid; i <= k_hi; i += kResampleThreads) {
      const float sg = s_conv[i - k_lo];
      const float* f = s_flux + (i - k_lo + H);
      float r = f[0];
      if (sg > 0.01f) {
        const int hw = (int)ceilf(sg * A.trunc);
        const float a2 = -0.72134752044448170368f / (sg * sg);   // -log2(e) / (2 sigma^2)
        float num = f[0], den = 1.f, xf = 1.f;
        if (staged) {
#pragma unroll 2
          for (int x = 1; x <= hw; ++x, xf += 1.f) {
            float w;
            const float arg = a2 * xf * xf;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(arg));
            num = fmaf(w, f[-x] + f[x], num);
            den = fmaf(2.f, w, den);
          }
        } else {
          for (int x = 1; x <= hw; ++x, xf += 1.f) {
            float w;
            const float arg = a2 * xf * xf;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(arg));
            num = fmaf(w, __ldg(row + max(i - x, 0)) + __ldg(row + min(i + x, A.n_lam - 1)), num);
            den = fmaf(2.f, w, den);
          }
        }
        r = num / den;
      }
      s_conv[i - k_lo] = r;
    }
    __syncthreads();
    // ---- 5. flux-conserving rebin
    const double* __restrict__ oe = A.old_edges;
    const double e_first = __ldg(oe) * s, e_last = __ldg(oe + A.n_lam) * s;
    for (int j = tid; j < A.n_px; j += kResampleThreads) {
      const double a = __ldg(A.new_edges + j), b = __ldg(A.new_edges + j + 1);
      float v = A.fill;
      if (!(a < e_first) && !(b > e_last)) {
        const int start = min(last_edge_le(A, s, inv_sf, a), A.n_lam - 1);
        int stop = start;                                   // spectres' own walk: the last bin that starts before b
        while (stop < A.n_lam - 1 && __ldg(oe + stop + 1) * s < b) ++stop;
        if (stop == start) {
          v = s_conv[start - k_lo];
        } else {
          // overlap widths: first and last bins partially, the others whole
          const float w0 = (float)(__ldg(oe + start + 1) * s - a);
          const float w1 = (float)(b - __ldg(oe + stop) * s);
          float num = w0 * s_conv[start - k_lo] + w1 * s_conv[stop - k_lo], den = w0 + w1;
          for (int k = start + 1; k < stop; ++k) {
            const float w = (float)((__ldg(oe + k + 1) - __ldg(oe + k)) * s);
            num = fmaf(w, s_conv[k - k_lo], num);
            den += w;
          }
          v = num / den;
        }
      }
      out[j] = v;
    }
    __syncthreads();   // the staging buffer is reused by the next galaxy
  }
}

}  // namespace sb2
