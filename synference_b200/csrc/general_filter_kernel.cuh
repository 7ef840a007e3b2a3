// Filter integration on a GENERAL (not constant-R) wavelength axis -- the fallback behind the shift-table epilogue.
//
// The fused epilogue of the contraction kernels needs grid and filters on one geometric axis (re-interpolating a filter
// onto the observed abscissa is then an integer shift plus one blend weight).  The reference's README and tests keep the
// SPS grid's native axis (README.md:100-102, tests/conftest.py:70,85); for such models the contraction kernel writes the
// observed-frame spectrum of a slice of galaxies to HBM (spec_out) and this kernel integrates every filter with the
// reference's general semantics (Sed.get_photo_fnu -> Filter.apply_filter):
//   x_i = c / (lam_i (1+z))  (or lam_obs in the 'lam' variant);  T_i = the filter's own table interpolated linearly IN x
//   at x_i, 0 outside its range;  flux = trapz(f T / x, x) / trapz(T / x, x) over the samples with T_i > 0 only --
//   consecutive KEPT samples are joined, so an interior zero is bridged exactly as numpy's compress + trapz does.
// One warp per (galaxy, filter): lanes take 32 consecutive wavelengths, a ballot finds each kept lane's previous kept
// sample (inside the 32, or the carry of earlier ones), float64 abscissae (the trapezoid widths are differences of
// neighbouring x: 1/R ~ 1e-3 relative, out of float32's reach).  ~15 KB per galaxy round-trip through HBM: a slow path by
// design, taken only when the axis is not geometric.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sb2 {

struct GeneralFilterArgs {
  const float* spectra;     // [n][n_lam] observed-frame f_nu at base mass on the rest axis (spec_out of the contraction kernel)
  const double* redshift;   // [n]
  const double* log_mass;   // [n] or nullptr
  double base_mass;
  long long n;
  int n_lam, n_filt, variant;   // variant 0: integrate in nu, 1: in lambda
  const double* lam;        // [n_lam] rest-frame axis, ascending
  const long long* off;     // [n_filt + 1] offsets into the filter tables
  const double* f_lam;      // filter tables: wavelengths (ascending) ...
  const double* f_t;        // ... and transmissions
  float* flux_base;         // [n][n_filt] or nullptr
  double* flux_scaled;      // [n][n_filt] or nullptr
};

constexpr double kCAngstrom = 2.99792458e18;

__global__ void __launch_bounds__(256) general_filter_kernel(const GeneralFilterArgs A) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const unsigned FULL = 0xffffffffu;
  for (long long job = warp; job < A.n * A.n_filt; job += n_warps) {
    const long long g = job / A.n_filt;
    const int f = (int)(job - g * A.n_filt);
    const double z = A.redshift[g];
    const double zp = 1.0 + z;
    const double* fl = A.f_lam + A.off[f];
    const double* ft = A.f_t + A.off[f];
    const int nt = (int)(A.off[f + 1] - A.off[f]);
    // rest-frame bins whose observed wavelength falls inside the filter's table: lam_i zp in [fl[0], fl[nt-1]]
    int i_lo = 0, i_hi = A.n_lam;     // first i with lam_i zp >= fl[0];  first i with lam_i zp > fl[nt-1]
    {
      int lo = 0, hi = A.n_lam;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (A.lam[mid] * zp < fl[0]) lo = mid + 1; else hi = mid; }
      i_lo = lo;
      lo = i_lo; hi = A.n_lam;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (A.lam[mid] * zp <= fl[nt - 1]) lo = mid + 1; else hi = mid; }
      i_hi = lo;
    }
    double num = 0.0, den = 0.0;
    double cx = 0.0, cy = 0.0, cw = 0.0;   // carry: the last kept sample of earlier chunks
    bool have = false;
    int k = 0;                             // table segment pointer (monotone in i within a lane's stride is not guaranteed: re-searched)
    const float* sp = A.spectra + (size_t)g * A.n_lam;
    for (int i0 = i_lo; i0 < i_hi; i0 += 32) {
      const int i = i0 + lane;
      double x = 0.0, y = 0.0, w = 0.0;
      bool keep = false;
      if (i < i_hi) {
        const double lobs = A.lam[i] * zp;
        // segment k with fl[k] <= lobs <= fl[k+1]
        int lo = 0, hi = nt - 1;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (fl[mid] <= lobs) lo = mid; else hi = mid; }
        k = lo;
        double t;
        if (A.variant == 0) {   // linear in nu between the table nodes (np.interp on the reversed table)
          const double xa = kCAngstrom / fl[k], xb = kCAngstrom / fl[k + 1];
          x = kCAngstrom / lobs;
          t = (xa == xb) ? ft[k] : ft[k + 1] + (ft[k] - ft[k + 1]) * ((x - xb) / (xa - xb));
        } else {
          x = lobs;
          t = (fl[k + 1] == fl[k]) ? ft[k] : ft[k] + (ft[k + 1] - ft[k]) * ((lobs - fl[k]) / (fl[k + 1] - fl[k]));
        }
        keep = t > 0.0;
        if (keep) { w = t / x; y = (double)sp[i] * w; }
      }
      const unsigned mask = __ballot_sync(FULL, keep);
      const unsigned below = mask & ((1u << lane) - 1u);
      const int src = below ? 31 - __clz(below) : 0;
      const double px = __shfl_sync(FULL, x, src), py = __shfl_sync(FULL, y, src), pw = __shfl_sync(FULL, w, src);
      if (keep) {
        if (below) { num += 0.5 * (y + py) * (x - px); den += 0.5 * (w + pw) * (x - px); }
        else if (have) { num += 0.5 * (y + cy) * (x - cx); den += 0.5 * (w + cw) * (x - cx); }
      }
      if (mask) {
        const int last = 31 - __clz(mask);
        cx = __shfl_sync(FULL, x, last); cy = __shfl_sync(FULL, y, last); cw = __shfl_sync(FULL, w, last);
        have = true;
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      num += __shfl_xor_sync(FULL, num, o);
      den += __shfl_xor_sync(FULL, den, o);
    }
    if (lane == 0) {
      // (no in-band sample: the reference raises; a lone sample has no trapezoid either -> NaN)
      const float fb = (have && den != 0.0 && z >= 0.0) ? (float)(num / den) : __int_as_float(0x7fc00000);
      if (A.flux_base) A.flux_base[g * A.n_filt + f] = fb;
      if (A.flux_scaled) {
        const double ms = A.log_mass ? pow(10.0, A.log_mass[g]) / A.base_mass : 1.0;
        A.flux_scaled[g * A.n_filt + f] = (double)fb * ms;
      }
    }
  }
}

}  // namespace sb2
