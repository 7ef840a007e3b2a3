// Weight construction ("prep") kernel: one warp per galaxy, float64.
//
// For the galaxy in sorted slot t it produces everything the contraction kernel needs:
//   * SFZH weights  w[iz*n_age + ia] = sf[ia] * zd[iz]  (Stars.__init__ in the
//     reference, library.py:1372-1379), written as a TF32 hi/lo pair (3xTF32 operand A);
//   * the Inoue+14 transmission row exp(-tau(z, lam_i (1+z))) for the bins blueward of Ly-alpha;
//   * per-galaxy scalars: integer/fractional filter shift (m, beta) on the geometric axis,
//     dust exponent scale, flux scale (1+z)/(4 pi d_L^2) from the cosmology table, mass scale.
// HBM-bound by design: algorithmic bytes per galaxy = 2 * 4 * k_pad (weights) + 4 * n_blue (IGM).
#pragma once
#include "ptx.cuh"
#include "../../include/synference_b200.h"

namespace sb2 {

struct PrepModel {
  int n_age, na_pad, n_z, K, k_pad, n_lam, n_filt, n_blue, n_lines, variant, igm_on;
  int delta;     // 1: DeltaConstant batch grouped by metallicity bracket -> weights row holds only the
                 //    two bracketing grid metallicities: [sf*(1-f) (na_pad) | sf*f (na_pad)], stride w_stride
  int w_stride;  // floats per weights row (k_pad, or 2*na_pad in delta mode)
  const double* ages;      // [n_age] yr
  const double* edges;     // [n_age] e_0..e_{n_age-1} (bin a spans [e_a, e_{a+1}], a < n_age-1)
  const double* zmet;      // [n_z]
  const double* log10zmet; // [n_z]
  double lam0, q, ln_q, grid_scale, base_mass;
  const int* filt_lo;
  const int* filt_hi;
  // igm tables
  const double* bin_pow;  // [8][n_blue]
  const int* nline;       // [n_blue]
  const int* lc_on;       // [n_blue]
  const double* thr;      // [3][64]
  const double* pre;      // [5][n_lines+1]
  // cosmology
  int cosmo_n;
  double cosmo_ds;
  const double* dc;
  const double* ddc;
  const double* age;
  const double* dage;
};

struct PrepParams {  // device pointers (sb2_params with device arrays)
  long long n;
  const double* redshift;
  const double* log_mass;
  const double* tau_v;
  int sfh_type;
  int sfh_stride;
  const double* sfh_rows;
  int max_age_from_z;
  unsigned norm_mask;
  double age_zmax_gyr;
  int zd_type;
  const double* zd_value;
  const double* zd_sigma;
  const double* coef_att;
  const double* coef_unatt;
};

struct PrepOut {
  float* w_hi;      // [n_pad][k_pad]
  float* w_lo;      // [n_pad][k_pad]
  double* w_f64;    // optional [n][K] in ORIGINAL order (parity hook), else nullptr
  float* igm;       // [n_tiles][n_blue][128]   (written by igm_kernel)
  double* zpow;     // [13][n_pad] powers of (1+z) for igm_kernel, then z
  int* g_m;         // [n_pad]
  float* g_beta;    // [n_pad]
  float* g_gamma;   // [n_pad] 1 - beta
  float* g_taut;    // [n_pad]
  float* g_scale;   // [n_pad]
  float* g_ca;      // [n_pad]
  float* g_cb;      // [n_pad]
  int* g_orig;      // [n_pad]  original index, -1 for padding rows
  double* g_mscale; // [n_pad]
  unsigned* g_trunc;// [n_pad]  bit f set: filter f not fully covered by the grid at this z
};

__device__ __forceinline__ double hermite_lut(const double* y, const double* dy, double ds, int n, double s) {
  double x = s / ds;
  x = fmin(fmax(x, 0.0), (double)n - 1e-9);
  int k = (int)x;
  double t = x - (double)k;
  double omt = 1.0 - t;
  double h00 = (1.0 + 2.0 * t) * omt * omt, h10 = t * omt * omt;
  double h01 = t * t * (3.0 - 2.0 * t), h11 = t * t * (t - 1.0);
  return h00 * y[k] + h10 * ds * dy[k] + h01 * y[k + 1] + h11 * ds * dy[k + 1];
}

// DeltaConstant: all mass is shared between the two bracketing grid metallicities.
// Returns the bracket j in [0, n_z-2] and the weight f of metallicity j+1 (1-f goes to j); values
// outside the grid clamp to the end bin.  Used by the grouping keys and by prep_kernel (same result).
__device__ __forceinline__ int delta_bracket(const double* __restrict__ zx, int n_z, double zv, double* f_out) {
  int j = 0;  // largest j with zx[j] <= zv
  for (int i = 1; i < n_z; ++i) j = (zx[i] <= zv) ? i : j;
  double f;
  if (zv <= zx[0]) { j = 0; f = 0.0; }
  else if (zv >= zx[n_z - 1]) { j = n_z - 2; f = 1.0; }
  else f = (zv - zx[j]) / (zx[j + 1] - zx[j]);
  *f_out = f;
  return j;
}

// Phi(uh) - Phi(ul) evaluated in whichever tail avoids cancellation.
__device__ __forceinline__ double phi_diff(double ul, double uh) {
  const double r = 0.70710678118654752440;
  if (ul + uh > 0.0) return 0.5 * (erfc(ul * r) - erfc(uh * r));
  return 0.5 * (erfc(-uh * r) - erfc(-ul * r));
}

// Mass formed between lookback ages [lo, hi] (already clipped to [min_age, max_age]).
__device__ double sfh_bin_mass(int type, const double* p, double mn, double mx, double lo, double hi,
                               double e_lo, double e_hi) {
  if (type == SB2_SFH_CONTINUITY) {
    const int nb = (int)p[0];
    const double* edges = p + 1;
    const double* ratios = p + 1 + nb + 1;
    double sfr = 1.0, m = 0.0;
    for (int j = 0; j < nb; ++j) {
      if (j > 0) sfr *= pow(10.0, -ratios[j - 1]);
      double ov = fmin(e_hi, edges[j + 1]) - fmax(e_lo, edges[j]);
      if (ov > 0.0) m += sfr * ov;
    }
    return m;
  }
  if (!(hi > lo)) return 0.0;
  switch (type) {
    case SB2_SFH_CONSTANT:
      return hi - lo;
    case SB2_SFH_GAUSSIAN: {
      double pk = p[0], s = p[1];
      return s * 2.50662827463100050242 * phi_diff((lo - pk) / s, (hi - pk) / s);
    }
    case SB2_SFH_EXPONENTIAL:
    case SB2_SFH_DECLINING_EXP: {
      double tau = (type == SB2_SFH_EXPONENTIAL) ? p[0] : -p[0];
      double shift = tau > 0.0 ? (mx - mn) / tau : 0.0;
      return tau * (exp((mx - lo) / tau - shift) - exp((mx - hi) / tau - shift));
    }
    case SB2_SFH_DELAYED_EXP: {
      double tau = p[0];
      double t1 = mx - lo, t2 = mx - hi;
      return -tau * (t1 + tau) * exp(-t1 / tau) + tau * (t2 + tau) * exp(-t2 / tau);
    }
    case SB2_SFH_LOGNORMAL: {
      double tau = p[0], pk = p[1];
      double t0 = log(mx - pk) + tau * tau;
      double u_hi = (log(fmax(mx - hi, 1e-300)) - t0) / tau;
      double u_lo = (log(fmax(mx - lo, 1e-300)) - t0) / tau;
      return tau * 2.50662827463100050242 * phi_diff(u_hi, u_lo);
    }
    default:
      return nan("");
  }
}

constexpr int kPrepWarps = 8;

__global__ void __launch_bounds__(kPrepWarps * 32)
prep_kernel(PrepModel M, PrepParams P, PrepOut O, const int* __restrict__ perm, long long n_pad) {
  extern __shared__ double prep_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * kPrepWarps + warp;
  if (t >= n_pad) return;
  double* sf = prep_smem + (size_t)warp * (M.n_age + M.n_z + SB2_SFH_ROW);
  double* zd = sf + M.n_age;
  double* prow = zd + M.n_z;
  const unsigned FULL = 0xffffffffu;

  const long long g = perm ? (long long)perm[t] : (t < P.n ? t : -1);
  if (g < 0) {  // padding row (end of a tile / of a metallicity group)
    for (int k = lane; k < M.w_stride; k += 32) {
      O.w_hi[t * M.w_stride + k] = 0.f;
      O.w_lo[t * M.w_stride + k] = 0.f;
    }
    if (M.igm_on && lane < 13) O.zpow[(size_t)lane * n_pad + t] = (lane < 12) ? 1.0 : 0.0;
    if (lane == 0) {
      O.g_m[t] = 0; O.g_beta[t] = 0.5f; O.g_gamma[t] = 0.5f; O.g_taut[t] = 0.f; O.g_scale[t] = 0.f; O.g_ca[t] = 0.f;
      O.g_cb[t] = 0.f; O.g_orig[t] = -1; O.g_mscale[t] = 0.0; O.g_trunc[t] = 0u;
    }
    return;
  }

  const double z = P.redshift[g];
  const double zp = 1.0 + z;
  const double s = log1p(z);

  // ---- SFH parameters ------------------------------------------------------------------
  if (lane < SB2_SFH_ROW) prow[lane] = (lane < P.sfh_stride) ? P.sfh_rows[g * P.sfh_stride + lane] : 0.0;
  __syncwarp();
  double mn = prow[0], mx = prow[1];
  if (P.max_age_from_z) {
    double age_gyr = hermite_lut(M.age, M.dage, M.cosmo_ds, M.cosmo_n, s);
    mx = (age_gyr - P.age_zmax_gyr) * 1.0e9;
    if (lane < SB2_SFH_ROW - 2 && ((P.norm_mask >> lane) & 1u)) prow[2 + lane] *= mx;
    __syncwarp();
  }
  const double* p = prow + 2;

  double part = 0.0;
  for (int a = lane; a < M.n_age; a += 32) {
    double m = 0.0;
    if (a < M.n_age - 1) {  // the last age bin receives no mass (A2)
      double e_lo = M.edges[a], e_hi = M.edges[a + 1];
      double lo = fmin(fmax(e_lo, mn), mx), hi = fmin(fmax(e_hi, mn), mx);
      m = sfh_bin_mass(P.sfh_type, p, mn, mx, lo, hi, e_lo, e_hi);
    }
    sf[a] = m;
    part += m;
  }
  for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
  const double inv_sf = 1.0 / part;

  // ---- metallicity weights ---------------------------------------------------------------
  const double zv = P.zd_value[g];
  const double zs = P.zd_sigma ? P.zd_sigma[g] : 0.0;
  const bool logz = (P.zd_type == SB2_ZD_DELTA_LOG10 || P.zd_type == SB2_ZD_NORMAL_LOG10);
  const double* zx = logz ? M.log10zmet : M.zmet;
  double zpart = 0.0;
  int zj = 0;
  double zf = 0.0;
  if (P.zd_type == SB2_ZD_DELTA_LINEAR || P.zd_type == SB2_ZD_DELTA_LOG10) {
    if (M.n_z >= 2) zj = delta_bracket(zx, M.n_z, zv, &zf);
    for (int i = lane; i < M.n_z; i += 32) zd[i] = (i == zj) ? 1.0 - zf : ((i == zj + 1) ? zf : 0.0);
    zpart = 1.0;
  } else {
    for (int i = lane; i < M.n_z; i += 32) {
      double u = (zx[i] - zv) / zs;
      double w = exp(-0.5 * u * u);
      zd[i] = w;
      zpart += w;
    }
    for (int o = 16; o; o >>= 1) zpart += __shfl_xor_sync(FULL, zpart, o);
  }
  __syncwarp();
  const double inv = inv_sf / zpart;

  // ---- weights row (TF32 hi/lo split) ------------------------------------------------------
  if (M.delta) {  // columns [0, na_pad): metallicity zj, [na_pad, 2 na_pad): zj+1 (grid columns zj*na_pad + k)
    for (int k = lane; k < M.w_stride; k += 32) {
      double w = 0.0;
      if (k < M.n_age) w = sf[k] * ((1.0 - zf) * inv);
      else if (k >= M.na_pad && k - M.na_pad < M.n_age) w = sf[k - M.na_pad] * (zf * inv);
      const float hi = to_tf32_rna((float)w);
      O.w_hi[t * M.w_stride + k] = hi;
      O.w_lo[t * M.w_stride + k] = to_tf32_rna((float)(w - (double)hi));
    }
  } else {
  for (int iz = 0; iz < M.n_z; ++iz) {  // column k = iz*na_pad + ia
    const double zw = zd[iz] * inv;
    for (int a = lane; a < M.na_pad; a += 32) {
      const int k = iz * M.na_pad + a;
      const double w = (a < M.n_age) ? sf[a] * zw : 0.0;
      const float hi = to_tf32_rna((float)w);
      O.w_hi[t * M.k_pad + k] = hi;
      O.w_lo[t * M.k_pad + k] = to_tf32_rna((float)(w - (double)hi));
      if (O.w_f64 && a < M.n_age) O.w_f64[g * M.K + iz * M.n_age + a] = w;
    }
  }
  for (int k = M.n_z * M.na_pad + lane; k < M.k_pad; k += 32) {
    O.w_hi[t * M.k_pad + k] = 0.f;
    O.w_lo[t * M.k_pad + k] = 0.f;
  }
  }

  // ---- per-galaxy scalars -------------------------------------------------------------------
  const double tq = s / M.ln_q;
  const int m = (int)floor(tq);
  const double r = exp(s - (double)m * M.ln_q);  // (1+z)/q^m in [1, q)
  const double beta = (M.variant == 0) ? (1.0 - 1.0 / r) / (1.0 - 1.0 / M.q) : (r - 1.0) / (M.q - 1.0);
  unsigned trunc = 0u;
  if (lane < M.n_filt) {
    int i_first = M.filt_lo[lane] - 1 - m, i_last = M.filt_hi[lane] - m;
    trunc = (i_first < 0 || i_last > M.n_lam - 1) ? 1u : 0u;
  }
  trunc = __ballot_sync(FULL, trunc != 0u);
  if (lane == 0) {
    const double dc = hermite_lut(M.dc, M.ddc, M.cosmo_ds, M.cosmo_n, s);
    const double dl_cm = zp * dc * 3.0856775814913673e24;
    const double scale = M.grid_scale * M.base_mass * zp / (4.0 * 3.14159265358979323846 * dl_cm * dl_cm) * 1.0e32;
    O.g_m[t] = m;
    O.g_beta[t] = (float)beta;
    O.g_gamma[t] = (float)((M.variant == 0) ? (1.0 / r - 1.0 / M.q) / (1.0 - 1.0 / M.q) : (M.q - r) / (M.q - 1.0));
    O.g_taut[t] = (float)((P.tau_v ? P.tau_v[g] : 0.0) * 1.44269504088896340736);
    O.g_scale[t] = (float)scale;
    O.g_ca[t] = (float)(P.coef_att ? P.coef_att[g] : 1.0);
    O.g_cb[t] = (float)(P.coef_unatt ? P.coef_unatt[g] : 1.0);
    O.g_orig[t] = (int)g;
    O.g_mscale[t] = P.log_mass ? pow(10.0, P.log_mass[g]) / M.base_mass : 1.0;
    O.g_trunc[t] = trunc;
  }

  // ---- (1+z)^p for the IGM kernel: p in (1.2, 2.1, 3.7, 5.5, -0.3, 2, 3, -0.9, 1.6, 3.4, 2.3, 3.3), then z
  if (M.igm_on) {
    const double zpw_exp[12] = {1.2, 2.1, 3.7, 5.5, -0.3, 2.0, 3.0, -0.9, 1.6, 3.4, 2.3, 3.3};
    if (lane < 12) O.zpow[(size_t)lane * n_pad + t] = pow(zp, zpw_exp[lane]);
    if (lane == 12) O.zpow[(size_t)12 * n_pad + t] = z;
  }
}

// Inoue+14 transmission exp(-tau(z, lam_i (1+z))) for the bins blueward of Ly-alpha.
// Block = one 128-galaxy tile x one strip of 64 wavelength bins; thread = galaxy, so the per-bin tables are
// warp-uniform loads, the (redshift-sorted) galaxies of a warp take the same branches, and the store of
// 128 consecutive floats per bin is coalesced in the tile-blocked layout the contraction epilogue reads.
// No pow(): every term is coef * (lam_i/911.8)^p * (1+z)^p with host-tabulated bin powers and the
// per-galaxy z powers from prep_kernel; "lines in regime k" are prefixes of the wavelength-sorted
// line list, tracked by pointers that only move down as the bin index grows.
constexpr int kIgmStrip = 64;

__global__ void __launch_bounds__(128) igm_kernel(PrepModel M, const double* __restrict__ zpow, float* __restrict__ igm,
                                                  int nb_pad, long long n_pad) {
  __shared__ double s_thr[3 * 64];
  __shared__ double s_pre[5 * 64];
  const int np1 = M.n_lines + 1;
  for (int i = threadIdx.x; i < 3 * 64; i += 128) s_thr[i] = M.thr[i];
  for (int i = threadIdx.x; i < 5 * np1; i += 128) s_pre[(i / np1) * 64 + (i % np1)] = M.pre[i];
  __syncthreads();
  const int nb = M.n_blue;
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
  double Z[12];
#pragma unroll
  for (int p = 0; p < 12; ++p) Z[p] = zpow[(size_t)p * n_pad + t];
  const double z = zpow[(size_t)12 * n_pad + t];
  const double zp = 1.0 + z;
  const int i0 = blockIdx.y * kIgmStrip;
  const int i1 = min(nb, i0 + kIgmStrip);
  float* out = igm + ((size_t)blockIdx.x * nb_pad) * 128 + threadIdx.x;
  for (int i = max(i0, nb); i < min(nb_pad, i0 + kIgmStrip); ++i) out[(size_t)i * 128] = 1.f;  // padding rows
  if (i0 >= nb) return;
  // lines (sorted by decreasing wavelength) still below the regime thresholds at the strip's first bin
  int n1 = 0, n2 = 0, nd = 0;
  {
    const double xl = __ldg(M.bin_pow + 7 * nb + i0) * zp;
#pragma unroll
    for (int step = 32; step; step >>= 1) {
      if (s_thr[0 * 64 + n1 + step - 1] > xl) n1 += step;
      if (s_thr[1 * 64 + n2 + step - 1] > xl) n2 += step;
      if (s_thr[2 * 64 + nd + step - 1] > xl) nd += step;
    }
  }
  for (int i = i0; i < i1; ++i) {
    const double xl = __ldg(M.bin_pow + 7 * nb + i) * zp;  // lam_obs / 911.8
    while (n1 > 0 && !(s_thr[0 * 64 + n1 - 1] > xl)) --n1;
    while (n2 > 0 && !(s_thr[1 * 64 + n2 - 1] > xl)) --n2;
    while (nd > 0 && !(s_thr[2 * 64 + nd - 1] > xl)) --nd;
    const double b12 = __ldg(M.bin_pow + 0 * nb + i) * Z[0], b37 = __ldg(M.bin_pow + 2 * nb + i) * Z[2];
    const double b55 = __ldg(M.bin_pow + 3 * nb + i) * Z[3];
    const double b2 = xl * xl, b3 = b2 * xl;
    const int J = __ldg(M.nline + i);
    const int a = min(J, n1), b = min(J, n2), c = min(J, nd);
    double tau = b12 * s_pre[0 * 64 + a] + b37 * (s_pre[1 * 64 + b] - s_pre[1 * 64 + a]) +
                 b55 * (s_pre[2 * 64 + J] - s_pre[2 * 64 + b]) + b2 * s_pre[3 * 64 + c] +
                 b3 * (s_pre[4 * 64 + J] - s_pre[4 * 64 + c]);
    if (__ldg(M.lc_on + i)) {
      const double b21 = __ldg(M.bin_pow + 1 * nb + i) * Z[1], bm3 = __ldg(M.bin_pow + 4 * nb + i) * Z[4];
      // Lyman continuum, DLA component
      if (z < 2.0) tau += 0.2113 * Z[5] - 0.07661 * Z[10] * bm3 - 0.1347 * b2;
      else if (xl >= 3.0) tau += 0.04696 * Z[6] - 0.01779 * Z[11] * bm3 - 0.02916 * b3;
      else tau += 0.6340 + 0.04696 * Z[6] - 0.01779 * Z[11] * bm3 - 0.1347 * b2 - 0.2905 * bm3;
      // Lyman continuum, LAF component
      if (z < 1.2) tau += 0.3248 * (b12 - Z[7] * b21);
      else if (z < 4.7) {
        if (xl >= 2.2) tau += 2.545e-2 * (Z[8] * b21 - b37);
        else tau += 2.545e-2 * Z[8] * b21 + 0.3248 * b12 - 0.2496 * b21;
      } else {
        if (xl > 5.7) tau += 5.221e-4 * (Z[9] * b21 - b55);
        else if (xl >= 2.2 && xl < 5.7) tau += 5.221e-4 * Z[9] * b21 + 0.2182 * b21 - 2.545e-2 * b37;
        else if (xl < 2.2) tau += 5.221e-4 * Z[9] * b21 + 0.3248 * b12 - 3.140e-2 * b21;
      }
    }
    out[(size_t)i * 128] = (float)exp(-tau);
  }
}

}  // namespace sb2
