// Weight construction ("prep") kernels, float64: scalars_kernel (one thread per galaxy), weights_kernel
// (one thread per (galaxy, age-bin edge)), igm_kernel (one thread per galaxy, strips of wavelength bins).
//
// For the galaxy in grouped slot t they produce everything the contraction kernel needs:
//   * SFZH weights  w[iz*n_age + ia] = sf[ia] * zd[iz]  (Stars.__init__ in the
//     reference, library.py:1372-1379), written as a TF32 hi/lo pair (3xTF32 operand A);
//   * the Inoue+14 transmission row exp(-tau(z, lam_i (1+z))) for the bins blueward of Ly-alpha;
//   * per-galaxy scalars: integer/fractional filter shift (m, beta) on the geometric axis,
//     dust exponent scale, flux scale (1+z)/(4 pi d_L^2) from the cosmology table, mass scale.
// HBM-bound by design: algorithmic bytes per galaxy = 2 * 4 * k_pad (weights) + 4 * n_blue (IGM).
#pragma once
#include <climits>
#include "ptx.cuh"
#include "../../include/synference_b200.h"

namespace sb2 {

struct PrepModel {
  int n_age, na_pad, n_z, K, k_pad, n_lam, n_filt, n_blue, n_lines, variant, igm_on, rest_frame;
  int delta;     // 1: DeltaConstant batch grouped by metallicity bracket -> weights row holds only the
                 //    two bracketing grid metallicities: [sf*(1-f) (na_pad) | sf*f (na_pad)], stride w_stride
  int w_stride;  // floats per weights row (k_pad, or 2*na_pad in delta mode)
  int cross;     // 1: w_lo holds, per group of 8 columns, the 16 bfloat16 [w_lo(0..7) | w_hi(0..7)] -- the K = 16 operand of
                 //    the ONE bfloat16 MMA that forms both small terms of the split product (synth_kernel, SynthArgs.cross)
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
  const double* lya_line;  // [n_z][n_age] line-continuum value of the Lyman-alpha bin (optional)
};

// Table-driven float64 log / exp / normal tail for weights_kernel (tables: synference_b200/fastmath.py, which also
// restates these functions in numpy for the CPU accuracy test).  The CUDA math library's erfc / log / exp cost ~5x
// the instructions (branchy range reduction, constants assembled from 32-bit immediates) and the kernel is
// instruction-issue bound.
struct FastMath {
  const double2* log_tab;   // [256] (ln m0, 1/m0)
  const double* exp_tab;    // [64]  2^(j/64)
  const double2* tail_tab;  // [tail_n][4] = 8 polynomial coefficients per interval
  double tail_w, tail_inv_w;
  int tail_n;
};

__device__ __forceinline__ double fm_log(const FastMath& F, double x) {   // x positive and normal
  const long long b = __double_as_longlong(x);
  const int e = (int)(b >> 52) - 1023;
  const double m = __longlong_as_double((b & 0x000FFFFFFFFFFFFFLL) | 0x3FF0000000000000LL);
  const double2 t = __ldg(F.log_tab + ((int)(b >> 44) & 0xFF));
  const double r = fma(m, t.y, -1.0);                    // |r| <= 2^-9
  double p = 0.2;
  p = fma(p, r, -0.25);
  p = fma(p, r, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  p = fma(p, r, 1.0);
  return fma((double)e, 0.6931471805599453, fma(r, p, t.x));
}

__device__ __forceinline__ double fm_exp(const FastMath& F, double y) {   // y <= ~0; underflow flushes to 0
  const double kf = rint(y * 92.33248261689366);         // 64 / ln 2
  const double r = fma(-kf, 2.531013593154441e-13, fma(-kf, 0.010830424695996044, y));   // Cody-Waite, ln2/64 = hi + lo
  const int k = (int)kf;
  const int e = k >> 6;
  double p = 1.0 / 120.0;
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double v = __ldg(F.exp_tab + (k & 63)) * p;      // in [1, 2]
  if (e < -1021) return 0.0;
  return __longlong_as_double(__double_as_longlong(v) + ((long long)e << 52));
}

// Q(u) = 1 - Phi(u), u >= 0:  exp(-u^2/2) * g(u), g from a local degree-7 polynomial
__device__ __forceinline__ double fm_tail(const FastMath& F, double u) {
  if (!(u < F.tail_w * (double)F.tail_n)) return 0.0;
  const int j = min((int)(u * F.tail_inv_w), F.tail_n - 1);
  const double s = u - ((double)j + 0.5) * F.tail_w;
  const double2* c = F.tail_tab + 4 * j;
  const double2 c01 = __ldg(c), c23 = __ldg(c + 1), c45 = __ldg(c + 2), c67 = __ldg(c + 3);
  double g = c67.y;
  g = fma(g, s, c67.x);
  g = fma(g, s, c45.y);
  g = fma(g, s, c45.x);
  g = fma(g, s, c23.y);
  g = fma(g, s, c23.x);
  g = fma(g, s, c01.y);
  g = fma(g, s, c01.x);
  return fm_exp(F, -0.5 * u * u) * g;
}

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
  const double* dust_slope;
  const double* dust_ampl;
  const double* fesc_lya;
  const double* tau_v_birth;
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
  float* g_slope;   // [n_pad] per-galaxy dust slope / bump amplitude (nullptr: global curve)
  float* g_ampl;    // [n_pad]
  float* g_taub;    // [n_pad] birth-cloud tau_V * log2(e) (nullptr: single screen)
  float* g_lya;     // [n_pad] fesc_lya_g * sum_k w_k lya_line[k] (nullptr: global Lyman-alpha escape fraction)
  int* g_orig;      // [n_pad]  original index, -1 for padding rows
  double* g_mscale; // [n_pad]
  unsigned* g_trunc;// [n_pad]  bit f set: filter f not fully covered by the grid at this z
};

// One column of a weights row: TF32 hi part, and the small part either as TF32 (three-pass product) or packed for the
// bfloat16 MMA (PrepModel.cross).  `base` = row offset in floats (a multiple of 8), `w` the float64 weight.
__device__ __forceinline__ void store_weight(const PrepOut& O, int cross, size_t base, int k, double w) {
  const float hi = to_tf32_rna((float)w);
  const float lo = (float)(w - (double)hi);
  O.w_hi[base + k] = hi;
  if (cross) {
    unsigned short* x = reinterpret_cast<unsigned short*>(O.w_lo + base + (k & ~7));
    unsigned short bl, bh;
    asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(bl) : "f"(lo));
    asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(bh) : "f"(hi));
    x[k & 7] = bl;
    x[8 + (k & 7)] = bh;
  } else {
    O.w_lo[base + k] = to_tf32_rna(lo);
  }
}

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

// ---- SFH bin masses from per-edge values ---------------------------------------------------------
// Every family's bin mass is F(e_{a+1}) - F(e_a) for an antiderivative F evaluated at the bin edges
// clipped to [min_age, max_age] (closed forms instead of scipy.quad), so each of the n_age
// edges is evaluated ONCE by its own thread and neighbouring threads difference them.  The normal-CDF
// families keep the small tail c = erfc(|u|/sqrt2)/2 and u itself, so the difference is formed in
// whichever tail avoids cancellation.
struct EdgeVal { double a, b; };

template <bool kFast>
__device__ __forceinline__ EdgeVal sfh_edge(const FastMath& F, int type, const double* __restrict__ p,
                                            const double* __restrict__ gc, double mn, double mx, double e_raw) {
  const double r = 0.70710678118654752440;
  const double t = fmin(fmax(e_raw, mn), mx);
  EdgeVal v{0.0, 0.0};
  switch (type) {
    case SB2_SFH_CONSTANT:
    case SB2_SFH_DOUBLE_POWERLAW:   // no antiderivative: the clipped edge itself, the bin is integrated numerically
      v.a = t;
      break;
    case SB2_SFH_GAUSSIAN: {
      const double u = (t - p[0]) / p[1];
      v.a = kFast ? fm_tail(F, fabs(u)) : 0.5 * erfc(fabs(u) * r);
      v.b = u;
      break;
    }
    case SB2_SFH_EXPONENTIAL:
    case SB2_SFH_DECLINING_EXP: {
      const double tau = (type == SB2_SFH_EXPONENTIAL) ? p[0] : -p[0];
      const double shift = tau > 0.0 ? (mx - mn) / tau : 0.0;
      const double arg = (mx - t) / tau - shift;
      v.a = -tau * ((kFast && arg <= 0.0) ? fm_exp(F, arg) : exp(arg));
      break;
    }
    case SB2_SFH_DELAYED_EXP: {
      const double tau = p[0], T = mx - t;
      const double arg = -T / tau;
      v.a = tau * (T + tau) * ((kFast && arg <= 0.0) ? fm_exp(F, arg) : exp(arg));
      break;
    }
    case SB2_SFH_LOGNORMAL: {  // gc[0] = ln(max_age - peak_age) + tau^2
      const double x = fmax(mx - t, 1e-300);
      const double u = ((kFast ? fm_log(F, x) : log(x)) - gc[0]) / p[0];
      v.a = kFast ? fm_tail(F, fabs(u)) : 0.5 * erfc(fabs(u) * r);
      v.b = u;
      break;
    }
    case SB2_SFH_CONTINUITY: {  // gc[j] = SFR of bin j; piecewise-constant SFR, edges not clipped
      const int nb = (int)p[0];
      const double* edges = p + 1;
      double F = 0.0;
      for (int j = 0; j < nb; ++j) {
        const double ov = fmin(e_raw, edges[j + 1]) - edges[j];
        if (ov > 0.0) F += gc[j] * ov;
      }
      v.a = F;
      break;
    }
    default:
      v.a = nan("");
  }
  return v;
}

// Phi(u_hi) - Phi(u_lo) from the tails c = 1 - Phi(|u|)
__device__ __forceinline__ double phi_between(EdgeVal lo, EdgeVal hi) {
  const bool pl = lo.b > 0.0, ph = hi.b > 0.0;
  if (pl && ph) return lo.a - hi.a;
  if (!pl && !ph) return hi.a - lo.a;
  if (!pl && ph) return 1.0 - hi.a - lo.a;
  return -(1.0 - lo.a - hi.a);
}

// DoublePowerLaw sfr(t) = 1 / ((t/peak)^alpha + (t/peak)^beta) over [lo, hi]: 16-point Gauss-Legendre (the reference
// integrates every bin with scipy.quad; a 0.1-dex age bin of this smooth integrand converges to ~1e-12)
__device__ __noinline__ double dpl_bin_mass(const double* __restrict__ p, double lo, double hi) {
  if (!(hi > lo)) return 0.0;
  const double x[8] = {0.0950125098376374, 0.2816035507792589, 0.4580167776572274, 0.6178762444026438,
                       0.7554044083550030, 0.8656312023878318, 0.9445750230732326, 0.9894009349916499};
  const double w[8] = {0.1894506104550685, 0.1826034150449236, 0.1691565193950025, 0.1495959888165767,
                       0.1246289712555339, 0.0951585116824928, 0.0622535239386479, 0.0271524594117541};
  const double mid = 0.5 * (lo + hi), half = 0.5 * (hi - lo), inv_pk = 1.0 / p[0];
  double acc = 0.0;
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int sgn = -1; sgn <= 1; sgn += 2) {
      const double l = log((mid + sgn * half * x[i]) * inv_pk);
      acc += w[i] / (exp(p[1] * l) + exp(p[2] * l));
    }
  }
  return acc * half;
}

// kDpl: the DoublePowerLaw quadrature is an out-of-line call; kernels for the other families are compiled without it (its
// mere presence cost the half-warp weight builder 12 %)
template <bool kDpl>
__device__ __forceinline__ double sfh_mass_from_edges(int type, const double* __restrict__ p, EdgeVal lo, EdgeVal hi) {
  const double s2pi = 2.50662827463100050242;
  if constexpr (kDpl) return dpl_bin_mass(p, lo.a, hi.a);
  switch (type) {
    case SB2_SFH_GAUSSIAN:
      return p[1] * s2pi * phi_between(lo, hi);
    case SB2_SFH_LOGNORMAL:
      return -p[0] * s2pi * phi_between(lo, hi);  // u decreases with lookback age
    default:
      return hi.a - lo.a;
  }
}

// ---- per-galaxy scalars: one THREAD per (padded, grouped) row ------------------------------------
__global__ void __launch_bounds__(256)
scalars_kernel(PrepModel M, PrepParams P, PrepOut O, const int* __restrict__ perm, long long n_pad) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pad) return;
  const long long g = perm ? (long long)perm[t] : (t < P.n ? t : -1);
  if (g < 0) {  // padding row (end of a tile / of a metallicity group)
    if (M.igm_on) {
      for (int p = 0; p < 12; ++p) O.zpow[(size_t)p * n_pad + t] = 1.0;
      O.zpow[(size_t)12 * n_pad + t] = 0.0;
    }
    O.g_m[t] = 0; O.g_beta[t] = 0.5f; O.g_gamma[t] = 0.5f; O.g_taut[t] = 0.f; O.g_scale[t] = 0.f; O.g_ca[t] = 0.f;
    O.g_cb[t] = 0.f; O.g_orig[t] = -1; O.g_mscale[t] = 0.0; O.g_trunc[t] = 0u;
    if (O.g_slope) { O.g_slope[t] = 0.f; O.g_ampl[t] = 0.f; }
    if (O.g_taub) O.g_taub[t] = 0.f;
    return;
  }
  double z = P.redshift[g];
  // a redshift that is not a finite number >= 0 cannot be placed on the wavelength axis: the galaxy gets NaN fluxes
  // (every filter flagged) and a harmless shift, instead of indexing the filter tables with garbage
  const bool z_ok = (z >= 0.0) && (z <= 1.0e6);
  if (!z_ok || M.rest_frame) z = 0.0;     // rest_frame: luminosities through the filters at their own wavelengths
  const double zp = 1.0 + z;
  const double s = log1p(z);
  const double tq = s / M.ln_q;
  const int m = (int)floor(tq);
  const double r = exp(s - (double)m * M.ln_q);  // (1+z)/q^m in [1, q)
  const double beta = (M.variant == 0) ? (1.0 - 1.0 / r) / (1.0 - 1.0 / M.q) : (r - 1.0) / (M.q - 1.0);
  unsigned trunc = z_ok ? 0u : 0xffffffffu;
  for (int f = 0; f < M.n_filt; ++f) {
    const int i_first = __ldg(M.filt_lo + f) - 1 - m, i_last = __ldg(M.filt_hi + f) - m;
    if (i_first < 0 || i_last > M.n_lam - 1) trunc |= 1u << f;
  }
  const double dc = hermite_lut(M.dc, M.ddc, M.cosmo_ds, M.cosmo_n, s);
  const double dl_cm = zp * dc * 3.0856775814913673e24;
  const double scale = M.rest_frame ? M.grid_scale * M.base_mass
                                    : M.grid_scale * M.base_mass * zp / (4.0 * 3.14159265358979323846 * dl_cm * dl_cm) * 1.0e32;
  O.g_m[t] = m;
  O.g_beta[t] = (float)beta;
  O.g_gamma[t] = (float)((M.variant == 0) ? (1.0 / r - 1.0 / M.q) / (1.0 - 1.0 / M.q) : (M.q - r) / (M.q - 1.0));
  O.g_taut[t] = (float)((P.tau_v ? P.tau_v[g] : 0.0) * 1.44269504088896340736);
  O.g_scale[t] = (float)scale;
  O.g_ca[t] = (float)(P.coef_att ? P.coef_att[g] : 1.0);
  O.g_cb[t] = (float)(P.coef_unatt ? P.coef_unatt[g] : 1.0);
  if (O.g_taub) O.g_taub[t] = (float)((P.tau_v_birth ? P.tau_v_birth[g] : 0.0) * 1.44269504088896340736);
  if (O.g_slope) {
    O.g_slope[t] = (float)(P.dust_slope ? P.dust_slope[g] : 0.0);
    O.g_ampl[t] = (float)(P.dust_ampl ? P.dust_ampl[g] : 0.0);
  }
  O.g_orig[t] = (int)g;
  O.g_mscale[t] = P.log_mass ? pow(10.0, P.log_mass[g]) / M.base_mass : 1.0;
  O.g_trunc[t] = trunc;
  // (1+z)^p for the IGM kernel: p in (1.2, 2.1, 3.7, 5.5, -0.3, 2, 3, -0.9, 1.6, 3.4, 2.3, 3.3), then z
  if (M.igm_on) {
    const double zpw_exp[12] = {1.2, 2.1, 3.7, 5.5, -0.3, 2.0, 3.0, -0.9, 1.6, 3.4, 2.3, 3.3};
#pragma unroll
    for (int p = 0; p < 12; ++p) O.zpow[(size_t)p * n_pad + t] = exp(zpw_exp[p] * s);
    O.zpow[(size_t)12 * n_pad + t] = z;
  }
}

// ---- SFZH weights: block = kWGal galaxies x 64 slots; thread (galaxy, edge) -------------------------
constexpr int kWGal = 4;
constexpr int kWSlots = 64;
constexpr int kGConst = 24;  // per-galaxy constants (ln-normal t0, continuity bin SFRs)

__host__ __device__ inline size_t weights_smem_doubles(int n_age, int n_z) {
  return (size_t)kWGal * (3 * (size_t)n_age + n_z + SB2_SFH_ROW + kGConst + 8);
}

template <bool kFast, bool kDpl>
__global__ void __launch_bounds__(kWGal * kWSlots)
weights_kernel(PrepModel M, FastMath F, PrepParams P, PrepOut O, const int* __restrict__ perm, long long n_pad) {
  extern __shared__ double prep_smem[];
  const int gq = threadIdx.x / kWSlots, slot = threadIdx.x % kWSlots;
  const long long t = (long long)blockIdx.x * kWGal + gq;
  double* base = prep_smem + (size_t)gq * (3 * M.n_age + M.n_z + SB2_SFH_ROW + kGConst + 8);
  double* eA = base;                  // [n_age] edge values
  double* eB = eA + M.n_age;
  double* sf = eB + M.n_age;          // [n_age] bin masses
  double* zd = sf + M.n_age;          // [n_z]
  double* prow = zd + M.n_z;          // [SB2_SFH_ROW]
  double* gc = prow + SB2_SFH_ROW;    // [kGConst]
  double* red = gc + kGConst;         // [8] reductions / scalars
  const unsigned FULL = 0xffffffffu;
  const bool in_range = t < n_pad;
  const long long g = !in_range ? -1 : (perm ? (long long)perm[t] : (t < P.n ? t : -1));
  const bool valid = g >= 0;

  // ---- SFH parameter row (+ max_age from redshift, *_norm scaling: library.py:1206, :1287-1289)
  if (valid && slot < SB2_SFH_ROW) prow[slot] = (slot < P.sfh_stride) ? P.sfh_rows[g * P.sfh_stride + slot] : 0.0;
  __syncthreads();
  if (valid && P.max_age_from_z) {
    if (slot == 0) {
      const double s = log1p(P.redshift[g]);
      red[0] = (hermite_lut(M.age, M.dage, M.cosmo_ds, M.cosmo_n, s) - P.age_zmax_gyr) * 1.0e9;
    }
  }
  __syncthreads();
  if (valid && P.max_age_from_z) {
    const double mxz = red[0];
    __syncwarp();
    if (slot < SB2_SFH_ROW - 2 && ((P.norm_mask >> slot) & 1u)) prow[2 + slot] *= mxz;
    if (slot == 0) prow[1] = mxz;
  }
  __syncthreads();
  const double mn = prow[0], mx = prow[1];
  const double* p = prow + 2;
  // ---- per-galaxy constants
  if (valid) {
    if (P.sfh_type == SB2_SFH_LOGNORMAL) {
      if (slot == 0) gc[0] = log(mx - p[1]) + p[0] * p[0];
    } else if (P.sfh_type == SB2_SFH_CONTINUITY) {
      const int nb = (int)p[0];
      const double* ratios = p + 1 + nb + 1;
      if (slot > 0 && slot < nb && slot < kGConst) gc[slot] = pow(10.0, -ratios[slot - 1]);
    }
  }
  __syncthreads();
  if (valid && P.sfh_type == SB2_SFH_CONTINUITY && slot == 0) {
    const int nb = min((int)p[0], kGConst);
    double sfr = 1.0;
    gc[0] = 1.0;
    for (int j = 1; j < nb; ++j) { sfr *= gc[j]; gc[j] = sfr; }
  }
  __syncthreads();
  // ---- edge values, bin masses (the last age bin receives no mass, A2)
  if (valid)
    for (int e = slot; e < M.n_age; e += kWSlots) {
      const EdgeVal v = sfh_edge<kFast>(F, P.sfh_type, p, gc, mn, mx, M.edges[e]);
      eA[e] = v.a; eB[e] = v.b;
    }
  __syncthreads();
  double part = 0.0;
  if (valid)
    for (int a = slot; a < M.n_age; a += kWSlots) {
      double mass = 0.0;
      if (a < M.n_age - 1) mass = sfh_mass_from_edges<kDpl>(P.sfh_type, p, EdgeVal{eA[a], eB[a]}, EdgeVal{eA[a + 1], eB[a + 1]});
      sf[a] = mass;
      part += mass;
    }
  // ---- metallicity weights
  const bool logz = (P.zd_type == SB2_ZD_DELTA_LOG10 || P.zd_type == SB2_ZD_NORMAL_LOG10);
  const double* zx = logz ? M.log10zmet : M.zmet;
  const bool zdelta = (P.zd_type == SB2_ZD_DELTA_LINEAR || P.zd_type == SB2_ZD_DELTA_LOG10);
  double zpart = 0.0;
  if (valid) {
    const double zv = P.zd_value[g];
    if (zdelta) {
      if (slot == 0) {
        double zf = 0.0;
        int zj = 0;
        if (M.n_z >= 2) zj = delta_bracket(zx, M.n_z, zv, &zf);
        red[4] = (double)zj; red[5] = zf;
      }
    } else {
      const double zs = P.zd_sigma[g];
      for (int i = slot; i < M.n_z; i += kWSlots) {
        const double u = (zx[i] - zv) / zs;
        const double w = exp(-0.5 * u * u);
        zd[i] = w;
        zpart += w;
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    part += __shfl_xor_sync(FULL, part, o);
    zpart += __shfl_xor_sync(FULL, zpart, o);
  }
  if ((slot & 31) == 0) { red[slot >> 5] = part; red[2 + (slot >> 5)] = zpart; }
  __syncthreads();
  if (!valid) {
    if (in_range)
      for (int k = slot; k < M.w_stride; k += kWSlots) {
        O.w_hi[t * M.w_stride + k] = 0.f;
        O.w_lo[t * M.w_stride + k] = 0.f;
      }
    return;
  }
  const double inv_sf = 1.0 / (red[0] + red[1]);
  // ---- weights row (TF32 hi/lo split)
  if (M.delta) {  // columns [0, na_pad): metallicity zj, [na_pad, 2 na_pad): zj+1 (grid columns zj*na_pad + k)
    const double zf = red[5];
    for (int k = slot; k < M.w_stride; k += kWSlots) {
      double w = 0.0;
      if (k < M.n_age) w = sf[k] * ((1.0 - zf) * inv_sf);
      else if (k >= M.na_pad && k - M.na_pad < M.n_age) w = sf[k - M.na_pad] * (zf * inv_sf);
      store_weight(O, M.cross, (size_t)t * M.w_stride, k, w);
    }
    return;
  }
  if (zdelta) {
    const int zj = (int)red[4];
    const double zf = red[5];
    for (int i = slot; i < M.n_z; i += kWSlots) zd[i] = (i == zj) ? 1.0 - zf : ((i == zj + 1) ? zf : 0.0);
  }
  const double inv = inv_sf / (zdelta ? 1.0 : red[2] + red[3]);
  __syncthreads();
  for (int k = slot; k < M.k_pad; k += kWSlots) {  // column k = iz*na_pad + ia
    const int iz = k / M.na_pad, a = k - iz * M.na_pad;
    const double w = (iz < M.n_z && a < M.n_age) ? sf[a] * (zd[iz] * inv) : 0.0;
    store_weight(O, M.cross, (size_t)t * M.k_pad, k, w);
    if (O.w_f64 && iz < M.n_z && a < M.n_age) O.w_f64[g * M.K + iz * M.n_age + a] = w;
  }
}

// ---- SFZH weights, half-warp per galaxy (n_age <= 64): lane l of the half-warp owns the 4 consecutive bin edges
// 4l..4l+3, so the per-galaxy work (parameter loads, bracket search, normalisation) is paid once per 16 lanes, the
// four edge evaluations of a lane are independent (ILP), neighbouring edges are differenced in registers / by one
// shuffle, and there is no block-level synchronisation at all.  ~6x fewer instructions per galaxy than
// weights_kernel above (which remains as the n_age > 64 fallback).
constexpr int kW2Gal = 8;   // galaxies (half-warps) per block of 128 threads
constexpr int kW2Smem = 64 + 64 + SB2_SFH_ROW + kGConst;   // doubles per galaxy: sf | zd | row | constants

__device__ __forceinline__ double shfl_down16(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d, 16); }
__device__ __forceinline__ double shfl_xor16(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m, 16); }

template <bool kFast, bool kDpl, bool kLya>
__global__ void __launch_bounds__(kW2Gal * 16)
weights2_kernel(PrepModel M, FastMath F, PrepParams P, PrepOut O, const int* __restrict__ perm, long long n_pad) {
  __shared__ double sm[kW2Gal * kW2Smem];
  const int hw = threadIdx.x >> 4, hl = threadIdx.x & 15;
  const long long t = (long long)blockIdx.x * kW2Gal + hw;
  double* sf = sm + hw * kW2Smem;      // [64] bin masses
  double* zd = sf + 64;                // [64] metallicity weights (dense rows only)
  double* prow = zd + 64;              // [SB2_SFH_ROW]
  double* gc = prow + SB2_SFH_ROW;     // [kGConst]
  const bool in_range = t < n_pad;
  const long long g = !in_range ? -1 : (perm ? (long long)perm[t] : (t < P.n ? t : -1));
  const bool valid = g >= 0;
  const long long gs = valid ? g : 0;  // padding rows run the same code on galaxy 0 and store zeros

  // ---- SFH parameter row (+ max_age from redshift, *_norm scaling: library.py:1206, :1287-1289)
  for (int i = hl; i < SB2_SFH_ROW; i += 16) prow[i] = (i < P.sfh_stride) ? P.sfh_rows[gs * P.sfh_stride + i] : 0.0;
  __syncwarp();
  if (P.max_age_from_z) {
    const double s = log1p(P.redshift[gs]);
    const double mxz = (hermite_lut(M.age, M.dage, M.cosmo_ds, M.cosmo_n, s) - P.age_zmax_gyr) * 1.0e9;
    __syncwarp();
    for (int i = hl; i < SB2_SFH_ROW - 2; i += 16)
      if ((P.norm_mask >> i) & 1u) prow[2 + i] *= mxz;
    if (hl == 0) prow[1] = mxz;
    __syncwarp();
  }
  const double mn = prow[0], mx = prow[1];
  const double* p = prow + 2;
  if (P.sfh_type == SB2_SFH_LOGNORMAL) {
    const double x = mx - p[1];
    if (hl == 0) gc[0] = ((kFast && x > 2.3e-308 && x < 1.7e308) ? fm_log(F, x) : log(x)) + p[0] * p[0];
  } else if (P.sfh_type == SB2_SFH_CONTINUITY) {
    const int nb = min((int)p[0], kGConst);
    const double* ratios = p + 1 + (int)p[0] + 1;
    for (int j = 1 + hl; j < nb; j += 16) gc[j] = pow(10.0, -ratios[j - 1]);
    __syncwarp();
    if (hl == 0) {
      double sfr = 1.0;
      gc[0] = 1.0;
      for (int j = 1; j < nb; ++j) { sfr *= gc[j]; gc[j] = sfr; }
    }
  }
  __syncwarp();

  // ---- four edges per lane, bin masses (the last age bin receives no mass, A2)
  EdgeVal ev[5];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int e = 4 * hl + q;
    ev[q] = (e < M.n_age) ? sfh_edge<kFast>(F, P.sfh_type, p, gc, mn, mx, M.edges[e]) : EdgeVal{0.0, 0.0};
  }
  ev[4].a = shfl_down16(ev[0].a, 1);
  ev[4].b = shfl_down16(ev[0].b, 1);
  double part = 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int a = 4 * hl + q;
    const double mass = (a < M.n_age - 1) ? sfh_mass_from_edges<kDpl>(P.sfh_type, p, ev[q], ev[q + 1]) : 0.0;
    sf[a] = mass;
    part += mass;
  }
#pragma unroll
  for (int o = 8; o; o >>= 1) part += shfl_xor16(part, o);
  const double inv_sf = 1.0 / part;

  // ---- metallicity weights
  const bool logz = (P.zd_type == SB2_ZD_DELTA_LOG10 || P.zd_type == SB2_ZD_NORMAL_LOG10);
  const double* zx = logz ? M.log10zmet : M.zmet;
  const bool zdelta = (P.zd_type == SB2_ZD_DELTA_LINEAR || P.zd_type == SB2_ZD_DELTA_LOG10);
  const double zv = P.zd_value[gs];
  int zj = 0;
  double zf = 0.0, zinv = 1.0;
  if (zdelta) {
    if (M.n_z >= 2) zj = delta_bracket(zx, M.n_z, zv, &zf);
  } else {
    const double zs = P.zd_sigma[gs];
    double zpart = 0.0;
    for (int i = hl; i < M.n_z; i += 16) {
      const double u = (zx[i] - zv) / zs;
      const double w = kFast ? fm_exp(F, -0.5 * u * u) : exp(-0.5 * u * u);
      zd[i] = w;
      zpart += w;
    }
#pragma unroll
    for (int o = 8; o; o >>= 1) zpart += shfl_xor16(zpart, o);
    zinv = 1.0 / zpart;
  }
  __syncwarp();
  if constexpr (kLya) {
    // per-galaxy Lyman-alpha line: the weighted sum of the line-continuum value of that one bin (a dot product over the
    // (age, Z) cells), times this galaxy's escape fraction; the contraction kernel adds it at the Lyman-alpha bin
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int a = 4 * hl + q;
      if (a < M.n_age) {
        double lz;
        if (zdelta) {
          lz = (1.0 - zf) * M.lya_line[zj * M.n_age + a] + (M.n_z >= 2 ? zf * M.lya_line[(zj + 1) * M.n_age + a] : 0.0);
        } else {
          lz = 0.0;
          for (int iz = 0; iz < M.n_z; ++iz) lz += zd[iz] * zinv * M.lya_line[iz * M.n_age + a];
        }
        acc += sf[a] * lz;
      }
    }
#pragma unroll
    for (int o = 8; o; o >>= 1) acc += shfl_xor16(acc, o);
    if (hl == 0 && in_range) O.g_lya[t] = valid ? (float)(P.fesc_lya[gs] * acc * inv_sf) : 0.f;
  }
  if (!in_range) return;

  // ---- weights row (TF32 hi/lo split)
  if (M.delta) {  // columns [0, na_pad): metallicity zj, [na_pad, 2 na_pad): zj+1 (grid columns zj*na_pad + k)
    const double s0 = valid ? (1.0 - zf) * inv_sf : 0.0, s1 = valid ? zf * inv_sf : 0.0;
    for (int k = hl; k < M.w_stride; k += 16) {
      double w = 0.0;
      if (k < M.n_age) w = sf[k] * s0;
      else if (k >= M.na_pad && k - M.na_pad < M.n_age) w = sf[k - M.na_pad] * s1;
      store_weight(O, M.cross, (size_t)t * M.w_stride, k, w);
    }
    return;
  }
  if (zdelta) {
    for (int i = hl; i < M.n_z; i += 16) zd[i] = (i == zj) ? 1.0 - zf : ((i == zj + 1) ? zf : 0.0);
    __syncwarp();
  }
  const double inv = valid ? inv_sf * zinv : 0.0;
  for (int iz = 0; iz < M.n_z; ++iz) {  // column k = iz*na_pad + ia
    const double zw = zd[iz] * inv;
    for (int a = hl; a < M.na_pad; a += 16) {
      const int k = iz * M.na_pad + a;
      const double w = (a < M.n_age) ? sf[a] * zw : 0.0;
      store_weight(O, M.cross, (size_t)t * M.k_pad, k, w);
      if (O.w_f64 && valid && a < M.n_age) O.w_f64[g * M.K + iz * M.n_age + a] = w;
    }
  }
  for (int k = M.n_z * M.na_pad + hl; k < M.k_pad; k += 16) {
    O.w_hi[t * M.k_pad + k] = 0.f;
    O.w_lo[t * M.k_pad + k] = 0.f;
  }
}

// ---- SFH bin masses, ONE GALAXY PER THREAD (bracket-grouped batches; feeds synth3_kernel) ---------------------------
// The half-warp builder above spends most of its ~675 warp instructions per galaxy on per-galaxy set-up replicated in 16
// lanes, shuffles and the strided hi/lo row write.  Here a thread walks its galaxy's age-bin edges in order (each edge
// evaluated once, differenced with the previous one in registers) and writes the RAW float64 bin masses tile-blocked,
// sf[(tile*n_age + a)*128 + t], so a warp's stores are contiguous; the two metallicity factors of a DeltaConstant galaxy
// travel as s0 = (1-f)/sum(sf), s1 = f/sum(sf).  The contraction kernel expands w[k] = sf[a]*s into the TF32 hi/lo pair
// on chip (408 B per galaxy through HBM instead of 896 B, and the same arithmetic as weights2_kernel: sf[a] * s).
struct SfOut {
  double* sf;   // [n_tiles][n_age][128] raw bin masses (rows of padding galaxies are 0)
  double* s0;   // [n_pad]
  double* s1;   // [n_pad]
  float* g_lya; // [n_pad] fesc_lya_g * sum_k w_k lya_line[k] (kLya instantiations; nullptr otherwise)
};

// families whose edge value needs at most two parameters (everything except DoublePowerLaw and Continuity)
template <bool kFast>
__device__ __forceinline__ EdgeVal sfh_edge_s(const FastMath& F, int type, double p0, double p1, double gc0, double mn, double mx,
                                              double e_raw) {
  const double r = 0.70710678118654752440;
  const double t = fmin(fmax(e_raw, mn), mx);
  EdgeVal v{0.0, 0.0};
  switch (type) {
    case SB2_SFH_CONSTANT:
      v.a = t;
      break;
    case SB2_SFH_GAUSSIAN: {
      const double u = (t - p0) / p1;
      v.a = kFast ? fm_tail(F, fabs(u)) : 0.5 * erfc(fabs(u) * r);
      v.b = u;
      break;
    }
    case SB2_SFH_EXPONENTIAL:
    case SB2_SFH_DECLINING_EXP: {
      const double tau = (type == SB2_SFH_EXPONENTIAL) ? p0 : -p0;
      const double shift = tau > 0.0 ? (mx - mn) / tau : 0.0;
      const double arg = (mx - t) / tau - shift;
      v.a = -tau * ((kFast && arg <= 0.0) ? fm_exp(F, arg) : exp(arg));
      break;
    }
    case SB2_SFH_DELAYED_EXP: {
      const double tau = p0, T = mx - t;
      const double arg = -T / tau;
      v.a = tau * (T + tau) * ((kFast && arg <= 0.0) ? fm_exp(F, arg) : exp(arg));
      break;
    }
    case SB2_SFH_LOGNORMAL: {
      const double x = fmax(mx - t, 1e-300);
      const double u = ((kFast ? fm_log(F, x) : log(x)) - gc0) / p0;
      v.a = kFast ? fm_tail(F, fabs(u)) : 0.5 * erfc(fabs(u) * r);
      v.b = u;
      break;
    }
    default:
      v.a = nan("");
  }
  return v;
}

__device__ __forceinline__ double sfh_mass_s(int type, double p0, double p1, EdgeVal lo, EdgeVal hi) {
  const double s2pi = 2.50662827463100050242;
  switch (type) {
    case SB2_SFH_GAUSSIAN: return p1 * s2pi * phi_between(lo, hi);
    case SB2_SFH_LOGNORMAL: return -p0 * s2pi * phi_between(lo, hi);
    default: return hi.a - lo.a;
  }
}

// kMode 0: two-parameter families (scalars in registers); 1: Continuity; 2: DoublePowerLaw (parameter row read from global)
// `put(a, mass)` receives the n_age bin masses in order; returns their sum.  s_edges: the grid's bin edges (shared memory).
template <bool kFast, int kMode, class Put>
__device__ __forceinline__ double sfh_masses_thread(const PrepModel& M, const FastMath& F, const PrepParams& P, long long g,
                                                    const double* __restrict__ s_edges, Put put) {
  const double* row = P.sfh_rows + g * P.sfh_stride;
  double mn = row[0], mx = row[1];
  double mxz = 1.0;
  if (P.max_age_from_z) {   // library.py:1206 / :1287-1289
    const double s = log1p(P.redshift[g]);
    mxz = (hermite_lut(M.age, M.dage, M.cosmo_ds, M.cosmo_n, s) - P.age_zmax_gyr) * 1.0e9;
    mx = mxz;
  }
  auto par = [&](int i) -> double {   // parameter i of the family, `_norm` ones scaled by max_age
    const double v = (2 + i < P.sfh_stride) ? row[2 + i] : 0.0;
    return (P.max_age_from_z && ((P.norm_mask >> i) & 1u)) ? v * mxz : v;
  };
  const int n_age = M.n_age;
  double part = 0.0;
  if constexpr (kMode == 0) {
    const int type = P.sfh_type;
    const double p0 = par(0), p1 = par(1);
    double gc0 = 0.0;
    if (type == SB2_SFH_LOGNORMAL) {
      const double x = mx - p1;
      gc0 = ((kFast && x > 2.3e-308 && x < 1.7e308) ? fm_log(F, x) : log(x)) + p0 * p0;
    }
    EdgeVal prev = sfh_edge_s<kFast>(F, type, p0, p1, gc0, mn, mx, s_edges[0]);
#pragma unroll 2
    for (int a = 0; a + 1 < n_age; ++a) {
      const EdgeVal cur = sfh_edge_s<kFast>(F, type, p0, p1, gc0, mn, mx, s_edges[a + 1]);
      const double mass = sfh_mass_s(type, p0, p1, prev, cur);
      put(a, mass);
      part += mass;
      prev = cur;
    }
  } else if constexpr (kMode == 1) {
    // piecewise-constant SFR over nb bins [edges_j, edges_j+1], SFR_j = prod_{i<=j} 10^(-ratio_i); F(e) = sum_j SFR_j * overlap
    const int nb = (int)par(0);
    double sfr[kGConst];
    {
      double cur = 1.0;
      sfr[0] = 1.0;
      for (int j = 1; j < nb && j < kGConst; ++j) { cur *= pow(10.0, -par(1 + nb + 1 + j - 1)); sfr[j] = cur; }
    }
    auto Fe = [&](double e) {
      double acc = 0.0;
      for (int j = 0; j < nb && j < kGConst; ++j) {
        const double ov = fmin(e, par(1 + j + 1)) - par(1 + j);
        if (ov > 0.0) acc += sfr[j] * ov;
      }
      return acc;
    };
    double prev = Fe(s_edges[0]);
    for (int a = 0; a + 1 < n_age; ++a) {
      const double cur = Fe(s_edges[a + 1]);
      const double mass = cur - prev;
      put(a, mass);
      part += mass;
      prev = cur;
    }
  } else {
    double p[3] = {par(0), par(1), par(2)};
    double prev = fmin(fmax(s_edges[0], mn), mx);
    for (int a = 0; a + 1 < n_age; ++a) {
      const double cur = fmin(fmax(s_edges[a + 1], mn), mx);
      const double mass = dpl_bin_mass(p, prev, cur);
      put(a, mass);
      part += mass;
      prev = cur;
    }
  }
  put(n_age - 1, 0.0);   // the last age bin receives no mass (A2)
  return part;
}

constexpr int kW3Threads = 128;   // one tile of the contraction kernel per block

template <bool kFast, int kMode, bool kLya = false>
__global__ void __launch_bounds__(kW3Threads)
weights3_kernel(PrepModel M, FastMath F, PrepParams P, SfOut O, double* __restrict__ w_f64, const int* __restrict__ perm,
                long long n_pad) {
  __shared__ double s_edges[64];
  if (threadIdx.x < M.n_age) s_edges[threadIdx.x] = M.edges[threadIdx.x];
  __syncthreads();
  const long long t = (long long)blockIdx.x * kW3Threads + threadIdx.x;
  if (t >= n_pad) return;
  const long long g = perm ? (long long)perm[t] : (t < P.n ? t : -1);
  double* sf = O.sf + (size_t)blockIdx.x * M.n_age * kW3Threads + threadIdx.x;
  if (g < 0) {
    for (int a = 0; a < M.n_age; ++a) sf[(size_t)a * kW3Threads] = 0.0;
    O.s0[t] = 0.0; O.s1[t] = 0.0;
    if constexpr (kLya) O.g_lya[t] = 0.f;
    return;
  }
  const bool logz = (P.zd_type == SB2_ZD_DELTA_LOG10);
  double zf = 0.0;
  int zj = 0;
  if (M.n_z >= 2) zj = delta_bracket(logz ? M.log10zmet : M.zmet, M.n_z, P.zd_value[g], &zf);
  // per-galaxy Lyman-alpha line: the weighted sum of the line-continuum value of that one bin over the (age, Z) cells
  double lya_acc = 0.0;
  const double part = sfh_masses_thread<kFast, kMode>(M, F, P, g, s_edges, [&](int a, double m) {
    sf[(size_t)a * kW3Threads] = m;
    if constexpr (kLya)
      lya_acc += m * ((1.0 - zf) * __ldg(M.lya_line + zj * M.n_age + a) + (M.n_z >= 2 ? zf * __ldg(M.lya_line + (zj + 1) * M.n_age + a) : 0.0));
  });
  const double inv_sf = 1.0 / part;
  if constexpr (kLya) O.g_lya[t] = (float)(P.fesc_lya[g] * lya_acc * inv_sf);
  const double s0 = (1.0 - zf) * inv_sf, s1 = zf * inv_sf;
  O.s0[t] = s0; O.s1[t] = s1;
  if (w_f64) {   // parity hook: the full row in the caller's order, k = iz*n_age + ia
    double* w = w_f64 + g * M.K;
    for (int k = 0; k < M.K; ++k) w[k] = 0.0;
    for (int a = 0; a < M.n_age; ++a) {
      const double m = sf[(size_t)a * kW3Threads];
      w[zj * M.n_age + a] = m * s0;
      if (M.n_z >= 2) w[(zj + 1) * M.n_age + a] = m * s1;
    }
  }
}

// First / last wavelength chunk (and bin) that any filter of a unit's galaxies can reach; one warp per unit
// (a tile of 128 rows, or a pair of tiles).
// Chunks outside the range are skipped by the contraction kernel, IGM bins below it are not evaluated.
__global__ void tile_range_kernel(const int* __restrict__ g_m, const int* __restrict__ g_orig, int n_tiles, int rows_per_unit,
                                  int lo_min, int hi_max, int n_lam, int lam_per_chunk, int all, int4* out) {
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (tile >= n_tiles) return;
  int mmin = INT_MAX, mmax = INT_MIN;
  for (int r = lane; r < rows_per_unit; r += 32) {
    const int row = tile * rows_per_unit + r;
    if (g_orig[row] >= 0) {
      mmin = min(mmin, g_m[row]);
      mmax = max(mmax, g_m[row]);
    }
  }
  for (int o = 16; o; o >>= 1) {
    mmin = min(mmin, __shfl_xor_sync(0xffffffffu, mmin, o));
    mmax = max(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
  }
  if (lane == 0) {
    int4 r;
    if (mmin > mmax) r = make_int4(0, -1, n_lam, -1);  // padding only
    else if (all) r = make_int4(0, (n_lam - 1) / lam_per_chunk, 0, n_lam - 1);
    else {
      const int i_lo = max(0, lo_min - 1 - mmax), i_hi = min(n_lam - 1, hi_max - mmin);
      r = (i_hi < i_lo) ? make_int4(0, -1, n_lam, -1) : make_int4(i_lo / lam_per_chunk, i_hi / lam_per_chunk, i_lo & ~31, i_hi);
    }
    out[tile] = r;
  }
}

// Inoue+14 transmission exp(-tau(z, lam_i (1+z))) for the bins blueward of Ly-alpha.
// Block = one 128-galaxy tile x one strip of 64 wavelength bins; thread = galaxy; the store of 128 consecutive floats per
// bin is coalesced in the tile-blocked layout the contraction epilogue reads.
//
// With x = lam_i / 911.8 every term of tau is coef * x^p * (1+z)^p' and WHICH terms apply (lines in each regime, the
// Lyman-continuum branches) depends on z only through a handful of comparisons.  The galaxies of a tile are redshift
// neighbours, so for almost every (tile, bin) all 128 take the same branches and
//       tau_g = c0 + sum_k c_k * (1+z_g)^{p_k},   p = (1.2, 2.1, 3.7, 5.5, -0.3, 2, 3)
// with coefficients that depend on the bin only.  Phase 1 (one thread per bin of the strip) evaluates the branch logic
// at the tile's lowest and highest redshift; if they agree it stores the 8 coefficients, else it flags the bin.  Phase 2
// (one thread per galaxy) is then a 7-term float64 dot product per bin -- ~25 instructions instead of ~100 -- and only
// flagged bins (a tile straddling a regime boundary at that wavelength) take the exact per-galaxy evaluation.
// No pow(): bin powers are host-tabulated, the (1+z) powers come from scalars_kernel; "lines in regime k" are prefixes of
// the wavelength-sorted line list, so per-regime sums are differences of prefix sums.
constexpr int kIgmStrip = 64;

struct IgmRegime { int a, b, c, dla, laf, xc; };   // line-regime prefixes, Lyman-continuum branch classes

__device__ __forceinline__ IgmRegime igm_regime(const double* __restrict__ s_thr, double xl, double z, int J) {
  int n1 = 0, n2 = 0, nd = 0;   // lines (sorted by decreasing wavelength) still below the regime thresholds
#pragma unroll
  for (int step = 32; step; step >>= 1) {
    if (s_thr[0 * 64 + n1 + step - 1] > xl) n1 += step;
    if (s_thr[1 * 64 + n2 + step - 1] > xl) n2 += step;
    if (s_thr[2 * 64 + nd + step - 1] > xl) nd += step;
  }
  IgmRegime r;
  r.a = min(J, n1); r.b = min(J, n2); r.c = min(J, nd);
  r.dla = z < 2.0 ? 0 : (xl >= 3.0 ? 1 : 2);
  r.laf = z < 1.2 ? 0 : (z < 4.7 ? (xl >= 2.2 ? 1 : 2) : (xl > 5.7 ? 3 : ((xl >= 2.2 && xl < 5.7) ? 4 : (xl < 2.2 ? 5 : 6))));
  r.xc = 0;
  return r;
}

// coefficients of (1, Z0..Z6) for one bin in regime R; q = per-bin powers x^1.2, x^2.1, x^3.7, x^5.5, x^-0.3, x
__device__ __forceinline__ void igm_coefficients(const double* __restrict__ s_pre, const IgmRegime& R, int J, bool lc_on,
                                                 double x12, double x21, double x37, double x55, double xm3, double x1, double* c) {
  const double x2 = x1 * x1, x3 = x2 * x1;
#pragma unroll
  for (int k = 0; k < 8; ++k) c[k] = 0.0;
  c[1] = x12 * s_pre[0 * 64 + R.a];
  c[3] = x37 * (s_pre[1 * 64 + R.b] - s_pre[1 * 64 + R.a]);
  c[4] = x55 * (s_pre[2 * 64 + J] - s_pre[2 * 64 + R.b]);
  c[6] = x2 * s_pre[3 * 64 + R.c];
  c[7] = x3 * (s_pre[4 * 64 + J] - s_pre[4 * 64 + R.c]);
  if (!lc_on) return;
  // Lyman continuum, DLA component
  if (R.dla == 0) c[6] += 0.2113 - 0.07661 * xm3 - 0.1347 * x2;
  else if (R.dla == 1) c[7] += 0.04696 - 0.01779 * xm3 - 0.02916 * x3;
  else { c[0] += 0.6340; c[7] += 0.04696 - 0.01779 * xm3; c[6] -= 0.1347 * x2; c[5] -= 0.2905 * xm3; }
  // Lyman continuum, LAF component
  switch (R.laf) {
    case 0: c[1] += 0.3248 * (x12 - x21); break;
    case 1: c[3] += 2.545e-2 * (x21 - x37); break;
    case 2: c[3] += 2.545e-2 * x21; c[1] += 0.3248 * x12; c[2] -= 0.2496 * x21; break;
    case 3: c[4] += 5.221e-4 * (x21 - x55); break;
    case 4: c[4] += 5.221e-4 * x21; c[2] += 0.2182 * x21; c[3] -= 2.545e-2 * x37; break;
    case 5: c[4] += 5.221e-4 * x21; c[1] += 0.3248 * x12; c[2] -= 3.140e-2 * x21; break;
    default: break;
  }
}

__device__ __forceinline__ float igm_exp(double tau) {
  // exp(-tau) = 2^n * 2^f: the split n = round(y), f = y - n is done in float64 (adding 1.5 * 2^52 leaves round(y) in the
  // low word), only 2^f with |f| <= 1/2 in float32 (ex2.approx: 2 ulp), so the result keeps float32 accuracy for any tau
  const double y = -tau * 1.44269504088896340736;
  const double t = y + 6755399441055744.0;
  const int n = __double2loint(t);
  const float f = (float)(y - (t - 6755399441055744.0));
  const int ex = min(n, 127) + 127;
  return (y > -150.0 && ex > 0) ? __int_as_float(ex << 23) * ex2_approx(f) : 0.f;
}

__global__ void __launch_bounds__(128) igm_kernel(PrepModel M, const double* __restrict__ zpow, const int* __restrict__ g_orig,
                                                  float* __restrict__ igm, const int4* __restrict__ tile_range, int range_shift,
                                                  int nb_pad, long long n_pad) {
  const int first_bin = tile_range ? tile_range[blockIdx.x >> range_shift].z : 0;  // bins below it are never integrated (warp-uniform)
  if ((int)(blockIdx.y + 1) * kIgmStrip <= first_bin) return;
  __shared__ double s_thr[3 * 64];
  __shared__ double s_pre[5 * 64];
  __shared__ __align__(16) double s_bin[kIgmStrip * 8];   // per bin: x^1.2, x^2.1, x^3.7, x^5.5, x^-0.3, x, nline, lc_on
  __shared__ __align__(16) double s_c[kIgmStrip * 8];     // per bin: coefficients of (1, Z0..Z6)
  __shared__ int s_uni[kIgmStrip];
  __shared__ double s_zr[8];
  const int np1 = M.n_lines + 1;
  const int nb = M.n_blue;
  const int strip0 = (int)blockIdx.y * kIgmStrip;
  for (int i = threadIdx.x; i < 3 * 64; i += 128) s_thr[i] = M.thr[i];
  for (int i = threadIdx.x; i < 5 * np1; i += 128) s_pre[(i / np1) * 64 + (i % np1)] = M.pre[i];
  for (int i = threadIdx.x; i < kIgmStrip * 8; i += 128) {
    const int b = strip0 + (i >> 3), k = i & 7;
    double v = 0.0;
    if (b < nb) {
      if (k < 5) v = M.bin_pow[k * nb + b];
      else if (k == 5) v = M.bin_pow[7 * nb + b];
      else if (k == 6) v = (double)M.nline[b];
      else v = (double)M.lc_on[b];
    }
    s_bin[i] = v;
  }
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
  const double z = zpow[(size_t)12 * n_pad + t];
  {  // the tile's redshift range over its real galaxies
    const bool valid = g_orig[t] >= 0;
    double zlo = valid ? z : 1e300, zhi = valid ? z : -1e300;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      zlo = fmin(zlo, __shfl_xor_sync(0xffffffffu, zlo, o));
      zhi = fmax(zhi, __shfl_xor_sync(0xffffffffu, zhi, o));
    }
    if ((threadIdx.x & 31) == 0) { s_zr[threadIdx.x >> 5] = zlo; s_zr[4 + (threadIdx.x >> 5)] = zhi; }
  }
  __syncthreads();
  const double zlo = fmin(fmin(s_zr[0], s_zr[1]), fmin(s_zr[2], s_zr[3])), zhi = fmax(fmax(s_zr[4], s_zr[5]), fmax(s_zr[6], s_zr[7]));
  const int i0 = max(strip0, first_bin);
  const int i1 = min(nb, strip0 + kIgmStrip);
  // ---- phase 1: one thread per bin of the strip
  if (threadIdx.x < kIgmStrip) {
    const int i = strip0 + threadIdx.x;
    int uni = 0;
    if (i >= i0 && i < i1 && zlo <= zhi) {
      const double* q = s_bin + threadIdx.x * 8;
      const int J = (int)q[6];
      const IgmRegime r0 = igm_regime(s_thr, q[5] * (1.0 + zlo), zlo, J), r1 = igm_regime(s_thr, q[5] * (1.0 + zhi), zhi, J);
      const bool lc = q[7] != 0.0;
      uni = (r0.a == r1.a && r0.b == r1.b && r0.c == r1.c && (!lc || (r0.dla == r1.dla && r0.laf == r1.laf))) ? 1 : 0;
      if (uni) igm_coefficients(s_pre, r0, J, lc, q[0], q[1], q[2], q[3], q[4], q[5], s_c + threadIdx.x * 8);
    }
    s_uni[threadIdx.x] = uni;
  }
  __syncthreads();
  // ---- phase 2: one thread per galaxy
  double Z[7];
#pragma unroll
  for (int p = 0; p < 7; ++p) Z[p] = zpow[(size_t)p * n_pad + t];
  float* out = igm + ((size_t)blockIdx.x * nb_pad) * 128 + threadIdx.x;
  for (int i = max(i0, nb); i < min(nb_pad, strip0 + kIgmStrip); ++i) out[(size_t)i * 128] = 1.f;  // padding rows
  if (i0 >= nb) return;
  float* op = out + (size_t)i0 * 128;
  for (int i = i0; i < i1; ++i, op += 128) {
    const int b = i - strip0;
    double tau;
    if (s_uni[b]) {
      const double2* c = reinterpret_cast<const double2*>(s_c + b * 8);
      const double2 c01 = c[0], c23 = c[1], c45 = c[2], c67 = c[3];
      tau = c01.x;
      tau = fma(c01.y, Z[0], tau);
      tau = fma(c23.x, Z[1], tau);
      tau = fma(c23.y, Z[2], tau);
      tau = fma(c45.x, Z[3], tau);
      tau = fma(c45.y, Z[4], tau);
      tau = fma(c67.x, Z[5], tau);
      tau = fma(c67.y, Z[6], tau);
    } else {   // the tile straddles a regime boundary at this wavelength: this galaxy's own branches
      const double* q = s_bin + b * 8;
      const int J = (int)q[6];
      const IgmRegime r = igm_regime(s_thr, q[5] * (1.0 + z), z, J);
      double c[8];
      igm_coefficients(s_pre, r, J, q[7] != 0.0, q[0], q[1], q[2], q[3], q[4], q[5], c);
      tau = c[0];
#pragma unroll
      for (int k = 0; k < 7; ++k) tau = fma(c[k + 1], Z[k], tau);
    }
    *op = igm_exp(tau);
  }
}

}  // namespace sb2
