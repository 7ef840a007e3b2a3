/*
 * CPU ORACLE (C / OpenMP restatement) -- TEST INFRASTRUCTURE ONLY.
 *
 * Same algorithm as oracle/oracle.py (float64 throughout), written the way Synthesizer's own
 * OpenMP C extensions do the work behind Pipeline.run (src/synference/library.py:2562-2576,
 * :2619): per galaxy a weighted sum over the (age, Z) grid cells, then attenuation, redshift,
 * IGM and last-axis trapezoid filter integration, threaded over galaxies.  It is the checker
 * for large parity cases and the CPU baseline bench.py times ("kind": "port"); the product
 * never links or calls it.  Validated against oracle.py in tests/test_oracle_golden.py.
 * Parity status: see the header of oracle.py ("parity unpinned" behind the Synthesizer boundary).
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC oracle_c.c -o _build/liboracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define C_ANG 2.99792458e18

/* ---- A7 Planck18 (same constants as oracle.py) ------------------------------------------- */
static const double H0 = 67.66, OM0 = 0.30966, TCMB = 2.7255, NEFF = 3.046, MNU = 0.06;
static double OGAM = -1.0, ODE;

static double nu_rel(double z) {
  double nu_y = MNU / (8.617333262e-5 * 0.7137658555036082 * TCMB);
  double rel = pow(1.0 + pow(0.3173 * nu_y / (1.0 + z), 1.83), 0.54644808743) + 2.0;
  return 0.22710731766 * (NEFF / 3.0) * rel;
}
static void cosmo_init(void) {
  if (OGAM >= 0.0) return;
  double h0 = H0 * 1e3 / 3.0856775814913673e22;
  double rho_c = 3.0 * h0 * h0 / (8.0 * M_PI * 6.6743e-11);
  double c = 299792.458e3;
  double og = 4.0 * 5.670374419e-8 / (c * c * c) * pow(TCMB, 4) / rho_c;
  ODE = 1.0 - OM0 - og * (1.0 + nu_rel(0.0));
  OGAM = og;
}
static double efunc(double z) {
  double zp = 1.0 + z;
  return sqrt(zp * zp * zp * (OGAM * (1.0 + nu_rel(z)) * zp + OM0) + ODE);
}
/* luminosity distance [cm]: composite 16-point Gauss-Legendre over ln(1+z) */
static const double GLX[8] = {0.0950125098376374, 0.2816035507792589, 0.4580167776572274, 0.6178762444026438,
                              0.7554044083550030, 0.8656312023878318, 0.9445750230732326, 0.9894009349916499};
static const double GLW[8] = {0.1894506104550685, 0.1826034150449236, 0.1691565193950025, 0.1495959888165767,
                              0.1246289712555339, 0.0951585116824928, 0.0622535239386479, 0.0271524594117541};
double oracle_luminosity_distance_cm(double z) {
  cosmo_init();
  double s1 = log1p(z), acc = 0.0;
  int panels = 8 + (int)(s1 * 8.0);
  for (int p = 0; p < panels; ++p) {
    double a = s1 * p / panels, b = s1 * (p + 1) / panels, mid = 0.5 * (a + b), half = 0.5 * (b - a);
    for (int i = 0; i < 8; ++i) {
      double sa = mid - half * GLX[i], sb = mid + half * GLX[i];
      acc += half * GLW[i] * (exp(sa) / efunc(expm1(sa)) + exp(sb) / efunc(expm1(sb)));
    }
  }
  return (1.0 + z) * acc * (299792.458 / H0) * 3.0856775814913673e24;
}

/* ---- A2 SFH bin masses (closed forms) ------------------------------------------------------ */
static double phi_diff(double ul, double uh) {
  const double r = 0.70710678118654752440;
  if (ul + uh > 0.0) return 0.5 * (erfc(ul * r) - erfc(uh * r));
  return 0.5 * (erfc(-uh * r) - erfc(-ul * r));
}
static double bin_mass(int type, const double* row, double e_lo, double e_hi) {
  double mn = row[0], mx = row[1];
  const double* p = row + 2;
  if (type == 7) {
    int nb = (int)p[0];
    const double *edges = p + 1, *ratios = p + 1 + nb + 1;
    double sfr = 1.0, m = 0.0;
    for (int j = 0; j < nb; ++j) {
      if (j > 0) sfr *= pow(10.0, -ratios[j - 1]);
      double ov = fmin(e_hi, edges[j + 1]) - fmax(e_lo, edges[j]);
      if (ov > 0.0) m += sfr * ov;
    }
    return m;
  }
  double lo = fmin(fmax(e_lo, mn), mx), hi = fmin(fmax(e_hi, mn), mx);
  if (!(hi > lo)) return 0.0;
  switch (type) {
    case 0: return hi - lo;
    case 1: return p[1] * 2.50662827463100050242 * phi_diff((lo - p[0]) / p[1], (hi - p[0]) / p[1]);
    case 2:
    case 3: {
      double tau = type == 2 ? p[0] : -p[0];
      double shift = tau > 0.0 ? (mx - mn) / tau : 0.0;
      return tau * (exp((mx - lo) / tau - shift) - exp((mx - hi) / tau - shift));
    }
    case 4: {
      double tau = p[0], t1 = mx - lo, t2 = mx - hi;
      return -tau * (t1 + tau) * exp(-t1 / tau) + tau * (t2 + tau) * exp(-t2 / tau);
    }
    case 5: {
      double tau = p[0], t0 = log(mx - p[1]) + tau * tau;
      double uh = (log(fmax(mx - hi, 1e-300)) - t0) / tau, ul = (log(fmax(mx - lo, 1e-300)) - t0) / tau;
      return tau * 2.50662827463100050242 * phi_diff(uh, ul);
    }
    default: return NAN;
  }
}

/* ---- A8 Inoue+14, written line by line like the upstream implementation ------------------- */
static double igm_tau(double z, double lobs, const double* laf, const double* dla, int nl) {
  const double z1l = 1.2, z2l = 4.7, z1d = 2.0, lamL = 911.8;
  double tau = 0.0, zp = 1.0 + z;
  for (int j = 0; j < nl; ++j) {
    double lj = laf[5 * j + 1];
    if (lobs < lj * zp) {
      double u = lobs / lj;
      if (lobs < lj * (1 + z1l)) tau += laf[5 * j + 2] * pow(u, 1.2);
      else if (lobs < lj * (1 + z2l)) tau += laf[5 * j + 3] * pow(u, 3.7);
      else tau += laf[5 * j + 4] * pow(u, 5.5);
      if (lobs < lj * (1 + z1d)) tau += dla[4 * j + 2] * u * u;
      else tau += dla[4 * j + 3] * u * u * u;
    }
  }
  if (lobs < lamL * zp) {
    double x = lobs / lamL;
    if (z < z1d) tau += 0.2113 * zp * zp - 0.07661 * pow(zp, 2.3) * pow(x, -0.3) - 0.1347 * x * x;
    else if (lobs >= lamL * (1 + z1d)) tau += 0.04696 * zp * zp * zp - 0.01779 * pow(zp, 3.3) * pow(x, -0.3) - 0.02916 * x * x * x;
    else tau += 0.6340 + 0.04696 * zp * zp * zp - 0.01779 * pow(zp, 3.3) * pow(x, -0.3) - 0.1347 * x * x - 0.2905 * pow(x, -0.3);
    if (z < z1l) tau += 0.3248 * (pow(x, 1.2) - pow(zp, -0.9) * pow(x, 2.1));
    else if (z < z2l) {
      if (lobs >= lamL * (1 + z1l)) tau += 2.545e-2 * (pow(zp, 1.6) * pow(x, 2.1) - pow(x, 3.7));
      else tau += 2.545e-2 * pow(zp, 1.6) * pow(x, 2.1) + 0.3248 * pow(x, 1.2) - 0.2496 * pow(x, 2.1);
    } else {
      if (lobs > lamL * (1 + z2l)) tau += 5.221e-4 * (pow(zp, 3.4) * pow(x, 2.1) - pow(x, 5.5));
      else if (lobs >= lamL * (1 + z1l) && lobs < lamL * (1 + z2l))
        tau += 5.221e-4 * pow(zp, 3.4) * pow(x, 2.1) + 0.2182 * pow(x, 2.1) - 2.545e-2 * pow(x, 3.7);
      else if (lobs < lamL * (1 + z1l))
        tau += 5.221e-4 * pow(zp, 3.4) * pow(x, 2.1) + 0.3248 * pow(x, 1.2) - 3.140e-2 * pow(x, 2.1);
    }
  }
  return tau;
}

/* np.interp(x, xp, fp, left=0, right=0) with xp ascending */
static double interp0(double x, const double* xp, const double* fp, int n) {
  if (x < xp[0] || x > xp[n - 1]) return 0.0;
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (xp[mid] <= x) lo = mid; else hi = mid;
  }
  if (x == xp[hi]) return fp[hi];
  double s = (fp[hi] - fp[lo]) / (xp[hi] - xp[lo]);
  return s * (x - xp[lo]) + fp[lo];
}

/*
 * The whole path.  Grid components g_att / g_un are [n_age][n_z][n_lam] (either may be NULL).
 * filters: concatenated own-axis tables (ascending wavelength) with offsets filt_off[n_filt+1].
 * variant 0: integrate in nu (filter table interpolated in nu); 1: in lambda.
 * Returns 0, or (1 + galaxy index) of the first galaxy for which a filter has no in-band sample.
 */
/* Extended form: optional second dust screen -- stars of the age bins flagged in `young` sit behind
 * exp(-tau_v_birth kappa_birth - tau_v kappa), the others behind exp(-tau_v kappa) -- and optional dust emission with
 * energy balance: + E_abs * dust_shape(nu), E_abs = trapezoid over nu of the light the screen(s) removed. */
int oracle_synthesize_ex(int64_t n_gal, const double* redshift, const double* tau_v, int sfh_type, int sfh_stride,
                      const double* sfh_rows, int zd_type, const double* zd_value, const double* zd_sigma,
                      int n_age, int n_z, int n_lam, const double* log10ages, const double* zmet,
                      const double* lam, const double* g_att, const double* g_un, const double* kappa,
                      int igm_on, const double* laf, const double* dla, int n_lines, int n_filt,
                      const int64_t* filt_off, const double* filt_lam, const double* filt_t, int variant,
                      double base_mass, int nthreads, double* out_flux, double* out_spec,
                      const double* kappa_birth, const double* tau_v_birth, const int* young, const double* dust_shape) {
  cosmo_init();
  int bad = 0;
  double* ages = (double*)malloc(sizeof(double) * n_age);
  double* edges = (double*)malloc(sizeof(double) * n_age);
  for (int i = 0; i < n_age; ++i) ages[i] = pow(10.0, log10ages[i]);
  edges[0] = 0.0;
  for (int i = 0; i + 1 < n_age; ++i) edges[i + 1] = 0.5 * (ages[i] + ages[i + 1]);
  /* filter tables in the integration variable, ascending */
  int64_t ntab = filt_off[n_filt];
  double* fx = (double*)malloc(sizeof(double) * ntab);
  double* ft = (double*)malloc(sizeof(double) * ntab);
  for (int f = 0; f < n_filt; ++f) {
    int64_t a = filt_off[f], b = filt_off[f + 1];
    for (int64_t i = a; i < b; ++i) {
      if (variant == 0) { fx[a + (b - 1 - i)] = C_ANG / filt_lam[i]; ft[a + (b - 1 - i)] = filt_t[i]; }
      else { fx[i] = filt_lam[i]; ft[i] = filt_t[i]; }
    }
  }
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    double* sf = (double*)malloc(sizeof(double) * n_age);
    double* zd = (double*)malloc(sizeof(double) * n_z);
    double* fnu = (double*)malloc(sizeof(double) * n_lam);
    double* att = (double*)malloc(sizeof(double) * n_lam);
    double* atty = (double*)malloc(sizeof(double) * n_lam);   /* young stars' share of the attenuated grid (two screens) */
    double* x = (double*)malloc(sizeof(double) * n_lam);
    double* tb = (double*)malloc(sizeof(double) * n_lam);
    double row[24];
#pragma omp for schedule(dynamic, 16)
    for (int64_t g = 0; g < n_gal; ++g) {
      double z = redshift[g];
      memset(row, 0, sizeof(row));
      memcpy(row, sfh_rows + g * sfh_stride, sizeof(double) * sfh_stride);
      double tot = 0.0;
      for (int a = 0; a < n_age; ++a) {
        sf[a] = (a < n_age - 1) ? bin_mass(sfh_type, row, edges[a], edges[a + 1]) : 0.0;
        tot += sf[a];
      }
      double ztot = 0.0;
      if (zd_type <= 1) {
        int j = 0;
        double v = zd_value[g], f = 0.0;
        double x0 = zd_type == 1 ? log10(zmet[0]) : zmet[0], xl = zd_type == 1 ? log10(zmet[n_z - 1]) : zmet[n_z - 1];
        for (int i = 0; i < n_z; ++i) zd[i] = 0.0;
        if (v <= x0) zd[0] = 1.0;
        else if (v >= xl) zd[n_z - 1] = 1.0;
        else {
          for (int i = 1; i < n_z; ++i) { double xi = zd_type == 1 ? log10(zmet[i]) : zmet[i]; if (xi <= v) j = i; }
          double xa = zd_type == 1 ? log10(zmet[j]) : zmet[j], xb = zd_type == 1 ? log10(zmet[j + 1]) : zmet[j + 1];
          f = (v - xa) / (xb - xa);
          zd[j] = 1.0 - f; zd[j + 1] = f;
        }
        ztot = 1.0;
      } else {
        for (int i = 0; i < n_z; ++i) {
          double xi = zd_type == 3 ? log10(zmet[i]) : zmet[i];
          double u = (xi - zd_value[g]) / zd_sigma[g];
          zd[i] = exp(-0.5 * u * u); ztot += zd[i];
        }
      }
      /* A4: grid-weighted sum over (age, Z) cells */
      for (int i = 0; i < n_lam; ++i) { fnu[i] = 0.0; att[i] = 0.0; atty[i] = 0.0; }
      for (int a = 0; a < n_age; ++a) {
        if (sf[a] == 0.0) continue;
        for (int iz = 0; iz < n_z; ++iz) {
          double w = sf[a] * zd[iz] / (tot * ztot);
          if (w == 0.0) continue;
          size_t base = ((size_t)a * n_z + iz) * n_lam;
          if (g_un) for (int i = 0; i < n_lam; ++i) fnu[i] += w * g_un[base + i];
          if (g_att) {
            double* dst = (kappa_birth && young && young[a]) ? atty : att;
            for (int i = 0; i < n_lam; ++i) dst[i] += w * g_att[base + i];
          }
        }
      }
      double dl = oracle_luminosity_distance_cm(z);
      double scale = base_mass * (1.0 + z) / (4.0 * M_PI * dl * dl) * 1e23 * 1e9;
      double tv = tau_v ? tau_v[g] : 0.0;
      double tvb = tau_v_birth ? tau_v_birth[g] : 0.0;
      if (g_att && kappa) {   /* screens, then energy balance on the rest-frame axis */
        double e_abs = 0.0, prev_d = 0.0;
        for (int i = 0; i < n_lam; ++i) {
          double before = att[i] + atty[i];
          double after = att[i] * exp(-tv * kappa[i]);
          if (kappa_birth) after += atty[i] * exp(-tvb * kappa_birth[i] - tv * kappa[i]);
          double d = before - after;
          if (i > 0) e_abs += 0.5 * (d + prev_d) * (C_ANG / lam[i - 1] - C_ANG / lam[i]);
          prev_d = d;
          att[i] = after;
        }
        if (dust_shape) for (int i = 0; i < n_lam; ++i) att[i] += e_abs * dust_shape[i];
      }
      for (int i = 0; i < n_lam; ++i) {
        double a_ = att[i];
        double lobs = lam[i] * (1.0 + z);
        double v = (fnu[i] + a_) * scale;
        if (igm_on) v *= exp(-igm_tau(z, lobs, laf, dla, n_lines));
        fnu[i] = v;
        x[i] = variant == 0 ? C_ANG / lobs : lobs;
      }
      if (out_spec) memcpy(out_spec + g * n_lam, fnu, sizeof(double) * n_lam);
      /* A9: filter integration over in-band (T > 0) samples */
      for (int f = 0; f < n_filt; ++f) {
        int64_t a = filt_off[f];
        int nt = (int)(filt_off[f + 1] - a);
        double num = 0.0, den = 0.0;
        int prev = -1, cnt = 0;
        for (int i = 0; i < n_lam; ++i) {
          tb[i] = interp0(x[i], fx + a, ft + a, nt);
          if (tb[i] > 0.0) {
            if (prev >= 0) {
              double dx = x[i] - x[prev];
              num += 0.5 * (fnu[i] * tb[i] / x[i] + fnu[prev] * tb[prev] / x[prev]) * dx;
              den += 0.5 * (tb[i] / x[i] + tb[prev] / x[prev]) * dx;
            }
            prev = i; ++cnt;
          }
        }
        if (cnt == 0) {
#pragma omp critical
          { if (!bad) bad = (int)(g + 1); }
          out_flux[g * n_filt + f] = NAN;
        } else out_flux[g * n_filt + f] = num / den;
      }
    }
    free(sf); free(zd); free(fnu); free(att); free(atty); free(x); free(tb);
  }
  free(ages); free(edges); free(fx); free(ft);
  return bad;
}

int oracle_synthesize(int64_t n_gal, const double* redshift, const double* tau_v, int sfh_type, int sfh_stride,
                      const double* sfh_rows, int zd_type, const double* zd_value, const double* zd_sigma,
                      int n_age, int n_z, int n_lam, const double* log10ages, const double* zmet,
                      const double* lam, const double* g_att, const double* g_un, const double* kappa,
                      int igm_on, const double* laf, const double* dla, int n_lines, int n_filt,
                      const int64_t* filt_off, const double* filt_lam, const double* filt_t, int variant,
                      double base_mass, int nthreads, double* out_flux, double* out_spec) {
  return oracle_synthesize_ex(n_gal, redshift, tau_v, sfh_type, sfh_stride, sfh_rows, zd_type, zd_value, zd_sigma, n_age, n_z,
                              n_lam, log10ages, zmet, lam, g_att, g_un, kappa, igm_on, laf, dla, n_lines, n_filt, filt_off,
                              filt_lam, filt_t, variant, base_mass, nthreads, out_flux, out_spec, NULL, NULL, NULL, NULL);
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
