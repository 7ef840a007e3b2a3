/*
 * synference_b200 -- C ABI of the B200-native mock-library hot path.
 *
 * The reference (synthesizer-project/synference) is pure Python and has no FFI of its
 * own: the path is reached through its Python API (src/synference/__init__.py:49-115)
 * and runs inside the third-party `synthesizer` package.  This header declares the
 * entry points a ctypes binding in the reference would call in place of those stages.
 * Each entry point names the reference interface it replaces (paths relative to the
 * reference repository).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a
 * negative sb2_status and records a message retrievable with sb2_last_error(); no
 * function throws or allocates caller-visible memory; `stream` is a cudaStream_t passed
 * as void* (NULL = default stream).  There is no CPU fallback: without a CUDA device the
 * compute entry points fail with SB2_ERR_CUDA.
 */
#ifndef SYNFERENCE_B200_H
#define SYNFERENCE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum sb2_status {
  SB2_OK = 0,
  SB2_ERR_INVALID = -1, /* bad argument / unsupported shape */
  SB2_ERR_CUDA = -2,    /* CUDA runtime or driver error     */
  SB2_ERR_CAPACITY = -3 /* batch larger than the context's workspace */
} sb2_status;

/* SFH families (synthesizer.parametric.SFH.* as constructed at
 * src/synference/library.py:1313). Row layout of sb2_params.sfh_rows:
 *   [0]=min_age [1]=max_age (yr, lookback) then
 *   GAUSSIAN: peak_age, sigma | EXPONENTIAL/DECLINING/DELAYED: tau | LOGNORMAL: tau, peak_age
 *   DOUBLE_POWERLAW: peak_age, alpha, beta (bins integrated by 16-point Gauss-Legendre)
 *   CONTINUITY: n_bins, edges[n_bins+1] (yr), logsfr_ratios[n_bins-1]                        */
enum { SB2_SFH_CONSTANT = 0, SB2_SFH_GAUSSIAN = 1, SB2_SFH_EXPONENTIAL = 2, SB2_SFH_DECLINING_EXP = 3,
       SB2_SFH_DELAYED_EXP = 4, SB2_SFH_LOGNORMAL = 5, SB2_SFH_DOUBLE_POWERLAW = 6, SB2_SFH_CONTINUITY = 7 };
#define SB2_SFH_ROW 24

/* Metallicity distributions (synthesizer.parametric.ZDist.*). */
enum { SB2_ZD_DELTA_LINEAR = 0, SB2_ZD_DELTA_LOG10 = 1, SB2_ZD_NORMAL_LINEAR = 2, SB2_ZD_NORMAL_LOG10 = 3 };

/* Static model: SPS grid, emission recipe, dust curve, filters, IGM and cosmology tables.
 * Replaces the objects handed to GalaxyBasis(...) (src/synference/library.py:1515-1531):
 * grid, emission_model, instrument, cosmo.  All pointers are HOST pointers, copied at
 * sb2_model_create time.                                                                  */
typedef struct sb2_model_desc {
  int32_t n_age, n_z, n_lam, n_comp, n_filt;
  int32_t n_age_pad; /* n_age rounded up to a multiple of 4 (8 recommended): grid column of (iz, ia) is
                      * iz*n_age_pad + ia, so every metallicity's columns start 16-byte (32-byte) aligned,
                      * which the TMA box origin of a bracket-grouped tile needs                        */
  int32_t k_pad;   /* n_age_pad*n_z rounded up to a multiple of 32                          */
  int32_t n_chunk; /* wavelength chunks of 256/n_comp bins (pseudo-bins included, see x_bins) */
  const double* log10ages;     /* [n_age]                                                   */
  const double* metallicities; /* [n_z]                                                     */
  /* Transposed, TF32 hi/lo-split grid in internal units, K-major:
   * row = chunk*256 + comp*(256/n_comp) + bin_in_chunk ; column k = iz*n_age_pad + ia      */
  const float* gt_hi; /* [n_chunk*256][k_pad]                                               */
  const float* gt_lo; /* [n_chunk*256][k_pad]                                               */
  double grid_scale;  /* erg/s/Hz/Msun per internal unit                                    */
  const float* kappa; /* [n_chunk*256/n_comp] tau(lambda)/tau_V, zero padded (NULL: no dust) */
  double lam0, q;     /* geometric wavelength axis lam_i = lam0*q^i [Angstrom]              */
  int32_t interp_variant; /* 0: filters interpolated/integrated in nu, 1: in lambda (A9)    */
  /* filters on the shared axis: band [filt_lo, filt_hi], packed (U, V) weight pairs        */
  const int32_t* filt_lo;
  const int32_t* filt_hi;
  const int32_t* filt_off;
  const float* filt_uv; /* float2[filt_uv_len]                                              */
  int32_t filt_uv_len;
  const double* filt_su;  /* [n_filt] sum(U): denominator = (1-beta)*su + beta*sdv          */
  const double* filt_sdv; /* [n_filt] sum(V)                                                */
  /* Inoue+14 tables (NULL igm_bin_pow: no IGM)                                             */
  int32_t n_blue, n_lines;
  const double* igm_bin_pow; /* [8][n_blue] */
  const int32_t* igm_nline;  /* [n_blue]    */
  const int32_t* igm_lc_on;  /* [n_blue]    */
  const double* igm_thr;     /* [3][64]     */
  const double* igm_pre;     /* [5][n_lines+1] */
  /* cosmology: cubic-Hermite tables over s = ln(1+z)                                       */
  int32_t cosmo_n;
  double cosmo_smax;
  const double* cosmo_dc;   /* comoving distance [Mpc] */
  const double* cosmo_ddc;  /* d/ds                    */
  const double* cosmo_age;  /* age [Gyr]               */
  const double* cosmo_dage; /* d/ds                    */
  double base_mass;         /* Msun the base photometry is quoted at (1e9, library.py:3217) */
  int64_t max_batch;        /* workspace capacity in galaxies                               */
  /* Optional tables for the weight builder's float64 special functions (normal-CDF tail, log, exp); NULL: the
   * CUDA math library is used instead (slower, same results to ~1e-12).  Built by synference_b200/fastmath.py:
   * fm_log_tab [256][2] = ln(m0), 1/m0 ; fm_exp_tab [64] = 2^(j/64) ; fm_tail_tab [fm_tail_n][8] = local
   * degree-7 polynomials of exp(u^2/2) Q(u) on intervals of width fm_tail_w.                              */
  /* Optional per-galaxy dust-curve shape (Calzetti2000(slope="...", ampl="...") of the reference's production script,
   * final_library_generation_multinode.py:496): tau(lambda)/tau_V = (kappa + ampl_g * dust_d0) * 2^(slope_g * dust_l2),
   * with `kappa` then holding the curve at slope = 0, ampl = 0; dust_d0 = bump profile / k(0.55 um) and
   * dust_l2 = log2(lambda / 0.55 um) on the same padded axis as kappa.  NULL: the curve is global (kappa only).   */
  const float* dust_d0;
  const float* dust_l2;
  /* Optional per-galaxy Lyman-alpha escape fraction (fesc_ly_alpha="fesc_lya" in the reference's production script): the
   * grids are then lowered WITHOUT the line-continuum value of the one bin nearest 1216 A (lya_bin), lya_line[iz*n_age+ia]
   * holds that value (internal units, times (1 - fesc) when fesc is global), and the kernel adds
   * fesc_lya_g * sum_k w_k lya_line[k] to the first component at that bin.  NULL: fesc_ly_alpha is global.          */
  const double* lya_line;
  int32_t lya_bin;
  /* Optional second dust screen (BimodalPacmanEmission of the reference's scripts, generate_library_full.py:221-231):
   * with n_comp = 2 the two grid components are then the YOUNG and the OLD stellar populations (split at age_pivot),
   * both attenuated: young by exp(-tau_v_birth * kappa_birth - tau_v * kappa), old by exp(-tau_v * kappa).
   * kappa_birth is on the same padded axis as kappa.  NULL: component 2 is unattenuated (escaped light).          */
  const float* kappa_birth;
  /* Optional dust emission with energy balance (spectrum "total" of PacmanEmission / TotalEmission / BimodalPacmanEmission
   * with a Greybody / Blackbody generator: min_example.py:110-120, final_library_generation_multinode.py:489-497):
   *   L_total(nu) = L_emergent(nu) + E_abs * g(nu),  E_abs = trapezoid over nu of (unattenuated - attenuated) light.
   * dust_wnu: the trapezoid weights (Hz / 1e14) on the same padded axis as kappa; dust_g: g(nu) * 1e14 on the n_lam axis (it
   * must vanish where the IGM acts); dust_duv[m][f] = sum_i dust_g[i] * (U_f, V_f)[i + m] for every integer redshift shift
   * m < dust_m_len: the emission's contribution to the filter numerators per unit E_abs.  With dust emission every
   * wavelength chunk is multiplied (the reduction needs the whole axis).  NULL: none.                                   */
  const float* dust_wnu;
  const float* dust_g;
  const float* dust_duv;
  int32_t dust_m_len;
  const double* fm_log_tab;
  const double* fm_exp_tab;
  const double* fm_tail_tab;
  int32_t fm_tail_n;
  double fm_tail_w;
  /* 1: rest-frame luminosities through the filters instead of observed fluxes (GalaxySimulator output_type "photo_lnu",
   * src/synference/library.py:5756-5761): no redshift shift, no distance factor -- results are L_nu in erg/s/Hz at
   * base_mass; create the model without IGM tables.  redshift is then only used for max_age_from_z.            */
  int32_t rest_frame;
  /* Optional PSEUDO-BINS for the absorbed-energy sum of the dust emission (needs dust_wnu; 0: the sum runs over the axis).
   * sum_i wnu_i L_i (1 - exp(-tau kappa_i)) depends on a wavelength only through kappa_i, so the axis can be projected onto
   * x_bins nodes in kappa: rows [x_bin0, x_bin0 + x_bins) of the padded axis (x_bin0 >= n_lam, both multiples of 192) hold
   * gt = sum_i wnu_i l_j(kappa_i) grid_i (l_j: Lagrange weights of node j), kappa = the node and dust_wnu = 1, and n_chunk
   * covers them.  The bracket-grouped kernel then multiplies these x_bins extra rows per tile instead of the whole axis and
   * keeps its wavelength-chunk skipping; other kernels ignore them.  Accurate to < 1e-6 of E_abs while
   * tau_V * (node spacing) < 0.42 -- beyond that the caller sets sb2_params.energy_full_axis.                          */
  int32_t x_bin0, x_bins;
} sb2_model_desc;

/* Per-galaxy parameters, struct of arrays (float64).  Replaces the per-galaxy object lists
 * (sfhs, metal_dists, redshifts, galaxy_params) of GalaxyBasis / create_galaxy
 * (src/synference/library.py:1340-1424, 2168-2261).                                        */
typedef struct sb2_params {
  int64_t n;
  const double* redshift; /* [n] */
  const double* log_mass; /* [n] log10(M/Msun) (NULL: base mass)                            */
  const double* tau_v;    /* [n] (NULL: 0)                                                  */
  int32_t sfh_type;
  int32_t sfh_stride;     /* doubles per row actually stored (<= SB2_SFH_ROW; rest are 0)   */
  const double* sfh_rows; /* [n][sfh_stride]                                                */
  /* optional: derive max_age = age(z) - age_zmax on device (library.py:1206) and scale
   * `_norm` parameters by it (library.py:1287-1289): bit i set => row[2+i] *= max_age       */
  int32_t max_age_from_z;
  uint32_t norm_mask;
  double age_zmax_gyr;
  int32_t zd_type;
  const double* zd_value; /* [n] */
  const double* zd_sigma; /* [n] (NULL for delta) */
  const double* coef_att;   /* [n] optional per-galaxy factor on the attenuated component   */
  const double* coef_unatt; /* [n] optional per-galaxy factor on the unattenuated component */
  const double* dust_slope; /* [n] per-galaxy power-law slope delta of the dust curve (requires dust_d0/dust_l2)  */
  const double* dust_ampl;  /* [n] per-galaxy UV-bump amplitude                       (requires dust_d0/dust_l2)  */
  const double* fesc_lya;   /* [n] per-galaxy Lyman-alpha escape fraction             (requires lya_line)         */
  const double* tau_v_birth;/* [n] birth-cloud optical depth of the young population (requires kappa_birth; tau_v = ISM) */
  /* HOST entry points only: 1 = every array above holds float32 values (the pointers are then const float* cast to
   * const double*), as draw_from_hypercube returns its draws (src/synference/library.py:1098).  They cross PCIe as
   * float32 -- half the bytes -- and are widened to float64 on the device; with max_age_from_z the SFH rows need no
   * host-side float64 arithmetic at all.  Must be 0 for the device entry points.                                   */
  int32_t host_f32;
  /* Layout of `flux_scaled`: 0 = [n][n_filt] (one row per galaxy).  > 0 = TRANSPOSED, [n_filt][scaled_ld] with galaxy g of
   * this call in column g (scaled_ld >= n): the (n_filters, n_galaxies) layout of a library's Grid/Photometry
   * (src/synference/library.py:4739-4742, :4074-4100), so that a batch lands in its column range of the library matrix
   * without a host-side transpose -- pass `matrix + first_column` and scaled_ld = the matrix's row length.
   * Not available together with spec_out on the host entry.                                                        */
  int64_t scaled_ld;
  /* 1: form the dust emission's absorbed energy over the whole wavelength axis even if the model holds pseudo-bins
   * (sb2_model_desc.x_bins): for optical depths beyond the pseudo-bins' accuracy range.                              */
  int32_t energy_full_axis;
} sb2_params;

typedef struct sb2_model sb2_model;

const char* sb2_last_error(void);
int sb2_device_count(void);
/* Number of kernels of THIS library launched by the calling process so far (the sort's CUB kernels are not counted).
 * Measurement hook for bench.py's `gpu_launches`; no reference counterpart.                                       */
long long sb2_kernel_launches(void);

/* Build / destroy the device-resident model.  `device` is the CUDA ordinal. */
int sb2_model_create(const sb2_model_desc* desc, int device, sb2_model** out);
int sb2_model_destroy(sb2_model* m);

/* Diagnostics: which in-kernel barrier waits timed out in the last failed launch ("" if none). */
const char* sb2_wait_debug(sb2_model* m);

/* SFZH weights only (parity hook for Stars.__init__ -> _get_sfzh, library.py:1372-1379).
 * params: DEVICE pointers. w_out: device float64 [n][n_age*n_z] (k = iz*n_age + ia).       */
int sb2_build_weights(sb2_model* m, const sb2_params* params, double* w_out, void* stream);

/* The fused path for one batch with DEVICE-resident parameters:
 * weights -> grid-weighted sum (tcgen05) -> dust -> redshift/IGM -> filter integration.
 * Replaces GalaxyBasis.process_galaxies -> Pipeline.run (library.py:2447-2694) and the mass
 * scaling loop of CombinedBasis.create_full_library (library.py:4567-4609).
 *   flux_base   device float32 [n][n_filt]  nJy at base_mass            (may be NULL)
 *   flux_scaled device float64 [n][n_filt]  float32(base)*10^logM/base  (may be NULL; transposed if params->scaled_ld > 0)
 *   spec_out    device float32 [n][n_lam]   observed-frame f_nu [nJy] at base_mass on the
 *               rest-frame axis (Pipeline.get_observed_spectra, library.py:2604) (may be NULL) */
int sb2_synth_photometry(sb2_model* m, const sb2_params* params, float* flux_base, double* flux_scaled,
                         float* spec_out, void* stream);

/* Device time of the stages of the most recent sb2_synth_photometry call on this model, from CUDA
 * events recorded on its stream: out3 = {sort [ms], weight/IGM kernel [ms], contraction kernel [ms]}.
 * Blocks until that call has finished.  (Measurement hook; no reference counterpart beyond the
 * wall-clock `pipeline_time` attribute, library.py:2617-2622.)                                   */
int sb2_last_stage_ms(sb2_model* m, float* out3);

/* Same, with HOST buffers: copies parameters in, runs, copies results out (synchronous).  With spec_out (host float32
 * [n][n_lam]: Pipeline.get_observed_spectra kept for the library, src/synference/library.py:4887-4919 / :4610-4617) the batch
 * is walked in slices through two device buffers, so that the 4*n_lam bytes per galaxy leaving the device overlap the kernels
 * of the next slice (pinned spec_out for real overlap).                                                              */
int sb2_synth_photometry_host(sb2_model* m, const sb2_params* params, float* flux_base, double* flux_scaled,
                              float* spec_out);

/* Asynchronous form for streams of batches (a library is generated batch by batch, library.py:2447-2694
 * `batch_size`): submit() enqueues copy-in, kernels and copy-out of one batch on the model's three streams using
 * staging slot 0 or 1 and returns; wait() blocks until that slot's results are in the caller's host buffers.
 * Alternating the slots overlaps the PCIe copies of one batch with the kernels of the next.  The host buffers
 * (pinned for real overlap) must stay valid and untouched until wait(); one model serves one host thread. */
int sb2_synth_photometry_host_submit(sb2_model* m, const sb2_params* params, float* flux_base, double* flux_scaled,
                                     int slot);
int sb2_synth_photometry_host_wait(sb2_model* m, int slot);

/* Depth-based scatter + flux->AB feature rows on DEVICE buffers.
 * Replaces SBI_Fitter._apply_depths (sbi_runner.py:580-691) and the AB branch of
 * create_feature_array_from_raw_photometry (sbi_runner.py:1698-1716, 1927-1932).
 *   flux      float64 [n_gal][n_filt] (nJy)        sigma float64 [n_filt] (nJy, depth/sigma level)
 *   normals   float64 [n_filt][n_gal*n_scatter] injected N(0,1) draws, or NULL => Philox4x32-10
 *             keyed by (seed, epoch) with counter (row, filter quad): one block = the normals of four filters
 *   out_flux  float64 [n_filt][n_rows] noisy flux (may be NULL)   -- bit-exact vs numpy for injected draws
 *   out_feat  float32 [n_rows][2*n_filt] = (mag..., mag_err...)   (may be NULL)
 * n_rows = n_gal*n_scatter, row r = g*n_scatter + s (np.repeat order).                      */
int sb2_depth_noise_features(const double* flux, int64_t n_gal, int32_t n_filt, int32_t n_scatter,
                             const double* sigma, double min_flux_pc_error, const double* normals,
                             uint64_t seed, uint64_t epoch, double norm_mag_limit, double* out_flux,
                             double* out_sigma, float* out_feat, void* stream);

/* Feature rows only, Philox draws, float32 fluxes (a library's Grid/Photometry as stored; widened on the device, which is
 * what numpy does with float32 flux + float64 noise): the per-epoch resampling of a training set, sbi_runner.py:580-691 +
 * :1698-1716.  flux float32 [n_gal][n_filt], out_feat float32 [n_rows][2*n_filt]; both 16-byte aligned.  Same draws and
 * rows as sb2_depth_noise_features(flux as float64, normals = NULL, out_flux = out_sigma = NULL).                        */
int sb2_depth_noise_features_f32(const float* flux, int64_t n_gal, int32_t n_filt, int32_t n_scatter, const double* sigma,
                                 double min_flux_pc_error, uint64_t seed, uint64_t epoch, double norm_mag_limit,
                                 float* out_feat, void* stream);

/* Same with several depth sets (2-D `depths` of SBI_Fitter._apply_depths, sbi_runner.py:626-647): sigma_sets is device
 * float64 [n_sets][n_filt]; set_index device int32 [n_filt][n_scatter] says which set the rows [j*n_gal, (j+1)*n_gal) of
 * filter f use (the reference draws it with np.random.randint and expands it with np.repeat(..., n, axis=1)).           */
int sb2_depth_noise_features_sets(const double* flux, int64_t n_gal, int32_t n_filt, int32_t n_scatter,
                                  const double* sigma_sets, int32_t n_sets, const int32_t* set_index,
                                  double min_flux_pc_error, const double* normals, uint64_t seed, uint64_t epoch,
                                  double norm_mag_limit, double* out_flux, double* out_sigma, float* out_feat, void* stream);

/* ---- Spectroscopic training path (cfg 5): library spectrum -> instrument-frame pixels ------------------------------
 * Replaces the per-galaxy Python loop of SBI_Fitter.create_feature_array_from_raw_spectra (sbi_runner.py:1322-1334) over
 * transform_spectrum (utils.py:185-254): redshift the axis, convolve with the per-pixel Gaussian of
 * convolve_variable_width_gaussian (utils.py:129-182; sigma from the instrument's R(lambda) in quadrature with the model's,
 * in units of the MEDIAN pixel of the redshifted axis, truncated at ceil(trunc*sigma), nearest-edge padding, sigma <= 0.01
 * copies), then the flux-conserving rebin of `spectres` onto observed_wave (pixels not fully covered get `fill`).        */
typedef struct sb2_resample_desc {
  int32_t n_lam, n_px, n_res;
  const double* theory_wave;   /* [n_lam] rest-frame wavelengths of the library spectra, increasing                */
  const double* observed_wave; /* [n_px]  pixel centres of the instrument, same unit, increasing                   */
  const double* res_wave;      /* [n_res] resolution curve abscissa (observed frame)                               */
  const double* res_r;         /* [n_res] R = lambda / FWHM                                                        */
  const double* theory_r;      /* [n_lam] resolution of the model spectra, or NULL: theory_r_scalar (inf = none)  */
  double theory_r_scalar;
  double trunc;                /* 4.0 in the reference                                                             */
  double fill;                 /* 0.0 in the reference                                                             */
} sb2_resample_desc;
typedef struct sb2_resampler sb2_resampler;
int sb2_resampler_create(const sb2_resample_desc* desc, int device, sb2_resampler** out);
int sb2_resampler_destroy(sb2_resampler* r);
/* spectra: device float32 [n][n_lam] (one row per galaxy, as sb2_synth_photometry's spec_out); redshift: device float64 [n];
 * out: device float32 [n][n_px].  A non-finite or <= -1 redshift gives a NaN row.                                        */
int sb2_resample_spectra(sb2_resampler* r, const float* spectra, const double* redshift, int64_t n, float* out, void* stream);
/* Same with HOST buffers (synchronous; copies in, runs, copies out). */
int sb2_resample_spectra_host(sb2_resampler* r, const float* spectra, const double* redshift, int64_t n, float* out);
/* Device time [ms] of the most recent sb2_resample_spectra call on this resampler (blocks until it has finished). */
int sb2_resample_last_ms(sb2_resampler* r, float* ms);

/* ---- Empirical uncertainty models on the device ---------------------------------------------------------------------
 * Replaces SBI_Fitter._apply_empirical_noise_models (sbi_runner.py:813-903: per filter, row copy ->
 * GeneralEmpiricalUncertaintyModel.apply_noise, noise_models.py:818-880) for all filters and rows in one launch.
 * One sb2_empirical_model per filter: the binned catalogue statistics the interpolators are built from
 * (noise_models.py:347-381), units, and the upper-limit rules with their replacement values resolved on the host
 * (noise_models.py:882-957).  Units: *_is_ab = 1 for AB magnitudes, else *_to_jy is the size of the linear unit in Jy.   */
#define SB2_EMP_MAX_BINS 64
typedef struct sb2_empirical_model {
  int32_t n_bins, extrapolate;
  int32_t internal_is_ab, in_is_ab, out_is_ab;
  int32_t observed_error;     /* error_type == "observed": sigma drawn again at the noisy flux                       */
  int32_t upper_limits;       /* SNR-based upper limits on                                                         */
  int32_t ul_active;          /* ... and an upper-limit value exists (upper_limit_value is not None)               */
  double internal_to_jy, in_to_jy, out_to_jy;
  double sigma_clip;          /* < 0: plain normal scatter, else truncated at +-sigma_clip sigma                   */
  double snr_threshold;       /* treat_as_upper_limits_below                                                       */
  double ul_flux;             /* replacement flux (interpolation unit): the limit, or a numeric behaviour          */
  double ul_scatter_std;      /* "scatter_limit": std of the +-3 sigma truncated normal added to ul_flux; < 0: none */
  double ul_err;              /* replacement error (interpolation unit)                                            */
  double min_err, max_err;    /* final clip of the error (output unit)                                             */
  /* AsinhEmpiricalUncertaintyModel (noise_models.py:443-635): outputs are asinh magnitudes with softening asinh_b [Jy]
   * (utils.py:647-704).  asinh_mode 0: general model above; 1: tables in asinh magnitudes; 2: tables in the linear unit
   * internal_to_jy.  observed_error = (error_type != "empirical"); the upper-limit fields and out_* are unused.         */
  int32_t asinh_mode, reserved_;
  double asinh_b;
  double centers[SB2_EMP_MAX_BINS], median[SB2_EMP_MAX_BINS], stdev[SB2_EMP_MAX_BINS];
} sb2_empirical_model;
/* flux: device float64 [n_filt][n] (the reference's (N_f, N_rows) layout); models: HOST array [n_filt];
 * draws: device float64 [4][n_filt][n] injected draws (sigma uniform, scatter normal -- uniform when sigma-clipped --,
 * re-draw uniform, limit-scatter uniform) or NULL => Philox4x32-10 keyed by (seed, epoch), counter (row, filter);
 * out_flux / out_sigma: device float64 [n_filt][n] (out_sigma may be NULL).                                          */
int sb2_empirical_noise(const double* flux, int64_t n, int32_t n_filt, const sb2_empirical_model* models,
                        const double* draws, uint64_t seed, uint64_t epoch, double* out_flux, double* out_sigma,
                        void* stream);

/* ---- filter integration on a GENERAL wavelength axis (fallback) ------------------------------------------------------------
 * The fused epilogue needs grid and filters on one constant-R axis.  For models on the SPS grid's native axis (the
 * reference's README and tests: README.md:100-102, tests/conftest.py:70,85) the spectra of a batch are written to device
 * memory (sb2_synth_photometry with spec_out) and integrated here with the general semantics of
 * Sed.get_photo_fnu -> Filter.apply_filter: every filter's OWN table (offsets[f] .. offsets[f+1] of
 * filt_lam / filt_t, ascending wavelengths) interpolated linearly in the integration variable (variant 0: nu, 1: lambda)
 * onto the observed abscissa, samples with T > 0 only, trapezoid of f T / x over trapezoid of T / x.                    */
typedef struct sb2_filterset sb2_filterset;
int sb2_filterset_create(int32_t n_filt, const int64_t* offsets, const double* filt_lam, const double* filt_t,
                         const double* grid_lam, int32_t n_lam, int32_t variant, int device, sb2_filterset** out);
int sb2_filterset_destroy(sb2_filterset* s);
/* spectra: device float32 [n][n_lam] (spec_out); redshift / log_mass: device float64 [n] (log_mass may be NULL);
 * flux_base device float32 [n][n_filt] and/or flux_scaled device float64 = float32(base) * 10^log_mass / base_mass.  */
int sb2_filter_integrate(sb2_filterset* s, const float* spectra, const double* redshift, const double* log_mass, double base_mass,
                         int64_t n, float* flux_base, double* flux_scaled, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SYNFERENCE_B200_H */
