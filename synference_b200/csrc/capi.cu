// C ABI of the B200-native mock-library hot path (see include/synference_b200.h).
// Owns the device-resident model, the per-batch workspace, the TMA descriptors and the launches.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include "../../include/synference_b200.h"
#include "noise_kernel.cuh"
#include "empirical_kernel.cuh"
#include "general_filter_kernel.cuh"
#include "prep_kernel.cuh"
#include "resample_kernel.cuh"
#include "synth_kernel.cuh"
#include "synth3_kernel.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU_TRY(expr)                                                                         \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return fail(SB2_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));          \
  } while (0)

// SB2_DEBUG_SYNC=1: synchronise after every launch and name the stage that failed.
bool debug_sync() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("SB2_DEBUG_SYNC");
    v = (e && e[0] && e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}
std::atomic<long long> g_launches{0};   // kernels of this library launched by this process (sb2_kernel_launches)
#define STAGE_CHECK(name, st)                                                                  \
  do {                                                                                         \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                        \
    cudaError_t e_ = cudaGetLastError();                                                       \
    if (e_ == cudaSuccess && debug_sync()) e_ = cudaStreamSynchronize(st);                     \
    if (e_ != cudaSuccess) return fail(SB2_ERR_CUDA, std::string(name) + ": " + cudaGetErrorString(e_)); \
  } while (0)

template <class T>
int upload(T** dst, const T* src, size_t n) {
  *dst = nullptr;
  if (n == 0 || src == nullptr) return SB2_OK;
  CU_TRY(cudaMalloc(reinterpret_cast<void**>(dst), n * sizeof(T)));
  CU_TRY(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return SB2_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2-D row-major float32 matrix [rows][cols]; box = 32 columns (128 B) x box_rows, SWIZZLE_128B.
int make_tmap(CUtensorMap* tm, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(SB2_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SB2_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return SB2_OK;
}

// Grouping keys: (metallicity bracket, redshift).  Tiles of the contraction kernel hold galaxies of ONE
// bracket (so a DeltaConstant batch multiplies only the 2*n_age grid columns it can touch) in redshift
// order (so the filter windows of a warp's galaxies coincide).  Dense batches use bracket 0 for everyone.
constexpr int kMaxGroups = 64;
// Sort key = bracket << 12 | ln(1 + z) in 4096 steps up to z = 100 (three steps per wavelength bin of an R = 300 axis: what a
// tile needs is neighbours in the integer redshift SHIFT, not in z to 17 mantissa bits).  With <= 16 brackets the key has 16
// bits and the radix sort takes two 8-bit passes instead of four.
constexpr int kKeyShift = 12;
constexpr float kKeyScale = 4095.f / 4.7f;

__global__ void group_keys_kernel(const double* __restrict__ z, const double* __restrict__ zv,
                                  const double* __restrict__ zx, int n_z, int delta, unsigned* keys, int* idx,
                                  int* counts, long long n) {
  __shared__ int h[kMaxGroups];
  if (threadIdx.x < kMaxGroups) h[threadIdx.x] = 0;
  __syncthreads();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int j = 0;
    if (delta) {
      double f;
      j = sb2::delta_bracket(zx, n_z, zv[i], &f);
    }
    const float zf = (z[i] >= 0.0 && z[i] <= 1.0e6) ? (float)z[i] : 0.f;   // unusable redshifts sort first (scalars_kernel flags them)
    keys[i] = ((unsigned)j << kKeyShift) | min(4095u, (unsigned)(log1pf(zf) * kKeyScale));
    idx[i] = (int)i;
    atomicAdd(&h[j], 1);
  }
  __syncthreads();
  if (threadIdx.x < kMaxGroups && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], h[threadIdx.x]);
}

// counts -> first sorted position (cum) and first padded row (pad_start) of every group, the number of units
// (a unit = one tile of 128 rows, or a pair of tiles for the CTA-pair kernel) and each unit's first grid column.
__global__ void group_layout_kernel(const int* __restrict__ counts, int n_groups, int cols_per_group, int rows_per_unit,
                                    int* cum, int* pad_start, int* tile_k0, int* n_tiles_out, int max_tiles) {
  __shared__ int s_first[kMaxGroups + 1];
  if (threadIdx.x == 0) {
    int c = 0, t = 0;
    for (int j = 0; j < n_groups; ++j) {
      cum[j] = c;
      pad_start[j] = t * rows_per_unit;
      s_first[j] = t;
      c += counts[j];
      t += (counts[j] + rows_per_unit - 1) / rows_per_unit;
    }
    s_first[n_groups] = t;
    *n_tiles_out = t;
  }
  __syncthreads();
  const int total = min(s_first[n_groups], max_tiles);
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    int j = 0;
    while (j + 1 < n_groups && s_first[j + 1] <= t) ++j;
    tile_k0[t] = j * cols_per_group;
  }
}

// synth3_kernel's bfloat16 operand: per row and group of 8 k-values the 32 bytes [bf16(g_hi[0..7]) | bf16(g_lo[0..7])], i.e. the
// K = 16 slice that meets the weights' [w_lo | w_hi] in one kind::f16 MMA (same bytes per row as the TF32 tiles)
__global__ void cross_pack_kernel(const float* __restrict__ hi, const float* __restrict__ lo, uint32_t* __restrict__ x, size_t n_groups) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_groups; i += (size_t)gridDim.x * blockDim.x) {
    const float* h = hi + 8 * i;
    const float* l = lo + 8 * i;
    uint32_t* o = x + 8 * i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j] = sb2::pack_bf16x2(h[2 * j], h[2 * j + 1]);
      o[4 + j] = sb2::pack_bf16x2(l[2 * j], l[2 * j + 1]);
    }
  }
}

// float32 parameters staged by the host entry (sb2_params.host_f32) -> the float64 arrays the kernels read
// (the float32 staging area mirrors the float64 one element for element, so a segment is one offset into both; all of a
//  slice's arrays go in ONE launch: blockIdx.y picks the array)
struct WidenSegs {
  long long off[12];
  long long cnt[12];
};
__global__ void widen_kernel(const float* __restrict__ src, double* __restrict__ dst, const __grid_constant__ WidenSegs sg) {
  const long long off = sg.off[blockIdx.y], n = sg.cnt[blockIdx.y];
  const float* __restrict__ s = src + off;
  double* __restrict__ d = dst + off;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((off & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    // (segment starts on 16-byte boundaries of both areas) four elements per thread: one 16-byte load, two 16-byte stores
    const long long n4 = n >> 2;
    for (; i < n4; i += stride) {
      const float4 v = reinterpret_cast<const float4*>(s)[i];
      reinterpret_cast<double2*>(d)[2 * i] = make_double2((double)v.x, (double)v.y);
      reinterpret_cast<double2*>(d)[2 * i + 1] = make_double2((double)v.z, (double)v.w);
    }
    for (long long j = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) d[j] = (double)s[j];
    return;
  }
  for (; i < n; i += stride) d[i] = (double)s[i];
}

__global__ void group_scatter_kernel(const unsigned* __restrict__ keys_sorted, const int* __restrict__ perm_sorted,
                                     const int* __restrict__ cum, const int* __restrict__ pad_start, int* perm_pad,
                                     long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int j = (int)(keys_sorted[i] >> kKeyShift);
    perm_pad[pad_start[j] + (int)(i - cum[j])] = perm_sorted[i];
  }
}

}  // namespace

// Experiment / fallback switches (environment), read ONCE when a model is created -- never on the per-call path.
struct Switches {
  bool n256 = false, cta_pair = false, libm = false, weights_v1 = false, one_pass = false, trace = false, no_synth3 = false, no_split = false;
  bool no_fuse = false, tf32x3 = false, bf16_dense = false;
  int dbg = 0;
  long long host_slices = 0;   // 0: automatic
  int dense_grid = 0;          // CTAs of the dense-K contraction; 0: automatic (see dense_grid_size)
  static bool on(const char* k) { const char* e = std::getenv(k); return e && e[0] && e[0] != '0'; }
  void read() {
    n256 = on("SB2_N256"); cta_pair = on("SB2_CTA_PAIR"); libm = on("SB2_LIBM"); weights_v1 = on("SB2_WEIGHTS_V1");
    one_pass = on("SB2_ONE_PASS"); trace = on("SB2_TRACE"); no_synth3 = on("SB2_NO_SYNTH3"); no_split = on("SB2_NO_SPLIT");
    no_fuse = on("SB2_NO_FUSE"); tf32x3 = on("SB2_TF32X3"); bf16_dense = on("SB2_BF16_DENSE");
    if (const char* e = std::getenv("SB2_DBG")) dbg = std::atoi(e);
    if (const char* e = std::getenv("SB2_HOST_SLICES")) host_slices = std::atoll(e);
    if (const char* e = std::getenv("SB2_DENSE_GRID")) dense_grid = std::atoi(e);
  }
};

const sb2_model* g_wait_owner[64] = {};   // per device: the model whose host-mapped buffer the watchdog symbol points to

struct sb2_model {
  int device = 0;
  int n_sm = 0;
  Switches sw;
  // synth3 (weights as the TMEM operand): raw SFH bin masses, tile-blocked, and the two metallicity factors
  double *sf = nullptr, *s0 = nullptr, *s1 = nullptr;
  CUtensorMap tm_g96_hi, tm_g96_lo;
  bool s3_ok = false;
  bool last_fused = false;   // the last synth3 launch wrote the fluxes itself (no finalize_kernel needed)
  sb2_model_desc d{};  // dims and scalars (pointers inside are NOT valid after create)
  long long cap = 0, cap_pad = 0;
  // model tables
  double *ages = nullptr, *edges = nullptr, *zmet = nullptr, *log10zmet = nullptr;
  float *gt_hi = nullptr, *gt_lo = nullptr, *kappa = nullptr, *filt_uv = nullptr;
  float* gt_x = nullptr;   // synth3_kernel's bfloat16 operand of the split product's small terms (cross_pack_kernel)
  float *dust_d0 = nullptr, *dust_l2 = nullptr, *g_slope = nullptr, *g_ampl = nullptr;   // per-galaxy dust shape (optional)
  double* lya_line = nullptr;   // per-galaxy Lyman-alpha escape (optional)
  float* g_lya = nullptr;
  float *kappa_birth = nullptr, *g_taub = nullptr;   // second dust screen (optional)
  float *dust_wnu = nullptr, *dust_g = nullptr, *dust_duv = nullptr, *e_part = nullptr;   // dust emission (optional)
  int *filt_lo = nullptr, *filt_hi = nullptr;
  double *bin_pow = nullptr, *thr = nullptr, *pre = nullptr;
  int *nline = nullptr, *lc_on = nullptr;
  double *dc = nullptr, *ddc = nullptr, *age = nullptr, *dage = nullptr;
  double *fm_log = nullptr, *fm_exp = nullptr, *fm_tail = nullptr;   // weight builder's special-function tables (optional)
  std::vector<int> h_lo, h_hi, h_off;
  std::vector<float> h_su, h_sdv;
  // workspace
  float *w_hi = nullptr, *w_lo = nullptr, *igm = nullptr;
  int *g_m = nullptr, *g_orig = nullptr, *perm = nullptr, *idx = nullptr;
  float *g_beta = nullptr, *g_gamma = nullptr, *g_taut = nullptr, *g_scale = nullptr, *g_ca = nullptr, *g_cb = nullptr;
  unsigned *keys = nullptr, *keys_sorted = nullptr;
  int *perm_pad = nullptr, *grp = nullptr, *tile_k0 = nullptr;  // grp: counts[64] | cum[64] | pad_start[64] | n_tiles
  int uv_len = 0;      // entries of the padded (U, V) tables
  int n_blue_pad = 0;  // IGM rows per tile (n_blue rounded up to 32; the extra rows hold 1)
  int4* tile_range = nullptr;
  float2* part = nullptr;  // [2][n_filt][cap_pad] partial filter numerators of the two epilogue groups
  int wd_stride = 0;  // floats per weights row in DeltaConstant (bracket-grouped) mode; 0: mode unavailable
  CUtensorMap tm_wd_hi, tm_wd_lo, tm_g2_hi, tm_g2_lo, tm_g160_hi, tm_g160_lo;  // g160: 160-row boxes (three-accumulator variant)
  size_t smem160_bytes = 0, smem_optin = 0;  // g2: 128-row boxes (one CTA's half of a chunk)
  size_t smem2_bytes = 0;
  double* g_mscale = nullptr;
  double* zpow = nullptr;
  unsigned* g_trunc = nullptr;
  void* cub_tmp = nullptr;
  size_t cub_bytes = 0;
  // host-entry staging (device side)
  // two staging slots, so the copies of one batch overlap the kernels of the next (sb2_synth_photometry_host_submit)
  double* stage_params[2] = {nullptr, nullptr};  // redshift | log_mass | tau_v | zd_value | zd_sigma | ca | cb | sfh rows
  float* stage_params32[2] = {nullptr, nullptr}; // the same arrays as float32 (host_f32 transport), widened on the device
  float* stage_spec[2] = {nullptr, nullptr};     // full-wavelength output of one slice (host entry with spec_out), allocated on first use
  long long stage_spec_rows = 0;
  cudaEvent_t ev_spec[2] = {nullptr, nullptr};   // slice's spectra are on the host
  float* stage_flux[2] = {nullptr, nullptr};
  double* stage_flux64[2] = {nullptr, nullptr};
  cudaEvent_t ev_slot[2] = {nullptr, nullptr};   // slot's results are on the host
  bool slot_busy[2] = {false, false};
  CUtensorMap tm_w_hi, tm_w_lo, tm_g_hi, tm_g_lo;
  CUtensorMap tm_g2_x, tm_g96_x, tm_g160_x, tm_g_x;
  size_t smem_bytes = 0;
  // host entry point: copy-in / compute / copy-out streams and per-slice events (slices are pipelined)
  cudaStream_t st_h2d = nullptr, st_comp = nullptr, st_d2h = nullptr;
  cudaEvent_t ev_in[8] = {}, ev_done[8] = {};
  unsigned int* wait_dbg = nullptr;  // host-mapped: who timed out in an mbarrier wait (protocol-bug watchdog)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // start | sorted | weights built | synthesised
  bool ev_valid = false;
};

extern "C" {

const char* sb2_last_error(void) { return g_err.c_str(); }

/* Watchdog record of the most recent mbarrier-wait timeout (diagnostics; empty string if none). */
const char* sb2_wait_debug(sb2_model* m) {
  static thread_local std::string out;
  out.clear();
  if (!m || !m->wait_dbg) return out.c_str();
  const unsigned n = m->wait_dbg[0] < 63u ? m->wait_dbg[0] : 63u;
  for (unsigned i = 0; i < n; ++i) {
    const unsigned int* e = m->wait_dbg + 4 + i * 4;
    char buf[128];
    std::snprintf(buf, sizeof buf, "id=0x%x block=%u thread=%u parity=%u; ", e[0], e[1], e[2], e[3]);
    out += buf;
  }
  return out.c_str();
}

long long sb2_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

int sb2_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int sb2_model_destroy(sb2_model* m) {
  if (!m) return SB2_OK;
  cudaSetDevice(m->device);
  void* ptrs[] = {m->ages, m->edges, m->zmet, m->log10zmet, m->gt_hi, m->gt_lo, m->gt_x, m->kappa, m->dust_d0, m->dust_l2, m->g_slope, m->g_ampl, m->lya_line, m->g_lya, m->kappa_birth, m->g_taub, m->dust_wnu, m->dust_g, m->dust_duv, m->e_part, m->filt_uv, m->filt_lo,
                  m->filt_hi, m->bin_pow, m->thr, m->pre, m->nline, m->lc_on, m->dc, m->ddc, m->age, m->dage, m->fm_log, m->fm_exp, m->fm_tail,
                  m->w_hi, m->w_lo, m->igm, m->g_m, m->g_orig, m->perm, m->idx, m->g_beta, m->g_gamma, m->g_taut, m->g_scale,
                  m->g_ca, m->g_cb, m->keys, m->keys_sorted, m->perm_pad, m->grp, m->tile_k0, m->tile_range, m->part, m->g_mscale, m->zpow, m->g_trunc, m->cub_tmp, m->stage_params[0],
                  m->stage_params[1], m->stage_flux[0], m->stage_flux[1], m->stage_flux64[0], m->stage_flux64[1], m->sf, m->s0, m->s1,
                  m->stage_params32[0], m->stage_params32[1], m->stage_spec[0], m->stage_spec[1]};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (cudaEvent_t e : m->ev)
    if (e) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (m->ev_slot[i]) cudaEventDestroy(m->ev_slot[i]);
    if (m->ev_spec[i]) cudaEventDestroy(m->ev_spec[i]);
  }
  for (int i = 0; i < 8; ++i) {
    if (m->ev_in[i]) cudaEventDestroy(m->ev_in[i]);
    if (m->ev_done[i]) cudaEventDestroy(m->ev_done[i]);
  }
  if (m->st_h2d) cudaStreamDestroy(m->st_h2d);
  if (m->st_comp) cudaStreamDestroy(m->st_comp);
  if (m->st_d2h) cudaStreamDestroy(m->st_d2h);
  if (m->wait_dbg) {
    if (m->device >= 0 && m->device < 64 && g_wait_owner[m->device] == m) {   // do not leave the symbol dangling
      unsigned int* none = nullptr;
      cudaMemcpyToSymbol(sb2::g_wait_dbg, &none, sizeof(none));
      g_wait_owner[m->device] = nullptr;
    }
    cudaFreeHost(m->wait_dbg);
  }
  delete m;
  return SB2_OK;
}

int sb2_model_create(const sb2_model_desc* d, int device, sb2_model** out) {
  if (!d || !out) return fail(SB2_ERR_INVALID, "null argument");
  *out = nullptr;
  if (d->n_comp != 1 && d->n_comp != 2) return fail(SB2_ERR_INVALID, "n_comp must be 1 or 2");
  if (d->n_filt < 1 || d->n_filt > sb2::kMaxFilt) return fail(SB2_ERR_INVALID, "n_filt must be in [1, 32]");
  if (d->n_age_pad < d->n_age || d->n_age_pad % 4 != 0) return fail(SB2_ERR_INVALID, "n_age_pad must be a multiple of 4 >= n_age");
  if (d->k_pad % 32 != 0 || d->k_pad < d->n_age_pad * d->n_z) return fail(SB2_ERR_INVALID, "bad k_pad");
  const int lch = sb2::kBN / d->n_comp;
  if (d->x_bins < 0 || (d->x_bins > 0 && (!d->dust_wnu || d->x_bins % 192 != 0 || d->x_bin0 % 192 != 0 || d->x_bin0 < d->n_lam)))
    return fail(SB2_ERR_INVALID, "pseudo-bins need dust_wnu; x_bin0 >= n_lam and x_bins must be multiples of 192");
  if (d->n_chunk != ((d->x_bins > 0 ? d->x_bin0 + d->x_bins : d->n_lam) + lch - 1) / lch) return fail(SB2_ERR_INVALID, "bad n_chunk");
  if (d->max_batch < 1) return fail(SB2_ERR_INVALID, "max_batch must be positive");
  if (d->rest_frame && d->igm_bin_pow) return fail(SB2_ERR_INVALID, "rest_frame models take no IGM tables");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(SB2_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(SB2_ERR_CUDA, "device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                                  "; this library carries sm_100a (B200) code only");
  sb2_model* m = new sb2_model();
  m->device = device;
  m->n_sm = prop.multiProcessorCount;
  m->d = *d;
  m->sw.read();
  int rc = SB2_OK;
#define UP(dst, src, n) if ((rc = upload(&m->dst, src, (size_t)(n))) != SB2_OK) { sb2_model_destroy(m); return rc; }
  // ages and bin edges (A2): e_0 = 0, e_{i+1} = (t_i + t_{i+1})/2
  std::vector<double> ages(d->n_age), edges(d->n_age), lz(d->n_z);
  for (int i = 0; i < d->n_age; ++i) ages[i] = std::pow(10.0, d->log10ages[i]);
  edges[0] = 0.0;
  for (int i = 0; i + 1 < d->n_age; ++i) edges[i + 1] = 0.5 * (ages[i] + ages[i + 1]);
  for (int i = 0; i < d->n_z; ++i) lz[i] = std::log10(d->metallicities[i]);
  UP(ages, ages.data(), d->n_age);
  UP(edges, edges.data(), d->n_age);
  UP(zmet, d->metallicities, d->n_z);
  UP(log10zmet, lz.data(), d->n_z);
  const size_t g_elems = (size_t)d->n_chunk * sb2::kBN * d->k_pad;
  UP(gt_hi, d->gt_hi, g_elems);
  UP(gt_lo, d->gt_lo, g_elems);
  if (cudaMalloc(&m->gt_x, g_elems * sizeof(float)) != cudaSuccess) { sb2_model_destroy(m); return fail(SB2_ERR_CUDA, "out of device memory (gt_x)"); }
  cross_pack_kernel<<<1024, 256>>>(m->gt_hi, m->gt_lo, reinterpret_cast<uint32_t*>(m->gt_x), g_elems / 8);
  if (cudaDeviceSynchronize() != cudaSuccess) { sb2_model_destroy(m); return fail(SB2_ERR_CUDA, "cross_pack_kernel failed"); }
  {
    std::vector<float> kap((size_t)d->n_chunk * lch, 0.f);
    if (d->kappa) std::memcpy(kap.data(), d->kappa, kap.size() * sizeof(float));
    UP(kappa, kap.data(), kap.size());
    if (d->kappa_birth) {
      if (!d->kappa || d->n_comp != 2) { sb2_model_destroy(m); return fail(SB2_ERR_INVALID, "kappa_birth needs kappa and n_comp = 2"); }
      std::vector<float> tb(kap.size(), 0.f);
      std::memcpy(tb.data(), d->kappa_birth, tb.size() * sizeof(float));
      UP(kappa_birth, tb.data(), tb.size());
    }
    if (d->dust_wnu) {
      if (!d->kappa || !d->dust_g || !d->dust_duv || d->dust_m_len < 1) {
        sb2_model_destroy(m);
        return fail(SB2_ERR_INVALID, "dust_wnu needs kappa, dust_g, dust_duv and dust_m_len");
      }
      UP(dust_wnu, d->dust_wnu, kap.size());
      UP(dust_g, d->dust_g, (size_t)d->n_lam);
      UP(dust_duv, d->dust_duv, (size_t)d->dust_m_len * d->n_filt * 2);
    }
    if (d->dust_d0 && d->dust_l2) {
      if (!d->kappa) { sb2_model_destroy(m); return fail(SB2_ERR_INVALID, "dust_d0/dust_l2 need kappa"); }
      std::vector<float> t0(kap.size(), 0.f), t1(kap.size(), 0.f);
      std::memcpy(t0.data(), d->dust_d0, t0.size() * sizeof(float));
      std::memcpy(t1.data(), d->dust_l2, t1.size() * sizeof(float));
      UP(dust_d0, t0.data(), t0.size());
      UP(dust_l2, t1.data(), t1.size());
    }
  }
  {
    // (U, V) tables with kUvPad zero entries on both sides of every filter, so the epilogue's shifted reads
    // need no index clamps (synth_kernel.cuh, fast path)
    std::vector<float> uvp;
    for (int f = 0; f < d->n_filt; ++f) {
      const int len = d->filt_hi[f] - d->filt_lo[f] + 4;
      if (len < 4 || d->filt_off[f] < 0 || d->filt_off[f] + len > d->filt_uv_len) {
        sb2_model_destroy(m);
        return fail(SB2_ERR_INVALID, "filter table offsets inconsistent with filt_lo/filt_hi");
      }
      m->h_off.push_back((int)(uvp.size() / 2));
      uvp.insert(uvp.end(), (size_t)2 * sb2::kUvPad, 0.f);
      uvp.insert(uvp.end(), d->filt_uv + (size_t)2 * d->filt_off[f], d->filt_uv + (size_t)2 * (d->filt_off[f] + len));
      uvp.insert(uvp.end(), (size_t)2 * sb2::kUvPad, 0.f);
    }
    m->uv_len = (int)(uvp.size() / 2);
    UP(filt_uv, uvp.data(), uvp.size());
  }
  UP(filt_lo, d->filt_lo, d->n_filt);
  UP(filt_hi, d->filt_hi, d->n_filt);
  m->h_lo.assign(d->filt_lo, d->filt_lo + d->n_filt);
  m->h_hi.assign(d->filt_hi, d->filt_hi + d->n_filt);
  for (int f = 0; f < d->n_filt; ++f) {
    m->h_su.push_back((float)d->filt_su[f]);
    m->h_sdv.push_back((float)d->filt_sdv[f]);
  }
  if (d->igm_bin_pow && d->n_blue > 0) {
    UP(bin_pow, d->igm_bin_pow, (size_t)8 * d->n_blue);
    UP(nline, d->igm_nline, d->n_blue);
    UP(lc_on, d->igm_lc_on, d->n_blue);
    UP(thr, d->igm_thr, 3 * 64);
    UP(pre, d->igm_pre, (size_t)5 * (d->n_lines + 1));
  } else {
    m->d.n_blue = 0;
  }
  m->n_blue_pad = (m->d.n_blue + 31) / 32 * 32;
  UP(dc, d->cosmo_dc, d->cosmo_n + 1);
  UP(ddc, d->cosmo_ddc, d->cosmo_n + 1);
  UP(age, d->cosmo_age, d->cosmo_n + 1);
  UP(dage, d->cosmo_dage, d->cosmo_n + 1);
  if (d->lya_line) {
    if (d->lya_bin < 0 || d->lya_bin >= d->n_lam) { sb2_model_destroy(m); return fail(SB2_ERR_INVALID, "lya_bin out of range"); }
    UP(lya_line, d->lya_line, (size_t)d->n_age * d->n_z);
  }
  if (d->fm_log_tab && d->fm_exp_tab && d->fm_tail_tab && d->fm_tail_n > 0 && d->fm_tail_w > 0.0) {
    UP(fm_log, d->fm_log_tab, 512);
    UP(fm_exp, d->fm_exp_tab, 64);
    UP(fm_tail, d->fm_tail_tab, (size_t)8 * d->fm_tail_n);
  }
#undef UP
  // workspace
  m->cap = d->max_batch;
  m->cap_pad = (m->cap + 255) / 256 * 256 + 256 * (long long)(d->n_z > 1 ? d->n_z - 1 : 1);
  // bracket-grouped (DeltaConstant) mode: a tile's grid columns start at bracket*n_age_pad -- TMA needs that
  // start 16-byte aligned, hence n_age_pad % 4 == 0 -- and span two metallicities
  m->wd_stride = (d->n_z >= 2 && d->n_z <= kMaxGroups) ? 2 * d->n_age_pad : 0;
  const size_t np = (size_t)m->cap_pad;
#define AL(ptr, bytes) { cudaError_t e_ = cudaMalloc(reinterpret_cast<void**>(&m->ptr), (bytes)); \
    if (e_ != cudaSuccess) { sb2_model_destroy(m); return fail(SB2_ERR_CUDA, std::string("cudaMalloc " #ptr ": ") + cudaGetErrorString(e_)); } }
  AL(w_hi, np * d->k_pad * 4);
  AL(w_lo, np * d->k_pad * 4);
  AL(igm, (np / 128) * (size_t)(m->n_blue_pad > 0 ? m->n_blue_pad : 1) * 128 * 4);
  AL(tile_range, (np / 128) * sizeof(int4));
  AL(part, (size_t)sb2::kMaxGroups * d->n_filt * np * sizeof(float2));
  AL(g_m, np * 4); AL(g_orig, np * 4); AL(perm, np * 4); AL(idx, np * 4);
  if (m->dust_d0) { AL(g_slope, np * 4); AL(g_ampl, np * 4); }
  if (m->lya_line) { AL(g_lya, np * 4); }
  if (m->kappa_birth) { AL(g_taub, np * 4); }
  if (m->dust_wnu) { AL(e_part, (size_t)sb2::kMaxGroups * np * 4); }
  AL(g_beta, np * 4); AL(g_gamma, np * 4); AL(g_taut, np * 4); AL(g_scale, np * 4); AL(g_ca, np * 4); AL(g_cb, np * 4);
  AL(keys, np * 4); AL(keys_sorted, np * 4); AL(perm_pad, np * 4); AL(tile_k0, (np / 128) * 4); AL(grp, (3 * kMaxGroups + 1) * 4);
  AL(g_mscale, np * 8); AL(g_trunc, np * 4); AL(zpow, np * 13 * 8);
  // synth3: bracket-grouped batches whose two metallicities' columns fit the TMEM weights region
  m->s3_ok = m->wd_stride > 0 && m->wd_stride <= sb2::kS3WCols && d->n_age <= 64 && d->n_age_pad % 8 == 0 && !m->sw.no_synth3;
  if (m->s3_ok) {   // ... and whose per-wavelength tables leave room for the operand ring in shared memory (long native axes do not)
    const int c = d->n_comp, kap_len = d->n_chunk * (sb2::kBN / c), uvl = (int)(d->filt_uv_len + 2 * sb2::kUvPad * d->n_filt);
    const int feat = (d->dust_d0 ? sb2::kFeatDustShape : 0) | (d->kappa_birth ? sb2::kFeatTwoScreens : 0) | (d->dust_wnu ? sb2::kFeatAbsorbed : 0);
    // (the least launch_synth3_t can run with: two ring stages and no transpose tiles for spectra.  Asking for more here --
    //  three stages AND the spectra tiles -- sent every two-component model to the round-1 kernel: 4.6 -> 5.6 ms per 1M.)
    if (sb2::synth3_smem_bytes(c == 1 ? 96 : 128, 2, d->n_age, uvl, kap_len, false, feat) > (size_t)prop.sharedMemPerBlockOptin) m->s3_ok = false;
  }
  if (m->s3_ok) { AL(sf, (np / 128) * (size_t)d->n_age * 128 * 8); AL(s0, np * 8); AL(s1, np * 8); }
  m->cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, m->cub_bytes, m->keys, m->keys_sorted, m->idx, m->perm, (int)m->cap);
  AL(cub_tmp, m->cub_bytes + 16);
  for (int sl = 0; sl < 2; ++sl) {
    AL(stage_params[sl], (size_t)m->cap * (11 + SB2_SFH_ROW) * 8);
    AL(stage_params32[sl], (size_t)m->cap * (11 + SB2_SFH_ROW) * 4);
    AL(stage_flux[sl], (size_t)m->cap * d->n_filt * 4);
    AL(stage_flux64[sl], (size_t)m->cap * d->n_filt * 8);
  }
#undef AL
  if ((rc = make_tmap(&m->tm_w_hi, m->w_hi, np, d->k_pad, sb2::kBM)) != SB2_OK ||
      (rc = make_tmap(&m->tm_w_lo, m->w_lo, np, d->k_pad, sb2::kBM)) != SB2_OK ||
      (m->wd_stride && (rc = make_tmap(&m->tm_wd_hi, m->w_hi, np, m->wd_stride, sb2::kBM)) != SB2_OK) ||
      (m->wd_stride && (rc = make_tmap(&m->tm_wd_lo, m->w_lo, np, m->wd_stride, sb2::kBM)) != SB2_OK) ||
      (rc = make_tmap(&m->tm_g2_hi, m->gt_hi, (uint64_t)d->n_chunk * sb2::kBN, d->k_pad, sb2::kBN2 / 2)) != SB2_OK ||
      (rc = make_tmap(&m->tm_g2_lo, m->gt_lo, (uint64_t)d->n_chunk * sb2::kBN, d->k_pad, sb2::kBN2 / 2)) != SB2_OK ||
      (rc = make_tmap(&m->tm_g160_hi, m->gt_hi, (uint64_t)d->n_chunk * sb2::kBN, d->k_pad, 160)) != SB2_OK ||
      (rc = make_tmap(&m->tm_g160_lo, m->gt_lo, (uint64_t)d->n_chunk * sb2::kBN, d->k_pad, 160)) != SB2_OK ||
      (rc = make_tmap(&m->tm_g96_hi, m->gt_hi, (uint64_t)d->n_chunk * sb2::kBN, d->k_pad, 96)) != SB2_OK ||
      (rc = make_tmap(&m->tm_g96_lo, m->gt_lo, (uint64_t)d->n_chunk * sb2::kBN, d->k_pad, 96)) != SB2_OK ||
      (rc = make_tmap(&m->tm_g2_x, m->gt_x, (uint64_t)d->n_chunk * sb2::kBN, d->k_pad, sb2::kBN2 / 2)) != SB2_OK ||
      (rc = make_tmap(&m->tm_g96_x, m->gt_x, (uint64_t)d->n_chunk * sb2::kBN, d->k_pad, 96)) != SB2_OK ||
      (rc = make_tmap(&m->tm_g160_x, m->gt_x, (uint64_t)d->n_chunk * sb2::kBN, d->k_pad, 160)) != SB2_OK ||
      (rc = make_tmap(&m->tm_g_x, m->gt_x, (uint64_t)d->n_chunk * sb2::kBN, d->k_pad, sb2::kBN)) != SB2_OK ||
      (rc = make_tmap(&m->tm_g_hi, m->gt_hi, (uint64_t)d->n_chunk * sb2::kBN, d->k_pad, sb2::kBN)) != SB2_OK ||
      (rc = make_tmap(&m->tm_g_lo, m->gt_lo, (uint64_t)d->n_chunk * sb2::kBN, d->k_pad, sb2::kBN)) != SB2_OK) {
    sb2_model_destroy(m);
    return rc;
  }
  m->smem_bytes = 1024 + (size_t)sb2::kStages * sb2::kStageBytes + (((size_t)m->uv_len * 8 + 15) & ~size_t(15)) + sb2::kBarBytes;
  m->smem_optin = (size_t)prop.sharedMemPerBlockOptin;
  m->smem160_bytes = 1024 + (size_t)sb2::kStages * sb2::SynthCfg<160>::kStageBytesN + (((size_t)m->uv_len * 8 + 15) & ~size_t(15)) + sb2::kBarBytes;
  m->smem2_bytes = 1024 + (size_t)sb2::kW2Bytes + (size_t)sb2::kG2Slots * sb2::kG2Slot + (((size_t)m->uv_len * 8 + 15) & ~size_t(15)) + sb2::kBarBytes;
  if (m->smem2_bytes > (size_t)prop.sharedMemPerBlockOptin || m->wd_stride > sb2::kW2Kb * sb2::kBK || (m->n_sm & 1)) m->smem2_bytes = 0;  // CTA-pair kernel unavailable
  {
    // The filter tables share the SM's shared memory with a kernel's operand ring.  The variants differ in what they need
    // (256-column chunks most, synth3_kernel least): a model is usable if ANY of them fits; a batch that needs one that does
    // not is refused at launch (launch_synth_t), not here -- 28 filters fit synth3_kernel but not the 256-column kernel.
    const size_t uvb = ((size_t)m->uv_len * 8 + 15) & ~size_t(15);
    const size_t split_bytes = 1024 + (size_t)sb2::kStages * sb2::SynthCfg<128>::kStageBytesN + uvb + sb2::kBarBytes;
    const bool any = m->smem_bytes <= m->smem_optin || (d->n_comp == 1 && m->smem160_bytes <= m->smem_optin) ||
                     split_bytes <= m->smem_optin || m->s3_ok;
    if (!any) {
      sb2_model_destroy(m);
      return fail(SB2_ERR_INVALID, "filter tables do not fit in shared memory next to the operand pipeline (" +
                                       std::to_string(std::min(m->smem_bytes, split_bytes)) + " B needed)");
    }
  }
  if (cudaHostAlloc(reinterpret_cast<void**>(&m->wait_dbg), 260 * sizeof(unsigned int), cudaHostAllocMapped) == cudaSuccess) {
    std::memset(m->wait_dbg, 0, 260 * sizeof(unsigned int));
    unsigned int* dptr = nullptr;
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&dptr), m->wait_dbg, 0) == cudaSuccess &&
        cudaMemcpyToSymbol(sb2::g_wait_dbg, &dptr, sizeof(dptr)) == cudaSuccess && device >= 0 && device < 64)
      g_wait_owner[device] = m;     // (one record buffer per device: the most recently created model's)
  }
  bool ok = cudaStreamCreateWithFlags(&m->st_h2d, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&m->st_comp, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&m->st_d2h, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 8 && ok; ++i)
    ok = cudaEventCreateWithFlags(&m->ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&m->ev_done[i], cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < 2 && ok; ++i) ok = cudaEventCreateWithFlags(&m->ev_slot[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    sb2_model_destroy(m);
    return fail(SB2_ERR_CUDA, "stream / event creation failed");
  }
  for (int i = 0; i < 4; ++i) {
    if (cudaEventCreate(&m->ev[i]) != cudaSuccess) {
      sb2_model_destroy(m);
      return fail(SB2_ERR_CUDA, "cudaEventCreate failed");
    }
  }
  *out = m;
  return SB2_OK;
}

}  // extern "C"

namespace {

// One spectral component: chunks of 160 columns and three TMEM accumulators (unless SB2_N256 asks for the two-accumulator form)
bool use_n160(const sb2_model* m) { return m->d.n_comp == 1 && !m->sw.n256; }

bool delta_mode(const sb2_model* m, const sb2_params* p) {
  return m->wd_stride > 0 && (p->zd_type == SB2_ZD_DELTA_LINEAR || p->zd_type == SB2_ZD_DELTA_LOG10);
}

// Per-galaxy emission extras of a model as a synth3_kernel feature set (sb2::kFeat*).
int model_features(const sb2_model* m) {
  return (m->dust_d0 ? sb2::kFeatDustShape : 0) | (m->lya_line ? sb2::kFeatLya : 0) | (m->kappa_birth ? sb2::kFeatTwoScreens : 0) |
         (m->dust_wnu ? sb2::kFeatAbsorbed : 0);
}
// feature sets synth3_kernel is instantiated for (what the reference's scripts combine): none | Lya | dust shape (+ Lya) |
// two screens | absorbed energy | two screens + absorbed | dust shape + Lya + absorbed (the production script's `total`)
bool s3_has_features(int f) {
  using namespace sb2;
  return f == 0 || f == kFeatLya || f == kFeatDustShape || f == (kFeatDustShape | kFeatLya) || f == kFeatTwoScreens ||
         f == kFeatAbsorbed || f == (kFeatTwoScreens | kFeatAbsorbed) || f == (kFeatDustShape | kFeatLya | kFeatAbsorbed) ||
         f == (kFeatLya | kFeatAbsorbed);
}
// Bracket-grouped batches take synth3_kernel (weights as the TMEM operand).
bool use_s3(const sb2_model* m, const sb2_params* p) {
  return m->s3_ok && delta_mode(m, p) && s3_has_features(model_features(m)) && !m->sw.cta_pair;
}
int s3_cols(const sb2_model* m) { return m->d.n_comp == 1 ? 96 : 128; }   // accumulator columns per chunk (3 x 96 | 2 x 128 beside the weights)

// Dense K (weights over every (age, Z) bin): the hi*hi terms of a chunk go to kDenseSplit accumulators of 128 columns by K
// range, which divides the truncation error of the tensor core's FP32 accumulation by as much (synth_kernel, kSplit).
constexpr int kDenseSplit = 3;
// Bracket-grouped batches of grids with many ages (BC03: 221) do not fit synth3_kernel's tensor-memory weights and have a
// long K too (2 x 224 columns): they take the same split -- 7.6e-6 -> see DESIGN 6.1 -- with the bracket's weight maps.
constexpr int kDeltaSplitMinK8 = 24;      // K >= 192 columns: below that one accumulator is already at ~3e-6
// The two small terms of the split product as ONE bfloat16 MMA (Synth3Args.cross | SynthArgs.cross + PrepModel.cross).
//   synth3_kernel: the default for launches that output photometry only -- the band integrals average the per-wavelength
//   rounding of the bfloat16 factors; launches that also write spectra keep three TF32 passes (per-wavelength values see the
//   worst case, 2^-17 of a product).  SB2_TF32X3=1 restores three TF32 passes everywhere.
//   synth_kernel (dense K and the fallbacks): opt-in with SB2_BF16_DENSE=1 -- the dense configuration's margin (<= 3e-6) is
//   worth more than the 3 % it gains there.
bool cross_s3(const sb2_model* m) { return !m->sw.tf32x3 && !m->sw.cta_pair; }
bool cross_mode(const sb2_model* m) { return m->sw.bf16_dense && !m->sw.tf32x3 && !m->sw.cta_pair; }
bool use_split(const sb2_model* m, bool delta) {
  if (m->sw.no_split) return false;
  if (delta) return m->wd_stride / 8 >= kDeltaSplitMinK8;
  return (m->d.k_pad / 8 + 3) / 4 >= kDenseSplit;   // every accumulator gets a k-block
}


template <int C, int NF, bool SPEC, bool PG>
int launch_synth_t(sb2_model* m, const sb2::SynthArgs& a, int grid, bool delta, cudaStream_t st) {
  if (use_split(m, delta)) {
    auto k = sb2::synth_kernel<C, NF, SPEC, 128, PG, kDenseSplit>;
    sb2::SynthArgs a2 = a;
    const size_t fixed = 1024 + (((size_t)m->uv_len * 8 + 15) & ~size_t(15)) + sb2::kBarBytes;
    size_t bytes = fixed + (size_t)sb2::kStages * sb2::SynthCfg<128>::kStageBytesN;
    if (bytes > m->smem_optin) return fail(SB2_ERR_INVALID, "filter tables do not fit in shared memory next to this batch's kernel (dense K)");
    if (SPEC && bytes + sb2::kSpecSmemBytes <= m->smem_optin) {
      bytes += sb2::kSpecSmemBytes;
      a2.spec_smem = 1;
    }
    if (a.cross && a.two_pass) {
      // half stages (one W tile + one G tile, 32 KB): as deep a ring as fits, at least the four that the full stages' space holds
      const size_t half_stage = (size_t)sb2::kABytes + sb2::SynthCfg<128>::kBBytesN, extra = a2.spec_smem ? sb2::kSpecSmemBytes : 0;
      int ns = 8;
      while (ns > 4 && fixed + extra + ns * half_stage > m->smem_optin) --ns;
      a2.n_stages = ns;
      bytes = fixed + extra + ns * half_stage;
    }
    CU_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    k<<<grid, sb2::kSynthThreads, bytes, st>>>(delta ? m->tm_wd_hi : m->tm_w_hi, delta ? m->tm_wd_lo : m->tm_w_lo, m->tm_g2_hi,
                                               a.cross ? m->tm_g2_x : m->tm_g2_lo, a2);
    STAGE_CHECK("synth_kernel (split accumulators)", st);
    return SB2_OK;
  }
  if constexpr (C == 1) {
    if (use_n160(m)) {
      auto k = sb2::synth_kernel<C, NF, SPEC, 160, PG>;
      sb2::SynthArgs a2 = a;
      size_t bytes = m->smem160_bytes;
      if (bytes > m->smem_optin) return fail(SB2_ERR_INVALID, "filter tables do not fit in shared memory next to this batch's kernel (160-column chunks)");
      if (SPEC && bytes + sb2::kSpecSmemBytes <= m->smem_optin) {   // room for the spectra transpose tiles
        bytes += sb2::kSpecSmemBytes;
        a2.spec_smem = 1;
      }
      CU_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      k<<<grid, sb2::kSynthThreads, bytes, st>>>(delta ? m->tm_wd_hi : m->tm_w_hi, delta ? m->tm_wd_lo : m->tm_w_lo,
                                                 m->tm_g160_hi, a.cross ? m->tm_g160_x : m->tm_g160_lo, a2);
      STAGE_CHECK("synth_kernel", st);
      return SB2_OK;
    }
  }
  auto k = sb2::synth_kernel<C, NF, SPEC, sb2::kBN, PG>;
  if (m->smem_bytes > m->smem_optin) return fail(SB2_ERR_INVALID, "filter tables do not fit in shared memory next to this batch's kernel (256-column chunks)");
  CU_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem_bytes));
  k<<<grid, sb2::kSynthThreads, m->smem_bytes, st>>>(delta ? m->tm_wd_hi : m->tm_w_hi, delta ? m->tm_wd_lo : m->tm_w_lo,
                                                     m->tm_g_hi, a.cross ? m->tm_g_x : m->tm_g_lo, a);
  STAGE_CHECK("synth_kernel", st);
  return SB2_OK;
}

template <int C, int NF, bool SPEC>
int launch_synth2_t(sb2_model* m, const sb2::SynthArgs& a, int grid, cudaStream_t st) {
  auto k = sb2::synth2_kernel<C, NF, SPEC>;
  CU_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem2_bytes));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(sb2::kSynthThreads);
  cfg.dynamicSmemBytes = m->smem2_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CU_TRY(cudaLaunchKernelEx(&cfg, k, m->tm_wd_hi, m->tm_wd_lo, m->tm_g2_hi, m->tm_g2_lo, a));
  STAGE_CHECK("synth2_kernel", st);
  return SB2_OK;
}

int launch_synth2(sb2_model* m, const sb2::SynthArgs& a, int grid, cudaStream_t st) {
  const int nf = m->d.n_filt, c = m->d.n_comp;
  const bool spec = a.out_spec != nullptr;
#define SB2_PICK(C, NF) (spec ? launch_synth2_t<C, NF, true>(m, a, grid, st) : launch_synth2_t<C, NF, false>(m, a, grid, st))
  if (c == 1) {
    if (nf <= 8) return SB2_PICK(1, 8);
    if (nf <= 24) return SB2_PICK(1, 24);
    return SB2_PICK(1, 32);
  }
  if (nf <= 8) return SB2_PICK(2, 8);
  if (nf <= 24) return SB2_PICK(2, 24);
  return SB2_PICK(2, 32);
#undef SB2_PICK
}

template <int C, int NF, bool SPEC, int N, int FEAT = 0>
int launch_synth3_t(sb2_model* m, const sb2::SynthArgs& a, int grid, cudaStream_t st) {
  auto k = sb2::synth3_kernel<C, NF, SPEC, N, FEAT>;
  sb2::SynthArgs a2 = a;
  sb2::Synth3Args x{};
  x.sf = m->sf; x.s0 = m->s0; x.s1 = m->s1; x.n_age = m->d.n_age; x.na_pad = m->d.n_age_pad; x.w_stride = m->wd_stride;
  x.kb_split = a.n_kb / 2;
  x.cross = (cross_s3(m) && a.out_spec == nullptr) ? 1 : 0;
  bool spec = false;
  const int kap_len = m->d.n_chunk * (sb2::kBN / C);
  const int feat_tab = a.x_count > 0 ? (FEAT & ~sb2::kFeatAbsorbed) : FEAT;    // pseudo-bins: the energy weights are 1, no table
  // fused output (SynthArgs.fuse_out): n_filt KB for the groups' exchange, if three ring stages still fit beside it
  int n_x = a.fuse_out ? m->d.n_filt : 0;
  if (n_x && sb2::synth3_smem_bytes(N, 3, x.n_age, m->uv_len, kap_len, false, feat_tab, n_x) > m->smem_optin) { n_x = 0; a2.fuse_out = 0; }
  m->last_fused = a2.fuse_out != 0;
  if (SPEC && sb2::synth3_smem_bytes(N, 3, x.n_age, m->uv_len, kap_len, true, feat_tab, n_x) <= m->smem_optin) { spec = true; a2.spec_smem = 1; }
  int ns = sb2::kS3MaxStages;
  while (ns > 1 && sb2::synth3_smem_bytes(N, ns, x.n_age, m->uv_len, kap_len, spec, feat_tab, n_x) > m->smem_optin) --ns;
  const size_t bytes = sb2::synth3_smem_bytes(N, ns, x.n_age, m->uv_len, kap_len, spec, feat_tab, n_x);
  if (ns < 2 || bytes > m->smem_optin) return fail(SB2_ERR_INVALID, "synth3_kernel: filter tables leave no room for the operand ring");
  x.n_stages = ns;
  CU_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  const CUtensorMap& th = C == 2 ? m->tm_g2_hi : m->tm_g96_hi;      // boxes of 64 rows (one component of a chunk) | 96 rows
  const CUtensorMap& tl = x.cross ? (C == 2 ? m->tm_g2_x : m->tm_g96_x) : (C == 2 ? m->tm_g2_lo : m->tm_g96_lo);
  k<<<grid, sb2::kS3Threads, bytes, st>>>(th, tl, a2, x);
  STAGE_CHECK("synth3_kernel", st);
  return SB2_OK;
}

// feature-set instantiations: one accumulator width (32 filters) to bound the number of kernels
template <int FEAT>
int launch_synth3_feat(sb2_model* m, const sb2::SynthArgs& a, int grid, cudaStream_t st) {
  const bool spec = a.out_spec != nullptr;
  if (m->d.n_comp == 1) {
    if constexpr ((FEAT & sb2::kFeatTwoScreens) != 0) return fail(SB2_ERR_INVALID, "two dust screens need two components");
    else return spec ? launch_synth3_t<1, 32, true, 96, FEAT>(m, a, grid, st) : launch_synth3_t<1, 32, false, 96, FEAT>(m, a, grid, st);
  }
  return spec ? launch_synth3_t<2, 32, true, 128, FEAT>(m, a, grid, st) : launch_synth3_t<2, 32, false, 128, FEAT>(m, a, grid, st);
}

int launch_synth3(sb2_model* m, const sb2::SynthArgs& a, int grid, cudaStream_t st) {
  const int nf = m->d.n_filt, c = m->d.n_comp;
  const bool spec = a.out_spec != nullptr;
  {
    using namespace sb2;
    switch (model_features(m)) {
      case 0: break;
      case kFeatLya: return launch_synth3_feat<kFeatLya>(m, a, grid, st);
      case kFeatDustShape: return launch_synth3_feat<kFeatDustShape>(m, a, grid, st);
      case kFeatDustShape | kFeatLya: return launch_synth3_feat<kFeatDustShape | kFeatLya>(m, a, grid, st);
      case kFeatTwoScreens: return launch_synth3_feat<kFeatTwoScreens>(m, a, grid, st);
      case kFeatAbsorbed: return launch_synth3_feat<kFeatAbsorbed>(m, a, grid, st);
      case kFeatTwoScreens | kFeatAbsorbed: return launch_synth3_feat<kFeatTwoScreens | kFeatAbsorbed>(m, a, grid, st);
      case kFeatLya | kFeatAbsorbed: return launch_synth3_feat<kFeatLya | kFeatAbsorbed>(m, a, grid, st);
      case kFeatDustShape | kFeatLya | kFeatAbsorbed: return launch_synth3_feat<kFeatDustShape | kFeatLya | kFeatAbsorbed>(m, a, grid, st);
      default: return fail(SB2_ERR_INVALID, "no synth3_kernel instantiation for this feature set");
    }
  }
#define SB2_PICK3(C, NF, N) (spec ? launch_synth3_t<C, NF, true, N>(m, a, grid, st) : launch_synth3_t<C, NF, false, N>(m, a, grid, st))
  if (c == 1) {
    if (nf <= 8) return SB2_PICK3(1, 8, 96);
    if (nf <= 24) return SB2_PICK3(1, 24, 96);
    return SB2_PICK3(1, 32, 96);
  }
  if (nf <= 8) return SB2_PICK3(2, 8, 128);
  if (nf <= 24) return SB2_PICK3(2, 24, 128);
  return SB2_PICK3(2, 32, 128);
#undef SB2_PICK3
}

int launch_synth(sb2_model* m, const sb2::SynthArgs& a, int grid, bool delta, cudaStream_t st) {
  const int nf = m->d.n_filt, c = m->d.n_comp;
  const bool spec = a.out_spec != nullptr;
  const bool pg = a.dust_d0 != nullptr || a.g_lya != nullptr || a.kappa_birth != nullptr || a.wnu != nullptr;   // per-galaxy emission extras: their own instantiation
#define SB2_PICK(C, NF)                                                                                             \
  (pg ? (spec ? launch_synth_t<C, NF, true, true>(m, a, grid, delta, st) : launch_synth_t<C, NF, false, true>(m, a, grid, delta, st)) \
      : (spec ? launch_synth_t<C, NF, true, false>(m, a, grid, delta, st) : launch_synth_t<C, NF, false, false>(m, a, grid, delta, st)))
  if (c == 1) {
    if (nf <= 8) return SB2_PICK(1, 8);
    if (nf <= 24) return SB2_PICK(1, 24);
    return SB2_PICK(1, 32);
  }
  if (nf <= 8) return SB2_PICK(2, 8);
  if (nf <= 24) return SB2_PICK(2, 24);
  return SB2_PICK(2, 32);
#undef SB2_PICK
}

sb2::PrepModel prep_model(const sb2_model* m) {
  sb2::PrepModel M{};
  const sb2_model_desc& d = m->d;
  M.n_age = d.n_age; M.n_z = d.n_z; M.na_pad = d.n_age_pad; M.K = d.n_age * d.n_z; M.k_pad = d.k_pad; M.n_lam = d.n_lam;
  M.n_filt = d.n_filt; M.n_blue = d.n_blue; M.n_lines = d.n_lines; M.variant = d.interp_variant;
  M.igm_on = d.n_blue > 0;
  M.rest_frame = d.rest_frame ? 1 : 0;
  M.ages = m->ages; M.edges = m->edges; M.zmet = m->zmet; M.log10zmet = m->log10zmet;
  M.lam0 = d.lam0; M.q = d.q; M.ln_q = std::log(d.q); M.grid_scale = d.grid_scale; M.base_mass = d.base_mass;
  M.filt_lo = m->filt_lo; M.filt_hi = m->filt_hi;
  M.bin_pow = m->bin_pow; M.nline = m->nline; M.lc_on = m->lc_on; M.thr = m->thr; M.pre = m->pre;
  M.cosmo_n = d.cosmo_n; M.cosmo_ds = d.cosmo_smax / d.cosmo_n;
  M.dc = m->dc; M.ddc = m->ddc; M.age = m->age; M.dage = m->dage; M.lya_line = m->lya_line;
  return M;
}

sb2::PrepParams prep_params(const sb2_params* p) {
  sb2::PrepParams P{};
  P.n = p->n; P.redshift = p->redshift; P.log_mass = p->log_mass; P.tau_v = p->tau_v;
  P.sfh_type = p->sfh_type; P.sfh_stride = p->sfh_stride; P.sfh_rows = p->sfh_rows;
  P.max_age_from_z = p->max_age_from_z; P.norm_mask = p->norm_mask; P.age_zmax_gyr = p->age_zmax_gyr;
  P.zd_type = p->zd_type; P.zd_value = p->zd_value; P.zd_sigma = p->zd_sigma;
  P.coef_att = p->coef_att; P.coef_unatt = p->coef_unatt; P.dust_slope = p->dust_slope; P.dust_ampl = p->dust_ampl; P.fesc_lya = p->fesc_lya; P.tau_v_birth = p->tau_v_birth;
  return P;
}

int check_params(const sb2_model* m, const sb2_params* p) {
  if (!m || !p) return fail(SB2_ERR_INVALID, "null argument");
  if (p->n < 1) return fail(SB2_ERR_INVALID, "empty batch");
  if (p->n > m->cap) return fail(SB2_ERR_CAPACITY, "batch of " + std::to_string(p->n) + " exceeds max_batch " + std::to_string(m->cap));
  if (!p->redshift || !p->sfh_rows || !p->zd_value) return fail(SB2_ERR_INVALID, "redshift, sfh_rows and zd_value are required");
  if (p->sfh_stride < 2 || p->sfh_stride > SB2_SFH_ROW) return fail(SB2_ERR_INVALID, "sfh_stride out of range");
  if (p->sfh_type < 0 || p->sfh_type > SB2_SFH_CONTINUITY)
    return fail(SB2_ERR_INVALID, "unsupported sfh_type");
  if (p->zd_type < 0 || p->zd_type > SB2_ZD_NORMAL_LOG10) return fail(SB2_ERR_INVALID, "bad zd_type");
  if (p->zd_type >= SB2_ZD_NORMAL_LINEAR && !p->zd_sigma) return fail(SB2_ERR_INVALID, "zd_sigma required for Normal");
  if ((p->dust_slope || p->dust_ampl) && !m->dust_d0)
    return fail(SB2_ERR_INVALID, "per-galaxy dust_slope / dust_ampl need a model created with dust_d0 and dust_l2");
  if (p->tau_v_birth && !m->kappa_birth) return fail(SB2_ERR_INVALID, "tau_v_birth needs a model created with kappa_birth");
  if (p->fesc_lya && !m->lya_line) return fail(SB2_ERR_INVALID, "per-galaxy fesc_lya needs a model created with lya_line");
  if (m->lya_line && !p->fesc_lya) return fail(SB2_ERR_INVALID, "this model reads fesc_lya per galaxy: params.fesc_lya is required");
  if (p->fesc_lya && (m->d.n_age > 64 || m->d.n_z > 64)) return fail(SB2_ERR_INVALID, "per-galaxy fesc_lya supports n_age, n_z <= 64");
  if (p->scaled_ld != 0 && p->scaled_ld < p->n) return fail(SB2_ERR_INVALID, "scaled_ld must be 0 or >= n");
  return SB2_OK;
}

int check_device_params(const sb2_model* m, const sb2_params* p) {
  int rc = check_params(m, p);
  if (rc == SB2_OK && p->host_f32) return fail(SB2_ERR_INVALID, "host_f32 is a host-entry transport option; device arrays are float64");
  return rc;
}

// A unit of work of the contraction kernel: one 128-galaxy tile, or a PAIR of tiles for the CTA-pair kernel
// (bracket-grouped batches).
int rows_per_unit_old(const sb2_model* m, bool delta) {
  // The CTA-pair kernel (synth2_kernel) is parity-tested but not yet faster than the single-CTA kernel on B200
  // (both sit on the same synchronisation/epilogue floor); it is opt-in: SB2_CTA_PAIR=1.
  return (delta && m->smem2_bytes > 0 && !m->dust_d0 && !m->lya_line && !m->kappa_birth && !m->dust_wnu && m->sw.cta_pair) ? 256 : 128;
}
int rows_per_unit(const sb2_model* m, const sb2_params* p, bool delta) {
  if (delta && use_s3(m, p)) return 128;
  return rows_per_unit_old(m, delta);
}
// Rows the grouped layout of a batch of n galaxies can occupy (every group is padded to whole units).
long long padded_rows(const sb2_model* m, const sb2_params* p, long long n, bool delta) {
  const long long rpu = rows_per_unit(m, p, delta);
  return (n + rpu - 1) / rpu * rpu + (delta ? rpu * (m->d.n_z - 1) : 0);
}

// group by (metallicity bracket, redshift) -> perm_pad ; prep kernel
int run_prep(sb2_model* m, const sb2_params* p, double* w_f64, bool sorted, bool all_lam, cudaStream_t st) {
  const long long n = p->n;
  const bool delta = sorted && delta_mode(m, p);
  const long long n_pad = padded_rows(m, p, n, delta);
  const int rpu = rows_per_unit(m, p, delta);
  const int* perm = nullptr;
  cudaEventRecord(m->ev[0], st);
  if (sorted) {
    const int tb = 256;
    const int n_groups = delta ? m->d.n_z - 1 : 1;
    const bool logz = p->zd_type == SB2_ZD_DELTA_LOG10;
    CU_TRY(cudaMemsetAsync(m->grp, 0, (3 * kMaxGroups + 1) * 4, st));
    CU_TRY(cudaMemsetAsync(m->perm_pad, 0xFF, (size_t)n_pad * 4, st));
    group_keys_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, st>>>(p->redshift, p->zd_value, logz ? m->log10zmet : m->zmet,
                                                                   m->d.n_z, delta ? 1 : 0, m->keys, m->idx, m->grp, n);
    STAGE_CHECK("group_keys_kernel", st);
    size_t bytes = m->cub_bytes;
    int key_bits = kKeyShift + 1;
    while ((1 << (key_bits - kKeyShift)) < n_groups) ++key_bits;
    CU_TRY(cub::DeviceRadixSort::SortPairs(m->cub_tmp, bytes, m->keys, m->keys_sorted, m->idx, m->perm, (int)n, 0, key_bits, st));
    STAGE_CHECK("radix sort", st);
    group_layout_kernel<<<1, 256, 0, st>>>(m->grp, n_groups, m->d.n_age_pad, rpu, m->grp + kMaxGroups, m->grp + 2 * kMaxGroups,
                                           m->tile_k0, m->grp + 3 * kMaxGroups, (int)(n_pad / rpu));
    STAGE_CHECK("group_layout_kernel", st);
    group_scatter_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, st>>>(m->keys_sorted, m->perm, m->grp + kMaxGroups,
                                                                      m->grp + 2 * kMaxGroups, m->perm_pad, n);
    STAGE_CHECK("group_scatter_kernel", st);
    perm = m->perm_pad;
  }
  cudaEventRecord(m->ev[1], st);
  sb2::PrepModel M = prep_model(m);
  M.delta = delta ? 1 : 0;
  M.w_stride = delta ? m->wd_stride : m->d.k_pad;
  M.cross = cross_mode(m) ? 1 : 0;
  sb2::PrepParams P = prep_params(p);
  sb2::PrepOut O{};
  O.w_hi = m->w_hi; O.w_lo = m->w_lo; O.w_f64 = w_f64; O.igm = m->igm; O.g_m = m->g_m; O.g_beta = m->g_beta; O.g_gamma = m->g_gamma;
  O.g_taut = m->g_taut; O.g_scale = m->g_scale; O.g_ca = m->g_ca; O.g_cb = m->g_cb; O.g_orig = m->g_orig;
  O.g_slope = m->g_slope; O.g_ampl = m->g_ampl; O.g_lya = m->g_lya; O.g_taub = m->g_taub;
  O.g_mscale = m->g_mscale; O.g_trunc = m->g_trunc; O.zpow = m->zpow;
  if (!w_f64) {  // the parity hook (sb2_build_weights) needs the weights only
    sb2::scalars_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, st>>>(M, P, O, perm, n_pad);
    STAGE_CHECK("scalars_kernel", st);
    const sb2_model_desc& d = m->d;
    int lo_min = m->h_lo[0], hi_max = m->h_hi[0];
    for (int f = 1; f < d.n_filt; ++f) { lo_min = std::min(lo_min, m->h_lo[f]); hi_max = std::max(hi_max, m->h_hi[f]); }
    const int wpb = 8, n_units = (int)(n_pad / rpu);
    const int cols = (delta && use_s3(m, p)) ? s3_cols(m) : (rpu == 256 || use_split(m, delta)) ? sb2::kBN2 : (use_n160(m) ? 160 : sb2::kBN);
    sb2::tile_range_kernel<<<(n_units + wpb - 1) / wpb, wpb * 32, 0, st>>>(m->g_m, m->g_orig, n_units, rpu, lo_min, hi_max, d.n_lam,
                                                                         cols / d.n_comp, all_lam ? 1 : 0, m->tile_range);
    STAGE_CHECK("tile_range_kernel", st);
  }
  const size_t sh = sb2::weights_smem_doubles(M.n_age, M.n_z) * sizeof(double);
  const unsigned blocks = (unsigned)((n_pad + sb2::kWGal - 1) / sb2::kWGal);
  sb2::FastMath F{};
  const bool fast = m->fm_tail && !m->sw.libm;
  if (fast) {
    F.log_tab = reinterpret_cast<const double2*>(m->fm_log); F.exp_tab = m->fm_exp;
    F.tail_tab = reinterpret_cast<const double2*>(m->fm_tail);
    F.tail_w = m->d.fm_tail_w; F.tail_inv_w = 1.0 / m->d.fm_tail_w; F.tail_n = m->d.fm_tail_n;
  }
  const bool dpl = p->sfh_type == SB2_SFH_DOUBLE_POWERLAW;
  if (use_s3(m, p) && (delta || w_f64)) {   // one galaxy per thread: raw bin masses, tile-blocked (expanded on chip by synth3_kernel)
    sb2::SfOut S{m->sf, m->s0, m->s1, m->g_lya};
    const unsigned blocks3 = (unsigned)((n_pad + sb2::kW3Threads - 1) / sb2::kW3Threads);
    const bool lya = p->fesc_lya != nullptr && !w_f64;
#define SB2_W3(FAST, MODE) (lya ? sb2::weights3_kernel<FAST, MODE, true><<<blocks3, sb2::kW3Threads, 0, st>>>(M, F, P, S, w_f64, perm, n_pad) \
                                : sb2::weights3_kernel<FAST, MODE, false><<<blocks3, sb2::kW3Threads, 0, st>>>(M, F, P, S, w_f64, perm, n_pad))
    if (dpl) SB2_W3(false, 2);
    else if (p->sfh_type == SB2_SFH_CONTINUITY) SB2_W3(false, 1);
    else if (fast) SB2_W3(true, 0);
    else SB2_W3(false, 0);
#undef SB2_W3
  } else if (M.n_age <= 64 && M.n_z <= 64 && !m->sw.weights_v1) {   // half-warp per galaxy
    const unsigned blocks2 = (unsigned)((n_pad + sb2::kW2Gal - 1) / sb2::kW2Gal);
    const bool lya = p->fesc_lya != nullptr && !w_f64;
    if (dpl && lya) sb2::weights2_kernel<false, true, true><<<blocks2, sb2::kW2Gal * 16, 0, st>>>(M, F, P, O, perm, n_pad);
    else if (dpl) sb2::weights2_kernel<false, true, false><<<blocks2, sb2::kW2Gal * 16, 0, st>>>(M, F, P, O, perm, n_pad);
    else if (fast && lya) sb2::weights2_kernel<true, false, true><<<blocks2, sb2::kW2Gal * 16, 0, st>>>(M, F, P, O, perm, n_pad);
    else if (fast) sb2::weights2_kernel<true, false, false><<<blocks2, sb2::kW2Gal * 16, 0, st>>>(M, F, P, O, perm, n_pad);
    else if (lya) sb2::weights2_kernel<false, false, true><<<blocks2, sb2::kW2Gal * 16, 0, st>>>(M, F, P, O, perm, n_pad);
    else sb2::weights2_kernel<false, false, false><<<blocks2, sb2::kW2Gal * 16, 0, st>>>(M, F, P, O, perm, n_pad);
  } else if (dpl) {
    sb2::weights_kernel<false, true><<<blocks, sb2::kWGal * sb2::kWSlots, sh, st>>>(M, F, P, O, perm, n_pad);
  } else if (fast) {
    sb2::weights_kernel<true, false><<<blocks, sb2::kWGal * sb2::kWSlots, sh, st>>>(M, F, P, O, perm, n_pad);
  } else {
    sb2::weights_kernel<false, false><<<blocks, sb2::kWGal * sb2::kWSlots, sh, st>>>(M, F, P, O, perm, n_pad);
  }
  STAGE_CHECK("weights_kernel", st);
  if (M.igm_on && !w_f64) {
    dim3 grid((unsigned)(n_pad / 128), (unsigned)((m->n_blue_pad + sb2::kIgmStrip - 1) / sb2::kIgmStrip));
    sb2::igm_kernel<<<grid, 128, 0, st>>>(M, m->zpow, m->g_orig, m->igm, m->tile_range, rpu == 256 ? 1 : 0, m->n_blue_pad, n_pad);
    STAGE_CHECK("igm_kernel", st);
  }
  cudaEventRecord(m->ev[2], st);
  return SB2_OK;
}

}  // namespace

extern "C" {

int sb2_build_weights(sb2_model* m, const sb2_params* p, double* w_out, void* stream) {
  int rc = check_device_params(m, p);
  if (rc != SB2_OK) return rc;
  if (!w_out) return fail(SB2_ERR_INVALID, "w_out is null");
  CU_TRY(cudaSetDevice(m->device));
  return run_prep(m, p, w_out, false, false, (cudaStream_t)stream);
}

int sb2_synth_photometry(sb2_model* m, const sb2_params* p, float* flux_base, double* flux_scaled,
                         float* spec_out, void* stream) {
  int rc = check_device_params(m, p);
  if (rc != SB2_OK) return rc;
  if (!flux_base && !flux_scaled && !spec_out) return fail(SB2_ERR_INVALID, "no output requested");
  CU_TRY(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  // (spectra and the absorbed-energy sum of the dust emission need every wavelength chunk, not just the filters')
  // ... unless the model holds pseudo-bins for that sum and the batch takes the kernel that reads them (synth3_kernel)
  const bool x_mode = m->d.x_bins > 0 && !p->energy_full_axis && use_s3(m, p);
  if ((rc = run_prep(m, p, nullptr, true, spec_out != nullptr || (m->dust_wnu != nullptr && !x_mode), st)) != SB2_OK) return rc;
  const sb2_model_desc& d = m->d;
  sb2::SynthArgs a{};
  const bool delta = delta_mode(m, p);
  a.n_gal = (int)p->n;
  const int rpu = rows_per_unit(m, p, delta);
  a.n_tiles = (int)(padded_rows(m, p, p->n, delta) / rpu);  // units
  a.n_tiles_dev = m->grp + 3 * kMaxGroups;
  a.tile_k0 = m->tile_k0;
  a.k8_total = (delta ? m->wd_stride : d.k_pad) / 8;
  a.dbg = m->sw.dbg;
  a.cross = cross_mode(m) ? 1 : 0;
  a.two_pass = ((!delta || (!use_s3(m, p) && use_split(m, delta))) && !m->sw.one_pass) ? 1 : 0;
  {
    const int lch = sb2::kBN / d.n_comp, lch3 = s3_cols(m) / d.n_comp;
    a.n_chunk = (d.n_lam + lch - 1) / lch;          // the real axis; the tables (kap_len) also cover the pseudo-bins
    a.kap_len = d.n_chunk * lch;
    a.x_first = x_mode ? d.x_bin0 / lch3 : 0;
    a.x_count = x_mode ? d.x_bins / lch3 : 0;
  }
  a.n_kb = (a.k8_total + 3) / 4; a.n_lam = d.n_lam; a.n_filt = d.n_filt;
  a.n_blue = d.n_blue; a.n_blue_pad = m->n_blue_pad; a.uv_len = m->uv_len;
  a.tile_range = m->tile_range;
  a.dust_d0 = m->dust_d0; a.dust_l2 = m->dust_l2; a.g_slope = m->g_slope; a.g_ampl = m->g_ampl;
  a.g_lya = m->g_lya; a.lya_bin = d.lya_bin; a.kappa_birth = m->kappa_birth; a.g_taub = m->g_taub;
  a.wnu = m->dust_wnu; a.e_part = m->e_part;
  a.kappa = m->kappa; a.filt_uv = reinterpret_cast<const float2*>(m->filt_uv); a.igm = m->igm;
  a.g_m = m->g_m; a.g_beta = m->g_beta; a.g_gamma = m->g_gamma; a.g_taut = m->g_taut; a.g_scale = m->g_scale; a.g_ca = m->g_ca;
  a.g_cb = m->g_cb; a.g_orig = m->g_orig; a.g_mscale = m->g_mscale; a.g_trunc = m->g_trunc;
  a.out_base = flux_base; a.out_scaled = flux_scaled; a.out_spec = spec_out;
  for (int f = 0; f < d.n_filt; ++f) {
    a.filt_lo[f] = m->h_lo[f]; a.filt_hi[f] = m->h_hi[f]; a.filt_off[f] = m->h_off[f];
    a.filt_su[f] = m->h_su[f]; a.filt_sdv[f] = m->h_sdv[f];
  }
  a.part = m->part;
  a.n_rows = (long long)a.n_tiles * rpu;
  // fused output: synth3_kernel's epilogue writes the fluxes, the dust emission's share of the filter sums included
  m->last_fused = false;
  //               (with dust emission AND spectra the emission's spectrum is added from the e_part planes: not fused)
  a.fuse_out = (rpu != 256 && use_s3(m, p) && !(m->dust_wnu && spec_out) && !m->sw.no_fuse) ? 1 : 0;
  a.scaled_ld = flux_scaled ? p->scaled_ld : 0;
  a.dust_duv = m->dust_wnu ? reinterpret_cast<const float2*>(m->dust_duv) : nullptr;
  a.dust_m_len = d.dust_m_len;
  if (rpu == 256) {
    const int grid = 2 * std::min(a.n_tiles, m->n_sm / 2);
    rc = launch_synth2(m, a, grid, st);
  } else if (use_s3(m, p)) {
    rc = launch_synth3(m, a, a.n_tiles < m->n_sm ? a.n_tiles : m->n_sm, st);
  } else {
    int grid = a.n_tiles < m->n_sm ? a.n_tiles : m->n_sm;
    if (m->sw.dense_grid > 0) grid = std::min(grid, m->sw.dense_grid);
    rc = launch_synth(m, a, grid, delta, st);
  }
  if (rc == SB2_OK) {
    sb2::FinalizeArgs fa{};
    fa.part = m->part; fa.n_rows = a.n_rows; fa.n_filt = d.n_filt; fa.n_comp = d.n_comp; fa.n_groups = (rpu == 256 && sb2::kT2Buf >= 3 && sb2::kMaxGroups >= 3) ? 3 : 2;
    fa.g_beta = m->g_beta; fa.g_gamma = m->g_gamma; fa.g_scale = m->g_scale; fa.g_ca = m->g_ca; fa.g_orig = m->g_orig;
    fa.g_mscale = m->g_mscale; fa.g_trunc = m->g_trunc; fa.out_base = flux_base; fa.out_scaled = flux_scaled;
    fa.scaled_ld = flux_scaled ? p->scaled_ld : 0;
    for (int f = 0; f < d.n_filt; ++f) { fa.filt_su[f] = m->h_su[f]; fa.filt_sdv[f] = m->h_sdv[f]; }
    fa.e_part = m->dust_wnu ? m->e_part : nullptr; fa.dust_duv = reinterpret_cast<const float2*>(m->dust_duv); fa.dust_g = m->dust_g;
    fa.g_m = m->g_m; fa.dust_m_len = d.dust_m_len; fa.n_lam = d.n_lam; fa.out_spec = spec_out;
    if (spec_out && fa.e_part) {
      sb2::dust_spec_kernel<<<(unsigned)((a.n_rows + 7) / 8), 256, 0, st>>>(fa, a.n_tiles_dev, rpu);
      STAGE_CHECK("dust_spec_kernel", st);
    }
    if ((flux_base || flux_scaled) && !m->last_fused) {
      sb2::finalize_kernel<<<(unsigned)((a.n_rows + 255) / 256), 256, 0, st>>>(fa, a.n_tiles_dev, rpu);
      STAGE_CHECK("finalize_kernel", st);
    }
  }
  cudaEventRecord(m->ev[3], st);
  m->ev_valid = (rc == SB2_OK);
  return rc;
}

int sb2_last_stage_ms(sb2_model* m, float* out3) {
  if (!m || !out3) return fail(SB2_ERR_INVALID, "null argument");
  if (!m->ev_valid) return fail(SB2_ERR_INVALID, "no completed sb2_synth_photometry call to time");
  CU_TRY(cudaEventSynchronize(m->ev[3]));
  for (int i = 0; i < 3; ++i) CU_TRY(cudaEventElapsedTime(&out3[i], m->ev[i], m->ev[i + 1]));
  return SB2_OK;
}

int sb2_synth_photometry_host_wait(sb2_model* m, int slot) {
  if (!m || slot < 0 || slot > 1) return fail(SB2_ERR_INVALID, "bad slot");
  if (!m->slot_busy[slot]) return SB2_OK;
  CU_TRY(cudaSetDevice(m->device));
  CU_TRY(cudaEventSynchronize(m->ev_slot[slot]));
  m->slot_busy[slot] = false;
  return SB2_OK;
}

// Host entry with full-wavelength output (cfg 5's write path, library.py:4887-4919): 4 n_lam bytes per galaxy leave the
// device, so the batch is walked in slices of kSpecSlice galaxies through TWO device buffers -- the device-to-host copy of
// one slice's spectra (pinned destination for real overlap) runs beside the kernels of the next.
namespace {
constexpr long long kSpecSlice = 32768;

int host_with_spectra(sb2_model* m, const sb2_params* p, float* flux_base, double* flux_scaled, float* spec_out) {
  int rc = check_params(m, p);
  if (rc != SB2_OK) return rc;
  if (p->host_f32) return fail(SB2_ERR_INVALID, "host_f32 transport is not available together with spec_out");
  if (p->scaled_ld) return fail(SB2_ERR_INVALID, "a transposed flux_scaled (scaled_ld) is not available together with spec_out");
  CU_TRY(cudaSetDevice(m->device));
  for (int s = 0; s < 2; ++s) {
    rc = sb2_synth_photometry_host_wait(m, s);     // the staging areas below are shared with the photometry-only entry
    if (rc != SB2_OK) return rc;
  }
  const long long n = p->n, rows = std::min<long long>(kSpecSlice, m->cap);
  const int nf = m->d.n_filt, nl = m->d.n_lam;
  if (m->stage_spec_rows < rows) {
    for (int s = 0; s < 2; ++s) {
      if (m->stage_spec[s]) cudaFree(m->stage_spec[s]);
      m->stage_spec[s] = nullptr;
      CU_TRY(cudaMalloc(reinterpret_cast<void**>(&m->stage_spec[s]), (size_t)rows * nl * sizeof(float)));
      if (!m->ev_spec[s]) CU_TRY(cudaEventCreateWithFlags(&m->ev_spec[s], cudaEventDisableTiming));
    }
    m->stage_spec_rows = rows;
  }
  constexpr int kNA = 12;
  const double* src[kNA] = {p->redshift, p->log_mass, p->tau_v, p->zd_value, p->zd_sigma, p->coef_att, p->coef_unatt,
                            p->dust_slope, p->dust_ampl, p->fesc_lya, p->tau_v_birth, p->sfh_rows};
  bool used[2] = {false, false};
  int k = 0;
  for (long long a = 0; a < n; a += rows, ++k) {
    const long long b = std::min(n, a + rows), cnt = b - a;
    const int s = k & 1;
    if (used[s]) CU_TRY(cudaEventSynchronize(m->ev_spec[s]));   // slot's previous spectra have left the device
    double* base = m->stage_params[s];
    sb2_params dp = *p;
    dp.n = cnt;
    double* dev[kNA];
    for (int i = 0; i < kNA; ++i) {
      const size_t w = (i == kNA - 1) ? (size_t)p->sfh_stride : 1;
      dev[i] = src[i] ? base : nullptr;
      if (src[i]) {
        CU_TRY(cudaMemcpyAsync(base, src[i] + a * w, (size_t)cnt * w * sizeof(double), cudaMemcpyHostToDevice, m->st_h2d));
        base += (size_t)cnt * w;
      }
    }
    CU_TRY(cudaEventRecord(m->ev_in[s], m->st_h2d));
    CU_TRY(cudaStreamWaitEvent(m->st_comp, m->ev_in[s], 0));
    dp.redshift = dev[0]; dp.log_mass = dev[1]; dp.tau_v = dev[2]; dp.zd_value = dev[3]; dp.zd_sigma = dev[4];
    dp.coef_att = dev[5]; dp.coef_unatt = dev[6]; dp.dust_slope = dev[7]; dp.dust_ampl = dev[8]; dp.fesc_lya = dev[9];
    dp.tau_v_birth = dev[10]; dp.sfh_rows = dev[11];
    rc = sb2_synth_photometry(m, &dp, flux_base ? m->stage_flux[s] : nullptr, flux_scaled ? m->stage_flux64[s] : nullptr,
                              m->stage_spec[s], m->st_comp);
    if (rc != SB2_OK) {
      const std::string msg = g_err;
      cudaStreamSynchronize(m->st_h2d); cudaStreamSynchronize(m->st_comp); cudaStreamSynchronize(m->st_d2h);
      g_err = msg;
      return rc;
    }
    CU_TRY(cudaEventRecord(m->ev_done[s], m->st_comp));
    CU_TRY(cudaStreamWaitEvent(m->st_d2h, m->ev_done[s], 0));
    CU_TRY(cudaMemcpyAsync(spec_out + (size_t)a * nl, m->stage_spec[s], (size_t)cnt * nl * sizeof(float), cudaMemcpyDeviceToHost, m->st_d2h));
    if (flux_base) CU_TRY(cudaMemcpyAsync(flux_base + (size_t)a * nf, m->stage_flux[s], (size_t)cnt * nf * 4, cudaMemcpyDeviceToHost, m->st_d2h));
    if (flux_scaled) CU_TRY(cudaMemcpyAsync(flux_scaled + (size_t)a * nf, m->stage_flux64[s], (size_t)cnt * nf * 8, cudaMemcpyDeviceToHost, m->st_d2h));
    CU_TRY(cudaEventRecord(m->ev_spec[s], m->st_d2h));
    used[s] = true;
    // the next slice must not overwrite this slot's parameter staging before its kernels have read it: slot s is reused two
    // slices later, after ev_spec[s] (recorded behind ev_done[s]) has been waited for above
  }
  CU_TRY(cudaStreamSynchronize(m->st_d2h));
  return SB2_OK;
}
}  // namespace

int sb2_synth_photometry_host(sb2_model* m, const sb2_params* p, float* flux_base, double* flux_scaled,
                              float* spec_out) {
  if (spec_out) return host_with_spectra(m, p, flux_base, flux_scaled, spec_out);
  int rc = sb2_synth_photometry_host_submit(m, p, flux_base, flux_scaled, 0);
  if (rc != SB2_OK) return rc;
  return sb2_synth_photometry_host_wait(m, 0);
}

int sb2_synth_photometry_host_submit(sb2_model* m, const sb2_params* p, float* flux_base, double* flux_scaled, int slot) {
  int rc = check_params(m, p);
  if (rc != SB2_OK) return rc;
  if (slot < 0 || slot > 1) return fail(SB2_ERR_INVALID, "slot must be 0 or 1");
  if (!flux_base && !flux_scaled) return fail(SB2_ERR_INVALID, "no output requested");
  CU_TRY(cudaSetDevice(m->device));
  if (m->slot_busy[slot]) {   // the slot's staging buffers are still in use by an unfinished submit
    rc = sb2_synth_photometry_host_wait(m, slot);
    if (rc != SB2_OK) return rc;
  }
  const size_t n = (size_t)p->n;
  const int nf = m->d.n_filt;
  // Device layout of the staged parameters: one full-length array per field; slices [a, b) of the batch are
  // copied in on st_h2d, synthesised on st_comp and copied out on st_d2h, so the PCIe copies of one slice
  // overlap the kernels of its neighbours (pinned host buffers are needed for the overlap, not for correctness).
  double* base = m->stage_params[slot];
  constexpr int kNA = 12;   // staged arrays; the last one is the SFH row table
  const double* src[kNA] = {p->redshift, p->log_mass, p->tau_v, p->zd_value, p->zd_sigma, p->coef_att, p->coef_unatt,
                            p->dust_slope, p->dust_ampl, p->fesc_lya, p->tau_v_birth, p->sfh_rows};
  double* dev[kNA];
  for (int i = 0; i < kNA; ++i) {
    const size_t w = (i == kNA - 1) ? (size_t)p->sfh_stride : 1;
    dev[i] = src[i] ? base : nullptr;
    if (src[i]) base += n * w;
  }
  int n_slice = 1;
  {
    // measured on B200: each extra slice costs ~0.3 ms of per-launch overhead, 2 slices per 1M galaxies is best for a
    // lone batch; when the other slot is in flight the neighbouring batch already provides the overlap
    const long long want = m->sw.host_slices > 0 ? m->sw.host_slices : (m->slot_busy[slot ^ 1] ? 1 : (long long)(n / 400000));
    n_slice = (int)std::min<long long>(8, std::max<long long>(1, want));
  }
  const size_t per = ((n + n_slice - 1) / n_slice + 255) / 256 * 256;
  const bool trace = m->sw.trace;   // per-slice timeline on stderr (diagnostics)
  cudaEvent_t tr[1 + 8 * 4] = {};
  if (trace) {
    for (auto& e : tr) cudaEventCreate(&e);
    cudaEventRecord(tr[0], m->st_h2d);
    cudaStreamWaitEvent(m->st_comp, tr[0], 0);
    cudaStreamWaitEvent(m->st_d2h, tr[0], 0);
  }
  for (int sl = 0; sl < n_slice; ++sl) {
    const size_t a = (size_t)sl * per, b = std::min(n, a + per);
    if (a >= b) break;
    for (int i = 0; i < kNA; ++i) {
      if (!src[i]) continue;
      const size_t w = (i == kNA - 1) ? (size_t)p->sfh_stride : 1;
      if (p->host_f32) {   // half the PCIe bytes; widened on the compute stream below
        float* d32 = m->stage_params32[slot] + (dev[i] - m->stage_params[slot]);
        CU_TRY(cudaMemcpyAsync(d32 + a * w, reinterpret_cast<const float*>(src[i]) + a * w, (b - a) * w * sizeof(float), cudaMemcpyHostToDevice, m->st_h2d));
      } else {
        CU_TRY(cudaMemcpyAsync(dev[i] + a * w, src[i] + a * w, (b - a) * w * sizeof(double), cudaMemcpyHostToDevice, m->st_h2d));
      }
    }
    CU_TRY(cudaEventRecord(m->ev_in[sl], m->st_h2d));
    if (trace) cudaEventRecord(tr[1 + sl * 4 + 0], m->st_h2d);
    CU_TRY(cudaStreamWaitEvent(m->st_comp, m->ev_in[sl], 0));
    if (trace) cudaEventRecord(tr[1 + sl * 4 + 1], m->st_comp);
    if (p->host_f32) {
      WidenSegs sg{};
      int n_seg = 0;
      long long longest = 0;
      for (int i = 0; i < kNA; ++i) {
        if (!src[i]) continue;
        const size_t w = (i == kNA - 1) ? (size_t)p->sfh_stride : 1;
        sg.off[n_seg] = (long long)((dev[i] - m->stage_params[slot]) + a * w);
        sg.cnt[n_seg] = (long long)((b - a) * w);
        longest = std::max(longest, sg.cnt[n_seg]);
        ++n_seg;
      }
      const unsigned bx = (unsigned)std::min<long long>((longest / 4 + 255) / 256 + 1, 1024);
      widen_kernel<<<dim3(bx, (unsigned)n_seg), 256, 0, m->st_comp>>>(m->stage_params32[slot], m->stage_params[slot], sg);
      STAGE_CHECK("widen_kernel", m->st_comp);
    }
    sb2_params dp = *p;
    dp.host_f32 = 0;
    dp.n = (int64_t)(b - a);
    dp.redshift = dev[0] + a; dp.log_mass = dev[1] ? dev[1] + a : nullptr; dp.tau_v = dev[2] ? dev[2] + a : nullptr;
    dp.zd_value = dev[3] + a; dp.zd_sigma = dev[4] ? dev[4] + a : nullptr;
    dp.coef_att = dev[5] ? dev[5] + a : nullptr; dp.coef_unatt = dev[6] ? dev[6] + a : nullptr;
    dp.dust_slope = dev[7] ? dev[7] + a : nullptr; dp.dust_ampl = dev[8] ? dev[8] + a : nullptr;
    dp.fesc_lya = dev[9] ? dev[9] + a : nullptr;
    dp.tau_v_birth = dev[10] ? dev[10] + a : nullptr;
    dp.sfh_rows = dev[11] + a * p->sfh_stride;
    // transposed scaled output: the staging area is [n_filt][n], this slice fills its columns [a, b)
    const bool transp = flux_scaled && p->scaled_ld > 0;
    dp.scaled_ld = transp ? (int64_t)n : 0;
    rc = sb2_synth_photometry(m, &dp, flux_base ? m->stage_flux[slot] + a * nf : nullptr,
                              flux_scaled ? m->stage_flux64[slot] + (transp ? a : a * nf) : nullptr, nullptr, m->st_comp);
    if (rc != SB2_OK) {   // earlier slices are still in flight on the three streams: drain them before reporting
      const std::string msg = g_err;
      cudaStreamSynchronize(m->st_h2d); cudaStreamSynchronize(m->st_comp); cudaStreamSynchronize(m->st_d2h);
      g_err = msg;
      return rc;
    }
    CU_TRY(cudaEventRecord(m->ev_done[sl], m->st_comp));
    if (trace) cudaEventRecord(tr[1 + sl * 4 + 2], m->st_comp);
    CU_TRY(cudaStreamWaitEvent(m->st_d2h, m->ev_done[sl], 0));
    if (flux_base)
      CU_TRY(cudaMemcpyAsync(flux_base + a * nf, m->stage_flux[slot] + a * nf, (b - a) * nf * 4, cudaMemcpyDeviceToHost, m->st_d2h));
    if (transp)
      CU_TRY(cudaMemcpy2DAsync(flux_scaled + a, (size_t)p->scaled_ld * 8, m->stage_flux64[slot] + a, n * 8, (b - a) * 8, (size_t)nf,
                               cudaMemcpyDeviceToHost, m->st_d2h));
    else if (flux_scaled)
      CU_TRY(cudaMemcpyAsync(flux_scaled + a * nf, m->stage_flux64[slot] + a * nf, (b - a) * nf * 8, cudaMemcpyDeviceToHost, m->st_d2h));
    if (trace) cudaEventRecord(tr[1 + sl * 4 + 3], m->st_d2h);
  }
  CU_TRY(cudaEventRecord(m->ev_slot[slot], m->st_d2h));
  m->slot_busy[slot] = true;
  if (trace) {
    cudaEventSynchronize(m->ev_slot[slot]);
    for (int sl = 0; sl < n_slice && (size_t)sl * per < n; ++sl) {
      float t[4];
      for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&t[k], tr[0], tr[1 + sl * 4 + k]);
      std::fprintf(stderr, "[sb2 trace] slice %d: h2d done %.3f ms | kernels %.3f -> %.3f ms | d2h done %.3f ms\n", sl, t[0], t[1], t[2], t[3]);
    }
    for (auto& e : tr) cudaEventDestroy(e);
  }
  return SB2_OK;
}

}  // extern "C"

namespace {
// Feature rows only with Philox draws: depth_noise_feat_kernel (one thread per row and filter quad).
template <typename T>
int launch_noise_feat(const sb2::NoiseArgs& a, cudaStream_t st) {
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const long long rows = a.n_gal * a.n_scatter;
  const int rows_pb = 256 / ((a.n_filt + 3) / 4);
  long long blocks = (rows + rows_pb - 1) / rows_pb;
  if (blocks > (long long)n_sm * 32) blocks = (long long)n_sm * 32;
  sb2::depth_noise_feat_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(a);
  STAGE_CHECK("depth_noise_feat_kernel", st);
  return SB2_OK;
}
}  // namespace

extern "C" {

int sb2_depth_noise_features(const double* flux, int64_t n_gal, int32_t n_filt, int32_t n_scatter,
                             const double* sigma, double min_flux_pc_error, const double* normals, uint64_t seed,
                             uint64_t epoch, double norm_mag_limit, double* out_flux, double* out_sigma,
                             float* out_feat, void* stream) {
  if (!flux || !sigma || n_gal < 1 || n_filt < 1 || n_scatter < 1) return fail(SB2_ERR_INVALID, "bad argument");
  if (!out_flux && !out_feat) return fail(SB2_ERR_INVALID, "no output requested");
  sb2::NoiseArgs a{};
  a.flux = flux; a.n_gal = n_gal; a.n_filt = n_filt; a.n_scatter = n_scatter; a.sigma = sigma;
  a.min_pc = min_flux_pc_error; a.normals = normals; a.seed = seed; a.epoch = epoch; a.mag_limit = norm_mag_limit;
  a.out_flux = out_flux; a.out_sigma = out_sigma; a.out_feat = out_feat;
  const long long rows = (long long)n_gal * n_scatter;
  if (!normals && !out_flux && !out_sigma && out_feat && n_filt <= 1024 && (reinterpret_cast<uintptr_t>(flux) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out_feat) & 15) == 0)
    return launch_noise_feat<double>(a, (cudaStream_t)stream);
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (rows + 255) / 256;
  const long long cap = (long long)n_sm * 16;
  if (blocks > cap) blocks = cap;
  sb2::depth_noise_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  STAGE_CHECK("depth_noise_kernel", (cudaStream_t)stream);
  return SB2_OK;
}

int sb2_depth_noise_features_f32(const float* flux, int64_t n_gal, int32_t n_filt, int32_t n_scatter, const double* sigma,
                                 double min_flux_pc_error, uint64_t seed, uint64_t epoch, double norm_mag_limit,
                                 float* out_feat, void* stream) {
  if (!flux || !sigma || !out_feat || n_gal < 1 || n_filt < 1 || n_filt > 1024 || n_scatter < 1) return fail(SB2_ERR_INVALID, "bad argument");
  if ((reinterpret_cast<uintptr_t>(flux) & 15) != 0 || (reinterpret_cast<uintptr_t>(out_feat) & 15) != 0)
    return fail(SB2_ERR_INVALID, "flux and out_feat must be 16-byte aligned");
  sb2::NoiseArgs a{};
  a.flux32 = flux; a.n_gal = n_gal; a.n_filt = n_filt; a.n_scatter = n_scatter; a.sigma = sigma;
  a.min_pc = min_flux_pc_error; a.seed = seed; a.epoch = epoch; a.mag_limit = norm_mag_limit; a.out_feat = out_feat;
  return launch_noise_feat<float>(a, (cudaStream_t)stream);
}

int sb2_depth_noise_features_sets(const double* flux, int64_t n_gal, int32_t n_filt, int32_t n_scatter,
                                  const double* sigma_sets, int32_t n_sets, const int32_t* set_index,
                                  double min_flux_pc_error, const double* normals, uint64_t seed, uint64_t epoch,
                                  double norm_mag_limit, double* out_flux, double* out_sigma, float* out_feat, void* stream) {
  if (!flux || !sigma_sets || !set_index || n_gal < 1 || n_filt < 1 || n_scatter < 1 || n_sets < 1)
    return fail(SB2_ERR_INVALID, "bad argument");
  if (!out_flux && !out_feat) return fail(SB2_ERR_INVALID, "no output requested");
  sb2::NoiseArgs a{};
  a.flux = flux; a.n_gal = n_gal; a.n_filt = n_filt; a.n_scatter = n_scatter; a.sigma = sigma_sets; a.set_index = set_index;
  a.min_pc = min_flux_pc_error; a.normals = normals; a.seed = seed; a.epoch = epoch; a.mag_limit = norm_mag_limit;
  a.out_flux = out_flux; a.out_sigma = out_sigma; a.out_feat = out_feat;
  const long long rows = (long long)n_gal * n_scatter;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = std::min<long long>((rows + 255) / 256, (long long)n_sm * 16);
  sb2::depth_noise_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  CU_TRY(cudaGetLastError());
  return SB2_OK;
}

// ---- spectroscopic path -------------------------------------------------------------------------------------------
struct sb2_resampler {
  int device = 0, n_sm = 148;
  sb2::ResampleArgs a{};
  double *old_edges = nullptr, *new_edges = nullptr;
  float *lam = nullptr, *lam_k = nullptr, *qth2 = nullptr, *res_wave = nullptr, *res_r = nullptr, *res_slope = nullptr;
  int *edge_lut = nullptr, *res_lut = nullptr;
  size_t smem = 0;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool ev_valid = false;
  float* stage_in = nullptr; double* stage_z = nullptr; float* stage_out = nullptr;   // host-buffer form
  long long stage_n = 0;
};

static void spectres_edges(const double* w, int n, std::vector<double>& e) {   // bin edges as `spectres` makes them
  e.resize((size_t)n + 1);
  e[0] = w[0] - (w[1] - w[0]) / 2;
  e[n] = w[n - 1] + (w[n - 1] - w[n - 2]) / 2;
  for (int i = 1; i < n; ++i) e[i] = (w[i] + w[i - 1]) / 2;
}

int sb2_resampler_create(const sb2_resample_desc* d, int device, sb2_resampler** out) {
  if (!d || !out) return fail(SB2_ERR_INVALID, "null argument");
  if (d->n_lam < 3 || d->n_px < 2 || d->n_res < 1 || !d->theory_wave || !d->observed_wave || !d->res_wave || !d->res_r)
    return fail(SB2_ERR_INVALID, "resampler: need >= 3 model wavelengths, >= 2 pixels and a resolution curve");
  for (int i = 1; i < d->n_lam; ++i)
    if (!(d->theory_wave[i] > d->theory_wave[i - 1])) return fail(SB2_ERR_INVALID, "theory_wave must increase");
  for (int i = 1; i < d->n_px; ++i)
    if (!(d->observed_wave[i] > d->observed_wave[i - 1])) return fail(SB2_ERR_INVALID, "observed_wave must increase");
  for (int i = 1; i < d->n_res; ++i)
    if (!(d->res_wave[i] > d->res_wave[i - 1])) return fail(SB2_ERR_INVALID, "resolution curve abscissa must increase");
  for (int i = 0; i < d->n_res; ++i)
    if (!(d->res_r[i] > 0)) return fail(SB2_ERR_INVALID, "resolution R must be positive");
  if (!(d->trunc > 0)) return fail(SB2_ERR_INVALID, "trunc must be positive");
  CU_TRY(cudaSetDevice(device));
  sb2_resampler* r = new sb2_resampler();
  r->device = device;
  cudaDeviceGetAttribute(&r->n_sm, cudaDevAttrMultiProcessorCount, device);
  std::vector<double> oe, ne, diff((size_t)d->n_lam - 1);
  spectres_edges(d->theory_wave, d->n_lam, oe);
  spectres_edges(d->observed_wave, d->n_px, ne);
  for (int i = 0; i + 1 < d->n_lam; ++i) diff[i] = d->theory_wave[i + 1] - d->theory_wave[i];
  std::sort(diff.begin(), diff.end());
  const size_t nd = diff.size();
  const double med = (nd & 1) ? diff[nd / 2] : 0.5 * (diff[nd / 2 - 1] + diff[nd / 2]);   // np.median
  std::vector<float> lam(d->n_lam), lam_k(d->n_lam), qth2(d->n_lam), rw(d->n_res), rr(d->n_res), rs(std::max(d->n_res - 1, 1), 0.f);
  const double fwhm = 2.0 * std::sqrt(2.0 * std::log(2.0));
  for (int i = 0; i < d->n_lam; ++i) {
    lam[i] = (float)d->theory_wave[i];
    // the reference divides with where=pixel_scale != 0 -> sigma 0
    lam_k[i] = med > 0 ? (float)(d->theory_wave[i] / (med * fwhm)) : 0.f;
    const double rt = d->theory_r ? d->theory_r[i] : d->theory_r_scalar;
    qth2[i] = (float)(1.0 / (rt * rt));        // inf -> 0
  }
  for (int i = 0; i < d->n_res; ++i) { rw[i] = (float)d->res_wave[i]; rr[i] = (float)d->res_r[i]; }
  for (int i = 0; i + 1 < d->n_res; ++i) rs[i] = (float)((d->res_r[i + 1] - d->res_r[i]) / (d->res_wave[i + 1] - d->res_wave[i]));
  int rc = SB2_OK;
#define RUP(dst, src, n) if (rc == SB2_OK) rc = upload(&r->dst, src, (size_t)(n));
  RUP(old_edges, oe.data(), oe.size());
  RUP(new_edges, ne.data(), ne.size());
  RUP(lam, lam.data(), lam.size());
  RUP(lam_k, lam_k.data(), lam_k.size());
  RUP(qth2, qth2.data(), qth2.size());
  RUP(res_wave, rw.data(), rw.size());
  RUP(res_r, rr.data(), rr.size());
  RUP(res_slope, rs.data(), rs.size());
  // search tables over uniform steps of log2(wavelength): entry b = last node at or below the left end of bucket b
  auto make_lut = [](const double* x, int n, int buckets, std::vector<int>& lut, float& u0, float& inv_du) {
    const double a = std::log2(x[0]), b = std::log2(x[n - 1]);
    const double du = (b - a) / buckets;
    u0 = (float)a; inv_du = (float)(1.0 / du);
    lut.resize(buckets);
    int k = 0;
    for (int i = 0; i < buckets; ++i) {
      // (a shade to the left of the bucket's nominal end, so that float rounding of u0 / inv_du cannot overshoot)
      const double left = std::exp2(a + (i - 0.5) * du);
      while (k + 1 < n && x[k + 1] <= left) ++k;
      lut[i] = k;
    }
  };
  std::vector<int> elut, rlut;
  sb2::ResampleArgs& a = r->a;
  if (oe[0] > 0) {
    make_lut(oe.data(), d->n_lam + 1, 4 * d->n_lam, elut, a.lut_u0, a.lut_inv_du);
    RUP(edge_lut, elut.data(), elut.size());
    a.lut_n = (int)elut.size();
    if (d->n_res >= 2 && d->res_wave[0] > 0) {
      make_lut(d->res_wave, d->n_res, 16 * d->n_res, rlut, a.res_u0, a.res_inv_du);
      RUP(res_lut, rlut.data(), rlut.size());
      a.res_lut_n = (int)rlut.size();
    }
  }
#undef RUP
  if (rc == SB2_OK && (cudaEventCreate(&r->ev[0]) != cudaSuccess || cudaEventCreate(&r->ev[1]) != cudaSuccess))
    rc = fail(SB2_ERR_CUDA, "cudaEventCreate");
  a.edge_lut = r->edge_lut; a.res_lut = r->res_lut;
  a.n_lam = d->n_lam; a.n_px = d->n_px; a.n_res = d->n_res;
  a.old_edges = r->old_edges; a.new_edges = r->new_edges; a.lam = r->lam; a.lam_k = r->lam_k; a.qth2 = r->qth2;
  a.res_wave = r->res_wave; a.res_r = r->res_r; a.res_slope = r->res_slope;
  a.trunc = (float)d->trunc; a.fill = (float)d->fill;
  // staging: the contributing bins (at most the whole axis) plus 2 * h_cap taps, plus the smoothed bins
  int max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  a.h_cap = 512;
  // the bins under the observed window: its rest-frame image [first edge / (1+z), last edge / (1+z)] has a fixed RATIO, so
  // the largest count over all redshifts is the largest count over all windows of that ratio along the axis
  int nb_max = d->n_lam;
  if (oe[0] > 0 && ne[0] > 0) {
    const double ratio = ne[d->n_px] / ne[0] * (1.0 + 1e-9);
    nb_max = 1;
    for (int k = 0, k2 = 0; k < d->n_lam; ++k) {
      while (k2 < d->n_lam && oe[k2] < oe[k] * ratio) ++k2;     // bins k .. k2-1 start inside a window that starts in bin k
      nb_max = std::max(nb_max, k2 - k + 1);
    }
    nb_max = std::min(nb_max + 1, d->n_lam);
  }
  a.nb_max = nb_max;
  r->smem = ((size_t)2 * nb_max + 2 * (size_t)a.h_cap) * sizeof(float);
  if (rc == SB2_OK && r->smem > (size_t)max_smem - 1024) rc = fail(SB2_ERR_INVALID, "resampler: model axis too long for one CTA's shared memory");
  if (rc == SB2_OK && cudaFuncSetAttribute(sb2::resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r->smem) != cudaSuccess)
    rc = fail(SB2_ERR_CUDA, "cudaFuncSetAttribute(resample_kernel)");
  if (rc != SB2_OK) { sb2_resampler_destroy(r); return rc; }
  *out = r;
  return SB2_OK;
}

int sb2_resampler_destroy(sb2_resampler* r) {
  if (!r) return SB2_OK;
  cudaSetDevice(r->device);
  void* ptrs[] = {r->old_edges, r->new_edges, r->lam, r->lam_k, r->qth2, r->res_wave, r->res_r, r->res_slope, r->edge_lut, r->res_lut, r->stage_in, r->stage_z, r->stage_out};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (auto& e : r->ev) if (e) cudaEventDestroy(e);
  delete r;
  return SB2_OK;
}

int sb2_resample_spectra(sb2_resampler* r, const float* spectra, const double* redshift, int64_t n, float* out, void* stream) {
  if (!r || !spectra || !redshift || !out) return fail(SB2_ERR_INVALID, "null argument");
  if (n < 0) return fail(SB2_ERR_INVALID, "negative n");
  if (n == 0) return SB2_OK;
  CU_TRY(cudaSetDevice(r->device));
  cudaStream_t st = (cudaStream_t)stream;
  sb2::ResampleArgs a = r->a;
  a.spectra = spectra; a.redshift = redshift; a.out = out; a.n = n;
  // persistent grid: whole waves of the CTAs that fit on the SMs
  int per_sm = 1;
  CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sb2::resample_kernel, sb2::kResampleThreads, r->smem));
  const long long grid = std::min<long long>(n, (long long)r->n_sm * std::max(per_sm, 1));
  cudaEventRecord(r->ev[0], st);
  sb2::resample_kernel<<<(unsigned)grid, sb2::kResampleThreads, r->smem, st>>>(a);
  CU_TRY(cudaGetLastError());
  cudaEventRecord(r->ev[1], st);
  r->ev_valid = true;
  return SB2_OK;
}

int sb2_resample_last_ms(sb2_resampler* r, float* ms) {
  if (!r || !ms) return fail(SB2_ERR_INVALID, "null argument");
  if (!r->ev_valid) return fail(SB2_ERR_INVALID, "no sb2_resample_spectra call to time");
  CU_TRY(cudaEventSynchronize(r->ev[1]));
  CU_TRY(cudaEventElapsedTime(ms, r->ev[0], r->ev[1]));
  return SB2_OK;
}

int sb2_resample_spectra_host(sb2_resampler* r, const float* spectra, const double* redshift, int64_t n, float* out) {
  if (!r || !spectra || !redshift || !out) return fail(SB2_ERR_INVALID, "null argument");
  if (n < 0) return fail(SB2_ERR_INVALID, "negative n");
  if (n == 0) return SB2_OK;
  CU_TRY(cudaSetDevice(r->device));
  if (n > r->stage_n) {
    for (void* p : {(void*)r->stage_in, (void*)r->stage_z, (void*)r->stage_out}) if (p) cudaFree(p);
    r->stage_in = nullptr; r->stage_z = nullptr; r->stage_out = nullptr; r->stage_n = 0;
    CU_TRY(cudaMalloc(&r->stage_in, (size_t)n * r->a.n_lam * sizeof(float)));
    CU_TRY(cudaMalloc(&r->stage_z, (size_t)n * sizeof(double)));
    CU_TRY(cudaMalloc(&r->stage_out, (size_t)n * r->a.n_px * sizeof(float)));
    r->stage_n = n;
  }
  CU_TRY(cudaMemcpy(r->stage_in, spectra, (size_t)n * r->a.n_lam * sizeof(float), cudaMemcpyHostToDevice));
  CU_TRY(cudaMemcpy(r->stage_z, redshift, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  int rc = sb2_resample_spectra(r, r->stage_in, r->stage_z, n, r->stage_out, nullptr);
  if (rc != SB2_OK) return rc;
  CU_TRY(cudaMemcpy(out, r->stage_out, (size_t)n * r->a.n_px * sizeof(float), cudaMemcpyDeviceToHost));
  return SB2_OK;
}

// ---- empirical uncertainty models ---------------------------------------------------------------------------------
static_assert(SB2_EMP_MAX_BINS == sb2::kEmpMaxBins, "header and kernel disagree on the table size");

int sb2_empirical_noise(const double* flux, int64_t n, int32_t n_filt, const sb2_empirical_model* models,
                        const double* draws, uint64_t seed, uint64_t epoch, double* out_flux, double* out_sigma,
                        void* stream) {
  if (!flux || !models || !out_flux || n < 0 || n_filt < 1) return fail(SB2_ERR_INVALID, "bad argument");
  if (n == 0) return SB2_OK;
  std::vector<sb2::EmpiricalModelDev> h((size_t)n_filt);
  for (int f = 0; f < n_filt; ++f) {
    const sb2_empirical_model& m = models[f];
    if (m.n_bins < 2 || m.n_bins > SB2_EMP_MAX_BINS) return fail(SB2_ERR_INVALID, "empirical model: 2 <= n_bins <= SB2_EMP_MAX_BINS");
    for (int i = 1; i < m.n_bins; ++i)
      if (!(m.centers[i] > m.centers[i - 1])) return fail(SB2_ERR_INVALID, "empirical model: bin centres must increase");
    if ((!m.internal_is_ab && m.asinh_mode != 1 && !(m.internal_to_jy > 0)) || (!m.in_is_ab && !(m.in_to_jy > 0)) || (!m.out_is_ab && !(m.out_to_jy > 0)))
      return fail(SB2_ERR_INVALID, "empirical model: linear units need a positive size in Jy");
    sb2::EmpiricalModelDev& d = h[f];
    d.n_bins = m.n_bins; d.extrapolate = m.extrapolate; d.internal_is_ab = m.internal_is_ab; d.in_is_ab = m.in_is_ab;
    d.out_is_ab = m.out_is_ab; d.observed_error = m.observed_error; d.upper_limits = m.upper_limits; d.ul_active = m.ul_active;
    d.internal_to_jy = m.internal_to_jy; d.in_to_jy = m.in_to_jy; d.out_to_jy = m.out_to_jy; d.sigma_clip = m.sigma_clip;
    d.snr_threshold = m.snr_threshold; d.ul_flux = m.ul_flux; d.ul_scatter_std = m.ul_scatter_std; d.ul_err = m.ul_err;
    d.min_err = m.min_err; d.max_err = m.max_err;
    d.asinh_mode = m.asinh_mode; d.pad_ = 0; d.asinh_b = m.asinh_b;
    if (m.asinh_mode < 0 || m.asinh_mode > 2 || (m.asinh_mode && !(m.asinh_b > 0)))
      return fail(SB2_ERR_INVALID, "empirical model: asinh_mode in {0, 1, 2}, asinh_b > 0");
    if (m.asinh_mode == 2 && !(m.internal_to_jy > 0)) return fail(SB2_ERR_INVALID, "empirical model: asinh_mode 2 needs a linear interpolation unit");
    std::memcpy(d.centers, m.centers, sizeof(d.centers));
    std::memcpy(d.median, m.median, sizeof(d.median));
    std::memcpy(d.stdev, m.stdev, sizeof(d.stdev));
  }
  cudaStream_t st = (cudaStream_t)stream;
  sb2::EmpiricalModelDev* dm = nullptr;
  CU_TRY(cudaMallocAsync(reinterpret_cast<void**>(&dm), h.size() * sizeof(sb2::EmpiricalModelDev), st));
  cudaError_t e = cudaMemcpyAsync(dm, h.data(), h.size() * sizeof(sb2::EmpiricalModelDev), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);    // h is pageable and goes out of scope
  if (e != cudaSuccess) { cudaFreeAsync(dm, st); return fail(SB2_ERR_CUDA, cudaGetErrorString(e)); }
  sb2::EmpiricalArgs a{};
  a.flux = flux; a.n = n; a.n_filt = n_filt; a.models = dm; a.draws = draws; a.seed = seed; a.epoch = epoch;
  a.out_flux = out_flux; a.out_sigma = out_sigma;
  for (int r = 0; r < 10; ++r) {   // key (seed_lo, seed_hi ^ epoch_hi), bumped per round by Philox's Weyl constants
    a.rk0[r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
    a.rk1[r] = ((uint32_t)(seed >> 32) ^ (uint32_t)(epoch >> 32)) + (uint32_t)r * 0xBB67AE85u;
  }
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  // parity mode: one thread per element; production mode: one thread per pair of rows
  long long bx = ((draws ? n : (n + 1) / 2) + 255) / 256;
  const long long cap = std::max<long long>(1, (long long)n_sm * 8 / n_filt);
  if (bx > cap) bx = cap;
  if (draws) sb2::empirical_noise_kernel<<<dim3((unsigned)bx, (unsigned)n_filt), 256, 0, st>>>(a);
  else sb2::empirical_noise_fast_kernel<<<dim3((unsigned)bx, (unsigned)n_filt), 256, 0, st>>>(a);
  e = cudaGetLastError();
  cudaFreeAsync(dm, st);
  if (e != cudaSuccess) return fail(SB2_ERR_CUDA, cudaGetErrorString(e));
  return SB2_OK;
}

// ---- filter integration on a general wavelength axis ----------------------------------------------------------------
struct sb2_filterset {
  int device = 0, n_sm = 148, n_filt = 0, n_lam = 0, variant = 0;
  double* lam = nullptr; long long* off = nullptr; double* f_lam = nullptr; double* f_t = nullptr;
};

int sb2_filterset_create(int32_t n_filt, const int64_t* offsets, const double* filt_lam, const double* filt_t,
                         const double* grid_lam, int32_t n_lam, int32_t variant, int device, sb2_filterset** out) {
  if (!offsets || !filt_lam || !filt_t || !grid_lam || !out || n_filt < 1 || n_lam < 2) return fail(SB2_ERR_INVALID, "bad argument");
  for (int f = 0; f < n_filt; ++f) {
    if (offsets[f + 1] - offsets[f] < 2) return fail(SB2_ERR_INVALID, "a filter table needs at least two samples");
    for (int64_t i = offsets[f] + 1; i < offsets[f + 1]; ++i)
      if (!(filt_lam[i] >= filt_lam[i - 1])) return fail(SB2_ERR_INVALID, "filter wavelengths must not decrease");
  }
  for (int i = 1; i < n_lam; ++i)
    if (!(grid_lam[i] > grid_lam[i - 1])) return fail(SB2_ERR_INVALID, "the wavelength axis must increase");
  CU_TRY(cudaSetDevice(device));
  sb2_filterset* s = new sb2_filterset();
  s->device = device; s->n_filt = n_filt; s->n_lam = n_lam; s->variant = variant;
  cudaDeviceGetAttribute(&s->n_sm, cudaDevAttrMultiProcessorCount, device);
  std::vector<long long> off(offsets, offsets + n_filt + 1);
  int rc = upload(&s->lam, grid_lam, (size_t)n_lam);
  if (rc == SB2_OK) rc = upload(&s->off, off.data(), off.size());
  if (rc == SB2_OK) rc = upload(&s->f_lam, filt_lam, (size_t)offsets[n_filt]);
  if (rc == SB2_OK) rc = upload(&s->f_t, filt_t, (size_t)offsets[n_filt]);
  if (rc != SB2_OK) { sb2_filterset_destroy(s); return rc; }
  *out = s;
  return SB2_OK;
}

int sb2_filterset_destroy(sb2_filterset* s) {
  if (!s) return SB2_OK;
  cudaSetDevice(s->device);
  for (void* p : {(void*)s->lam, (void*)s->off, (void*)s->f_lam, (void*)s->f_t}) if (p) cudaFree(p);
  delete s;
  return SB2_OK;
}

int sb2_filter_integrate(sb2_filterset* s, const float* spectra, const double* redshift, const double* log_mass, double base_mass,
                         int64_t n, float* flux_base, double* flux_scaled, void* stream) {
  if (!s || !spectra || !redshift || n < 0) return fail(SB2_ERR_INVALID, "bad argument");
  if (!flux_base && !flux_scaled) return fail(SB2_ERR_INVALID, "no output requested");
  if (n == 0) return SB2_OK;
  CU_TRY(cudaSetDevice(s->device));
  sb2::GeneralFilterArgs a{};
  a.spectra = spectra; a.redshift = redshift; a.log_mass = log_mass; a.base_mass = base_mass; a.n = n;
  a.n_lam = s->n_lam; a.n_filt = s->n_filt; a.variant = s->variant; a.lam = s->lam; a.off = s->off; a.f_lam = s->f_lam; a.f_t = s->f_t;
  a.flux_base = flux_base; a.flux_scaled = flux_scaled;
  const long long warps = (long long)n * s->n_filt;
  const long long blocks = std::min<long long>((warps + 7) / 8, (long long)s->n_sm * 16);
  cudaStream_t st = (cudaStream_t)stream;
  sb2::general_filter_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  STAGE_CHECK("general_filter_kernel", st);
  return SB2_OK;
}

}  // extern "C"
