// Contraction + fused epilogue kernel (the hot path of the hot path).
//
//   S[g, lam] = sum_k W[g, k] * G[k, lam]      (N_gal x K) . (K x N_lam)
//
// is the one dense contraction of the reference path (grid-weighted spectral sum, SURVEY A4;
// Pipeline.run at library.py:2619).  It runs on the 5th-gen tensor cores as 3xTF32
// (W_hi*G_hi + W_hi*G_lo + W_lo*G_hi, FP32 accumulate in TMEM) so the result carries ~2^-22
// relative error instead of TF32's 2^-11.  Structure (one persistent CTA per SM):
//
//   warp 0      TMA producer   : 128-galaxy x 32-k tile of W (hi, lo) and 256-row x 32-k tile
//                                of G^T (hi, lo) per stage, SWIZZLE_128B, mbarrier completion
//   warp 1      MMA issuer     : 12 tcgen05.mma (M128 N256 K8, kind::tf32) per stage into one of
//                                two 256-column TMEM accumulators
//   warp 2      TMEM allocator
//   warps 4-11  fused epilogue : two warpgroups, one per TMEM accumulator (even / odd chunks), so each
//                                scheduler has two warps to hide TMEM / L1 / MUFU latency.  Thread t of
//                                a group owns galaxy t of the tile (= TMEM lane t); it streams its
//                                spectrum out of TMEM 32 wavelengths at a time and applies dust
//                                attenuation exp(-tau_V kappa), component mixing, the IGM row, and
//                                accumulates the trapezoidal filter numerators with the per-galaxy
//                                (m, beta) shift of the filter tables -- the spectrum never goes to
//                                shared or global memory.  Each group writes its partial numerators
//                                (numU, numV per filter); finalize_kernel adds the two and scales.
//
// A wavelength chunk is 256 accumulator columns: 256 wavelengths for one spectral component, or
// 128 wavelengths x 2 components (attenuated | unattenuated) when the emission recipe needs both.
// Tiles hold galaxies of one metallicity bracket in redshift order (capi.cu: group_*_kernel), so
//   * a DeltaConstant tile multiplies only the 2*n_age_pad grid columns it can touch (tile_k0),
//   * the filter windows of a warp's 32 galaxies coincide (warp-uniform window tests), and
//   * wavelength chunks no filter of the tile reaches are skipped altogether (tile_range).
#pragma once
#include <climits>
#include "ptx.cuh"

namespace sb2 {

constexpr int kBM = 128;                       // galaxies per tile
constexpr int kBN = 256;                       // accumulator columns per chunk
constexpr int kBK = 32;                        // k per stage (32 x 4 B = one 128 B swizzle row)
constexpr int kStages = 2;
constexpr int kABytes = kBM * kBK * 4;         // 16 KiB
constexpr int kBBytes = kBN * kBK * 4;         // 32 KiB
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;  // hi + lo of both operands = 96 KiB
// Bottleneck experiments (skip filter sums / dust exp / IGM / TMEM loads / MMAs / G or W traffic): compiled in only
// with -DSB2_EXPERIMENTS, selected at run time with the SB2_DBG bit mask.
#ifdef SB2_EXPERIMENTS
#define SB2_DBG_BITS(A) ((A).dbg)
#else
#define SB2_DBG_BITS(A) 0
#endif

#ifndef SB2_GROUPS
#define SB2_GROUPS 2
#endif
constexpr int kMaxGroups = SB2_GROUPS;         // epilogue warpgroups in the CTA (partial-numerator planes)
constexpr int kEpiWarp0 = 3;                   // first epilogue warp (warps 0-2: TMA producer, MMA issuer, TMEM allocator)
constexpr int kSynthThreads = 32 * kEpiWarp0 + 128 * kMaxGroups;
constexpr int kSpecSmemBytes = 8 * kMaxGroups / 2 * 32 * 33 * 4;   // one 32 x 33 float tile per epilogue warp
constexpr int kBarBytes = 256;                 // barriers + TMEM slot
constexpr int kTfPerGroup = 4;                 // "accumulator ready" barriers per epilogue group: one per chunk that can be
                                               // outstanding (<= number of TMEM accumulators), so no barrier is ever committed
                                               // twice before its group has seen the first completion
constexpr int kMaxFilt = 32;
constexpr int kUvPad = 40;       // zero entries on both sides of every filter's (U, V) table
constexpr int kFastSpread = 10;  // max (m_max - m_min) within a warp for the unclamped table reads

struct SynthArgs {
  int n_gal, n_tiles, n_chunk, n_kb, n_lam, n_filt, n_blue, n_blue_pad, uv_len;
  int k8_total;              // K/8 MMA steps actually needed (the last k-block may be partial)
  int dbg;                   // experiments only (SB2_DBG): 1 skip filter sums, 2 skip dust exp, 4 two ring slots
  int two_pass;              // 1: long K, cross terms summed before the hi*hi terms (single-CTA kernel only)
  int n_stages;              // split-accumulator kernel with cross and two_pass: ring depth of HALF stages (see synth_kernel); else unused
  int cross;                 // 1: the W_lo / G_lo operands hold the packed bfloat16 pairs [lo | hi] x [hi ; lo] of the small
                             //    terms, multiplied by ONE kind::f16 MMA per 8 k-values (synth_kernel; PrepModel.cross)
  const int* n_tiles_dev;    // actual tile count (<= n_tiles) when the batch was grouped on device, else nullptr
  const int* tile_k0;        // [n_tiles] first grid column (k) of each tile's weights, nullptr: 0
  const int4* tile_range;    // [n_tiles] {first, last wavelength chunk any filter of the tile needs, first bin & ~31, last bin}; nullptr: all
  const float* kappa;        // [n_chunk * lam_per_chunk], zero padded
  const float* dust_d0;      // per-galaxy dust shape: tau/tau_V = (kappa + ampl_g dust_d0) 2^(slope_g dust_l2); nullptr: global
  const float* dust_l2;
  const float* g_slope;
  const float* g_ampl;
  const float* kappa_birth;  // second screen (young population): nullptr = component 2 is unattenuated
  const float* g_taub;
  const float* wnu;          // dust emission: trapezoid weights in frequency for the absorbed-energy sum; nullptr: none
  float* e_part;             // [kMaxGroups][n_rows] each epilogue group's share of sum_i (unattenuated - attenuated)_i wnu_i
  const float* g_lya;        // per-galaxy Lyman-alpha line term, added to the first component at bin lya_bin; nullptr: none
  int lya_bin;
  // Absorbed energy from PSEUDO-BINS (synth3_kernel only; x_count = 0: none).  The absorbed-energy sum over the whole axis,
  // sum_i wnu_i v_i (1 - exp(-tau kappa_i)), depends on a wavelength only through kappa_i, so the host projects the grid
  // onto x_bins nodes in kappa (degree-7 Lagrange weights; engine.py) and appends them to the axis as extra rows of the
  // grid with kappa = the node and wnu = 1.  Every tile then multiplies x_count extra chunks [x_first, x_first + x_count)
  // (units of the launched kernel's chunk) after the chunks its filters need, the energy sum runs on those chunks only,
  // and the real chunks keep their skipping and the plain epilogue.
  int x_first, x_count;
  int kap_len;               // length of the kappa / wnu / ... tables (real axis padded + pseudo-bins)
  // synth3_kernel, fused output: the two epilogue groups exchange their filter numerators through shared memory at the end
  // of a tile and group 0 writes the fluxes itself (what finalize_kernel does from the `part` planes in HBM otherwise)
  int fuse_out;
  long long scaled_ld;       // layout of out_scaled for the fused output (FinalizeArgs.scaled_ld)
  const float2* dust_duv;    // fused output with dust emission: [dust_m_len][n_filt] (FinalizeArgs.dust_duv); nullptr: none
  int dust_m_len;
  const float2* filt_uv;     // padded tables, uv_len entries
  const float* igm;          // [n_tiles][n_blue_pad][128]
  const int* g_m;
  const float* g_beta;   // blend weight of the filter sample n+1
  const float* g_gamma;  // 1 - beta, rounded from float64 (no cancellation at the band edges)
  const float* g_taut;
  const float* g_scale;
  const float* g_ca;
  const float* g_cb;
  const int* g_orig;
  const double* g_mscale;
  const unsigned* g_trunc;
  float2* part;          // [kMaxGroups][n_filt][n_rows] partial filter numerators of the epilogue groups
  long long n_rows;      // padded rows (n_tiles * 128)
  float* out_base;
  double* out_scaled;
  float* out_spec;
  int spec_smem;         // 1: the launch reserved kSpecSmemBytes after the barriers for the spectra transpose tiles
  int filt_lo[kMaxFilt], filt_hi[kMaxFilt], filt_off[kMaxFilt];  // filt_off: start of the PADDED table
  float filt_su[kMaxFilt], filt_sdv[kMaxFilt];
};

__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 r;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr));
  return r;
}
// acc.{x,y} += s * uv.{x,y}   (one FFMA2 on sm_100)
__device__ __forceinline__ void ffma2_bcast(float2& acc, float s, float2 uv) {
  uint64_t a, b, c;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(s));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(uv.x), "f"(uv.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc.x), "f"(acc.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(c) : "l"(a), "l"(b), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(c));
}

// tau(lambda)/tau_V of a galaxy with its own slope and bump amplitude (Noll+09 form): the helper curve is
// linear in the amplitude, the slope is a power law in lambda / 0.55 um.
__device__ __forceinline__ float4 dust_shape(float4 k, float4 d, float4 l, float slope, float ampl) {
  return make_float4(fmaf(ampl, d.x, k.x) * ex2_approx(slope * l.x), fmaf(ampl, d.y, k.y) * ex2_approx(slope * l.y),
                     fmaf(ampl, d.z, k.z) * ex2_approx(slope * l.z), fmaf(ampl, d.w, k.w) * ex2_approx(slope * l.w));
}

// 1 - 2^y given T = ex2(y): the series where 1 - T would cancel (small optical depths), so that the absorbed energy of a
// nearly transparent galaxy keeps float32 relative accuracy.  (y > 0 happens: the pinned Calzetti curve is extrapolated
// linearly beyond 2.2 um and goes negative there, SURVEY A6 -- "absorption" is then negative, as in the reference.)
__device__ __forceinline__ float one_minus_ex2(float y, float T) {
  const float t = y * 0.6931471805599453f;
  float p = fmaf(t, 2.48015873e-5f, 1.98412698e-4f);
  p = fmaf(p, t, 1.38888889e-3f);
  p = fmaf(p, t, 8.33333333e-3f);
  p = fmaf(p, t, 4.16666667e-2f);
  p = fmaf(p, t, 1.66666667e-1f);
  p = fmaf(p, t, 0.5f);
  p = fmaf(p, t, 1.f);
  return fabsf(y) < 1.f ? -t * p : 1.f - T;
}

// Chunk visiting order: ascending.  (Walking the chunks cyclically from a different start per CTA, to keep SMs that work
// on neighbouring tiles off the same L2 lines, was tried: no measurable gain, and it makes the order in which a galaxy's
// chunks are summed depend on which CTA it landed on -- results would no longer be bit-identical across batch
// compositions.)
__device__ __forceinline__ int chunk_at(int c_first, int n_c, int rot, int j) { return c_first + j; }
__device__ __forceinline__ int chunk_rot(int n_c, unsigned who) { return 0; }

// Fused epilogue of one CTA (warps 3-14): see the header comment.  kCta = 2: the CTA is one half of a pair that
// shares the MMA (cta_group::2); `unit` is then a pair of tiles, this CTA owns tile 2*unit + rank and hands its
// accumulators back through the LEADER's tempty barriers (tempty_addr, shared::cluster).
// Work split: warpgroup g takes the wavelength chunks with c % kGroups == g (a fixed function of the chunk, so a
// galaxy's partial sums do not depend on what else is in the batch); the chunk's accumulator is TMEM buffer it % kBuf.
// "Accumulator ready" barriers are PER GROUP (tfull_bar[kTfPerGroup g + k % kTfPerGroup] for the group's k-th chunk): an mbarrier
// wait only carries one parity bit, so a waiter must see every phase of a barrier in order -- which a group does
// for its own pair, but would not for per-buffer barriers that other groups also consume.
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}

// Per-galaxy emission extras of the epilogue, as a compile-time feature set (synth3_kernel: one instantiation per set that
// the models actually use) -- or kFeatRuntime: the pointers in SynthArgs decide at run time (synth_kernel's kPgDust blob).
constexpr int kFeatDustShape = 1;   // per-galaxy dust-curve slope / bump amplitude (dust_d0, dust_l2, g_slope, g_ampl)
constexpr int kFeatLya = 2;         // per-galaxy Lyman-alpha line term (g_lya at lya_bin)
constexpr int kFeatTwoScreens = 4;  // birth cloud + ISM (kappa_birth, g_taub); two components, both attenuated
constexpr int kFeatAbsorbed = 8;    // energy balance: absorbed-energy sum for the dust emission (wnu, e_part)
constexpr int kFeatRuntime = -1;

// Shared-memory copies of the per-wavelength tables (byte addresses; 0: read the global array instead).  With ~225 KB of
// the SM's 228 KB configured as shared memory there is no L1 left, so every __ldg of a table was an L2 round trip per
// sub-chunk with two epilogue warps per scheduler to hide it.
struct EpiTables { uint32_t kappa, d0, l2, kappa_birth, wnu; };

template <bool kTabS>
__device__ __forceinline__ float4 epi_tab4(const float* g, uint32_t saddr, int i0, int j4) {
  if constexpr (kTabS) return lds_f4(saddr + (uint32_t)(i0 + 4 * j4) * 4u);
  else return __ldg(reinterpret_cast<const float4*>(g + i0) + j4);
}

template <int kComp, int kNF, bool kSpec, int kCta, int kN, int kGroups, bool kPgDust, int kWarp0 = kEpiWarp0, int kBufT = 512 / kN,
          int kFeat = kFeatRuntime, bool kTabS = false, int kSplit = 1>
__device__ __forceinline__ void epilogue_loop(const SynthArgs& A, const float2* s_uv, float* s_spec, uint64_t* tfull_bar,
                                              uint64_t* tempty_bar, uint32_t tempty_addr, uint32_t tmem_base, int unit0, int unit_stride,
                                              int n_units, uint32_t cta_rank, EpiTables T = EpiTables{0u, 0u, 0u, 0u, 0u},
                                              float2* s_x = nullptr) {
  constexpr int kLch = kN / kComp;      // wavelengths per chunk
  constexpr int kSub = kLch / 32;       // 32-wavelength sub-chunks per chunk
  constexpr uint32_t kBuf = kBufT;      // TMEM accumulators (2 x 256, 3 x 160 or 4 x 128 columns; synth3: what W leaves free)
  // kSplit > 1 (dense K): a chunk's sum is spread over kSplit accumulators of kN columns (K ranges; see synth_kernel), which
  // the epilogue adds in FP32 (round to nearest) as it reads them.  The chunk with sequence number `it` owns accumulators
  // (3 it + p) mod 4, p < kSplit = 3 -- the fourth is where the MMA warp already sums the NEXT chunk's cross terms -- and
  // the groups share every chunk instead of alternating: group g takes the 32-wavelength sub-chunks with sub % kGroups == g.
  constexpr bool kShare = kSplit > 1;
  static_assert(!kShare || (kSplit == 3 && kN == 128), "rotating accumulators: three of four 128-column accumulators per chunk");
  static_assert((int)kBuf <= kTfPerGroup && (kShare || kGroups <= (int)kBuf) && kGroups <= kMaxGroups, "a group must own a whole accumulator while it drains it");
  const int warp = warp_uniform((int)(threadIdx.x >> 5)), lane = threadIdx.x & 31;
  const int c_all_last = (A.n_chunk * kBN / kComp + kLch - 1) / kLch - 1;   // last chunk of kLch wavelengths on the padded axis
  const uint32_t grp = (uint32_t)(warp - kWarp0) >> 2;
  if (grp >= (uint32_t)kGroups) return;
  {
    const int et = (warp & 3) * 32 + lane;                     // galaxy within tile == TMEM lane (a warp may only touch
                                                               // the TMEM lane quarter warp % 4)
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t uv_base = smem_u32(s_uv);
    const unsigned FULL = 0xffffffffu;
    uint32_t it = 0, gk = 0;   // chunks seen by the CTA / chunks taken by this group
    bool x_seen = false;       // fused output: group 1 has handed a tile's numerators over before
    for (int unit = unit0; unit < n_units; unit += unit_stride) {
      const int tile = unit * kCta + (int)cta_rank;
      const int4 cr = A.tile_range ? __ldg(A.tile_range + unit) : make_int4(0, c_all_last, 0, 0);
      const int row = tile * kBM + et;
      const int orig = A.g_orig[row];
      int m = A.g_m[row];
      const float ntaut = -A.g_taut[row];
      constexpr bool kStatic = kFeat != kFeatRuntime;   // feature set fixed at compile time (folds every test below)
      constexpr bool pg_dust = kStatic ? (kFeat != 0) : kPgDust;   // compile-time: a run-time test here cost 10 % of the kernel (register pressure)
      // (kPgDust instantiations carry all per-galaxy emission extras: dust-curve shape and/or the Lyman-alpha line)
      const bool dust_pg = kStatic ? bool(kFeat & kFeatDustShape) : (pg_dust && A.dust_d0 != nullptr);
      const float slope = dust_pg ? A.g_slope[row] : 0.f, ampl = dust_pg ? A.g_ampl[row] : 0.f;
      const bool lya_on = kStatic ? bool(kFeat & kFeatLya) : (pg_dust && A.g_lya != nullptr);
      const float lya = lya_on ? A.g_lya[row] : 0.f;
      const bool two_screens = kStatic ? bool(kFeat & kFeatTwoScreens) && kComp == 2 : (pg_dust && kComp == 2 && A.kappa_birth != nullptr);
      const float ntaub = two_screens ? -A.g_taub[row] : 0.f;
      const bool absorbed = kStatic ? bool(kFeat & kFeatAbsorbed) : (pg_dust && A.wnu != nullptr);   // energy balance: what the dust removes, summed over the axis
      float e_abs = 0.f;
      // redshift-shift range of this warp's real galaxies (padding rows follow the others)
      int mmin = orig >= 0 ? m : INT_MAX, mmax = orig >= 0 ? m : INT_MIN;
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        mmin = min(mmin, __shfl_xor_sync(FULL, mmin, o));
        mmax = max(mmax, __shfl_xor_sync(FULL, mmax, o));
      }
      if (mmin > mmax) mmin = mmax = 0;
      if (orig < 0) m = mmin;
      const bool fast = (mmax - mmin) <= kFastSpread;
      float2 acc[kNF];
#pragma unroll
      for (int f = 0; f < kNF; ++f) acc[f] = make_float2(0.f, 0.f);

      // lane f looks after filter f: sub-chunk starting at i0 overlaps its shifted window iff
      // lo_f - 32 - mmax <= i0 <= hi_f - mmin  (warp-uniform bounds)
      const int w_lo = (lane < A.n_filt ? A.filt_lo[lane < kMaxFilt ? lane : 0] : INT_MAX / 2) - 32 - mmax;
      const int w_hi = (lane < A.n_filt ? A.filt_hi[lane < kMaxFilt ? lane : 0] : -1) - mmin;
      const int n_real = max(cr.y - cr.x + 1, 0), n_c = n_real + A.x_count;
      for (int j = 0; j < n_c; ++j, ++it) {
        const bool pseudo = j >= n_real;                       // an absorbed-energy chunk (warp-uniform)
        const int c = pseudo ? A.x_first + (j - n_real) : cr.x + j;
        const bool absorb_here = absorbed && (A.x_count == 0 || pseudo);
        if (!kShare && (uint32_t)c % (uint32_t)kGroups != grp) continue;
        const uint32_t buf = it % kBuf;
        mbar_wait(&tfull_bar[kTfPerGroup * grp + (gk % kTfPerGroup)], (gk / kTfPerGroup) & 1u, 0x600u + (it << 12));
        ++gk;
        tc_fence_after();
        if (SB2_DBG_BITS(A) & 2048) {   // experiment: hand the accumulator straight back, no epilogue work at all
          __syncwarp();
          if (lane == 0) { if constexpr (kCta == 2) mbar_arrive_cluster(tempty_addr + buf * 8u); else mbar_arrive(tempty_bar + buf); }
          continue;
        }
        const uint32_t t_lane = tmem_base + lane_base;
        const uint32_t t_acc = t_lane + (kShare ? ((3u * it) & 3u) : buf) * kN;
        const int sub_step = kShare ? kGroups : 1;
        // sub-chunks of this chunk that hold wavelengths (>= 1), and this group's first one
        const int n_sub = kShare ? max(1, min(kSub, (A.n_lam - c * kLch + 31) >> 5)) : kSub;
        if (kShare && (int)grp >= n_sub) {   // nothing of this chunk for the group: hand its accumulators back at once
          __syncwarp();
          if (lane == 0)
            for (int p = 0; p < kSplit; ++p) mbar_arrive(tempty_bar + ((3u * it + p) & 3u));
          continue;
        }
#pragma unroll 1
        for (int sub = kShare ? (int)grp : 0; sub < kSub; sub += sub_step) {
          const int i0 = c * kLch + sub * 32;
          const bool last_sub = kShare ? (sub + sub_step >= n_sub) : ((sub == kSub - 1) || (!pseudo && i0 + 32 >= A.n_lam));
          float s[32];
          float e_sub = 0.f;   // absorbed energy of this sub-chunk (blocked summation: 4 -> 32 -> axis)
          {
            // (prefetching the next sub-chunk's TMEM columns behind the filter work was measured 6% SLOWER: the
            //  tcgen05.ld latency is already covered by the other epilogue group on the same scheduler)
            uint32_t v[32];
            if (SB2_DBG_BITS(A) & 16) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0x3f800000u;
            } else {
              tmem_ld_32x32b_x32(t_acc + sub * 32, v);
              if constexpr (kSplit > 1) {   // + the other K ranges' partial sums
#pragma unroll 1
                for (int p = 1; p < kSplit; ++p) {
                  uint32_t w[32];
                  tmem_ld_32x32b_x32(t_lane + ((3u * it + p) & 3u) * kN + sub * 32, w);
                  tmem_ld_wait();
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
                }
              }
            }
            if constexpr (kComp == 2) {
              const float ca = A.g_ca[row], cb = A.g_cb[row];
              uint32_t u[32];
              tmem_ld_32x32b_x32(t_acc + kLch + sub * 32, u);
              if constexpr (kSplit > 1) {
#pragma unroll 1
                for (int p = 1; p < kSplit; ++p) {
                  uint32_t w[32];
                  tmem_ld_32x32b_x32(t_lane + ((3u * it + p) & 3u) * kN + kLch + sub * 32, w);
                  tmem_ld_wait();
#pragma unroll
                  for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(__uint_as_float(u[j]) + __uint_as_float(w[j]));
                }
              }
              tmem_ld_wait();
              if (lya_on && A.lya_bin >= i0 && A.lya_bin < i0 + 32) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (i0 + j == A.lya_bin) v[j] = __float_as_uint(__uint_as_float(v[j]) + lya);
              }
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                float4 k4 = epi_tab4<kTabS>(A.kappa, T.kappa, i0, j4);
                if (pg_dust && dust_pg) k4 = dust_shape(k4, epi_tab4<kTabS>(A.dust_d0, T.d0, i0, j4), epi_tab4<kTabS>(A.dust_l2, T.l2, i0, j4), slope, ampl);
                if (two_screens) {   // young: birth cloud + ISM, old: ISM only (both components are attenuated)
                  const float4 b4 = epi_tab4<kTabS>(A.kappa_birth, T.kappa_birth, i0, j4);
                  const float4 w4 = !absorb_here ? make_float4(0.f, 0.f, 0.f, 0.f) : pseudo ? make_float4(1.f, 1.f, 1.f, 1.f) : epi_tab4<kTabS>(A.wnu, T.wnu, i0, j4);
                  float e4 = 0.f;
#define SB2_TS(q, K, B, W) { const float yo = ntaut * (K), yy = fmaf(ntaub, (B), yo), To = ex2_approx(yo), Ty = ex2_approx(yy); \
                             const float vy = ca * __uint_as_float(v[4 * j4 + q]), vo = cb * __uint_as_float(u[4 * j4 + q]); \
                             s[4 * j4 + q] = vy * Ty + vo * To; \
                             if (absorb_here) e4 = fmaf(fmaf(vy, one_minus_ex2(yy, Ty), vo * one_minus_ex2(yo, To)), (W), e4); }
                  SB2_TS(0, k4.x, b4.x, w4.x) SB2_TS(1, k4.y, b4.y, w4.y) SB2_TS(2, k4.z, b4.z, w4.z) SB2_TS(3, k4.w, b4.w, w4.w)
#undef SB2_TS
                  e_sub += e4;
                  continue;
                }
                s[4 * j4 + 0] = ca * (__uint_as_float(v[4 * j4 + 0]) * ex2_approx(ntaut * k4.x)) + cb * __uint_as_float(u[4 * j4 + 0]);
                s[4 * j4 + 1] = ca * (__uint_as_float(v[4 * j4 + 1]) * ex2_approx(ntaut * k4.y)) + cb * __uint_as_float(u[4 * j4 + 1]);
                s[4 * j4 + 2] = ca * (__uint_as_float(v[4 * j4 + 2]) * ex2_approx(ntaut * k4.z)) + cb * __uint_as_float(u[4 * j4 + 2]);
                s[4 * j4 + 3] = ca * (__uint_as_float(v[4 * j4 + 3]) * ex2_approx(ntaut * k4.w)) + cb * __uint_as_float(u[4 * j4 + 3]);
                if (absorb_here) {
                  const float4 w4 = pseudo ? make_float4(1.f, 1.f, 1.f, 1.f) : epi_tab4<kTabS>(A.wnu, T.wnu, i0, j4);
                  float e4 = ca * __uint_as_float(v[4 * j4 + 0]) * one_minus_ex2(ntaut * k4.x, ex2_approx(ntaut * k4.x)) * w4.x;
                  e4 = fmaf(ca * __uint_as_float(v[4 * j4 + 1]) * one_minus_ex2(ntaut * k4.y, ex2_approx(ntaut * k4.y)), w4.y, e4);
                  e4 = fmaf(ca * __uint_as_float(v[4 * j4 + 2]) * one_minus_ex2(ntaut * k4.z, ex2_approx(ntaut * k4.z)), w4.z, e4);
                  e4 = fmaf(ca * __uint_as_float(v[4 * j4 + 3]) * one_minus_ex2(ntaut * k4.w, ex2_approx(ntaut * k4.w)), w4.w, e4);
                  e_sub += e4;
                }
              }
            } else {
              tmem_ld_wait();
              if (lya_on && A.lya_bin >= i0 && A.lya_bin < i0 + 32) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (i0 + j == A.lya_bin) v[j] = __float_as_uint(__uint_as_float(v[j]) + lya);
              }
              if (SB2_DBG_BITS(A) & 2) {
#pragma unroll
                for (int j = 0; j < 32; ++j) s[j] = __uint_as_float(v[j]);
              } else
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {  // ca goes into the final scale
                float4 k4 = epi_tab4<kTabS>(A.kappa, T.kappa, i0, j4);
                if (pg_dust && dust_pg) k4 = dust_shape(k4, epi_tab4<kTabS>(A.dust_d0, T.d0, i0, j4), epi_tab4<kTabS>(A.dust_l2, T.l2, i0, j4), slope, ampl);
                s[4 * j4 + 0] = __uint_as_float(v[4 * j4 + 0]) * ex2_approx(ntaut * k4.x);
                s[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) * ex2_approx(ntaut * k4.y);
                s[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) * ex2_approx(ntaut * k4.z);
                s[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) * ex2_approx(ntaut * k4.w);
                if (absorb_here) {   // (the lone component's coefficient is applied with the final scale, like the fluxes')
                  const float4 w4 = pseudo ? make_float4(1.f, 1.f, 1.f, 1.f) : epi_tab4<kTabS>(A.wnu, T.wnu, i0, j4);
                  float e4 = __uint_as_float(v[4 * j4 + 0]) * one_minus_ex2(ntaut * k4.x, ex2_approx(ntaut * k4.x)) * w4.x;
                  e4 = fmaf(__uint_as_float(v[4 * j4 + 1]) * one_minus_ex2(ntaut * k4.y, ex2_approx(ntaut * k4.y)), w4.y, e4);
                  e4 = fmaf(__uint_as_float(v[4 * j4 + 2]) * one_minus_ex2(ntaut * k4.z, ex2_approx(ntaut * k4.z)), w4.z, e4);
                  e4 = fmaf(__uint_as_float(v[4 * j4 + 3]) * one_minus_ex2(ntaut * k4.w, ex2_approx(ntaut * k4.w)), w4.w, e4);
                  e_sub += e4;
                }
              }
            }
          }
          if (lya_on && absorb_here && pseudo && j == n_real && sub == 0) {
            // the per-galaxy Lyman-alpha term is not part of the grid, hence not of the pseudo-bins: its absorbed share
            const float yl = fmaf(two_screens ? ntaub : 0.f, two_screens ? __ldg(A.kappa_birth + A.lya_bin) : 0.f,
                                  ntaut * __ldg(A.kappa + A.lya_bin));
            e_sub += (kComp == 2 ? A.g_ca[row] : 1.f) * lya * one_minus_ex2(yl, ex2_approx(yl)) * __ldg(A.wnu + A.lya_bin);
          }
          e_abs += e_sub;
          if (last_sub) {  // all TMEM reads of this accumulator are done: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (kCta == 2) mbar_arrive_cluster(tempty_addr + buf * 8u);   // the leader CTA's barrier
              else if constexpr (kShare) { for (int p = 0; p < kSplit; ++p) mbar_arrive(tempty_bar + ((3u * it + p) & 3u)); }
              else mbar_arrive(tempty_bar + buf);
            }
          }
          if (i0 < A.n_blue && !(SB2_DBG_BITS(A) & 8)) {  // rows [n_blue, n_blue_pad) of the table hold 1
            const float* ig = A.igm + ((size_t)tile * A.n_blue_pad + i0) * 128 + et;
#pragma unroll
            for (int j = 0; j < 32; ++j) s[j] *= __ldg(ig + j * 128);
          }
          if constexpr (kSpec) {
            if (A.out_spec != nullptr && !pseudo) {
              const float sc = orig < 0 ? 0.f : ((kComp == 1) ? A.g_scale[row] * A.g_ca[row] : A.g_scale[row]);
              if (s_spec != nullptr) {
                // a thread holds 32 wavelengths of ONE galaxy; transposed through a 32 x 33 tile so that each store
                // instruction writes 128 contiguous bytes of one output row (per-thread stores reached 2 % of HBM)
                float* tile = s_spec + (warp - kWarp0) * (32 * 33);
#pragma unroll
                for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = s[j] * sc;
                __syncwarp();
                const bool in_lam = i0 + lane < A.n_lam;
#pragma unroll 4
                for (int r = 0; r < 32; ++r) {
                  const int o = __shfl_sync(FULL, orig, r);
                  if (o >= 0 && in_lam) A.out_spec[(size_t)o * A.n_lam + i0 + lane] = tile[r * 33 + lane];
                }
                __syncwarp();
              } else if (orig >= 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (i0 + j < A.n_lam) A.out_spec[(size_t)orig * A.n_lam + i0 + j] = s[j] * sc;
              }
            }
          }
          // filter numerators: (numU_f, numV_f) += s_i * (U_f[n], V_f[n]),  n = i + m ; loop over the filters whose
          // window overlaps this sub-chunk (warp-uniform), compact code: one body, accumulator picked by a switch
          unsigned fm = __ballot_sync(FULL, !pseudo && i0 >= w_lo && i0 <= w_hi);
          if (SB2_DBG_BITS(A) & 1) fm = 0u;
#pragma unroll 1
          while (fm != 0u) {
            const int f = __ffs(fm) - 1;
            fm &= fm - 1u;
            float2 t0 = make_float2(0.f, 0.f), t1 = make_float2(0.f, 0.f);
            // padded table entry p <-> n = filt_lo - 2 - kUvPad + p
            const int p0 = kUvPad + i0 + m - (A.filt_lo[f] - 2);
            const uint32_t tab = uv_base + (uint32_t)A.filt_off[f] * 8u;
            if (fast) {  // all 32 reads stay inside the zero padding
              const uint32_t addr = tab + (uint32_t)p0 * 8u;
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                ffma2_bcast(t0, s[j], lds_f2(addr + j * 8));
                ffma2_bcast(t1, s[j + 1], lds_f2(addr + j * 8 + 8));
              }
            } else {     // same association as the fast path, indices clamped into the padding
              const int pmax = A.filt_hi[f] - A.filt_lo[f] + 3 + 2 * kUvPad;
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                ffma2_bcast(t0, s[j], lds_f2(tab + (uint32_t)min(max(p0 + j, 0), pmax) * 8u));
                ffma2_bcast(t1, s[j + 1], lds_f2(tab + (uint32_t)min(max(p0 + j + 1, 0), pmax) * 8u));
              }
            }
            t0.x += t1.x;
            t0.y += t1.y;
            switch (f) {
#define SB2_ACC(i) case i: if (i < kNF) { acc[i < kNF ? i : 0].x += t0.x; acc[i < kNF ? i : 0].y += t0.y; } break;
              SB2_ACC(0) SB2_ACC(1) SB2_ACC(2) SB2_ACC(3) SB2_ACC(4) SB2_ACC(5) SB2_ACC(6) SB2_ACC(7)
              SB2_ACC(8) SB2_ACC(9) SB2_ACC(10) SB2_ACC(11) SB2_ACC(12) SB2_ACC(13) SB2_ACC(14) SB2_ACC(15)
              SB2_ACC(16) SB2_ACC(17) SB2_ACC(18) SB2_ACC(19) SB2_ACC(20) SB2_ACC(21) SB2_ACC(22) SB2_ACC(23)
              SB2_ACC(24) SB2_ACC(25) SB2_ACC(26) SB2_ACC(27) SB2_ACC(28) SB2_ACC(29) SB2_ACC(30) SB2_ACC(31)
#undef SB2_ACC
              default: break;
            }
          }
          if (last_sub) break;
        }
      }
      if constexpr (kTabS && kGroups == 2 && !kShare && kCta == 1) {
        if (s_x != nullptr) {
          // ---- fused output: group 1 parks its numerators in shared memory (named barrier 1: "full", 2: "read"), group 0
          //      adds its own in the order finalize_kernel uses (group 0 + group 1) and writes the fluxes
          if (grp == 1) {
            if (x_seen) named_bar_sync(2, 256);
            x_seen = true;
#pragma unroll
            for (int f = 0; f < kNF; ++f)
              if (f < A.n_filt) s_x[f * kBM + et] = acc[f];
            if (absorbed) reinterpret_cast<float*>(s_x + A.n_filt * kBM)[et] = e_abs;   // (one more row of the exchange area)
            __threadfence_block();
            named_bar_arrive(1, 256);
          } else {
            named_bar_sync(1, 256);
            if (orig >= 0) {
              const float beta = A.g_beta[row], gamma = A.g_gamma[row];
              const float scf = (kComp == 1) ? A.g_scale[row] * A.g_ca[row] : A.g_scale[row];
              const unsigned trunc = A.g_trunc[row];
              const double mscale = A.out_scaled ? A.g_mscale[row] : 0.0;
              const bool tr = A.scaled_ld > 0;
              const bool vec = (A.n_filt % 4) == 0 &&
                               ((reinterpret_cast<uintptr_t>(A.out_base) | (tr ? 0 : reinterpret_cast<uintptr_t>(A.out_scaled))) & 15) == 0;
              // dust emission: E_abs (group 0 + group 1, finalize_kernel's order) times the emission's filter numerators
              // at this galaxy's integer redshift shift
              float e_tot = 0.f;
              const float2* duv = nullptr;
              if (absorbed && A.dust_duv != nullptr) {
                e_tot = e_abs + reinterpret_cast<const float*>(s_x + A.n_filt * kBM)[et];
                const int mm = A.g_m[row];
                if (mm >= 0 && mm < A.dust_m_len) duv = A.dust_duv + (size_t)mm * A.n_filt;
              }
#pragma unroll
              for (int f0 = 0; f0 < kNF; f0 += 4) {
                if (f0 < A.n_filt) {
                  float fl[4];
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    constexpr int kLast = kNF - 1;
                    const int f = f0 + q;
                    fl[q] = 0.f;
                    if (f < A.n_filt) {
                      const float2 o = s_x[f * kBM + et];
                      float nu = acc[f <= kLast ? f : kLast].x + o.x, nv = acc[f <= kLast ? f : kLast].y + o.y;
                      if (duv != nullptr) {
                        const float2 dd = __ldg(duv + f);
                        nu = fmaf(e_tot, dd.x, nu); nv = fmaf(e_tot, dd.y, nv);
                      }
                      float flux = fmaf(beta, nv, gamma * nu) / fmaf(beta, A.filt_sdv[f], gamma * A.filt_su[f]) * scf;
                      if ((trunc >> f) & 1u) flux = __int_as_float(0x7fc00000);
                      fl[q] = flux;
                    }
                  }
                  if (vec) {
                    if (A.out_base) *reinterpret_cast<float4*>(A.out_base + (size_t)orig * A.n_filt + f0) = make_float4(fl[0], fl[1], fl[2], fl[3]);
                    if (A.out_scaled && tr) {
#pragma unroll
                      for (int q = 0; q < 4; ++q) A.out_scaled[(size_t)(f0 + q) * A.scaled_ld + orig] = (double)fl[q] * mscale;
                    } else if (A.out_scaled) {
                      double2* o2 = reinterpret_cast<double2*>(A.out_scaled + (size_t)orig * A.n_filt + f0);
                      o2[0] = make_double2((double)fl[0] * mscale, (double)fl[1] * mscale);
                      o2[1] = make_double2((double)fl[2] * mscale, (double)fl[3] * mscale);
                    }
                  } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                      if (f0 + q < A.n_filt) {
                        if (A.out_base) A.out_base[(size_t)orig * A.n_filt + f0 + q] = fl[q];
                        if (A.out_scaled)
                          A.out_scaled[tr ? (size_t)(f0 + q) * A.scaled_ld + orig : (size_t)orig * A.n_filt + f0 + q] = (double)fl[q] * mscale;
                      }
                  }
                }
              }
            }
            if (unit + unit_stride < n_units) named_bar_arrive(2, 256);
          }
          continue;
        }
      }
      // ---- this group's partial numerators (finalize_kernel adds the two groups and scales)
#pragma unroll
      for (int f = 0; f < kNF; ++f)
        if (f < A.n_filt) A.part[((size_t)grp * A.n_filt + f) * A.n_rows + row] = acc[f];
      if (absorbed) A.e_part[(size_t)grp * A.n_rows + row] = e_abs;
    }
  }
}

// kN = accumulator columns per chunk: 256 (two TMEM accumulators) or, for one spectral component, 160 (THREE
// accumulators: the MMA warp always has a free one, so an epilogue group that hands an accumulator back finds its next
// chunk already multiplied instead of waiting one MMA time for it).
template <int kN>
struct SynthCfg {
  static constexpr int kBBytesN = kN * kBK * 4;
  static constexpr int kStageBytesN = 2 * kABytes + 2 * kBBytesN;
  static constexpr int kBufN = 512 / kN;
};

// kSplit = 3 (dense K, kN = 128): the W_hi*G_hi terms of a chunk are spread over three accumulators by K range (the small
// cross terms go to the first), so each running sum -- and with it the truncation of every tensor-core accumulation -- is
// three times smaller; the epilogue adds the partial sums.  TMEM holds four such accumulators and chunk `it` takes numbers
// (3 it + p) mod 4: the one left over is the FIRST accumulator of chunk it + 1, whose cross-term pass (two thirds of a
// chunk's MMAs) therefore runs while the epilogue drains chunk it -- the tensor pipe never waits for the epilogue.
// G is fetched as 64-row boxes (the pair kernel's maps).
template <int kComp, int kNF, bool kSpec, int kN, bool kPgDust, int kSplit = 1>
__global__ void __launch_bounds__(kSynthThreads, 1)
synth_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
             const __grid_constant__ CUtensorMap tm_g_hi, const __grid_constant__ CUtensorMap tm_g_lo,
             const __grid_constant__ SynthArgs A) {
  static_assert(kN == kBN || kComp == 1 || kSplit > 1, "the grid's two-component row layout is built for 256-column chunks");
  static_assert(kSplit == 1 || kN == 128, "split accumulators: 64-row boxes of G, two per operand");
  constexpr int kBBytes = SynthCfg<kN>::kBBytesN;
  constexpr int kStageBytes = SynthCfg<kN>::kStageBytesN;
  constexpr uint32_t kBuf = SynthCfg<kN>::kBufN;
  constexpr int kLch = kN / kComp;  // wavelengths per chunk
  constexpr int kSub = kLch / 32;    // 32-wavelength sub-chunks per chunk
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve-up: [stages][W_hi | W_lo | G_hi | G_lo], filter table, barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // Half stages (split-accumulator kernel, bfloat16 small terms, two passes): a pass needs only ONE tile per operand -- the
  // packed pair tiles in pass 0, the hi tiles in pass 1 -- so a stage is [W tile | G tile] and the same shared memory holds
  // twice as many k-blocks in flight (the two-stage ring of four-tile stages left this kernel waiting on L2 latency).
  const bool half = kSplit > 1 && A.cross != 0 && A.two_pass != 0;
  const int n_stages = half ? A.n_stages : kStages;
  const int stage_bytes = half ? kABytes + kBBytes : kStageBytes;
  constexpr int kMaxStages = kSplit > 1 ? 8 : kStages;
  float2* s_uv = reinterpret_cast<float2*>(smem + n_stages * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + n_stages * stage_bytes + ((A.uv_len * 8 + 15) & ~15));
  uint64_t* full_bar = bars;                    // [kMaxStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kMaxStages;      // [kMaxStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kMaxStages;                    // [kTfPerGroup * kMaxGroups] MMA -> epilogue group (see epilogue_loop)
  uint64_t* tempty_bar = tfull_bar + kTfPerGroup * kMaxGroups;    // [kBuf]           epilogue -> MMA, per TMEM accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 4);

  const int warp = warp_uniform((int)(threadIdx.x >> 5)), lane = threadIdx.x & 31;   // uniform: see synth3_kernel

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_w_hi); prefetch_tmap(&tm_w_lo); prefetch_tmap(&tm_g_hi); prefetch_tmap(&tm_g_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < kTfPerGroup * kMaxGroups; ++b) mbar_init(&tfull_bar[b], 1);
    for (int b = 0; b < (kSplit > 1 ? 4 : (int)kBuf); ++b) mbar_init(&tempty_bar[b], kSplit > 1 ? 8 : 4);  // 4 warps per epilogue group (split: both groups drain every chunk)
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < A.uv_len; i += kSynthThreads) s_uv[i] = A.filt_uv[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = A.n_tiles_dev ? min(A.n_tiles, __ldg(A.n_tiles_dev)) : A.n_tiles;
  const int c_all_last = (A.n_chunk * kBN / kComp + kLch - 1) / kLch - 1;

  if (warp == 0) {
    // ===================================================================== TMA producer
    // warp-uniform loop, one elected lane issues (see ptx.cuh).  two_pass (dense K): pass 0 streams all four
    // operand tiles of a k-block for the two small cross terms, pass 1 streams the hi tiles again for hi*hi.
    {
      const uint32_t elected = elect_one() ? 1u : 0u;
      const uint32_t s_addr = smem_u32(smem), full0 = smem_u32(full_bar);
      const int n_tiles_u = warp_uniform(n_tiles), n_kb = A.n_kb, n_pass = A.two_pass ? 2 : 1;
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles_u; tile += gridDim.x) {
        const int k0 = warp_uniform(A.tile_k0 ? __ldg(A.tile_k0 + tile) : 0);
        const int4 cr = A.tile_range ? __ldg(A.tile_range + tile) : make_int4(0, c_all_last, 0, 0);
        const int c_first = warp_uniform(cr.x), c_last = warp_uniform(cr.y);
        const int n_c = c_last - c_first + 1, rot = chunk_rot(n_c, blockIdx.x);
        for (int j = 0; j < n_c; ++j) {
          const int c = chunk_at(c_first, n_c, rot, j);
          for (int pass = 0; pass < n_pass; ++pass) {
            const bool lo_tiles = (pass == 0);   // single pass: everything in pass 0
            // bfloat16 small terms (SynthArgs.cross): the packed "lo" tiles carry both factors of the small terms, so the
            // first of two passes does not fetch the hi tiles at all (4 instead of 6 tile fetches per k-block and chunk)
            const bool hi_tiles = !(A.cross && n_pass == 2 && pass == 0);
            for (int kb = 0; kb < n_kb; ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              const uint32_t st = s_addr + (uint32_t)(stage * stage_bytes), fb = full0 + (uint32_t)stage * 8u;
              if (SB2_DBG_BITS(A) & 64) {   // experiment: no operand traffic
                mbar_expect_tx_e(elected, &full_bar[stage], 0);
              } else if (kSplit > 1 && half) {
                // half stage: [W tile | G tile (two 64-row boxes)], packed pair tiles in pass 0, hi tiles in pass 1
                const CUtensorMap* tw = pass == 0 ? &tm_w_lo : &tm_w_hi;
                const CUtensorMap* tg = pass == 0 ? &tm_g_lo : &tm_g_hi;
                const int r0 = kComp == 1 ? c * kN : (c * kLch / (kBN / 2)) * kBN + (c * kLch) % (kBN / 2);
                const int r1 = kComp == 1 ? r0 + 64 : r0 + kBN / 2;
                mbar_expect_tx_e(elected, &full_bar[stage], kABytes + kBBytes);
                tma_load_2d_e(elected, st, tw, fb, kb * kBK, tile * kBM, kEvictNormal);
                tma_load_2d_e(elected, st + kABytes, tg, fb, k0 + kb * kBK, r0, kEvictLast);
                tma_load_2d_e(elected, st + kABytes + kBBytes / 2, tg, fb, k0 + kb * kBK, r1, kEvictLast);
              } else {
              mbar_expect_tx_e(elected, &full_bar[stage], (lo_tiles && hi_tiles) ? kStageBytes : kABytes + kBBytes);
              if (hi_tiles) tma_load_2d_e(elected, st, &tm_w_hi, fb, kb * kBK, tile * kBM, kEvictNormal);
              if (lo_tiles) tma_load_2d_e(elected, st + kABytes, &tm_w_lo, fb, kb * kBK, tile * kBM, kEvictNormal);
              if constexpr (kSplit > 1) {
                // rows of the chunk: one component = kLch consecutive wavelengths; with two components the grid's rows come
                // in blocks of 256 [component 0: 128 wavelengths | component 1: the same 128] (as in synth3_kernel)
                const int r0 = kComp == 1 ? c * kN : (c * kLch / (kBN / 2)) * kBN + (c * kLch) % (kBN / 2);
                const int r1 = kComp == 1 ? r0 + 64 : r0 + kBN / 2;
                if (hi_tiles) {
                  tma_load_2d_e(elected, st + 2 * kABytes, &tm_g_hi, fb, k0 + kb * kBK, r0, kEvictLast);
                  tma_load_2d_e(elected, st + 2 * kABytes + kBBytes / 2, &tm_g_hi, fb, k0 + kb * kBK, r1, kEvictLast);
                }
                if (lo_tiles) {
                  tma_load_2d_e(elected, st + 2 * kABytes + kBBytes, &tm_g_lo, fb, k0 + kb * kBK, r0, kEvictLast);
                  tma_load_2d_e(elected, st + 2 * kABytes + kBBytes + kBBytes / 2, &tm_g_lo, fb, k0 + kb * kBK, r1, kEvictLast);
                }
              } else {
                if (hi_tiles) tma_load_2d_e(elected, st + 2 * kABytes, &tm_g_hi, fb, k0 + kb * kBK, c * kN, kEvictLast);
                if (lo_tiles) tma_load_2d_e(elected, st + 2 * kABytes + kBBytes, &tm_g_lo, fb, k0 + kb * kBK, c * kN, kEvictLast);
              }
              }
              if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // FP32 accumulation in the tensor core truncates, so the error grows with the number of accumulations made
    // onto the (large) running sum: with a long K the small cross terms W_lo*G_hi + W_hi*G_lo are summed FIRST
    // (two_pass), while the accumulator is still small, and the W_hi*G_hi terms after.
    {
      constexpr uint32_t idesc = make_idesc_tf32(kBM, kN), idesc_x = make_idesc_bf16(kBM, kN);
      const bool cross = A.cross != 0;
      const uint32_t elected = elect_one() ? 1u : 0u;
      const uint32_t s_addr = smem_u32(smem);
      const uint64_t desc0 = make_kmajor_sw128_desc(0);
      const uint32_t tmem_u = (uint32_t)warp_uniform((int)tmem_base);
      const int n_tiles_u = warp_uniform(n_tiles), n_kb = A.n_kb, k8_total = A.k8_total, two_pass = A.two_pass;
      int stage = 0; uint32_t phase = 0; uint32_t it = 0;
      uint32_t gk0 = 0, gk1 = 0;   // chunks handed to epilogue group 0 / 1 so far
      uint32_t acc_par = 0;        // split accumulators: bit x = parity of the uses of accumulator x so far
      for (int tile = blockIdx.x; tile < n_tiles_u; tile += gridDim.x) {
        const int4 cr = A.tile_range ? __ldg(A.tile_range + tile) : make_int4(0, c_all_last, 0, 0);
        const int c_first = warp_uniform(cr.x), c_last = warp_uniform(cr.y);
        const int n_c = c_last - c_first + 1, rot = chunk_rot(n_c, blockIdx.x);
        for (int j = 0; j < n_c; ++j, ++it) {
          const int c = chunk_at(c_first, n_c, rot, j);
          const uint32_t buf = kSplit > 1 ? ((3u * it) & 3u) : it % kBuf;
          if constexpr (kSplit > 1) {
            mbar_wait(&tempty_bar[buf], ((acc_par >> buf) & 1u) ^ 1u);
            acc_par ^= 1u << buf;
          } else {
            mbar_wait(&tempty_bar[buf], ((it / kBuf) & 1u) ^ 1u);  // epilogue has drained this accumulator
          }
          tc_fence_after();
          const uint32_t d_tmem = tmem_u + buf * kN;
          for (int pass = 0; pass <= two_pass; ++pass) {
            for (int kb = 0; kb < n_kb; ++kb) {
              // split accumulators: k-block kb's hi*hi terms go to accumulator kb * kSplit / n_kb; the first MMA into
              // accumulators 1.. overwrites (accumulator 0 already holds cross terms by then)
              const int part = kSplit > 1 ? (kb * kSplit) / n_kb : 0;
              const bool part_first = kSplit > 1 && part > 0 && ((kb - 1) * kSplit) / n_kb != part && pass == two_pass;
              const uint32_t acc_hh = (buf + (uint32_t)part) & 3u;
              const uint32_t d_hh = kSplit > 1 ? tmem_u + acc_hh * kN : d_tmem;
              if (part_first) {   // (the previous chunk's epilogue has long finished with it: see the header comment)
                mbar_wait(&tempty_bar[acc_hh], ((acc_par >> acc_hh) & 1u) ^ 1u);
                acc_par ^= 1u << acc_hh;
              }
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint64_t da = desc0 + (uint64_t)(((s_addr + stage * stage_bytes) & 0x3FFFF) >> 4);
              const uint64_t db = da + (uint64_t)(((half ? 1 : 2) * kABytes) >> 4);
              const int k4n = min(kBK / 8, k8_total - kb * (kBK / 8));
#pragma unroll
              for (int k4 = 0; k4 < kBK / 8; ++k4) {
                if (k4 < k4n) {
                  // (half stages: the one W tile and the one G tile of the pass, whichever they are)
                  const uint64_t a_hi = da + (uint64_t)(k4 * 2), a_lo = half ? a_hi : da + (uint64_t)((kABytes >> 4) + k4 * 2);
                  const uint64_t b_hi = db + (uint64_t)(k4 * 2), b_lo = half ? b_hi : db + (uint64_t)((kBBytes >> 4) + k4 * 2);
                  if (!(SB2_DBG_BITS(A) & 32)) {
                  if (pass == 0) {
                    if (cross) {
                      umma_bf16_e(elected, d_tmem, a_lo, b_lo, idesc_x, (kb | k4) != 0);  // [w_lo | w_hi] x [g_hi ; g_lo], K = 16
                    } else {
                      umma_tf32_e(elected, d_tmem, a_lo, b_hi, idesc, (kb | k4) != 0);  // small terms first
                      umma_tf32_e(elected, d_tmem, a_hi, b_lo, idesc, 1u);
                    }
                  }
                  if (pass == two_pass) umma_tf32_e(elected, d_hh, a_hi, b_hi, idesc, (part_first && k4 == 0) ? 0u : 1u);
                  }
                }
              }
              umma_commit_e(elected, &empty_bar[stage]);  // smem slot reusable once these MMAs retire
              if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
          }
          if (kSplit > 1) {   // both groups drain every chunk
            umma_commit_e(elected, &tfull_bar[gk0 % kTfPerGroup]); ++gk0;
            umma_commit_e(elected, &tfull_bar[kTfPerGroup + gk1 % kTfPerGroup]); ++gk1;
          } else
          if ((c & 1) == 0) { umma_commit_e(elected, &tfull_bar[gk0 % kTfPerGroup]); ++gk0; }                 // accumulator complete:
          else              { umma_commit_e(elected, &tfull_bar[kTfPerGroup + gk1 % kTfPerGroup]); ++gk1; }   // wake the group that owns chunk c
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    float* s_spec = (kSpec && A.spec_smem) ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes) : nullptr;
    epilogue_loop<kComp, kNF, kSpec, 1, kN, 2, kPgDust, kEpiWarp0, (int)kBuf, kFeatRuntime, false, kSplit>(A, s_uv, s_spec, tfull_bar, tempty_bar, 0u, tmem_base, (int)blockIdx.x, (int)gridDim.x, n_tiles, 0u);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------
// CTA-pair variant for bracket-grouped (DeltaConstant) batches: K = 2*n_age_pad <= 128.
//
// A cluster of two CTAs works on a PAIR of tiles (256 galaxies of one metallicity bracket) with
// tcgen05.mma.cta_group::2 (M = 256 over the pair, N = 128): each CTA keeps the hi/lo weights of ITS 128
// galaxies resident in shared memory for the whole pair (loaded once instead of once per chunk) and streams
// only ITS half (64 rows) of every G k-block, so the L2 -> SM traffic per flop is a third of the
// single-CTA kernel's, which was L2-bandwidth bound.  The leader CTA issues all MMAs; completion is
// multicast to both CTAs' barriers; both CTAs run the fused epilogue on their own 128 TMEM lanes.
#ifndef SB2_BN2
#define SB2_BN2 128
#endif
constexpr int kW2Kb = 4;                          // resident k-blocks (K <= 128)
constexpr int kW2Bytes = kW2Kb * 2 * kABytes;     // [kb][hi | lo] x 16 KiB = 128 KiB
constexpr int kBN2 = SB2_BN2;                         // accumulator columns per chunk: 4 TMEM accumulators, two per epilogue
                                                  // group, so the MMA of a group's next chunk overlaps its current one
constexpr int kG2Half = (kBN2 / 2) * kBK * 4;     // 8 KiB: this CTA's 64 rows of one G k-block
constexpr int kG2Slot = 2 * kG2Half;              // lo | hi
constexpr int kG2Slots = 512 / kBN2;        // 64 KiB ring either way
constexpr int kT2Buf = 512 / kBN2;

template <int kComp, int kNF, bool kSpec>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kSynthThreads, 1)
synth2_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
              const __grid_constant__ CUtensorMap tm_g_hi, const __grid_constant__ CUtensorMap tm_g_lo,
              const __grid_constant__ SynthArgs A) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;                              // resident weights
  uint8_t* s_g = smem + kW2Bytes;                   // G ring
  float2* s_uv = reinterpret_cast<float2*>(s_g + kG2Slots * kG2Slot);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_uv) + ((A.uv_len * 8 + 15) & ~15));
  uint64_t* full_bar = bars;                        // [kG2Slots] TMA (both CTAs) -> MMA   (leader's copy is used)
  uint64_t* empty_bar = bars + kG2Slots;            // [kG2Slots] MMA -> TMA               (multicast to both)
  uint64_t* tfull_bar = bars + 2 * kG2Slots;                            // [kTfPerGroup kMaxGroups] MMA -> epilogue group (multicast to both)
  uint64_t* tempty_bar = tfull_bar + kTfPerGroup * kMaxGroups;          // [kT2Buf] epilogues of both CTAs -> MMA (leader's copy)
  uint64_t* wfull_bar = tempty_bar + kT2Buf;                            //          weights landed (both CTAs) -> MMA (leader's copy)
  uint64_t* wempty_bar = wfull_bar + 1;                                 //          MMA -> TMA: weights buffer free (multicast)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_w_hi); prefetch_tmap(&tm_w_lo); prefetch_tmap(&tm_g_hi); prefetch_tmap(&tm_g_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kG2Slots; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < kTfPerGroup * kMaxGroups; ++b) mbar_init(&tfull_bar[b], 1);
    for (int b = 0; b < kT2Buf; ++b) mbar_init(&tempty_bar[b], 8);  // 4 warps x 2 CTAs
    mbar_init(wfull_bar, 2);
    mbar_init(wempty_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm<512>(tmem_slot);
    tmem_relinquish_2sm();
  }
  for (int i = threadIdx.x; i < A.uv_len; i += kSynthThreads) s_uv[i] = A.filt_uv[i];
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_units = A.n_tiles_dev ? min(A.n_tiles, __ldg(A.n_tiles_dev)) : A.n_tiles;  // units = tile pairs
  const int unit0 = (int)(blockIdx.x >> 1), unit_stride = (int)(gridDim.x >> 1);
  const int n_slots = (SB2_DBG_BITS(A) & 4) ? 2 : kG2Slots;
  const int c_all_last = A.n_chunk * (kBN / kBN2) - 1;   // tile_range counts 128-column chunks for this kernel

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs)
    // warp-uniform loop, one elected lane issues (see ptx.cuh)
    {
      const uint32_t elected = elect_one() ? 1u : 0u;
      const uint32_t rk = (uint32_t)warp_uniform((int)rank);
      const uint32_t wfull_l = (uint32_t)warp_uniform((int)mapa_u32(smem_u32(wfull_bar), 0));
      const uint32_t full_l0 = (uint32_t)warp_uniform((int)mapa_u32(smem_u32(full_bar), 0));
      const uint32_t w_addr = smem_u32(s_w), g_addr = smem_u32(s_g);
      const int n_units_u = warp_uniform(n_units), n_kb = A.n_kb;
      int slot = 0; uint32_t phase = 0, wphase = 0;
      for (int unit = unit0; unit < n_units_u; unit += unit_stride) {
        const int tile = unit * 2 + (int)rk;
        const int k0 = warp_uniform(A.tile_k0 ? __ldg(A.tile_k0 + unit) : 0);
        const int4 cr = A.tile_range ? __ldg(A.tile_range + unit) : make_int4(0, c_all_last, 0, 0);
        const int c_first = warp_uniform(cr.x), c_last = warp_uniform(cr.y);
        // this CTA's weights, resident for the whole unit
        mbar_wait(wempty_bar, wphase ^ 1, 0x100u + (uint32_t)unit);
        if (rk == 0) mbar_expect_tx_e(elected, wfull_bar, (SB2_DBG_BITS(A) & 128) ? 0 : 2 * kW2Bytes); else mbar_arrive_cluster_e(elected, wfull_l);
        if (!(SB2_DBG_BITS(A) & 128))
#pragma unroll
        for (int kb = 0; kb < kW2Kb; ++kb) {
          tma_load_2d_2sm_e(elected, w_addr + kb * 2 * kABytes, &tm_w_hi, wfull_l, kb * kBK, tile * kBM, kEvictFirst);
          tma_load_2d_2sm_e(elected, w_addr + kb * 2 * kABytes + kABytes, &tm_w_lo, wfull_l, kb * kBK, tile * kBM, kEvictFirst);
        }
        wphase ^= 1;
        const int n_c = c_last - c_first + 1, rot = chunk_rot(n_c, blockIdx.x >> 1);
        for (int j = 0; j < n_c; ++j) {
          const int c = chunk_at(c_first, n_c, rot, j);
          // this CTA's 64 rows of the chunk's G^T block.  The grid is laid out in 256-row blocks
          // [comp][256/comp wavelengths]; accumulator columns are [rank 0 rows | rank 1 rows]
          const int g_row = (kComp == 1) ? c * kBN2 + (int)rk * (kBN2 / 2)
                                         : (c >> 1) * kBN + (int)rk * (kBN / 2) + (c & 1) * (kBN2 / 2);
          for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(&empty_bar[slot], phase ^ 1, 0x200u + (uint32_t)slot);
            const uint32_t fl = full_l0 + (uint32_t)slot * 8u;
            const uint32_t st = g_addr + (uint32_t)slot * kG2Slot;
            if (SB2_DBG_BITS(A) & 64) {   // experiment: no G traffic
              if (rk == 0) mbar_expect_tx_e(elected, &full_bar[slot], 0); else mbar_arrive_cluster_e(elected, fl);
            } else {
            if (rk == 0) mbar_expect_tx_e(elected, &full_bar[slot], 2 * kG2Slot); else mbar_arrive_cluster_e(elected, fl);
            tma_load_2d_2sm_e(elected, st, &tm_g_lo, fl, k0 + kb * kBK, g_row, kEvictLast);
            tma_load_2d_2sm_e(elected, st + kG2Half, &tm_g_hi, fl, k0 + kb * kBK, g_row, kEvictLast);
            }
            if (++slot == n_slots) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA only)
    // The whole warp runs this loop with warp-uniform values; one elected lane issues (see umma_tf32_2sm_e).
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(2 * kBM, kBN2);
      const uint32_t elected = elect_one() ? 1u : 0u;
      const uint32_t w_base = smem_u32(s_w), g_base = smem_u32(s_g);
      const uint64_t desc0 = make_kmajor_sw128_desc(0);
      const int n_kb = A.n_kb, k8_total = A.k8_total;
      const uint32_t tmem_u = (uint32_t)warp_uniform((int)tmem_base);
      const int n_units_u = warp_uniform(n_units);
      int slot = 0; uint32_t phase = 0, wphase = 0, it = 0;
      constexpr int kG = (kT2Buf >= 3 && kMaxGroups >= 3 ? 3 : 2);   // epilogue groups (must match the epilogue_loop instantiation below)
      uint32_t gk0 = 0, gk1 = 0, gk2 = 0;         // chunks handed to each epilogue group so far
      for (int unit = unit0; unit < n_units_u; unit += unit_stride) {
        int4 cr = A.tile_range ? __ldg(A.tile_range + unit) : make_int4(0, c_all_last, 0, 0);
        const int c_first = warp_uniform(cr.x), c_last = warp_uniform(cr.y);
        mbar_wait(wfull_bar, wphase, 0x300u + (uint32_t)unit);
        wphase ^= 1;
        tc_fence_after();
        const int n_c = c_last - c_first + 1, rot = chunk_rot(n_c, blockIdx.x >> 1);
        for (int j = 0; j < n_c; ++j, ++it) {
          const int c = chunk_at(c_first, n_c, rot, j);
          const uint32_t buf = it % kT2Buf;
          mbar_wait(&tempty_bar[buf], ((it / kT2Buf) & 1u) ^ 1u, 0x400u + (it << 12));  // both CTAs' epilogues have drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_u + buf * kBN2;
          for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(&full_bar[slot], phase, 0x500u + (uint32_t)slot + (it << 12));
            tc_fence_after();
            // descriptors differ only in the 14-bit start-address field (bytes >> 4)
            const uint64_t da = desc0 + (uint64_t)(((w_base + kb * 2 * kABytes) & 0x3FFFF) >> 4);
            const uint64_t db = desc0 + (uint64_t)(((g_base + slot * kG2Slot) & 0x3FFFF) >> 4);
            const int k4n = min(kBK / 8, k8_total - kb * (kBK / 8));
#pragma unroll
            for (int k4 = 0; k4 < kBK / 8; ++k4) {
              if (k4 < k4n) {
                const uint64_t a_hi = da + (uint64_t)(k4 * 2), a_lo = da + (uint64_t)((kABytes >> 4) + k4 * 2);
                const uint64_t b_lo = db + (uint64_t)(k4 * 2), b_hi = db + (uint64_t)((kG2Half >> 4) + k4 * 2);
                if (!(SB2_DBG_BITS(A) & 32)) {
                umma_tf32_2sm_e(elected, d_tmem, a_lo, b_hi, idesc, (kb | k4) != 0);  // small terms first
                umma_tf32_2sm_e(elected, d_tmem, a_hi, b_lo, idesc, 1u);
                umma_tf32_2sm_e(elected, d_tmem, a_hi, b_hi, idesc, 1u);
                }
              }
            }
            umma_commit_2sm_e(elected, &empty_bar[slot], 3);   // ring slot reusable in both CTAs
            if (++slot == n_slots) { slot = 0; phase ^= 1; }
          }
          const int g = c % kG;                                // accumulator complete in both CTAs: wake the owners of chunk c
          if (g == 0)      { umma_commit_2sm_e(elected, &tfull_bar[gk0 % kTfPerGroup], 3); ++gk0; }
          else if (g == 1) { umma_commit_2sm_e(elected, &tfull_bar[kTfPerGroup + gk1 % kTfPerGroup], 3); ++gk1; }
          else             { umma_commit_2sm_e(elected, &tfull_bar[(kMaxGroups >= 3 ? 2 * kTfPerGroup : 0) + gk2 % kTfPerGroup], 3); ++gk2; }
        }
        umma_commit_2sm_e(elected, wempty_bar, 3);             // weights buffers reusable in both CTAs
      }
    }
  } else if (warp >= kEpiWarp0) {
    epilogue_loop<kComp, kNF, kSpec, 2, kBN2, (kT2Buf >= 3 && kMaxGroups >= 3 ? 3 : 2), false>(A, s_uv, nullptr, tfull_bar, tempty_bar, mapa_u32(smem_u32(tempty_bar), 0), tmem_base, unit0,
                                        unit_stride, n_units, rank);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

// flux_f = (gamma numU + beta numV) / (gamma su_f + beta sv_f) * scale, numerators = sum of the two epilogue
// groups' partials; one thread per (padded) row, results scattered back to the caller's galaxy order.
struct FinalizeArgs {
  const float2* part;
  long long n_rows;
  int n_filt, n_comp, n_groups;
  const float *g_beta, *g_gamma, *g_scale, *g_ca;
  const int* g_orig;
  const double* g_mscale;
  const unsigned* g_trunc;
  float* out_base;
  double* out_scaled;
  long long scaled_ld;       // 0: out_scaled is [galaxy][filter]; > 0: [filter][scaled_ld] (a library's Grid/Photometry layout)
  const float* e_part;       // dust emission (nullptr: none): absorbed-energy partials of the epilogue groups,
  const float2* dust_duv;    // [dust_m_len][n_filt] filter numerators of the emission per unit absorbed energy, by shift m
  const float* dust_g;       // [n_lam] the emission's spectrum per unit absorbed energy (spectra output)
  const int* g_m;
  int dust_m_len, n_lam;
  float* out_spec;
  float filt_su[kMaxFilt], filt_sdv[kMaxFilt];
};

// absorbed energy of a row: the epilogue groups' partials in fixed order
__device__ __forceinline__ float absorbed_energy(const FinalizeArgs& A, long long row) {
  float e = 0.f;
  for (int g = 0; g < A.n_groups; ++g) e += A.e_part[(size_t)g * A.n_rows + row];
  return e;
}

__global__ void __launch_bounds__(256) finalize_kernel(const __grid_constant__ FinalizeArgs A, const int* __restrict__ n_units_dev,
                                                       int rows_per_unit) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long rows = n_units_dev ? min(A.n_rows, (long long)__ldg(n_units_dev) * rows_per_unit) : A.n_rows;
  if (row >= rows) return;
  const int orig = A.g_orig[row];
  if (orig < 0) return;
  const float beta = A.g_beta[row], gamma = A.g_gamma[row];
  const float sc = (A.n_comp == 1) ? A.g_scale[row] * A.g_ca[row] : A.g_scale[row];
  const unsigned trunc = A.g_trunc[row];
  const double mscale = A.g_mscale[row];
  float e_abs = 0.f;
  const float2* duv = nullptr;
  if (A.e_part != nullptr) {
    e_abs = absorbed_energy(A, row);
    const int m = A.g_m[row];
    if (m >= 0 && m < A.dust_m_len) duv = A.dust_duv + (size_t)m * A.n_filt;   // beyond: no filter reaches the axis any more
  }
  // a galaxy's n_filt fluxes are contiguous in the output: build them four at a time and store 16 bytes at once
  // (the rows land in the caller's galaxy order, i.e. scattered -- scalar stores would touch each 32-byte sector 8x)
  const bool tr = A.scaled_ld > 0;   // transposed scaled output: scalar stores into n_filt rows (neighbouring rows of a tile are
                                     // redshift neighbours, not galaxy neighbours, so the columns are scattered either way)
  const bool vec = (A.n_filt % 4) == 0 && ((reinterpret_cast<uintptr_t>(A.out_base) | (tr ? 0 : reinterpret_cast<uintptr_t>(A.out_scaled))) & 15) == 0;
  for (int f0 = 0; f0 < A.n_filt; f0 += 4) {
    float fl[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int f = f0 + q;
      fl[q] = 0.f;
      if (f < A.n_filt) {
        float nu = 0.f, nv = 0.f;
        for (int g = 0; g < A.n_groups; ++g) {   // fixed order: plane g holds the chunks with c % n_groups == g
          const float2 a = A.part[((size_t)g * A.n_filt + f) * A.n_rows + row];
          nu += a.x; nv += a.y;
        }
        if (duv != nullptr) {
          const float2 dd = __ldg(duv + f);
          nu = fmaf(e_abs, dd.x, nu); nv = fmaf(e_abs, dd.y, nv);
        }
        float flux = fmaf(beta, nv, gamma * nu) / fmaf(beta, A.filt_sdv[f], gamma * A.filt_su[f]) * sc;
        if ((trunc >> f) & 1u) flux = __int_as_float(0x7fc00000);
        fl[q] = flux;
      }
    }
    if (vec) {
      if (A.out_base) *reinterpret_cast<float4*>(A.out_base + (size_t)orig * A.n_filt + f0) = make_float4(fl[0], fl[1], fl[2], fl[3]);
      if (A.out_scaled && tr) {
#pragma unroll
        for (int q = 0; q < 4; ++q) A.out_scaled[(size_t)(f0 + q) * A.scaled_ld + orig] = (double)fl[q] * mscale;
      } else if (A.out_scaled) {
        double2* o = reinterpret_cast<double2*>(A.out_scaled + (size_t)orig * A.n_filt + f0);
        o[0] = make_double2((double)fl[0] * mscale, (double)fl[1] * mscale);
        o[1] = make_double2((double)fl[2] * mscale, (double)fl[3] * mscale);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (f0 + q < A.n_filt) {
          if (A.out_base) A.out_base[(size_t)orig * A.n_filt + f0 + q] = fl[q];
          if (A.out_scaled) A.out_scaled[tr ? (size_t)(f0 + q) * A.scaled_ld + orig : (size_t)orig * A.n_filt + f0 + q] = (double)fl[q] * mscale;
        }
    }
  }
}

// Spectra output with dust emission: out_spec[galaxy][i] += E_abs * g_i * scale (the emission vanishes where the IGM acts,
// so it is added after the fact).  One block per 8 rows, threads along the wavelength axis.
__global__ void __launch_bounds__(256) dust_spec_kernel(const __grid_constant__ FinalizeArgs A, const int* __restrict__ n_units_dev,
                                                        int rows_per_unit) {
  const long long rows = n_units_dev ? min(A.n_rows, (long long)__ldg(n_units_dev) * rows_per_unit) : A.n_rows;
  for (int r = 0; r < 8; ++r) {
    const long long row = (long long)blockIdx.x * 8 + r;
    if (row >= rows) return;
    const int orig = A.g_orig[row];
    if (orig < 0) continue;
    const float sc = (A.n_comp == 1) ? A.g_scale[row] * A.g_ca[row] : A.g_scale[row];
    const float e = absorbed_energy(A, row) * sc;
    float* o = A.out_spec + (size_t)orig * A.n_lam;
    for (int i = threadIdx.x; i < A.n_lam; i += blockDim.x) o[i] = fmaf(e, __ldg(A.dust_g + i), o[i]);
  }
}

}  // namespace sb2
