// Contraction + fused epilogue kernel (the hot path of the hot path).
//
//   S[g, lam] = sum_k W[g, k] * G[k, lam]      (N_gal x K) . (K x N_lam), K = n_age*n_z
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
//   warps 4-7   fused epilogue : thread t owns galaxy t of the tile (= TMEM lane t); it streams its
//                                spectrum out of TMEM 32 wavelengths at a time and applies dust
//                                attenuation exp(-tau_V kappa), component mixing, the IGM row,
//                                and accumulates the trapezoidal filter numerators with the
//                                per-galaxy (m, beta) shift of the filter tables -- the spectrum
//                                never goes to shared or global memory.
//
// A wavelength chunk is 256 accumulator columns: 256 wavelengths for one spectral component, or
// 128 wavelengths x 2 components (attenuated | unattenuated) when the emission recipe needs both.
// Galaxies arrive sorted by redshift so the filter windows of a warp's 32 galaxies nearly coincide.
#pragma once
#include "ptx.cuh"

namespace sb2 {

constexpr int kBM = 128;                       // galaxies per tile
constexpr int kBN = 256;                       // accumulator columns per chunk
constexpr int kBK = 32;                        // k per stage (32 x 4 B = one 128 B swizzle row)
constexpr int kStages = 2;
constexpr int kABytes = kBM * kBK * 4;         // 16 KiB
constexpr int kBBytes = kBN * kBK * 4;         // 32 KiB
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;  // hi + lo of both operands = 96 KiB
constexpr int kSynthThreads = 256;
constexpr int kMaxFilt = 32;

struct SynthArgs {
  int n_gal, n_tiles, n_chunk, n_kb, n_lam, n_filt, n_blue, uv_len;
  const float* kappa;   // [n_chunk * lam_per_chunk], zero padded
  const float2* filt_uv;
  const float* igm;     // [n_tiles][n_blue][128]
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
  float* out_base;
  double* out_scaled;
  float* out_spec;
  int filt_lo[kMaxFilt], filt_hi[kMaxFilt], filt_off[kMaxFilt];
  float filt_su[kMaxFilt], filt_sdv[kMaxFilt];
};

template <int kComp, int kNF>
__global__ void __launch_bounds__(kSynthThreads, 1)
synth_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
             const __grid_constant__ CUtensorMap tm_g_hi, const __grid_constant__ CUtensorMap tm_g_lo,
             const __grid_constant__ SynthArgs A) {
  constexpr int kLch = kBN / kComp;  // wavelengths per chunk
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve-up: [stages][W_hi | W_lo | G_hi | G_lo], filter table, barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float2* s_uv = reinterpret_cast<float2*>(smem + kStages * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + ((A.uv_len * 8 + 15) & ~15));
  uint64_t* full_bar = bars;                    // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;         // [kStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;     // [2]        MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * kStages + 2;// [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_w_hi); prefetch_tmap(&tm_w_lo); prefetch_tmap(&tm_g_hi); prefetch_tmap(&tm_g_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
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

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
        for (int c = 0; c < A.n_chunk; ++c) {
          for (int kb = 0; kb < A.n_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* st = smem + stage * kStageBytes;
            mbar_expect_tx(&full_bar[stage], kStageBytes);
            tma_load_2d(st, &tm_w_hi, &full_bar[stage], kb * kBK, tile * kBM, kEvictNormal);
            tma_load_2d(st + kABytes, &tm_w_lo, &full_bar[stage], kb * kBK, tile * kBM, kEvictNormal);
            tma_load_2d(st + 2 * kABytes, &tm_g_hi, &full_bar[stage], kb * kBK, c * kBN, kEvictLast);
            tma_load_2d(st + 2 * kABytes + kBBytes, &tm_g_lo, &full_bar[stage], kb * kBK, c * kBN, kEvictLast);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(kBM, kBN);
      int stage = 0; uint32_t phase = 0; uint32_t it = 0;
      for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
        for (int c = 0; c < A.n_chunk; ++c, ++it) {
          const uint32_t buf = it & 1u;
          mbar_wait(&tempty_bar[buf], ((it >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * kBN;
          for (int kb = 0; kb < A.n_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + stage * kStageBytes);
#pragma unroll
            for (int k4 = 0; k4 < kBK / 8; ++k4) {
              const uint64_t a_hi = make_kmajor_sw128_desc(st + k4 * 32);
              const uint64_t a_lo = make_kmajor_sw128_desc(st + kABytes + k4 * 32);
              const uint64_t b_hi = make_kmajor_sw128_desc(st + 2 * kABytes + k4 * 32);
              const uint64_t b_lo = make_kmajor_sw128_desc(st + 2 * kABytes + kBBytes + k4 * 32);
              umma_tf32(d_tmem, a_lo, b_hi, idesc, (kb | k4) != 0);  // small terms first
              umma_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
              umma_tf32(d_tmem, a_hi, b_hi, idesc, 1u);
            }
            umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull_bar[buf]);      // accumulator complete
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== fused epilogue
    const int et = threadIdx.x - 128;            // galaxy within tile == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const unsigned FULL = 0xffffffffu;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
      const int row = tile * kBM + et;
      const int m = A.g_m[row];
      const float beta = A.g_beta[row], gamma = A.g_gamma[row], taut = A.g_taut[row], ca = A.g_ca[row], cb = A.g_cb[row];
      const int orig = A.g_orig[row];
      const float scale = A.g_scale[row];
      int mmin = m, mmax = m;
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        mmin = min(mmin, __shfl_xor_sync(FULL, mmin, o));
        mmax = max(mmax, __shfl_xor_sync(FULL, mmax, o));
      }
      float acc[kNF];
#pragma unroll
      for (int f = 0; f < kNF; ++f) acc[f] = 0.f;

      for (int c = 0; c < A.n_chunk; ++c, ++it) {
        const uint32_t buf = it & 1u;
        mbar_wait(&tfull_bar[buf], (it >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_acc = tmem_base + lane_base + buf * kBN;
#pragma unroll 1
        for (int sub = 0; sub < kLch / 32; ++sub) {
          const int i0 = c * kLch + sub * 32;
          const bool last_sub = (sub == kLch / 32 - 1) || (i0 + 32 >= A.n_lam);
          float s[32];
          {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_acc + sub * 32, v);
            if constexpr (kComp == 2) {
              uint32_t u[32];
              tmem_ld_32x32b_x32(t_acc + kLch + sub * 32, u);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float att = ex2_approx(-taut * __ldg(A.kappa + i0 + j));
                s[j] = ca * (__uint_as_float(v[j]) * att) + cb * __uint_as_float(u[j]);
              }
            } else {
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float att = ex2_approx(-taut * __ldg(A.kappa + i0 + j));
                s[j] = ca * (__uint_as_float(v[j]) * att);
              }
            }
          }
          if (last_sub) {  // all TMEM reads of this accumulator are done: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
          }
          if (i0 < A.n_blue) {
            const float* ig = A.igm + ((size_t)tile * A.n_blue + i0) * 128 + et;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (i0 + j < A.n_blue) s[j] *= __ldg(ig + j * 128);
          }
          if (A.out_spec != nullptr && orig >= 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (i0 + j < A.n_lam) A.out_spec[(size_t)orig * A.n_lam + i0 + j] = s[j] * scale;
          }
          // filter numerators: num_f += s_i * (gamma * U_f[n] + beta * V_f[n]),  n = i + m
#pragma unroll
          for (int f = 0; f < kNF; ++f) {
            if (f < A.n_filt) {
              const int lo = A.filt_lo[f], hi = A.filt_hi[f];
              if (i0 + mmin <= hi && i0 + 31 + mmax >= lo - 1) {  // warp-uniform
                const float2* tab = s_uv + A.filt_off[f];
                const int kmax = hi - lo + 3;
                const int k0 = i0 + m - (lo - 2);
                float a = acc[f];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float2 uv = tab[min(max(k0 + j, 0), kmax)];
                  a = fmaf(s[j], fmaf(beta, uv.y, gamma * uv.x), a);
                }
                acc[f] = a;
              }
            }
          }
          if (last_sub) break;
        }
      }
      // ---- finalize: flux_f = num_f / den_f * scale ; den_f = gamma * su_f + beta * sv_f
      if (orig >= 0) {
        const unsigned trunc = A.g_trunc[row];
        const double mscale = A.g_mscale[row];
#pragma unroll
        for (int f = 0; f < kNF; ++f) {
          if (f < A.n_filt) {
            float flux = acc[f] / fmaf(beta, A.filt_sdv[f], gamma * A.filt_su[f]) * scale;
            if ((trunc >> f) & 1u) flux = __int_as_float(0x7fc00000);
            if (A.out_base) A.out_base[(size_t)orig * A.n_filt + f] = flux;
            if (A.out_scaled) A.out_scaled[(size_t)orig * A.n_filt + f] = (double)flux * mscale;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace sb2
