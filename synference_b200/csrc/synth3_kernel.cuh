// Contraction + fused epilogue, third form: the per-galaxy weights are the A operand IN TENSOR MEMORY.
//
// Why: with both operands in shared memory (synth_kernel), a kind::tf32 MMA of M128 x N160 x K8 lasts 80 cycles and reads
// 4 KB (A) + 5 KB (B) of shared memory, while the TMA engine writes another 6 KB per MMA into the same banks (weights and
// grid tiles re-streamed per chunk): ~190 B/cycle against the 128 B/cycle the shared-memory crossbar moves -- the tensor
// pipe idled at 60 % (ncu, round 1) because shared memory was saturated, whatever the L2 did.  Here
//   * a tile's weights (hi | lo, 2 x 112 TMEM columns) stay RESIDENT in tensor memory for all its wavelength chunks and
//     feed the MMA directly (tcgen05.mma, A from TMEM): shared memory only carries the grid operand, once per chunk;
//   * they are expanded on chip from the compact float64 SFH bin masses (weights3_kernel: 408 B per galaxy through HBM
//     instead of 896 B of TF32 pairs) by a WRITER warpgroup: w[k] = sf[a] * s_z -> (tf32 hi, tf32 lo) -> tcgen05.st;
//   * the freed shared memory is a deep ring of grid k-blocks (5-6 stages instead of 2), so L2 latency is covered;
//   * registers are handed between warpgroups with setmaxnreg (16 warps: 8 epilogue, 4 writers, 4 service).  The
//     schedulers favour the highest warp id among eligible warps, so the MMA issuer is warp 15 and the TMA producer
//     warp 14: they never queue behind epilogue or writer warps of their sub-partition.
// The weights buffer is single (TMEM: 224 columns of weights + 288 of accumulators) but split in two k-regions: the
// MMA warp releases region A after the first half of the tile's LAST chunk and region B after the second, the writers
// refill each as it is released, and the next tile's first chunk starts on region A while B is still being written.
//
// Bracket-grouped (DeltaConstant) batches with 2*n_age_pad <= 112 and n_age <= 64; everything else takes synth_kernel.
#pragma once
#include "synth_kernel.cuh"

namespace sb2 {

constexpr int kS3Threads = 512;
constexpr int kS3EpiWarp0 = 0;                         // warps 0-7: two epilogue warpgroups; 8-11 writers; 12-15 service
constexpr int kS3WCols = 112;                          // TMEM columns of W_hi (and of W_lo)
constexpr int kS3WBase = 512 - 2 * kS3WCols;           // 288: accumulators live in [0, 288)
constexpr int kS3BarBytes = 512;
constexpr int kS3MaxStages = 8;

struct Synth3Args {
  const double* sf;   // [n_tiles][n_age][128] raw SFH bin masses (weights3_kernel)
  const double* s0;   // [n_rows] factor of the lower bracketing metallicity
  const double* s1;   // [n_rows] factor of the upper one
  int n_age, na_pad, w_stride, n_stages, kb_split;
  int cross;          // 1: the two small terms of the split product as ONE bfloat16 MMA (see the header); 0: 3 x TF32
};

template <int kN>
struct Synth3Cfg {
  static constexpr int kBuf = kS3WBase / kN;                  // 3 x 96 or 2 x 128 accumulator columns
  static constexpr int kStageBytes = 2 * kN * kBK * 4;        // hi | lo tile of one grid k-block
};

// per-wavelength tables the epilogue reads from shared memory: the attenuation curve plus what the feature set adds
__host__ __device__ inline int synth3_n_tables(int feat) {
  return 1 + ((feat & kFeatDustShape) ? 2 : 0) + ((feat & kFeatTwoScreens) ? 1 : 0) + ((feat & kFeatAbsorbed) ? 1 : 0);
}
// n_x: filters whose numerators the epilogue groups exchange for the fused output (0: finalize_kernel does it), 1 KB each
__host__ __device__ inline size_t synth3_smem_bytes(int kn, int n_stages, int n_age, int uv_len, int kap_len, bool spec, int feat = 0,
                                                    int n_x = 0) {
  return 1024 + (size_t)n_stages * (2 * kn * kBK * 4) + (size_t)n_age * 1024 + (((size_t)uv_len * 8 + 15) & ~size_t(15)) +
         (size_t)kap_len * 4 * synth3_n_tables(feat) + kS3BarBytes + (spec ? kSpecSmemBytes : 0) +
         (n_x ? (size_t)(n_x + 1) * kBM * 8 : 0);      // (+ one row for the absorbed energy)
}

template <int kComp, int kNF, bool kSpec, int kN, int kFeat = 0>
__global__ void __launch_bounds__(kS3Threads, 1)
synth3_kernel(const __grid_constant__ CUtensorMap tm_g_hi, const __grid_constant__ CUtensorMap tm_g_lo,
              const __grid_constant__ SynthArgs A, const __grid_constant__ Synth3Args X) {
  constexpr int kStageBytes = Synth3Cfg<kN>::kStageBytes;
  constexpr int kHalf = kN * kBK * 4;
  constexpr uint32_t kBuf = Synth3Cfg<kN>::kBuf;
  constexpr int kLch = kN / kComp;
  static_assert(kLch % 32 == 0 && kN % 16 == 0, "chunks are whole 32-wavelength sub-chunks");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_stages = X.n_stages;
  double* s_sf = reinterpret_cast<double*>(smem + (size_t)n_stages * kStageBytes);           // [n_age][128]
  float2* s_uv = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(s_sf) + (size_t)X.n_age * 1024);
  float* s_kap = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_uv) + ((A.uv_len * 8 + 15) & ~15));   // attenuation curve (or zeros)
  const int kap_len = A.kap_len;
  // feature-set tables follow the attenuation curve: dust_d0 | dust_l2 | kappa_birth | wnu (only those the set uses)
  float* s_tab = s_kap + kap_len;
  float* s_d0 = nullptr; float* s_l2 = nullptr; float* s_kapb = nullptr; float* s_wnu = nullptr;
  if constexpr ((kFeat & kFeatDustShape) != 0) { s_d0 = s_tab; s_l2 = s_tab + kap_len; s_tab += 2 * kap_len; }
  if constexpr ((kFeat & kFeatTwoScreens) != 0) { s_kapb = s_tab; s_tab += kap_len; }
  // (with pseudo-bins the energy weights are 1 on the chunks that use them: no table, see launch_synth3_t)
  if constexpr ((kFeat & kFeatAbsorbed) != 0) { if (A.x_count == 0) { s_wnu = s_tab; s_tab += kap_len; } }
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_tab);
  uint64_t* full_bar = bars;                                   // [kS3MaxStages] TMA -> MMA
  uint64_t* empty_bar = full_bar + kS3MaxStages;               // [kS3MaxStages] MMA -> TMA
  uint64_t* tfull_bar = empty_bar + kS3MaxStages;              // [kTfPerGroup * 2] MMA -> epilogue group
  uint64_t* tempty_bar = tfull_bar + kTfPerGroup * 2;          // [4] epilogue -> MMA, per accumulator
  uint64_t* wready_bar = tempty_bar + 4;                       // [2] writers -> MMA, per weights region
  uint64_t* wfree_bar = wready_bar + 2;                        // [2] MMA -> writers
  uint64_t* sfull_bar = wfree_bar + 2;                         //     staging loaded (bulk copy) -> writers
  uint64_t* sempty_bar = sfull_bar + 1;                        //     writers -> loader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty_bar + 1);

  // (the shuffle tells ptxas the role branches below are warp-uniform: inside a branch it cannot prove uniform it keeps
  //  every loop variable and descriptor in vector registers and pays an R2UR per MMA operand)
  const int warp = warp_uniform((int)(threadIdx.x >> 5)), lane = threadIdx.x & 31;
  constexpr int kWMma = 15, kWMma2 = 14, kWTma = 13, kWLoad = 12, kWAlloc = 12;

  if (warp == kWTma && lane == 0) { prefetch_tmap(&tm_g_hi); prefetch_tmap(&tm_g_lo); }
  if (warp == kWMma && lane == 0) {
    for (int s = 0; s < kS3MaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 2); }   // empty: both issuers
    for (int b = 0; b < kTfPerGroup * 2; ++b) mbar_init(&tfull_bar[b], 1);
    for (int b = 0; b < 4; ++b) mbar_init(&tempty_bar[b], 4);
    for (int r = 0; r < 2; ++r) { mbar_init(&wready_bar[r], 4); mbar_init(&wfree_bar[r], 2); }   // wfree: both issuers
    mbar_init(sfull_bar, 1);
    mbar_init(sempty_bar, 4);
    fence_barrier_init();
  }
  if (warp == kWAlloc) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < A.uv_len; i += kS3Threads) s_uv[i] = A.filt_uv[i];
  for (int i = threadIdx.x; i < kap_len; i += kS3Threads) {
    s_kap[i] = A.kappa[i];
    if constexpr ((kFeat & kFeatDustShape) != 0) { s_d0[i] = A.dust_d0[i]; s_l2[i] = A.dust_l2[i]; }
    if constexpr ((kFeat & kFeatTwoScreens) != 0) s_kapb[i] = A.kappa_birth[i];
    if constexpr ((kFeat & kFeatAbsorbed) != 0) { if (s_wnu) s_wnu[i] = A.wnu[i]; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = A.n_tiles_dev ? min(A.n_tiles, __ldg(A.n_tiles_dev)) : A.n_tiles;
  const int c_all_last = (A.n_chunk * kBN / kComp + kLch - 1) / kLch - 1;
  const int n_kb = A.n_kb, kb_split = X.kb_split;

  if (warp >= 12) {
    reg_dec<56>();
    if (warp == kWTma) {
      // ===================================================================== TMA producer: grid k-blocks only
      const uint32_t elected = elect_one() ? 1u : 0u;
      const uint32_t s_addr = smem_u32(smem), full0 = smem_u32(full_bar);
      const int n_tiles_u = warp_uniform(n_tiles);
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles_u; tile += gridDim.x) {
        const int k0 = warp_uniform(A.tile_k0 ? __ldg(A.tile_k0 + tile) : 0);
        const int4 cr = A.tile_range ? __ldg(A.tile_range + tile) : make_int4(0, c_all_last, 0, 0);
        const int c_first = warp_uniform(cr.x), c_last = warp_uniform(cr.y);
        const int n_real = max(c_last - c_first + 1, 0), n_c = n_real + A.x_count;   // + the absorbed-energy chunks (SynthArgs)
        for (int j = 0; j < n_c; ++j) {
          const int c = j < n_real ? c_first + j : A.x_first + (j - n_real);
          for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1, 0x3100u + (uint32_t)stage);
            const uint32_t st = s_addr + (uint32_t)stage * kStageBytes, fb = full0 + (uint32_t)stage * 8u;
            if (SB2_DBG_BITS(A) & 1024) {   // experiment: no operand traffic
              mbar_expect_tx_e(elected, &full_bar[stage], 0);
              if (++stage == n_stages) { stage = 0; phase ^= 1; }
              continue;
            }
            mbar_expect_tx_e(elected, &full_bar[stage], kStageBytes);
            if constexpr (kComp == 1) {
              tma_load_2d_e(elected, st, &tm_g_hi, fb, k0 + kb * kBK, c * kN, kEvictLast);
              tma_load_2d_e(elected, st + kHalf, &tm_g_lo, fb, k0 + kb * kBK, c * kN, kEvictLast);
            } else {
              // the grid's rows come in blocks of 256: [component 0: 128 wavelengths | component 1: the same 128]; a
              // chunk's accumulator columns are [component 0: kLch wavelengths | component 1: the same]
              const int r0 = (c * kLch / (kBN / 2)) * kBN + (c * kLch) % (kBN / 2);
              tma_load_2d_e(elected, st, &tm_g_hi, fb, k0 + kb * kBK, r0, kEvictLast);
              tma_load_2d_e(elected, st + kHalf / 2, &tm_g_hi, fb, k0 + kb * kBK, r0 + kBN / 2, kEvictLast);
              tma_load_2d_e(elected, st + kHalf, &tm_g_lo, fb, k0 + kb * kBK, r0, kEvictLast);
              tma_load_2d_e(elected, st + kHalf + kHalf / 2, &tm_g_lo, fb, k0 + kb * kBK, r0 + kBN / 2, kEvictLast);
            }
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == kWMma || warp == kWMma2) {
      // ===================================================================== MMA issuers (A from TMEM), one per epilogue group
      // Issuing one M128 x N96 x K8 MMA costs this warp ~10 instructions of descriptor arithmetic, i.e. more than the 48 cycles
      // the tensor pipe needs for it: a single issuer warp was busy 80 % of the time for 64 % tensor-pipe activity (ncu).
      // Two issuers split the chunks by parity -- issuer g feeds epilogue group g -- each with its own commit stream.
      constexpr uint32_t idesc = make_idesc_tf32(kBM, kN), idesc_x = make_idesc_bf16(kBM, kN);
      const bool cross = X.cross != 0;
      const uint32_t me = warp == kWMma2 ? 1u : 0u;
      const uint32_t elected = elect_one() ? 1u : 0u;
      const uint32_t s_lo0 = ((smem_u32(smem) & 0x3FFFFu) >> 4) | kDescLoBase;
      const uint32_t tmem_u = (uint32_t)warp_uniform((int)tmem_base);
      const uint32_t a_hi0 = tmem_u + kS3WBase, a_lo0 = a_hi0 + kS3WCols;
      const int n_tiles_u = warp_uniform(n_tiles), k8_total = A.k8_total;
      const int kb_free0 = max(kb_split, 1) - 1;
      int stage = 0; uint32_t phase = 0, it = 0, ti = 0, gk = 0;
      for (int tile = blockIdx.x; tile < n_tiles_u; tile += gridDim.x, ++ti) {
        const int4 cr = A.tile_range ? __ldg(A.tile_range + tile) : make_int4(0, c_all_last, 0, 0);
        const int c_first = warp_uniform(cr.x), c_last = warp_uniform(cr.y);
        const int n_real = max(c_last - c_first + 1, 0), n_c = n_real + A.x_count;
        auto chunk = [&](int j) { return j < n_real ? c_first + j : A.x_first + (j - n_real); };
        // first / last position of the tile's chunk list that is mine (chunks go to the issuers by the parity of c)
        int my_first = 0, my_last = n_c - 1;
        while (my_first < n_c && (((uint32_t)chunk(my_first) ^ me) & 1u)) ++my_first;
        while (my_last >= 0 && (((uint32_t)chunk(my_last) ^ me) & 1u)) --my_last;
        if (my_last < my_first) {
          // no chunk of this tile is mine: keep the weights hand-shake in step.  The wait comes first -- this tile's weights
          // are only written once BOTH issuers released the previous tile's, so neither can arrive twice in one phase.
          mbar_wait(&wready_bar[0], ti & 1u, 0x3200u);
          mbar_wait(&wready_bar[1], ti & 1u, 0x3201u);
          umma_commit_e(elected, &wfree_bar[0]);
          umma_commit_e(elected, &wfree_bar[1]);
        }
        for (int j = 0; j < n_c; ++j, ++it) {
          const int c = chunk(j);
          // Both issuers walk EVERY chunk and observe every phase of the barriers they share (an mbarrier wait carries one
          // parity bit: a waiter that skipped a phase would alias).  The other issuer's stages are released as soon as they
          // have been seen -- the ring slot is refilled once both have arrived (count 2).
          const uint32_t buf = it % kBuf;
          mbar_wait(&tempty_bar[buf], ((it / kBuf) & 1u) ^ 1u, 0x3300u + (it << 12));
          if (((uint32_t)c & 1u) != me) {
            for (int kb = 0; kb < n_kb; ++kb) {
              mbar_wait(&full_bar[stage], phase, 0x3580u + (uint32_t)stage);
              if (elected) mbar_arrive(&empty_bar[stage]);
              if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
            continue;
          }
          tc_fence_after();
          const uint32_t d_tmem = tmem_u + buf * kN;
          for (int kb = 0; kb < n_kb; ++kb) {
            if (j == my_first) {   // this tile's weights: region A before the first k-block, region B before k-block kb_split
              if (kb == 0) { mbar_wait(&wready_bar[0], ti & 1u, 0x3400u); if (kb_split == 0) mbar_wait(&wready_bar[1], ti & 1u, 0x3401u); tc_fence_after(); }
              else if (kb == kb_split) { mbar_wait(&wready_bar[1], ti & 1u, 0x3402u); tc_fence_after(); }
            }
            mbar_wait(&full_bar[stage], phase, 0x3500u + (uint32_t)stage);
            tc_fence_after();
            const uint32_t b_lo32 = s_lo0 + (uint32_t)stage * (uint32_t)(kStageBytes >> 4);
            const uint32_t a_hi = a_hi0 + (uint32_t)(kb * kBK), a_lo = a_lo0 + (uint32_t)(kb * kBK);
            const int k4n = min(kBK / 8, k8_total - kb * (kBK / 8));
            if (cross) {
              // w g = w_hi g_hi + (w_lo g_hi + w_hi g_lo): the bracket is 2^-11 of the product, so bfloat16 factors (2^-9
              // each) leave 2^-19 of it -- one K16 kind::f16 MMA over [w_lo | w_hi] x [g_hi ; g_lo] at twice the TF32 rate
#pragma unroll
              for (int k4 = 0; k4 < kBK / 8; ++k4) {
                if (k4 < k4n) {
                  umma_bf16_ts_lo_e(elected, d_tmem, a_lo + k4 * 8, b_lo32 + (kHalf >> 4) + k4 * 2, idesc_x, (kb | k4) != 0);  // small terms first
                  umma_tf32_ts_lo_e(elected, d_tmem, a_hi + k4 * 8, b_lo32 + k4 * 2, idesc, 1u);
                }
              }
            } else {
#pragma unroll
              for (int k4 = 0; k4 < kBK / 8; ++k4) {
                if (k4 < k4n) {
                  umma_tf32_ts_lo_e(elected, d_tmem, a_lo + k4 * 8, b_lo32 + k4 * 2, idesc, (kb | k4) != 0);  // small terms first
                  umma_tf32_ts_lo_e(elected, d_tmem, a_hi + k4 * 8, b_lo32 + (kHalf >> 4) + k4 * 2, idesc, 1u);
                  umma_tf32_ts_lo_e(elected, d_tmem, a_hi + k4 * 8, b_lo32 + k4 * 2, idesc, 1u);
                }
              }
            }
            umma_commit_e(elected, &empty_bar[stage]);
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
            if (j == my_last) {    // my last chunk of the tile: hand the weights regions back as their MMAs retire
              if (kb == kb_free0) umma_commit_e(elected, &wfree_bar[0]);
              if (kb == n_kb - 1) umma_commit_e(elected, &wfree_bar[1]);
            }
          }
          umma_commit_e(elected, &tfull_bar[kTfPerGroup * me + gk % kTfPerGroup]);   // wake epilogue group `me`
          ++gk;
        }
      }
    } else if (warp == kWLoad) {
      // ===================================================================== staging loader: a tile's bin masses, one bulk copy
      const uint32_t elected = elect_one() ? 1u : 0u;
      const int n_tiles_u = warp_uniform(n_tiles);
      const uint32_t dst = smem_u32(s_sf), fb = smem_u32(sfull_bar), bytes = (uint32_t)X.n_age * 1024u;
      uint32_t ti = 0;
      for (int tile = blockIdx.x; tile < n_tiles_u; tile += gridDim.x, ++ti) {
        mbar_wait(sempty_bar, (ti & 1u) ^ 1u, 0x3600u);
        mbar_expect_tx_e(elected, sfull_bar, bytes);
        bulk_load_e(elected, dst, X.sf + (size_t)tile * X.n_age * 128, bytes, fb);
      }
    }
  } else if (warp >= 8) {
    // ===================================================================== writers: bin masses -> TF32 hi/lo weights in TMEM
    reg_dec<104>();
    const int t = (warp & 3) * 32 + lane;                         // galaxy of the tile == TMEM lane
    const uint32_t w_hi0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kS3WBase, w_lo0 = w_hi0 + kS3WCols;
    const int split_col = min(kb_split * kBK, X.w_stride);
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const double s0 = X.s0[(size_t)tile * kBM + t], s1 = X.s1[(size_t)tile * kBM + t];
      if (SB2_DBG_BITS(A) & 4096) {
        // experiment: the issue slots a fused weight builder would take on these warps -- (dbg >> 16) * 16 iterations of two
        // dependent float64 FMAs and a shared-memory read per galaxy and tile (the real builder: ~8 k instructions)
        double x0 = s0 + 1.0, x1 = s1 + 2.0;
        const int reps = (SB2_DBG_BITS(A) >> 16) * 16;
        for (int i = 0; i < reps; ++i) {
          x0 = fma(x0, 1.0000001, 1e-9) + s_sf[(i & 31) * 128 + t];
          x1 = fma(x1, 0.9999999, 1e-9);
        }
        if (x0 + x1 == 12345.678) s_sf[t] = x0;
      }
      mbar_wait(sfull_bar, ti & 1u, 0x3700u);
      for (int r = 0; r < 2; ++r) {
        mbar_wait(&wfree_bar[r], (ti & 1u) ^ 1u, 0x3800u + (uint32_t)r);
        tc_fence_after();
        const int c_lo = r ? split_col : 0, c_hi = ((SB2_DBG_BITS(A) & 512) && ti > 0) ? 0 : (r ? X.w_stride : split_col);   // experiment 512: hand-shake only
        for (int c = c_lo; c < c_hi; c += 8) {    // na_pad is a multiple of 8: a group of 8 columns has one metallicity
          const bool up = c >= X.na_pad;
          const int a0 = c - (up ? X.na_pad : 0);
          const double s = up ? s1 : s0;
          uint32_t hi[8], lo[8];
          float hf[8], lf[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const double w = (a0 + j < X.n_age) ? s_sf[(a0 + j) * 128 + t] * s : 0.0;
            hf[j] = to_tf32_rna((float)w);
            lf[j] = (float)(w - (double)hf[j]);
            hi[j] = __float_as_uint(hf[j]);
          }
          if (X.cross) {
            // the K16 bfloat16 operand of this group of 8 bins: [w_lo(0..7) | w_hi(0..7)], two elements per column
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              lo[j] = pack_bf16x2(lf[2 * j], lf[2 * j + 1]);
              lo[4 + j] = pack_bf16x2(hf[2 * j], hf[2 * j + 1]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) lo[j] = __float_as_uint(to_tf32_rna(lf[j]));
          }
          tmem_st_32x32b_x8(w_hi0 + (uint32_t)c, hi);
          tmem_st_32x32b_x8(w_lo0 + (uint32_t)c, lo);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&wready_bar[r]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sempty_bar);   // staging may be overwritten with the next tile's bin masses
    }
  } else {
    reg_inc<176>();
    float* s_spec = (kSpec && A.spec_smem) ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kS3BarBytes) : nullptr;
    const EpiTables tabs{smem_u32(s_kap), s_d0 ? smem_u32(s_d0) : 0u, s_l2 ? smem_u32(s_l2) : 0u, s_kapb ? smem_u32(s_kapb) : 0u,
                         s_wnu ? smem_u32(s_wnu) : 0u};
    float2* s_x = A.fuse_out ? reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + kS3BarBytes + (A.spec_smem ? kSpecSmemBytes : 0))
                             : nullptr;
    epilogue_loop<kComp, kNF, kSpec, 1, kN, 2, false, kS3EpiWarp0, (int)kBuf, kFeat, true>(A, s_uv, s_spec, tfull_bar, tempty_bar, 0u,
                                                                                          tmem_base, (int)blockIdx.x, (int)gridDim.x,
                                                                                          n_tiles, 0u, tabs, s_x);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWAlloc) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace sb2
