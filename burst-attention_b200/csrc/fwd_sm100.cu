// ba_fwd_chunk: one ring round of the forward on sm_100a.
//
// Replaces, for flash="cuda"/"triton", the reference's per-round
//   flash_attn_2_cuda.fwd  +  cuda_scale_out_lse_helper          (burst_utils.py:149-177, :20-33)
// and the Triton LAO tile with carried state                       (lao.py:66-244)
// by ONE kernel: the carried (O fp32 normalised, lse) state is loaded in the
// prologue as the initial online-softmax state (m = lse, l = 1, acc = O) and the
// merged state is written in the epilogue -- no separate merge pass over HBM.
//
// Structure (one CTA = two 128-row Q tiles of one (batch, head), ping-ponged):
//   warp 9      TMA producer: Q once, then K_i / V_i tiles into 2-stage rings
//   warp 8      single-thread tcgen05.mma issuer:
//                 S_w = Q_w K_i^T      (SS, both operands K-major SW128 smem)
//                 O_w += P_w V_i       (TS, P bf16/fp16 in TMEM aliasing S_w, V MN-major smem)
//   warps 0-3   softmax for Q tile 0 (thread t owns row t == TMEM lane t)
//   warps 4-7   softmax for Q tile 1
// TMEM (512 cols): S0 [0,128)  S1 [128,256)  O0 [256,384)  O1 [384,512).  Each S_w is split into two
// 64-key sub-tile buffers so QK^T for sub-tile j+1 is already in flight / finished while the softmax
// warps work on sub-tile j (S double buffering); P (16-bit) aliases the first 32 columns of its buffer.
// The O rescale is lazy (only when a row max grows by more than 2^8), done in
// place by the softmax warps, so it is off the steady-state critical path.
#include <math.h>
#include <stdlib.h>

#include "fwd_common.cuh"
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace ba {

struct __align__(8) FwdBarriers {
  uint64_t q_full;
  uint64_t k_full[kKStages], k_empty[kKStages];
  uint64_t b_full[kKStages];  // kBias: the key-bias operand tile of this K stage has been written
  uint64_t v_full[kVStages], v_empty[kVStages];
  uint64_t s_full[2][2];   // MMA -> softmax: S_w sub-tile buffer b ready in TMEM
  uint64_t p_ready[2][2];  // softmax -> MMA: P_w (buffer b) written (and O_w rescaled)
  uint64_t o_done[2];      // MMA -> softmax: P_w V accumulated into O_w (one completion per sub-tile)
  uint64_t o_final[2];     // MMA -> softmax: the LAST P_w V of this CTA has completed (single completion)
  uint32_t tmem_base;
};

constexpr int kSub = 64;  // keys per softmax sub-tile (S is double-buffered per Q tile in 64-column halves)
static_assert(kKStages == 2 && kVStages == 2, "the unrolled MMA issue loop assumes 2-stage K/V rings");
// shared-memory carve-up for head dim kD (64 or 128): a tile is [128 rows][kD] 16-bit = kD/64 SW128 boxes of
// [128 rows][64 cols] (16 KiB each)
template <int kD>
struct FwdLayout {
  static_assert(kD == 64 || kD == 128, "head dim 64 or 128");
  static constexpr uint32_t kTileB = 128 * kD * 2;
  static constexpr int kBoxes = kD / 64;
  static constexpr uint32_t kOffQ = 0;
  static constexpr uint32_t kOffK = 2 * kTileB;
  static constexpr uint32_t kOffV = kOffK + kKStages * kTileB;
  // kBias: per K stage a [128 keys][16 B] no-swizzle operand tile (bias / scale in three 16-bit parts), and the
  // constant A operand of the bias K step: [ones core matrix 128 B][zero core matrix 128 B]
  static constexpr uint32_t kOffBias = kOffV + kVStages * kTileB;
  static constexpr uint32_t kOffOnes = kOffBias + kKStages * 2048;
  static constexpr uint32_t kOffBars = kOffOnes + 256;
  static constexpr int kSmemBytes = kOffBars + 256 /*barriers*/;
};

// One step (sub-tile j, with U = j & 3 known at compile time so that every TMEM address, smem
// descriptor and barrier address below is a constant): PV for both Q tiles on sub-tile j, then the
// QK^T of sub-tile j+2 into the S buffers that PV just released.  sb16 = (smem base address) >> 4.
// S_w (+)= 1[q] (x) (bias[key] / scale): one extra K = 16 step (sm100_ptx.cuh, "fold"); A = one core matrix of
// identical rows [1 1 1 0 ...] shared by all row groups (SBO = 0) + a zero core matrix for k = 8..15, B = the
// stage's bias tile, 64 keys of it (second K chunk re-reads the first, LBO = 0: A is zero there)
template <bool kBF16, int kD>
__device__ __forceinline__ void fwd_issue_bias(uint32_t sb16, uint32_t tS, int stage, int sub) {
  using L = FwdLayout<kD>;
  constexpr uint32_t idesc_qk = make_idesc(kBF16, kBlockM, kSub, false, false);
  const uint32_t a_lo = sb16 + (L::kOffOnes >> 4) + desc_lo_lbo(128);
  const uint32_t b_lo = sb16 + ((L::kOffBias + stage * 2048 + sub * 1024) >> 4) + desc_lo_lbo(0);
  umma_ss_lh(tS, a_lo, desc_hi_noswz(0), b_lo, desc_hi_noswz(128), idesc_qk, 1u);
}

template <bool kBF16, int kD, bool kBias, int U>
__device__ __forceinline__ void fwd_mma_step(int j, uint32_t sb16, FwdBarriers* bars, int n_s0, int n_s1, int n_sub,
                                             bool load_state) {
  using L = FwdLayout<kD>;
  constexpr uint32_t kOffQ = L::kOffQ, kOffK = L::kOffK, kOffV = L::kOffV, kTileBytes = L::kTileB;
  constexpr int sub = U & 1, st = U >> 1, st_next = st ^ 1;
  constexpr uint32_t idesc_qk = make_idesc(kBF16, kBlockM, kSub, false, false);
  constexpr uint32_t idesc_pv = make_idesc(kBF16, kBlockM, kD, false, true);
  constexpr uint32_t hi = desc_hi(1024);
  const int tile = j >> 1;
  if (sub == 0) {
    mbar_wait(&bars->v_full[st], (tile >> 1) & 1);
    tc_fence_after();
  }
  const bool next_qk = (j + 2) < n_sub;  // sub-tile j+2 lives in K tile (tile + 1), stage st_next
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    const int n_w = w ? n_s1 : n_s0;
    const uint32_t tS = w * 128 + sub * kSub;  // == P buffer
    if (j < n_w) {
      mbar_wait(&bars->p_ready[w][sub], tile & 1);
      tc_fence_after();
      // O_w (+)= P_w * V[j]: V tile is [keys][d] -> MN-major B, LBO 16 KiB (64-wide d blocks), SBO 1 KiB
      const uint32_t b_lo = sb16 + ((kOffV + st * kTileBytes + sub * kSub * 128) >> 4) + desc_lo_lbo(kBoxBytes);
      const uint32_t acc = (j > 0 || load_state) ? 1u : 0u;
#pragma unroll
      for (int kk = 0; kk < kSub / 16; ++kk)
        umma_ts_lh(256 + w * 128, tS + kk * 8, b_lo + kk * (16 * 128 / 16), hi, idesc_pv, kk > 0 ? 1u : acc);
      umma_commit(&bars->o_done[w]);
      if (j == n_w - 1) umma_commit(&bars->o_final[w]);
    }
    if (w == 1 && sub == 1) umma_commit(&bars->v_empty[st]);
    if (next_qk) {
      if (w == 0 && sub == 0) {
        mbar_wait(&bars->k_full[st_next], ((tile + 1) >> 1) & 1);
        if constexpr (kBias) mbar_wait(&bars->b_full[st_next], ((tile + 1) >> 1) & 1);
        tc_fence_after();
      }
      if (j + 2 < n_w) {
        // S_w = Q_w K^T for the 64 keys of sub-tile j+2 (8 KiB into each 64-wide box of the K tile)
        const uint32_t a_lo = sb16 + ((kOffQ + w * kTileBytes) >> 4) + desc_lo_lbo(16);
        const uint32_t k_lo = sb16 + ((kOffK + st_next * kTileBytes + sub * kSub * 128) >> 4) + desc_lo_lbo(16);
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          const uint32_t off = ((kk >> 2) * kBoxBytes + (kk & 3) * 32) >> 4;
          umma_ss_lh(tS, a_lo + off, hi, k_lo + off, hi, idesc_qk, kk > 0 ? 1u : 0u);
        }
        if constexpr (kBias) fwd_issue_bias<kBF16, kD>(sb16, tS, st_next, sub);
        umma_commit(&bars->s_full[w][sub]);
      }
      if (w == 1 && sub == 1) umma_commit(&bars->k_empty[st_next]);
    }
  }
}

// number of 64-key sub-tiles a 128-row Q tile starting at r0 must visit
__device__ __forceinline__ int fwd_trip_count(int r0, const FwdParams& p) {
  if (r0 >= p.Sq) return 0;
  int r_last = min(r0 + kBlockM - 1, p.Sq - 1);
  int max_limit = p.causal ? min(r_last + p.causal_off, p.Sk - 1) : p.Sk - 1;
  return max_limit < 0 ? 0 : max_limit / kSub + 1;
}

template <bool kBF16, int kD, bool kBias>
__global__ void __launch_bounds__(kFwdThreads, 1)
fwd_chunk_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();      // SWIZZLE_128B atoms need a 1 KiB-aligned base
  using L = FwdLayout<kD>;
  constexpr uint32_t kOffQ = L::kOffQ, kOffK = L::kOffK, kOffV = L::kOffV, kTileBytes = L::kTileB;
  constexpr int kBoxes = L::kBoxes, kChunks = kD / 32;  // 32-column fp32 chunks of an O row
  uint8_t* sQ = smem + kOffQ;                       // [2][tile]
  uint8_t* sK = smem + kOffK;                       // [kKStages][tile]
  uint8_t* sV = smem + kOffV;                       // [kVStages][tile]
  FwdBarriers* bars = reinterpret_cast<FwdBarriers*>(smem + L::kOffBars);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  // causal: the last Q tiles see the most keys -- schedule them first (longest-processing-time order)
  const int row0 = (p.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * (2 * kBlockM);
  const int n_s0 = fwd_trip_count(row0, p);            // sub-tiles for Q tile 0 / 1
  const int n_s1 = fwd_trip_count(row0 + kBlockM, p);
  const int n_sub = max(n_s0, n_s1);
  const int n_tiles = (n_sub + 1) >> 1;                // 128-key K/V tiles to stream

  // ---------------------------------------------------------------- setup
  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 8) {
    if (lane == 0) {
      mbar_init(&bars->q_full, 1);
      for (int i = 0; i < kKStages; ++i) {
        mbar_init(&bars->k_full[i], 1);
        mbar_init(&bars->k_empty[i], 1);
        mbar_init(&bars->b_full[i], 1);
      }
      for (int i = 0; i < kVStages; ++i) {
        mbar_init(&bars->v_full[i], 1);
        mbar_init(&bars->v_empty[i], 1);
      }
      for (int w = 0; w < 2; ++w) {
        for (int bf = 0; bf < 2; ++bf) {
          mbar_init(&bars->s_full[w][bf], 1);
          mbar_init(&bars->p_ready[w][bf], 4);  // one elected arrive per softmax warp
        }
        mbar_init(&bars->o_done[w], 1);
        mbar_init(&bars->o_final[w], 1);
      }
      fence_mbar_init();
    }
    if (kBias && lane < 16) {  // constant A operand of the bias K step: ones in slots 0..2 | zero core matrix
      constexpr uint32_t one = kBF16 ? 0x3F80u : 0x3C00u;
      *reinterpret_cast<uint4*>(smem + L::kOffOnes + lane * 16) =
          lane < 8 ? make_uint4(one | (one << 16), one, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
      fence_proxy_async_smem();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // all 512 columns are allocated, so the allocation starts at TMEM address 0; the MMA issue path
  // relies on that to keep every TMEM address a compile-time constant
  if (bars->tmem_base != 0) __trap();
  constexpr uint32_t tmem_base = 0;

  if (warp == 9) {
    // ============================================================ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars->q_full, 2 * kTileBytes);
      for (int w = 0; w < 2; ++w)
        for (int half = 0; half < kBoxes; ++half)
          tma_load_4d(sQ + w * kTileBytes + half * kBoxBytes, &tmQ, &bars->q_full, half * 64, h,
                      row0 + w * kBlockM, b);
    }
    for (int i = 0; i < n_tiles; ++i) {
      const int ks = i % kKStages, kph = (i / kKStages) & 1;
      if (kBias || lane == 0) mbar_wait(&bars->k_empty[ks], kph ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars->k_full[ks], kTileBytes);
        for (int half = 0; half < kBoxes; ++half)
          tma_load_4d(sK + ks * kTileBytes + half * kBoxBytes, &tmK, &bars->k_full[ks], half * 64, h,
                      i * kBlockN, b);
      }
      if constexpr (kBias) {  // the stage's key-bias operand tile: lane handles keys lane + 32 j
        constexpr float kBig = kBF16 ? 1e30f : 60000.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = lane + 32 * j, key = i * kBlockN + r;
          float x = 0.f;
          if (key < p.Sk) x = __ldg(p.bias + (int64_t)b * p.bias_sb + (int64_t)h * p.bias_sh + key) * p.inv_scale;
          x = fminf(fmaxf(x, -kBig), kBig);
          uint32_t h0, h1, h2;
          split3<kBF16>(x, h0, h1, h2);
          *reinterpret_cast<uint4*>(smem + L::kOffBias + ks * 2048 + (r >> 3) * 128 + (r & 7) * 16) =
              make_uint4(h0 | (h1 << 16), h2, 0u, 0u);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->b_full[ks]);
      }
      if (lane == 0) {
        const int vs = i % kVStages, vph = (i / kVStages) & 1;
        mbar_wait(&bars->v_empty[vs], vph ^ 1);
        mbar_arrive_expect_tx(&bars->v_full[vs], kTileBytes);
        for (int half = 0; half < kBoxes; ++half)
          tma_load_4d(sV + vs * kTileBytes + half * kBoxBytes, &tmV, &bars->v_full[vs], half * 64, h,
                      i * kBlockN, b);
      }
      __syncwarp();
    }
  } else if (warp == 8) {
    // ============================================================ MMA issuer (whole warp, elected lane issues)
    const uint32_t sb16 = smem_u32(smem) >> 4;
    if (n_sub > 0) {
      constexpr uint32_t idesc_qk = make_idesc(kBF16, kBlockM, kSub, false, false);
      constexpr uint32_t hi = desc_hi(1024);
      mbar_wait(&bars->q_full, 0);
      mbar_wait(&bars->k_full[0], 0);
      if constexpr (kBias) mbar_wait(&bars->b_full[0], 0);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          if (j < (w ? n_s1 : n_s0)) {
            const uint32_t a_lo = sb16 + ((kOffQ + w * kTileBytes) >> 4) + desc_lo_lbo(16);
            const uint32_t k_lo = sb16 + ((kOffK + j * kSub * 128) >> 4) + desc_lo_lbo(16);
#pragma unroll
            for (int kk = 0; kk < kD / 16; ++kk) {
              const uint32_t off = ((kk >> 2) * kBoxBytes + (kk & 3) * 32) >> 4;
              umma_ss_lh(w * 128 + j * kSub, a_lo + off, hi, k_lo + off, hi, idesc_qk, kk > 0 ? 1u : 0u);
            }
            if constexpr (kBias) fwd_issue_bias<kBF16, kD>(sb16, w * 128 + j * kSub, 0, j);
            umma_commit(&bars->s_full[w][j]);
          }
        }
      }
      umma_commit(&bars->k_empty[0]);
    }
    const bool ls = p.load_state != 0;
    for (int j0 = 0; j0 < n_sub; j0 += 4) {
      fwd_mma_step<kBF16, kD, kBias, 0>(j0, sb16, bars, n_s0, n_s1, n_sub, ls);
      if (j0 + 1 < n_sub) fwd_mma_step<kBF16, kD, kBias, 1>(j0 + 1, sb16, bars, n_s0, n_s1, n_sub, ls);
      if (j0 + 2 < n_sub) fwd_mma_step<kBF16, kD, kBias, 2>(j0 + 2, sb16, bars, n_s0, n_s1, n_sub, ls);
      if (j0 + 3 < n_sub) fwd_mma_step<kBF16, kD, kBias, 3>(j0 + 3, sb16, bars, n_s0, n_s1, n_sub, ls);
    }
  } else {
    // ============================================================ softmax warps
    const int w = warp >> 2;                  // Q tile handled by this warpgroup
    const int t = threadIdx.x & 127;          // row within the tile == TMEM lane
    const int r0 = row0 + w * kBlockM;
    const int n = w == 0 ? n_s0 : n_s1;
    if (r0 < p.Sq) {
      const int row = r0 + t;
      const bool valid_row = row < p.Sq;
      const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
      const uint32_t tSw = tmem_base + lane_base + w * 128;
      const uint32_t tO = tmem_base + lane_base + 256 + w * 128;
      const float scale_log2 = p.scale_log2;
      const int limit = p.causal ? min(row + p.causal_off, p.Sk - 1) : p.Sk - 1;
      const int tile_min_limit = p.causal ? min(r0 + p.causal_off, p.Sk - 1) : p.Sk - 1;

      float m = -INFINITY, l = 0.f;
      if (p.load_state) {
        float lse_prev = -INFINITY;
        if (valid_row) lse_prev = p.lse[(int64_t)b * p.lse_sb + (int64_t)h * p.lse_sh + row];
        if (lse_prev != -INFINITY) {
          m = lse_prev * kLog2e;
          l = 1.f;
        }
        const float* src = p.o_acc + (int64_t)b * p.oacc_sb + (int64_t)row * p.oacc_ss + (int64_t)h * p.oacc_sh;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 f = valid_row ? __ldg(reinterpret_cast<const float4*>(src + c * 32 + j * 4))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            v[j * 4 + 0] = __float_as_uint(f.x);
            v[j * 4 + 1] = __float_as_uint(f.y);
            v[j * 4 + 2] = __float_as_uint(f.z);
            v[j * 4 + 3] = __float_as_uint(f.w);
          }
          tmem_st_x32(tO + c * 32, v);
        }
        tmem_wait_st();
      }

      for (int j = 0; j < n; ++j) {
        const int bf = j & 1;
        const uint32_t tS = tSw + bf * kSub;
        mbar_wait(&bars->s_full[w][bf], (j >> 1) & 1);
        tc_fence_after();
        uint32_t sr[kSub];
        tmem_ld_x32(tS, sr);
        tmem_ld_x32(tS + 32, sr + 32);
        tmem_wait_ld();
        float* s = reinterpret_cast<float*>(sr);

        const int kbase = j * kSub;
        if (kbase + kSub - 1 > tile_min_limit) {  // warpgroup-uniform
#pragma unroll
          for (int c = 0; c < kSub; ++c)
            if (kbase + c > limit) s[c] = -INFINITY;
        }
        float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
        for (int c = 4; c < kSub; c += 4) {
          mx0 = fmaxf(mx0, s[c]);
          mx1 = fmaxf(mx1, s[c + 1]);
          mx2 = fmaxf(mx2, s[c + 2]);
          mx3 = fmaxf(mx3, s[c + 3]);
        }
        const float m_new = fmaxf(m, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2);
        const bool grow = m_new > m + kRescaleThreshold;  // also true for m == -inf, m_new finite
        if (__any_sync(0xffffffffu, grow)) {
          const bool o_live = (j > 0) || p.load_state;
          if (o_live) {
            if (j > 0) {
              mbar_wait(&bars->o_done[w], (j - 1) & 1);
              tc_fence_after();
            }
            const float f = (m == -INFINITY) ? 0.f : ex2(m - m_new);
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
              uint32_t v[32];
              tmem_ld_x32(tO + c * 32, v);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
              tmem_st_x32(tO + c * 32, v);
            }
            l *= f;
          }
          m = m_new;
        }
        const float neg_m = (m == -INFINITY) ? 0.f : -m;
        float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
        uint32_t pk[kSub / 2];
#pragma unroll
        for (int c = 0; c < kSub; c += 4) {
          const float p0 = ex2(fmaf(s[c], scale_log2, neg_m));
          const float p1 = ex2(fmaf(s[c + 1], scale_log2, neg_m));
          const float p2 = ex2(fmaf(s[c + 2], scale_log2, neg_m));
          const float x3 = fmaf(s[c + 3], scale_log2, neg_m);
          const float p3 = ex2(x3);
          sum0 += p0;
          sum1 += p1;
          sum2 += p2;
          sum3 += p3;
          pk[c / 2] = pack2<kBF16>(p0, p1);  // element 2c in the low half
          pk[c / 2 + 1] = pack2<kBF16>(p2, p3);
        }
        l += (sum0 + sum1) + (sum2 + sum3);
        tmem_st_x32(tS, pk);  // P aliases the first 32 columns of its S buffer
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_ready[w][bf]);
      }

      // ---------------------------------------------------------- epilogue
      // A parity wait on o_done would be ambiguous here (with S double-buffered the last two PVs may
      // both be pending or both be done), so the last PV signals a dedicated single-shot barrier.
      if (n > 0) {
        mbar_wait(&bars->o_final[w], 0);
        tc_fence_after();
      }
      const bool o_live = (n > 0) || p.load_state;
      const float inv_l = l > 0.f ? 1.f / l : 0.f;
      const float lse_out = l > 0.f ? (m + lg2(l)) * kLn2 : -INFINITY;
      if (valid_row) p.lse[(int64_t)b * p.lse_sb + (int64_t)h * p.lse_sh + row] = lse_out;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        uint32_t v[32];
        if (o_live) {
          tmem_ld_x32(tO + c * 32, v);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        if (valid_row) {
          if (p.store_lowp) {
            uint16_t* dst = reinterpret_cast<uint16_t*>(p.o_out) + (int64_t)b * p.oout_sb +
                            (int64_t)row * p.oout_ss + (int64_t)h * p.oout_sh + c * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 o;
              o.x = pack2<kBF16>(__uint_as_float(v[i * 8 + 0]) * inv_l, __uint_as_float(v[i * 8 + 1]) * inv_l);
              o.y = pack2<kBF16>(__uint_as_float(v[i * 8 + 2]) * inv_l, __uint_as_float(v[i * 8 + 3]) * inv_l);
              o.z = pack2<kBF16>(__uint_as_float(v[i * 8 + 4]) * inv_l, __uint_as_float(v[i * 8 + 5]) * inv_l);
              o.w = pack2<kBF16>(__uint_as_float(v[i * 8 + 6]) * inv_l, __uint_as_float(v[i * 8 + 7]) * inv_l);
              *reinterpret_cast<uint4*>(dst + i * 8) = o;
            }
          } else {
            float* dst = p.o_acc + (int64_t)b * p.oacc_sb + (int64_t)row * p.oacc_ss + (int64_t)h * p.oacc_sh +
                         c * 32;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 o;
              o.x = __uint_as_float(v[i * 4 + 0]) * inv_l;
              o.y = __uint_as_float(v[i * 4 + 1]) * inv_l;
              o.z = __uint_as_float(v[i * 4 + 2]) * inv_l;
              o.w = __uint_as_float(v[i * 4 + 3]) * inv_l;
              *reinterpret_cast<float4*>(dst + i * 4) = o;
            }
          }
        }
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}


template <bool kBF16, int kD, bool kBias>
static int launch_fwd(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const FwdParams& p,
                      cudaStream_t stream) {
  auto kern = fwd_chunk_kernel<kBF16, kD, kBias>;
  constexpr int smem = FwdLayout<kD>::kSmemBytes;
  BA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((p.Sq + 2 * kBlockM - 1) / (2 * kBlockM), p.H, p.B);
  kern<<<grid, kFwdThreads, smem, stream>>>(tmQ, tmK, tmV, p);
  BA_CHECK_CUDA(cudaGetLastError());
  return BA_OK;
}

}  // namespace ba

template <int kD, bool kBias>
static int launch_fwd_dt(int dtype, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                         const ba::FwdParams& p, cudaStream_t st) {
  return dtype == BA_DTYPE_BF16 ? ba::launch_fwd<true, kD, kBias>(tmQ, tmK, tmV, p, st)
                                : ba::launch_fwd<false, kD, kBias>(tmQ, tmK, tmV, p, st);
}

extern "C" int ba_fwd_chunk(ba_tensor4 q, ba_tensor4 k, ba_tensor4 v, ba_tensor4 o_acc, ba_rowstat lse,
                            ba_tensor4 o_out, int B, int Sq, int Sk, int H, int D, float scale, int mask_mode,
                            int causal_offset, int flags, int dtype, void* stream) {
  ba_rowstat none = {nullptr, 0, 0};
  return ba_fwd_chunk_bias(q, k, v, none, o_acc, lse, o_out, B, Sq, Sk, H, D, scale, mask_mode, causal_offset, flags,
                           dtype, stream);
}

extern "C" int ba_fwd_chunk_bias(ba_tensor4 q, ba_tensor4 k, ba_tensor4 v, ba_rowstat key_bias, ba_tensor4 o_acc,
                                 ba_rowstat lse, ba_tensor4 o_out, int B, int Sq, int Sk, int H, int D, float scale,
                                 int mask_mode, int causal_offset, int flags, int dtype, void* stream) {
  using namespace ba;
  BA_REQUIRE(D == 128 || D == 64, "ba_fwd_chunk: head dim %d unsupported (64 or 128)", D);
  BA_REQUIRE(B > 0 && Sq > 0 && Sk > 0 && H > 0, "ba_fwd_chunk: empty problem B=%d Sq=%d Sk=%d H=%d", B, Sq, Sk, H);
  BA_REQUIRE(dtype == BA_DTYPE_FP16 || dtype == BA_DTYPE_BF16, "ba_fwd_chunk: bad dtype %d", dtype);
  BA_REQUIRE(mask_mode == BA_MASK_NONE || mask_mode == BA_MASK_CAUSAL, "ba_fwd_chunk: bad mask mode %d", mask_mode);
  BA_REQUIRE(scale > 0.f && isfinite(scale), "ba_fwd_chunk: softmax scale must be positive and finite");
  BA_REQUIRE(q.ptr && k.ptr && v.ptr && lse.ptr, "ba_fwd_chunk: null q/k/v/lse");
  const bool first = flags & BA_FWD_FIRST, last = flags & BA_FWD_LAST;
  BA_REQUIRE(!last || o_out.ptr, "ba_fwd_chunk: BA_FWD_LAST needs o_out");
  BA_REQUIRE((first && last) || o_acc.ptr, "ba_fwd_chunk: fp32 state o_acc required unless FIRST|LAST");
  BA_REQUIRE(H <= 65535 && B <= 65535, "ba_fwd_chunk: H and B must be <= 65535");
  if (o_acc.ptr)
    BA_REQUIRE((reinterpret_cast<uintptr_t>(o_acc.ptr) & 15) == 0 && o_acc.stride_s % 4 == 0 &&
                   o_acc.stride_h % 4 == 0 && o_acc.stride_b % 4 == 0,
               "ba_fwd_chunk: o_acc must be 16-byte aligned with strides multiple of 4 elements");
  if (o_out.ptr)
    BA_REQUIRE((reinterpret_cast<uintptr_t>(o_out.ptr) & 15) == 0 && o_out.stride_s % 8 == 0 &&
                   o_out.stride_h % 8 == 0 && o_out.stride_b % 8 == 0,
               "ba_fwd_chunk: o_out must be 16-byte aligned with strides multiple of 8 elements");

  CUtensorMap tmQ, tmK, tmV;
  const CUtensorMapDataType dt = lowp_dtype(dtype);
  int rc;
  if ((rc = make_tensor_map(&tmQ, q, B, Sq, H, D, dt, 2, 64, kBlockM, true))) return rc;
  if ((rc = make_tensor_map(&tmK, k, B, Sk, H, D, dt, 2, 64, kBlockN, true))) return rc;
  if ((rc = make_tensor_map(&tmV, v, B, Sk, H, D, dt, 2, 64, kBlockN, true))) return rc;

  FwdParams p;
  p.o_acc = static_cast<float*>(o_acc.ptr);
  p.oacc_sb = o_acc.stride_b, p.oacc_ss = o_acc.stride_s, p.oacc_sh = o_acc.stride_h;
  p.lse = lse.ptr;
  p.lse_sb = lse.stride_b, p.lse_sh = lse.stride_h;
  p.o_out = o_out.ptr;
  p.oout_sb = o_out.stride_b, p.oout_ss = o_out.stride_s, p.oout_sh = o_out.stride_h;
  p.B = B, p.Sq = Sq, p.Sk = Sk, p.H = H;
  p.scale_log2 = scale * kLog2e;
  p.causal = mask_mode == BA_MASK_CAUSAL;
  p.causal_off = causal_offset;
  p.load_state = first ? 0 : 1;
  p.store_lowp = last ? 1 : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p.bias = key_bias.ptr, p.bias_sb = key_bias.stride_b, p.bias_sh = key_bias.stride_h;
  p.inv_scale = 1.f / scale;
  if (key_bias.ptr)
    return D == 64 ? launch_fwd_dt<64, true>(dtype, tmQ, tmK, tmV, p, st) : launch_fwd_dt<128, true>(dtype, tmQ, tmK, tmV, p, st);
  return D == 64 ? launch_fwd_dt<64, false>(dtype, tmQ, tmK, tmV, p, st) : launch_fwd_dt<128, false>(dtype, tmQ, tmK, tmV, p, st);
}
