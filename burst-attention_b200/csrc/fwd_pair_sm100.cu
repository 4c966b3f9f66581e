// ba_fwd_chunk, CTA-pair variant (opt-in: BA_FWD_IMPL=5) -- the design profiles/README.md points to:
// tcgen05.mma fetches shared-memory operands at only ~64-80 B/clk/SM, so the single-CTA kernels are
// bound by operand fetch (SS-form QK^T) rather than by tensor math.  Here a cluster of TWO CTAs owns
// 256 query rows and issues ONE tcgen05.mma.cta_group::2 (M = 256) per k-step:
//   * each CTA keeps its own 128-row Q tile in TMEM (A operand, no smem fetch at all),
//   * K and V tiles are SPLIT across the pair (CTA r loads 64 of the 128 keys of K, and the d-half
//     [64r,64r+64) of V), so per MMA each SM fetches 2 KiB instead of 6-8 KiB,
//   * S is 128 keys wide AND double-buffered (one Q tile per CTA leaves the TMEM room):
//     TMEM (512 cols per CTA): Q [0,64)  S0 [64,192)  S1 [192,320)  O [320,448).
// Roles per CTA: warps 0-3 softmax for S columns [0,64) ("A"), warps 4-7 for [64,128) ("B") -- the two
// warpgroups exchange row maxima / row sums through smem; warp 8 MMA issuer (leader CTA only issues),
// warp 9 TMA producer (each CTA loads its halves; bytes complete on the leader's barriers).
// Signals: MMA -> both CTAs by multicast tcgen05.commit; softmax -> leader by remote mbarrier arrive.
//
// STATUS (round 1): numerically correct on B200 for every forward test case (ragged, causal, fp16,
// carried state), but only 775 TFLOP/s at S=16384 vs 1200 for the single-CTA kernel; a variant without
// the per-tile inter-warpgroup exchange (full-row max per warpgroup, P single-buffered in the spare
// TMEM columns) measured 686.  Not yet profiled -- the first item for the next round (ncu: leader issue
// rate with 5 multicast commits per tile, remote-arrive latency of the 16-way p_ready, MUFU phases of
// the two warpgroups running in lockstep).
#include <math.h>

#include "fwd_common.cuh"
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace ba {

constexpr int kPStages = 4;                       // K-half / V-half stages (16 KiB each)
constexpr int kHalfBytes = 64 * kHeadDim * 2;     // 16 KiB: 64 keys x 128 d (K half) or 128 keys x 64 d (V half)
constexpr uint32_t kPOffK = 0;
constexpr uint32_t kPOffV = kPStages * kHalfBytes;
constexpr uint32_t kPOffX = 2 * kPStages * kHalfBytes;  // fp32 [2][128] row-stat exchange
constexpr uint32_t kPOffBars = kPOffX + 2 * 128 * 4;
constexpr int kPairSmemBytes = kPOffBars + 512;
constexpr uint32_t kTQ = 0, kTS0 = 64, kTO = 320;       // TMEM columns (S_b at kTS0 + 128 b)

struct __align__(8) PairBarriers {
  uint64_t k_full[kPStages], k_empty[kPStages];   // full: leader's copy collects both CTAs' bytes
  uint64_t v_full[kPStages], v_empty[kPStages];
  uint64_t q_ready;      // leader: 16 warp arrivals (both CTAs) -- Q (and carried O) staged in TMEM
  uint64_t s_full[2];    // per CTA (multicast commit): S_b ready
  uint64_t p_ready[2];   // leader: 16 warp arrivals -- P_b written in both CTAs (and O rescaled)
  uint64_t o_done;       // per CTA (multicast): PV accumulated, one completion per tile
  uint64_t o_final;      // per CTA (multicast): last PV done
  uint32_t tmem_base;
};

__device__ __forceinline__ int pair_trip_count(int r0, const FwdParams& p) {  // 128-key tiles for the pair's 256 rows
  if (r0 >= p.Sq) return 0;
  int r_last = min(r0 + 2 * kBlockM - 1, p.Sq - 1);
  int max_limit = p.causal ? min(r_last + p.causal_off, p.Sk - 1) : p.Sk - 1;
  return max_limit < 0 ? 0 : max_limit / kBlockN + 1;
}

template <bool kBF16, int B, int ST>
__device__ __forceinline__ void pair_issue_qk(uint32_t sb16) {  // S_b[256 x 128] = Q K^T ; K half: 64 keys, K-major
  constexpr uint32_t idesc = make_idesc(kBF16, 256, kBlockN, false, false), hi = desc_hi(1024);
  const uint32_t k_lo = sb16 + ((kPOffK + ST * kHalfBytes) >> 4) + desc_lo_lbo(16);
#pragma unroll
  for (int kk = 0; kk < kHeadDim / 16; ++kk) {
    const uint32_t off = ((kk >> 2) * (kHalfBytes / 2) + (kk & 3) * 32) >> 4;  // two 64x64 boxes of 8 KiB
    umma_ts_2cta_lh(kTS0 + B * 128, kTQ + kk * 8, k_lo + off, hi, idesc, kk > 0 ? 1u : 0u);
  }
}
template <bool kBF16, int B, int ST>
__device__ __forceinline__ void pair_issue_pv(uint32_t sb16, uint32_t acc) {  // O += P_b V ; V half: 128 keys x 64 d
  constexpr uint32_t idesc = make_idesc(kBF16, 256, kHeadDim, false, true), hi = desc_hi(1024);
  const uint32_t v_lo = sb16 + ((kPOffV + ST * kHalfBytes) >> 4) + desc_lo_lbo(kHalfBytes);
#pragma unroll
  for (int kk = 0; kk < kBlockN / 16; ++kk)  // P: keys 0..63 at S_b cols [0,32), keys 64..127 at cols [64,96)
    umma_ts_2cta_lh(kTO, kTS0 + B * 128 + (kk >> 2) * 64 + (kk & 3) * 8, v_lo + kk * (16 * 128 / 16), hi, idesc,
                    kk > 0 ? 1u : acc);
}

// one tile of the leader's MMA warp; U = i % 4 (S buffer = U & 1, K/V stage = U) is compile-time
template <bool kBF16, int U>
__device__ __forceinline__ void pair_mma_tile(int i, int n, uint32_t sb16, PairBarriers* bars, bool load_state) {
  constexpr int B = U & 1, ST = U % kPStages, STN = (U + 2) % kPStages;
  mbar_wait(&bars->v_full[ST], (i / kPStages) & 1);
  mbar_wait(&bars->p_ready[B], (i >> 1) & 1);
  tc_fence_after();
  pair_issue_pv<kBF16, B, ST>(sb16, (i > 0 || load_state) ? 1u : 0u);
  umma_commit_2cta(&bars->o_done, 0x3);
  umma_commit_2cta(&bars->v_empty[ST], 0x3);
  if (i == n - 1) umma_commit_2cta(&bars->o_final, 0x3);
  if (i + 2 < n) {  // QK^T of tile i+2 into the S buffer PV(i) has just released
    mbar_wait(&bars->k_full[STN], ((i + 2) / kPStages) & 1);
    tc_fence_after();
    pair_issue_qk<kBF16, B, STN>(sb16);
    umma_commit_2cta(&bars->s_full[B], 0x3);
    umma_commit_2cta(&bars->k_empty[STN], 0x3);
  }
}

template <bool kBF16>
__global__ void __launch_bounds__(kFwdThreads, 1)
fwd_pair_kernel(const __grid_constant__ CUtensorMap tmK64, const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_pair[];
  uint8_t* smem = smem_pair;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sK = smem + kPOffK;
  uint8_t* sV = smem + kPOffV;
  float* sX = reinterpret_cast<float*>(smem + kPOffX);
  PairBarriers* bars = reinterpret_cast<PairBarriers*>(smem + kPOffBars);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int h = blockIdx.y, b = blockIdx.z;
  const int pair = blockIdx.x >> 1;
  const int row0 = pair * (2 * kBlockM);             // first row of the pair
  const int r0 = row0 + (int)rank * kBlockM;          // first row of this CTA
  const int n = pair_trip_count(row0, p);            // tiles the PAIR visits (both CTAs run all of them)

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&tmK64);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < kPStages; ++i) {
        mbar_init(&bars->k_full[i], 1);
        mbar_init(&bars->k_empty[i], 1);
        mbar_init(&bars->v_full[i], 1);
        mbar_init(&bars->v_empty[i], 1);
      }
      mbar_init(&bars->q_ready, 16);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars->s_full[i], 1);
        mbar_init(&bars->p_ready[i], 16);
      }
      mbar_init(&bars->o_done, 1);
      mbar_init(&bars->o_final, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2cta(&bars->tmem_base, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (bars->tmem_base != 0) __trap();

  if (warp == 9) {
    // ============================================================ TMA producer (this CTA's halves)
    if (lane == 0) {
      for (int i = 0; i < n; ++i) {
        const int st = i % kPStages, ph = (i / kPStages) & 1;
        mbar_wait(&bars->k_empty[st], ph ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&bars->k_full[st], 2 * kHalfBytes);
        for (int half = 0; half < 2; ++half)  // keys [128 i + 64 r, +64): two 64x64 boxes
          tma_load_4d_2cta(sK + st * kHalfBytes + half * (kHalfBytes / 2), &tmK64, &bars->k_full[st], half * 64, h,
                           i * kBlockN + (int)rank * 64, b);
        mbar_wait(&bars->v_empty[st], ph ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&bars->v_full[st], 2 * kHalfBytes);
        tma_load_4d_2cta(sV + st * kHalfBytes, &tmV, &bars->v_full[st], (int)rank * 64, h, i * kBlockN, b);
      }
    }
  } else if (warp == 8) {
    // ============================================================ MMA issuer: the leader CTA only
    if (rank == 0 && n > 0) {
      const uint32_t sb16 = smem_u32(smem) >> 4;
      mbar_wait(&bars->q_ready, 0);
      mbar_wait(&bars->k_full[0], 0);
      tc_fence_after();
      pair_issue_qk<kBF16, 0, 0>(sb16);
      umma_commit_2cta(&bars->s_full[0], 0x3);
      umma_commit_2cta(&bars->k_empty[0], 0x3);
      if (n > 1) {
        mbar_wait(&bars->k_full[1], 0);
        tc_fence_after();
        pair_issue_qk<kBF16, 1, 1>(sb16);
        umma_commit_2cta(&bars->s_full[1], 0x3);
        umma_commit_2cta(&bars->k_empty[1], 0x3);
      }
      const bool ls = p.load_state != 0;
      for (int i0 = 0; i0 < n; i0 += 4) {
        pair_mma_tile<kBF16, 0>(i0, n, sb16, bars, ls);
        if (i0 + 1 < n) pair_mma_tile<kBF16, 1>(i0 + 1, n, sb16, bars, ls);
        if (i0 + 2 < n) pair_mma_tile<kBF16, 2>(i0 + 2, n, sb16, bars, ls);
        if (i0 + 3 < n) pair_mma_tile<kBF16, 3>(i0 + 3, n, sb16, bars, ls);
      }
    }
  } else if (n > 0) {
    // ============================================================ softmax warps (both CTAs)
    const int g = warp >> 2;                  // column half of S handled by this warpgroup
    const int t = threadIdx.x & 127;          // row within the CTA's tile == TMEM lane
    const int row = r0 + t;
    const bool valid_row = row < p.Sq;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float scale_log2 = p.scale_log2;
    const int limit = p.causal ? min(row + p.causal_off, p.Sk - 1) : p.Sk - 1;
    const int tile_min_limit = p.causal ? min(r0 + p.causal_off, p.Sk - 1) : p.Sk - 1;

    // ---- stage this thread's half of its Q row (64 x 16 bit = 32 packed columns) into TMEM
    {
      const uint4* src = reinterpret_cast<const uint4*>(p.q + (int64_t)b * p.q_sb + (int64_t)row * p.q_ss +
                                                        (int64_t)h * p.q_sh + g * 64);
      uint32_t v[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 x = valid_row ? __ldg(src + i) : make_uint4(0u, 0u, 0u, 0u);
        v[i * 4 + 0] = x.x, v[i * 4 + 1] = x.y, v[i * 4 + 2] = x.z, v[i * 4 + 3] = x.w;
      }
      tmem_st_x32(lane_base + kTQ + g * 32, v);
    }
    // carried state: both warpgroups track the same running max m; the row sum is split (l_A + l_B)
    float m = -INFINITY, l = 0.f;
    if (p.load_state) {
      float lse_prev = -INFINITY;
      if (valid_row) lse_prev = p.lse[(int64_t)b * p.lse_sb + (int64_t)h * p.lse_sh + row];
      if (lse_prev != -INFINITY) {
        m = lse_prev * kLog2e;
        l = (g == 0) ? 1.f : 0.f;
      }
      const float* src = p.o_acc + (int64_t)b * p.oacc_sb + (int64_t)row * p.oacc_ss + (int64_t)h * p.oacc_sh + g * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 f = valid_row ? __ldg(reinterpret_cast<const float4*>(src + c * 32 + i * 4))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
          v[i * 4 + 0] = __float_as_uint(f.x);
          v[i * 4 + 1] = __float_as_uint(f.y);
          v[i * 4 + 2] = __float_as_uint(f.z);
          v[i * 4 + 3] = __float_as_uint(f.w);
        }
        tmem_st_x32(lane_base + kTO + g * 64 + c * 32, v);
      }
    }
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(&bars->q_ready, 0);

    for (int i = 0; i < n; ++i) {
      const int bf = i & 1;
      const uint32_t tS = lane_base + kTS0 + bf * 128 + g * 64;  // this warpgroup's 64 S columns
      mbar_wait(&bars->s_full[bf], (i >> 1) & 1);
      tc_fence_after();
      uint32_t sr[64];
      tmem_ld_x32(tS, sr);
      tmem_ld_x32(tS + 32, sr + 32);
      tmem_wait_ld();
      float* s = reinterpret_cast<float*>(sr);
      const int kbase = i * kBlockN + g * 64;
      if (i * kBlockN + kBlockN - 1 > tile_min_limit) {  // CTA-uniform
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (kbase + c > limit) s[c] = -INFINITY;
      }
      float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
      for (int c = 4; c < 64; c += 4) {
        mx0 = fmaxf(mx0, s[c]);
        mx1 = fmaxf(mx1, s[c + 1]);
        mx2 = fmaxf(mx2, s[c + 2]);
        mx3 = fmaxf(mx3, s[c + 3]);
      }
      // row maximum over all 128 keys: exchange the two warpgroups' partial maxima through smem
      const float mx_mine = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      sX[g * 128 + t] = mx_mine;
      named_bar_sync(2, 256);
      const float mx = fmaxf(mx_mine, sX[(g ^ 1) * 128 + t]);
      named_bar_sync(3, 256);  // both have read before the next tile overwrites the slots
      const float m_new = fmaxf(m, mx * scale_log2);
      const bool grow = m_new > m + kRescaleThreshold;
      if (__any_sync(0xffffffffu, grow)) {  // same outcome in the partner warp (same rows, same m, same mx)
        const bool o_live = (i > 0) || p.load_state;
        if (o_live) {
          if (i > 0) {
            mbar_wait(&bars->o_done, (i - 1) & 1);  // PV(i-1) finished (S(i) was only issued behind PV(i-2))
            tc_fence_after();
          }
          const float f = (m == -INFINITY) ? 0.f : ex2(m - m_new);
#pragma unroll
          for (int c = 0; c < 2; ++c) {  // each warpgroup rescales its half of the O columns
            uint32_t v[32];
            tmem_ld_x32(lane_base + kTO + g * 64 + c * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * f);
            tmem_st_x32(lane_base + kTO + g * 64 + c * 32, v);
          }
          l *= f;
        }
        m = m_new;
      }
      const float neg_m = (m == -INFINITY) ? 0.f : -m;
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
        const float p0 = ex2(fmaf(s[c], scale_log2, neg_m));
        const float p1 = ex2(fmaf(s[c + 1], scale_log2, neg_m));
        const float p2 = ex2(fmaf(s[c + 2], scale_log2, neg_m));
        const float p3 = ex2(fmaf(s[c + 3], scale_log2, neg_m));
        sum0 += p0;
        sum1 += p1;
        sum2 += p2;
        sum3 += p3;
        pk[c / 2] = pack2<kBF16>(p0, p1);
        pk[c / 2 + 1] = pack2<kBF16>(p2, p3);
      }
      l += (sum0 + sum1) + (sum2 + sum3);
      tmem_st_x32(tS, pk);  // P (16-bit) over the first 32 of this warpgroup's own S columns
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&bars->p_ready[bf], 0);
    }

    // ---------------------------------------------------------- epilogue: each warpgroup writes its d half
    mbar_wait(&bars->o_final, 0);
    tc_fence_after();
    sX[g * 128 + t] = l;
    named_bar_sync(2, 256);
    const float l_tot = l + sX[(g ^ 1) * 128 + t];
    const float inv_l = l_tot > 0.f ? 1.f / l_tot : 0.f;
    if (g == 0 && valid_row)
      p.lse[(int64_t)b * p.lse_sb + (int64_t)h * p.lse_sh + row] = l_tot > 0.f ? (m + lg2(l_tot)) * kLn2 : -INFINITY;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_x32(lane_base + kTO + g * 64 + c * 32, v);
      tmem_wait_ld();
      if (valid_row) {
        if (p.store_lowp) {
          uint16_t* dst = reinterpret_cast<uint16_t*>(p.o_out) + (int64_t)b * p.oout_sb + (int64_t)row * p.oout_ss +
                          (int64_t)h * p.oout_sh + g * 64 + c * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack2<kBF16>(__uint_as_float(v[j * 8 + 0]) * inv_l, __uint_as_float(v[j * 8 + 1]) * inv_l);
            o.y = pack2<kBF16>(__uint_as_float(v[j * 8 + 2]) * inv_l, __uint_as_float(v[j * 8 + 3]) * inv_l);
            o.z = pack2<kBF16>(__uint_as_float(v[j * 8 + 4]) * inv_l, __uint_as_float(v[j * 8 + 5]) * inv_l);
            o.w = pack2<kBF16>(__uint_as_float(v[j * 8 + 6]) * inv_l, __uint_as_float(v[j * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(dst + j * 8) = o;
          }
        } else {
          float* dst = p.o_acc + (int64_t)b * p.oacc_sb + (int64_t)row * p.oacc_ss + (int64_t)h * p.oacc_sh + g * 64 +
                       c * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o;
            o.x = __uint_as_float(v[j * 4 + 0]) * inv_l;
            o.y = __uint_as_float(v[j * 4 + 1]) * inv_l;
            o.z = __uint_as_float(v[j * 4 + 2]) * inv_l;
            o.w = __uint_as_float(v[j * 4 + 3]) * inv_l;
            *reinterpret_cast<float4*>(dst + j * 4) = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 8) tmem_dealloc_2cta(0, 512);
}

template <bool kBF16>
static int launch_pair(const CUtensorMap& tmK64, const CUtensorMap& tmV, const FwdParams& p, cudaStream_t stream) {
  auto kern = fwd_pair_kernel<kBF16>;
  BA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * ((p.Sq + 2 * kBlockM - 1) / (2 * kBlockM)), p.H, p.B);
  cfg.blockDim = dim3(kFwdThreads);
  cfg.dynamicSmemBytes = kPairSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmK64, tmV, p));
  return BA_OK;
}

int launch_fwd_pair(int dtype, const CUtensorMap& tmK64, const CUtensorMap& tmV, const FwdParams& p,
                    cudaStream_t stream) {
  // n == 0 (no visible key for a whole pair) with a carried state would need a pass-through; the
  // single-CTA kernel handles that case -- not reachable through the ring drivers (offsets 0 / -1)
  return dtype == BA_DTYPE_BF16 ? launch_pair<true>(tmK64, tmV, p, stream) : launch_pair<false>(tmK64, tmV, p, stream);
}


// =====================================================================================================
// Variant 6 (BA_FWD_IMPL=6; numerically correct on B200 on every forward diag case, 819 TFLOP/s at
// S=16384 -- faster than variant 5 (775) but far from the single-CTA kernel (1200)): same CTA pair, but the
// two softmax warpgroups of a CTA own DIFFERENT key tiles (even / odd) with their own O accumulator and
// running (m, l), merged once in the epilogue -- so they never synchronise per tile and their MUFU
// phases interleave (in variant 5 both warpgroups work on the same tile in lockstep: MUFU is saturated
// during the exp phase and idle otherwise).  TMEM has no room left for Q, so QK^T is SS-form with the B
// half split across the pair (6 KiB per MMA instead of 8).
// TMEM: S_A [0,128) S_B [128,256) O_A [256,384) O_B [384,512); P_w = first 64 columns of S_w.
// =====================================================================================================
constexpr int kP6Stages = 4;
constexpr uint32_t kP6OffQ = 0;                                   // 32 KiB: this CTA's Q tile
constexpr uint32_t kP6OffK = kTileBytes;                          // 4 x 16 KiB K halves
constexpr uint32_t kP6OffV = kP6OffK + kP6Stages * kHalfBytes;    // 4 x 16 KiB V halves
constexpr uint32_t kP6OffX = kP6OffV + kP6Stages * kHalfBytes;    // fp32 [2][2][128]: (m, l) of each stream
constexpr uint32_t kP6OffBars = kP6OffX + 4 * 128 * 4;
constexpr int kPair6SmemBytes = kP6OffBars + 512;

struct __align__(8) Pair6Barriers {
  uint64_t q_full;                                  // leader: both CTAs' Q tiles landed
  uint64_t k_full[kP6Stages], k_empty[kP6Stages];
  uint64_t v_full[kP6Stages], v_empty[kP6Stages];
  uint64_t st_ready;     // leader: 16 warp arrivals -- carried O state staged in TMEM (both CTAs)
  uint64_t s_full[2];    // per CTA (multicast): S_w ready
  uint64_t p_ready[2];   // leader: 8 warp arrivals (4 warps x 2 CTAs) -- P_w written, O_w rescaled
  uint64_t o_final[2];   // per CTA (multicast): last PV of stream w done
  uint32_t tmem_base;
};

template <bool kBF16, int W, int ST>
__device__ __forceinline__ void pair6_issue_qk(uint32_t sb16) {  // S_w[256 x 128] = Q K^T (SS; K half: 64 keys)
  constexpr uint32_t idesc = make_idesc(kBF16, 256, kBlockN, false, false), hi = desc_hi(1024);
  const uint32_t q_lo = sb16 + (kP6OffQ >> 4) + desc_lo_lbo(16);
  const uint32_t k_lo = sb16 + ((kP6OffK + ST * kHalfBytes) >> 4) + desc_lo_lbo(16);
#pragma unroll
  for (int kk = 0; kk < kHeadDim / 16; ++kk) {
    const uint32_t offq = ((kk >> 2) * kBoxBytes + (kk & 3) * 32) >> 4;
    const uint32_t offk = ((kk >> 2) * (kHalfBytes / 2) + (kk & 3) * 32) >> 4;
    if (elect_one())
      asm volatile(
          "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
          "setp.ne.b32 p, %6, 0;\n\t"
          "mov.b64 da, {%1, %2};\n\t"
          "mov.b64 db, {%3, %4};\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n"
          :
          : "r"(W * 128u), "r"(q_lo + offq), "r"(hi), "r"(k_lo + offk), "r"(hi), "r"(idesc), "r"(kk > 0 ? 1u : 0u)
          : "memory");
  }
}
template <bool kBF16, int W, int ST>
__device__ __forceinline__ void pair6_issue_pv(uint32_t sb16, uint32_t acc) {  // O_w += P_w V (TS; V half: 64 d)
  constexpr uint32_t idesc = make_idesc(kBF16, 256, kHeadDim, false, true), hi = desc_hi(1024);
  const uint32_t v_lo = sb16 + ((kP6OffV + ST * kHalfBytes) >> 4) + desc_lo_lbo(kHalfBytes);
#pragma unroll
  for (int kk = 0; kk < kBlockN / 16; ++kk)
    umma_ts_2cta_lh(256 + W * 128, W * 128 + kk * 8, v_lo + kk * (16 * 128 / 16), hi, idesc, kk > 0 ? 1u : acc);
}

template <bool kBF16, int U>
__device__ __forceinline__ void pair6_mma_tile(int i, int n, uint32_t sb16, Pair6Barriers* bars, bool load_state) {
  constexpr int W = U & 1, ST = U % kP6Stages, STN = (U + 2) % kP6Stages;
  mbar_wait(&bars->v_full[ST], (i / kP6Stages) & 1);
  mbar_wait(&bars->p_ready[W], (i >> 1) & 1);
  tc_fence_after();
  // the first PV of a stream starts its accumulator, except stream A when a carried state was staged
  pair6_issue_pv<kBF16, W, ST>(sb16, (i >= 2 || (W == 0 && load_state)) ? 1u : 0u);
  umma_commit_2cta(&bars->v_empty[ST], 0x3);
  if (i + 2 >= n) umma_commit_2cta(&bars->o_final[W], 0x3);  // last tile of this stream
  if (i + 2 < n) {
    mbar_wait(&bars->k_full[STN], ((i + 2) / kP6Stages) & 1);
    tc_fence_after();
    pair6_issue_qk<kBF16, W, STN>(sb16);
    umma_commit_2cta(&bars->s_full[W], 0x3);
    umma_commit_2cta(&bars->k_empty[STN], 0x3);
  }
}

template <bool kBF16>
__global__ void __launch_bounds__(kFwdThreads, 1)
fwd_pair6_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK64,
                 const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_pair6[];
  uint8_t* smem = smem_pair6;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem + kP6OffQ;
  uint8_t* sK = smem + kP6OffK;
  uint8_t* sV = smem + kP6OffV;
  float* sX = reinterpret_cast<float*>(smem + kP6OffX);
  Pair6Barriers* bars = reinterpret_cast<Pair6Barriers*>(smem + kP6OffBars);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int h = blockIdx.y, b = blockIdx.z;
  const int row0 = (blockIdx.x >> 1) * (2 * kBlockM);
  const int r0 = row0 + (int)rank * kBlockM;
  const int n = pair_trip_count(row0, p);

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK64);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 8) {
    if (lane == 0) {
      mbar_init(&bars->q_full, 1);
      for (int i = 0; i < kP6Stages; ++i) {
        mbar_init(&bars->k_full[i], 1);
        mbar_init(&bars->k_empty[i], 1);
        mbar_init(&bars->v_full[i], 1);
        mbar_init(&bars->v_empty[i], 1);
      }
      mbar_init(&bars->st_ready, 16);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars->s_full[i], 1);
        mbar_init(&bars->p_ready[i], 8);
        mbar_init(&bars->o_final[i], 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2cta(&bars->tmem_base, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (bars->tmem_base != 0) __trap();

  if (warp == 9) {
    if (lane == 0 && n > 0) {
      if (rank == 0) mbar_arrive_expect_tx(&bars->q_full, 2 * kTileBytes);
      for (int half = 0; half < 2; ++half)
        tma_load_4d_2cta(sQ + half * kBoxBytes, &tmQ, &bars->q_full, half * 64, h, r0, b);
      for (int i = 0; i < n; ++i) {
        const int st = i % kP6Stages, ph = (i / kP6Stages) & 1;
        mbar_wait(&bars->k_empty[st], ph ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&bars->k_full[st], 2 * kHalfBytes);
        for (int half = 0; half < 2; ++half)
          tma_load_4d_2cta(sK + st * kHalfBytes + half * (kHalfBytes / 2), &tmK64, &bars->k_full[st], half * 64, h,
                           i * kBlockN + (int)rank * 64, b);
        mbar_wait(&bars->v_empty[st], ph ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&bars->v_full[st], 2 * kHalfBytes);
        tma_load_4d_2cta(sV + st * kHalfBytes, &tmV, &bars->v_full[st], (int)rank * 64, h, i * kBlockN, b);
      }
    }
  } else if (warp == 8) {
    if (rank == 0 && n > 0) {
      const uint32_t sb16 = smem_u32(smem) >> 4;
      mbar_wait(&bars->q_full, 0);
      if (p.load_state) mbar_wait(&bars->st_ready, 0);
      mbar_wait(&bars->k_full[0], 0);
      tc_fence_after();
      pair6_issue_qk<kBF16, 0, 0>(sb16);
      umma_commit_2cta(&bars->s_full[0], 0x3);
      umma_commit_2cta(&bars->k_empty[0], 0x3);
      if (n > 1) {
        mbar_wait(&bars->k_full[1], 0);
        tc_fence_after();
        pair6_issue_qk<kBF16, 1, 1>(sb16);
        umma_commit_2cta(&bars->s_full[1], 0x3);
        umma_commit_2cta(&bars->k_empty[1], 0x3);
      }
      const bool ls = p.load_state != 0;
      for (int i0 = 0; i0 < n; i0 += 4) {
        pair6_mma_tile<kBF16, 0>(i0, n, sb16, bars, ls);
        if (i0 + 1 < n) pair6_mma_tile<kBF16, 1>(i0 + 1, n, sb16, bars, ls);
        if (i0 + 2 < n) pair6_mma_tile<kBF16, 2>(i0 + 2, n, sb16, bars, ls);
        if (i0 + 3 < n) pair6_mma_tile<kBF16, 3>(i0 + 3, n, sb16, bars, ls);
      }
    }
  } else if (n > 0) {
    // ============================================================ softmax: warpgroup w owns tiles w, w+2, ...
    const int w = warp >> 2;
    const int t = threadIdx.x & 127;
    const int row = r0 + t;
    const bool valid_row = row < p.Sq;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = lane_base + w * 128;
    const uint32_t tO = lane_base + 256 + w * 128;
    const float scale_log2 = p.scale_log2;
    const int limit = p.causal ? min(row + p.causal_off, p.Sk - 1) : p.Sk - 1;
    const int tile_min_limit = p.causal ? min(r0 + p.causal_off, p.Sk - 1) : p.Sk - 1;
    const int n_w = (n - w + 1) >> 1;  // tiles of this stream

    float m = -INFINITY, l = 0.f;
    if (p.load_state) {  // the carried state goes to stream A; both warpgroups stage half of its O columns
      if (w == 0) {
        float lse_prev = -INFINITY;
        if (valid_row) lse_prev = p.lse[(int64_t)b * p.lse_sb + (int64_t)h * p.lse_sh + row];
        if (lse_prev != -INFINITY) {
          m = lse_prev * kLog2e;
          l = 1.f;
        }
      }
      const float* src = p.o_acc + (int64_t)b * p.oacc_sb + (int64_t)row * p.oacc_ss + (int64_t)h * p.oacc_sh + w * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
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
        tmem_st_x32(lane_base + 256 + w * 64 + c * 32, v);  // into O_A
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&bars->st_ready, 0);
    }

    for (int jj = 0; jj < n_w; ++jj) {
      const int i = 2 * jj + w;  // global tile index
      mbar_wait(&bars->s_full[w], jj & 1);
      tc_fence_after();
      uint32_t sr[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(tS + c * 32, sr + c * 32);
      tmem_wait_ld();
      float* s = reinterpret_cast<float*>(sr);
      const int kbase = i * kBlockN;
      if (kbase + kBlockN - 1 > tile_min_limit) {
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (kbase + c > limit) s[c] = -INFINITY;
      }
      float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
      for (int c = 4; c < 128; c += 4) {
        mx0 = fmaxf(mx0, s[c]);
        mx1 = fmaxf(mx1, s[c + 1]);
        mx2 = fmaxf(mx2, s[c + 2]);
        mx3 = fmaxf(mx3, s[c + 3]);
      }
      const float m_new = fmaxf(m, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2);
      const bool grow = m_new > m + kRescaleThreshold;
      if (__any_sync(0xffffffffu, grow)) {
        // S_w of this tile was issued behind the previous PV of this stream: O_w is up to date
        const bool o_live = (jj > 0) || (w == 0 && p.load_state);
        if (o_live) {
          const float f = (m == -INFINITY) ? 0.f : ex2(m - m_new);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld_x32(tO + c * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * f);
            tmem_st_x32(tO + c * 32, v);
          }
          l *= f;
        }
        m = m_new;
      }
      const float neg_m = (m == -INFINITY) ? 0.f : -m;
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
      for (int c = 0; c < 128; c += 4) {
        const float p0 = ex2(fmaf(s[c], scale_log2, neg_m));
        const float p1 = ex2(fmaf(s[c + 1], scale_log2, neg_m));
        const float p2 = ex2(fmaf(s[c + 2], scale_log2, neg_m));
        const float p3 = ex2(fmaf(s[c + 3], scale_log2, neg_m));
        sum0 += p0;
        sum1 += p1;
        sum2 += p2;
        sum3 += p3;
        sr[c / 2] = pack2<kBF16>(p0, p1);      // in place: slot c/2 <= c already consumed
        sr[c / 2 + 1] = pack2<kBF16>(p2, p3);
      }
      l += (sum0 + sum1) + (sum2 + sum3);
      tmem_st_x32(tS, sr);
      tmem_st_x32(tS + 32, sr + 32);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&bars->p_ready[w], 0);
    }

    // ---------------------------------------------------------- epilogue: merge the two streams
    if (n_w > 0) {
      mbar_wait(&bars->o_final[w], 0);
      tc_fence_after();
    }
    sX[(w * 2 + 0) * 128 + t] = m;
    sX[(w * 2 + 1) * 128 + t] = l;
    named_bar_sync(2, 256);
    const float m_o = sX[((w ^ 1) * 2 + 0) * 128 + t], l_o = sX[((w ^ 1) * 2 + 1) * 128 + t];
    const float m_t = fmaxf(m, m_o);
    const float f_me = (m == -INFINITY) ? 0.f : ex2(m - m_t);
    const float f_ot = (m_o == -INFINITY) ? 0.f : ex2(m_o - m_t);
    const float l_tot = l * f_me + l_o * f_ot;
    const float inv_l = l_tot > 0.f ? 1.f / l_tot : 0.f;
    const float fa = (w == 0 ? f_me : f_ot) * inv_l, fb = (w == 0 ? f_ot : f_me) * inv_l;  // weights of O_A, O_B
    // a stream that never ran a PV (n == 1 -> stream B, without state) has an uninitialised accumulator
    const bool live_a = true, live_b = n > 1;
    if (w == 0 && valid_row)
      p.lse[(int64_t)b * p.lse_sb + (int64_t)h * p.lse_sh + row] = l_tot > 0.f ? (m_t + lg2(l_tot)) * kLn2 : -INFINITY;
#pragma unroll
    for (int c = 0; c < 2; ++c) {  // warpgroup w merges and writes d columns [64 w, 64 w + 64)
      uint32_t va[32], vb[32];
      tmem_ld_x32(lane_base + 256 + w * 64 + c * 32, va);
      tmem_ld_x32(lane_base + 384 + w * 64 + c * 32, vb);
      tmem_wait_ld();
      float o[32];
#pragma unroll
      for (int j = 0; j < 32; ++j)
        o[j] = (live_a ? __uint_as_float(va[j]) * fa : 0.f) + (live_b ? __uint_as_float(vb[j]) * fb : 0.f);
      if (valid_row) {
        if (p.store_lowp) {
          uint16_t* dst = reinterpret_cast<uint16_t*>(p.o_out) + (int64_t)b * p.oout_sb + (int64_t)row * p.oout_ss +
                          (int64_t)h * p.oout_sh + w * 64 + c * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 q4;
            q4.x = pack2<kBF16>(o[j * 8 + 0], o[j * 8 + 1]);
            q4.y = pack2<kBF16>(o[j * 8 + 2], o[j * 8 + 3]);
            q4.z = pack2<kBF16>(o[j * 8 + 4], o[j * 8 + 5]);
            q4.w = pack2<kBF16>(o[j * 8 + 6], o[j * 8 + 7]);
            *reinterpret_cast<uint4*>(dst + j * 8) = q4;
          }
        } else {
          float* dst = p.o_acc + (int64_t)b * p.oacc_sb + (int64_t)row * p.oacc_ss + (int64_t)h * p.oacc_sh + w * 64 +
                       c * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(dst + j * 4) = make_float4(o[j * 4], o[j * 4 + 1], o[j * 4 + 2], o[j * 4 + 3]);
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 8) tmem_dealloc_2cta(0, 512);
}

template <bool kBF16>
static int launch_pair6(const CUtensorMap& tmQ, const CUtensorMap& tmK64, const CUtensorMap& tmV, const FwdParams& p,
                        cudaStream_t stream) {
  auto kern = fwd_pair6_kernel<kBF16>;
  BA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPair6SmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * ((p.Sq + 2 * kBlockM - 1) / (2 * kBlockM)), p.H, p.B);
  cfg.blockDim = dim3(kFwdThreads);
  cfg.dynamicSmemBytes = kPair6SmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmQ, tmK64, tmV, p));
  return BA_OK;
}

int launch_fwd_pair6(int dtype, const CUtensorMap& tmQ, const CUtensorMap& tmK64, const CUtensorMap& tmV,
                     const FwdParams& p, cudaStream_t stream) {
  return dtype == BA_DTYPE_BF16 ? launch_pair6<true>(tmQ, tmK64, tmV, p, stream)
                                : launch_pair6<false>(tmQ, tmK64, tmV, p, stream);
}

}  // namespace ba
