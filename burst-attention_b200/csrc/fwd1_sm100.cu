// fwd1_chunk_kernel: forward tile kernel with ONE 128-row Q tile per CTA and BOTH GEMMs in TS form.
//
// Why: on this part one K = 16 step of tcgen05.mma costs N/2 clk when A comes from TMEM but
// max(74, N/2) clk when A comes from shared memory (tools/ubench.py).  The two-Q-tiles-per-CTA kernel
// (fwd_sm100.cu) must keep Q in shared memory -- two Q tiles, two double-buffered S and two O accumulators do
// not leave 128 TMEM columns for Q -- and issues its QK^T as N = 64 SS steps at 74 clk for 32 clk of math: it is
// bound by exactly that (3392 tensor-pipe clk per 2 x 128 x 128 tile pair, as measured).  With one Q tile per CTA
// everything fits: Q (16-bit, 64 columns) + S double-buffered 128 keys wide (256) + O (128) = 448 columns, QK^T
// and PV are both TS-form at N = 128: 8 x 64 + 8 x 64 = 1024 clk per 128 x 128 tile = the MUFU time of its
// 16384 exponentials.  One Q tile per CTA halves the reuse of a K/V tile, so K/V tiles are loaded ONCE PER
// CLUSTER OF TWO CTAs: each CTA's producer fetches half of every tile (64 rows) and TMA multicasts it into both
// CTAs' shared memory; a stage is recycled when BOTH CTAs' MMAs have consumed it (tcgen05.commit multicast on a
// count-2 barrier).  The MMAs themselves are plain cta_group::1.
//
//   warp 5      producer: K_i / V_i half tiles, multicast, 3-stage rings
//   warp 4      MMA issuer:  S[j&1] = Q K_j^T (TS, Q in TMEM)   O += P_j V_j (TS, P in TMEM over S[j&1])
//   warps 0-3   softmax (thread = row = TMEM lane): load Q row -> TMEM once; per key tile: S -> max / exp2 / P
// TMEM (512 columns): Q [0,64)  S0 [128,256)  S1 [256,384)  O [384,512)
// Same carried-state / mask / epilogue semantics as fwd_sm100.cu (head dim 128, no bias; other cases use that kernel).
#include <math.h>
#include <stdlib.h>

#include "fwd_common.cuh"
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace ba {

constexpr int kF1Threads = 192;
constexpr int kF1Stages = 3;
constexpr int kF1Tile = 128 * 128 * 2;  // 32 KiB: [128 keys][128 d] 16-bit = 2 SW128 boxes of [128][64]
constexpr uint32_t kF1OffK = 0;
constexpr uint32_t kF1OffV = kF1Stages * kF1Tile;
constexpr uint32_t kF1OffBars = 2 * kF1Stages * kF1Tile;
constexpr int kF1Smem = kF1OffBars + 256;
constexpr uint32_t kF1TQ = 0, kF1TS0 = 128, kF1TO = 384;  // TMEM columns; S[b] = kF1TS0 + 128 b

struct __align__(8) F1Barriers {
  uint64_t k_full[kF1Stages], k_empty[kF1Stages];  // empty: count 2 (this CTA's and the peer's MMA warp)
  uint64_t v_full[kF1Stages], v_empty[kF1Stages];
  uint64_t q_ready;     // softmax warps -> MMA: Q staged in TMEM (count 4)
  uint64_t s_full[2];   // MMA -> softmax
  uint64_t p_ready[2];  // softmax -> MMA (count 4)
  uint64_t o_done;      // MMA -> softmax: a PV has been accumulated (one completion per key tile)
  uint64_t o_final;     // MMA -> softmax: the last PV of this CTA
  uint32_t tmem_base;
};

// multicast TMA load: the box lands at the same shared-memory offset in every CTA of `mask`, and completes
// `bytes` on the barrier at the same offset in each of them
BA_DEVICE void tma_load_4d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3,
                              uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "h"(mask)
      : "memory");
}
// MMA completion -> arrive on the barrier at this offset in every CTA of `mask` (cta_group::1 MMAs)
BA_DEVICE void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  if (elect_one())
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
}

// number of 128-key tiles a 128-row Q tile starting at r0 must visit
__device__ __forceinline__ int f1_trip_count(int r0, const FwdParams& p) {
  if (r0 >= p.Sq) return 0;
  const int r_last = min(r0 + kBlockM - 1, p.Sq - 1);
  const int max_limit = p.causal ? min(r_last + p.causal_off, p.Sk - 1) : p.Sk - 1;
  return max_limit < 0 ? 0 : max_limit / kBlockN + 1;
}

template <bool kBF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kF1Threads, 1)
fwd1_chunk_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sK = smem + kF1OffK;
  uint8_t* sV = smem + kF1OffV;
  F1Barriers* bars = reinterpret_cast<F1Barriers*>(smem + kF1OffBars);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const uint32_t rank = cluster_ctarank();
  const int n_pairs = gridDim.x >> 1;
  const int pair = blockIdx.x >> 1;
  // causal: the last Q tiles see the most keys -- schedule them first (longest-processing-time order)
  const int row_pair = (p.causal ? n_pairs - 1 - pair : pair) * (2 * kBlockM);
  const int r0 = row_pair + (int)rank * kBlockM;
  const int n_own = f1_trip_count(r0, p);
  const int n_all = max(f1_trip_count(row_pair, p), f1_trip_count(row_pair + kBlockM, p));  // tiles the PAIR streams

  if (warp == 5 && lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 4) {
    if (lane == 0) {
      for (int i = 0; i < kF1Stages; ++i) {
        mbar_init(&bars->k_full[i], 1);
        mbar_init(&bars->k_empty[i], 2);
        mbar_init(&bars->v_full[i], 1);
        mbar_init(&bars->v_empty[i], 2);
      }
      mbar_init(&bars->q_ready, 4);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars->s_full[i], 1);
        mbar_init(&bars->p_ready[i], 4);
      }
      mbar_init(&bars->o_done, 1);
      mbar_init(&bars->o_final, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before anyone multicasts into them
  tc_fence_after();
  if (bars->tmem_base != 0) __trap();  // all 512 columns: the allocation starts at TMEM address 0

  if (warp == 5) {
    // ============================================================ producer: my half (64 keys) of every tile, to both CTAs
    if (lane == 0) {
      for (int i = 0; i < n_all; ++i) {
        const int st = i % kF1Stages, ph = (i / kF1Stages) & 1;
        const int key0 = i * kBlockN + (int)rank * 64;
        mbar_wait(&bars->k_empty[st], ph ^ 1);  // both CTAs have consumed the tile that was here
        mbar_arrive_expect_tx(&bars->k_full[st], kF1Tile);
        for (int half = 0; half < 2; ++half)
          tma_load_4d_mc(sK + st * kF1Tile + half * kBoxBytes + rank * 8192, &tmK, &bars->k_full[st], half * 64, h, key0,
                         b, 0x3);
        mbar_wait(&bars->v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&bars->v_full[st], kF1Tile);
        for (int half = 0; half < 2; ++half)
          tma_load_4d_mc(sV + st * kF1Tile + half * kBoxBytes + rank * 8192, &tmV, &bars->v_full[st], half * 64, h, key0,
                         b, 0x3);
      }
    }
  } else if (warp == 4) {
    // ============================================================ MMA issuer
    const uint32_t sb16 = smem_u32(smem) >> 4;
    constexpr uint32_t idesc_qk = make_idesc(kBF16, kBlockM, kBlockN, false, false);  // A TMEM, B K-major
    constexpr uint32_t idesc_pv = make_idesc(kBF16, kBlockM, 128, false, true);       // A TMEM, B MN-major
    constexpr uint32_t hi = desc_hi(1024);
    auto issue_qk = [&](int j) {  // S[j&1] = Q K_j^T
      const int st = j % kF1Stages;
      const uint32_t k_lo = sb16 + ((kF1OffK + st * kF1Tile) >> 4) + desc_lo_lbo(16);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = ((kk >> 2) * kBoxBytes + (kk & 3) * 32) >> 4;
        umma_ts_lh(kF1TS0 + (j & 1) * 128, kF1TQ + kk * 8, k_lo + off, hi, idesc_qk, kk > 0 ? 1u : 0u);
      }
    };
    if (n_own > 0) {
      mbar_wait(&bars->q_ready, 0);
      tc_fence_after();
    }
    // K tiles are consumed two ahead of the softmax (S double buffer), V tiles in step with it
    for (int j = 0; j < min(2, n_all); ++j) {
      mbar_wait(&bars->k_full[j % kF1Stages], 0);
      tc_fence_after();
      if (j < n_own) {
        issue_qk(j);
        umma_commit(&bars->s_full[j & 1]);
      }
      umma_commit_mc(&bars->k_empty[j % kF1Stages], 0x3);
    }
    const bool ls = p.load_state != 0;
    for (int j = 0; j < n_all; ++j) {
      const int st = j % kF1Stages;
      mbar_wait(&bars->v_full[st], (j / kF1Stages) & 1);
      if (j < n_own) {
        mbar_wait(&bars->p_ready[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t v_lo = sb16 + ((kF1OffV + st * kF1Tile) >> 4) + desc_lo_lbo(kBoxBytes);
        const uint32_t acc = (j > 0 || ls) ? 1u : 0u;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ts_lh(kF1TO, kF1TS0 + (j & 1) * 128 + kk * 8, v_lo + kk * (16 * 128 / 16), hi, idesc_pv,
                     kk > 0 ? 1u : acc);
        umma_commit(&bars->o_done);
        if (j == n_own - 1) umma_commit(&bars->o_final);
      }
      umma_commit_mc(&bars->v_empty[st], 0x3);
      if (j + 2 < n_all) {
        const int st2 = (j + 2) % kF1Stages;
        mbar_wait(&bars->k_full[st2], ((j + 2) / kF1Stages) & 1);
        tc_fence_after();
        if (j + 2 < n_own) {
          issue_qk(j + 2);  // overwrites S[j&1] = P_j: in order behind the PV above
          umma_commit(&bars->s_full[j & 1]);
        }
        umma_commit_mc(&bars->k_empty[st2], 0x3);
      }
    }
  } else {
    // ============================================================ softmax warps (thread = row)
    const int t = threadIdx.x;  // 0..127 == TMEM lane
    const int row = r0 + t;
    const bool valid_row = row < p.Sq;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    if (n_own > 0) {
      // ---- Q row -> TMEM (16-bit pairs, 64 columns): the A operand of every QK^T of this CTA
      const uint16_t* qrow = p.q + (int64_t)b * p.q_sb + (int64_t)row * p.q_ss + (int64_t)h * p.q_sh;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 x = valid_row ? __ldg(reinterpret_cast<const uint4*>(qrow + c * 64 + i * 8)) : make_uint4(0, 0, 0, 0);
          v[i * 4 + 0] = x.x, v[i * 4 + 1] = x.y, v[i * 4 + 2] = x.z, v[i * 4 + 3] = x.w;
        }
        tmem_st_x32(lane_base + kF1TQ + c * 32, v);
      }
    }
    const uint32_t tO = lane_base + kF1TO;
    const float scale_log2 = p.scale_log2;
    const int limit = p.causal ? min(row + p.causal_off, p.Sk - 1) : p.Sk - 1;
    const int tile_min_limit = p.causal ? min(r0 + p.causal_off, p.Sk - 1) : p.Sk - 1;
    float m = -INFINITY, l = 0.f;
    if (p.load_state && r0 < p.Sq) {
      float lse_prev = -INFINITY;
      if (valid_row) lse_prev = p.lse[(int64_t)b * p.lse_sb + (int64_t)h * p.lse_sh + row];
      if (lse_prev != -INFINITY) {
        m = lse_prev * kLog2e;
        l = 1.f;
      }
      const float* src = p.o_acc + (int64_t)b * p.oacc_sb + (int64_t)row * p.oacc_ss + (int64_t)h * p.oacc_sh;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 f = valid_row ? __ldg(reinterpret_cast<const float4*>(src + c * 32 + j * 4))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
          v[j * 4 + 0] = __float_as_uint(f.x), v[j * 4 + 1] = __float_as_uint(f.y);
          v[j * 4 + 2] = __float_as_uint(f.z), v[j * 4 + 3] = __float_as_uint(f.w);
        }
        tmem_st_x32(tO + c * 32, v);
      }
    }
    if (n_own > 0) {
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->q_ready);
    } else if (p.load_state) {
      tmem_wait_st();
    }

    for (int j = 0; j < n_own; ++j) {
      const int bf = j & 1;
      const uint32_t tS = lane_base + kF1TS0 + bf * 128;
      mbar_wait(&bars->s_full[bf], (j >> 1) & 1);
      tc_fence_after();
      uint32_t sr[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(tS + c * 32, sr + c * 32);
      tmem_wait_ld();
      float* s = reinterpret_cast<float*>(sr);
      const int kbase = j * kBlockN;
      if (kbase + kBlockN - 1 > tile_min_limit) {  // warp-group-uniform
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
      const bool grow = m_new > m + kRescaleThreshold;  // also true for m == -inf, m_new finite
      if (__any_sync(0xffffffffu, grow)) {
        const bool o_live = (j > 0) || p.load_state;
        if (o_live) {
          if (j > 0) {
            mbar_wait(&bars->o_done, (j - 1) & 1);
            tc_fence_after();
          }
          const float f = (m == -INFINITY) ? 0.f : ex2(m - m_new);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
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
#pragma unroll
      for (int c = 0; c < 128; c += 4) {
        const float p0 = ex2(fmaf(s[c], scale_log2, neg_m));
        const float p1 = ex2(fmaf(s[c + 1], scale_log2, neg_m));
        const float p2 = ex2(fmaf(s[c + 2], scale_log2, neg_m));
        const float p3 = ex2(fmaf(s[c + 3], scale_log2, neg_m));
        sum0 += p0, sum1 += p1, sum2 += p2, sum3 += p3;
        sr[c / 2] = pack2<kBF16>(p0, p1);  // in place: slot c/2 <= c has been consumed
        sr[c / 2 + 1] = pack2<kBF16>(p2, p3);
      }
      l += (sum0 + sum1) + (sum2 + sum3);
      tmem_st_x32(tS, sr);  // P (16-bit) over the first 64 columns of its S buffer
      tmem_st_x32(tS + 32, sr + 32);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_ready[bf]);
    }

    // ---------------------------------------------------------- epilogue
    if (r0 < p.Sq) {
      if (n_own > 0) {
        mbar_wait(&bars->o_final, 0);
        tc_fence_after();
      }
      const bool o_live = (n_own > 0) || p.load_state;
      const float inv_l = l > 0.f ? 1.f / l : 0.f;
      const float lse_out = l > 0.f ? (m + lg2(l)) * kLn2 : -INFINITY;
      if (valid_row) p.lse[(int64_t)b * p.lse_sb + (int64_t)h * p.lse_sh + row] = lse_out;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
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
            uint16_t* dst = reinterpret_cast<uint16_t*>(p.o_out) + (int64_t)b * p.oout_sb + (int64_t)row * p.oout_ss +
                            (int64_t)h * p.oout_sh + c * 32;
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
            float* dst = p.o_acc + (int64_t)b * p.oacc_sb + (int64_t)row * p.oacc_ss + (int64_t)h * p.oacc_sh + c * 32;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 o;
              o.x = __uint_as_float(v[i * 4 + 0]) * inv_l, o.y = __uint_as_float(v[i * 4 + 1]) * inv_l;
              o.z = __uint_as_float(v[i * 4 + 2]) * inv_l, o.w = __uint_as_float(v[i * 4 + 3]) * inv_l;
              *reinterpret_cast<float4*>(dst + i * 4) = o;
            }
          }
        }
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  cluster_sync_all();  // the peer may still be multicasting into / arriving on this CTA's shared memory
  if (warp == 4) tmem_dealloc(0, 512);
}

int launch_fwd1(int dtype, const CUtensorMap& tmK64, const CUtensorMap& tmV64, const FwdParams& p, cudaStream_t stream) {
  dim3 grid(2 * ((p.Sq + 2 * kBlockM - 1) / (2 * kBlockM)), p.H, p.B);
  if (dtype == BA_DTYPE_BF16) {
    auto kern = fwd1_chunk_kernel<true>;
    BA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kF1Smem));
    kern<<<grid, kF1Threads, kF1Smem, stream>>>(tmK64, tmV64, p);
  } else {
    auto kern = fwd1_chunk_kernel<false>;
    BA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kF1Smem));
    kern<<<grid, kF1Threads, kF1Smem, stream>>>(tmK64, tmV64, p);
  }
  BA_CHECK_CUDA(cudaGetLastError());
  return BA_OK;
}

}  // namespace ba
