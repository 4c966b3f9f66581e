// ba_bwd_chunk: one ring round of the backward on sm_100a.
//
// Replaces the reference's per-round flash_attn_2_cuda.bwd call
// (burst_utils.py:180-249; Triton twin lao.py:295-595) AND the three full-tensor
// "dq += buf; dk += buf; dv += buf" passes of burst_attn_interface.py:379-390:
// the kernel accumulates straight into fp32 dQ / dK / dV accumulators.
// delta = rowsum(O*dO) and the final lse are inputs (they travel with the
// Q-bundle), so O itself is never read here.
//
// One CTA owns one 128-key block of the home K/V chunk for one (batch, head)
// and loops over the 128-row blocks of the visiting Q-bundle:
//   S^T  = K Q_i^T - lse/scale   (SS; TMEM rows = keys, cols = queries; the row statistic is one extra K = 16 step)
//   dP^T = V dO_i^T - delta      (SS; same)
//   P^T  = exp2(S^T*c [+ bias]) , dS^T = P^T o dP^T             (8 compute warps, thread = key row)
//   dV  += P^T dO_i          (TS; P^T 16-bit in TMEM aliasing S^T)
//   dK  += dS^T Q_i          (TS; dS^T 16-bit in TMEM over the consumed dP^T columns)
//   dQ_i = dS K              (SS; dS^T also staged once in smem, read MN-major) -> TMEM (aliasing dP^T)
//   dQ_i -> 4 reduce warps -> smem -> cp.reduce.async.bulk.tensor (fp32 add in L2) -> dq_acc
// TMEM (512 cols): S^T/P^T [0,128)  dP^T/dS^T/dQ [128,256)  dK [256,256+D)  dV [384,384+D).
// smem (D = 128): K 32K, V 32K, Q 2x32K, dO 32K, dS 32K, dQ staging 2x16K, row-stat operand tile 2K (+640 B constants).
// Head dim D = 64 or 128 (template): the operand tiles shrink to one SW128 box, N = D for dV / dK / dQ.
#include <math.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "host_common.h"
#include "sm100_ptx.cuh"

namespace ba {

constexpr int kBwdThreads = 512;  // warps 0-7 compute, 8-11 dQ reduce, 12 MMA, 13 load, 14-15 idle (register donors)
constexpr int kTile = 128;
constexpr int kBoxB = kTile * 64 * 2;       // 16 KiB: 128 rows x 64 cols SW128 box
constexpr int kDsTileB = kTile * kTile * 2;  // 32 KiB: the dS^T tile [128 keys][128 q], two boxes
constexpr int kDqStageB = kTile * 32 * 4;  // 16 KiB: 128 rows x 32 fp32 cols SW128 box
constexpr float kBwdLog2e = 1.4426950408889634f;

template <int N>
__device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// ---- row statistics folded into the GEMMs ("fold") --------------------------------------------------------
// P^T = exp2(S^T c - lse2[q]) and dS^T = P^T o (dP^T - delta[q]) need a per-COLUMN value in a thread = key-row
// layout; reading them from shared memory costs 512 broadcast-LDS wavefronts per Q block.  Instead the tensor core
// adds them: one extra K = 16 step per GEMM,  S^T += 1[key] (x) (-lse/scale)[q],  dP^T += 1[key] (x) (-delta)[q],
// with the fp32 value split into two 16-bit parts (hi + lo: 16 significant bits with bf16, 22 with fp16 -- an
// error of <= 2^-17 |lse|/scale in the exponent, two orders of magnitude below the 16-bit rounding of P; the
// products with 1.0 are exact and the accumulation is fp32).
// Operand tiles are K-major, NO swizzle (core matrix = 8 rows x 16 B, contiguous; sm100_ptx.cuh):
//   B (written by the loader warp, 2 KiB): row q = 8 slots = [l0 l1 d0 d1 | l0' l1' d0' d1']: the first four belong
//     to even Q blocks, the last four to odd ones, so the statistics of block it+1 are written (8-byte stores) while
//     MMAs of block it still read their half; the tile is zeroed at kernel start (a slot nobody has written must
//     not hold a NaN pattern: it meets a 0 of the A operand).  Its second K chunk (k = 8..15) re-reads the first
//     (LBO = 0): harmless, A is zero there;
//   A (constant): ONE 128-byte core matrix of identical rows per pattern -- ones in the two slots to pick, zeros
//     elsewhere -- shared by all 16 row groups (SBO = 0): lse/even, delta/even, lse/odd, delta/odd, and a zero core
//     matrix that serves as every pattern's second K chunk.
template <bool kBF16>
BA_DEVICE void split2(float x, uint32_t& h0, uint32_t& h1) {
  h0 = to16<kBF16>(x);
  h1 = to16<kBF16>(x - from16<kBF16>(h0));
}

struct BwdParams {
  const float* lse;
  int64_t lse_sb, lse_sh;
  const float* delta;
  int64_t dl_sb, dl_sh;
  float* dk_acc;
  int64_t dk_sb, dk_ss, dk_sh;
  float* dv_acc;
  int64_t dv_sb, dv_ss, dv_sh;
  int B, Sq, Sk, H;
  float scale, scale_log2, inv_scale;
  int causal, causal_off;
  const float* bias;  // optional additive bias per key [B|1, H, Sk] (fp32), or null
  int64_t bias_sb, bias_sh;
  int* sem;     // deterministic mode: [B][H][nQ] turn counters ordering the dQ reductions by key block; else null
  int* ticket;  // deterministic mode: [B][H] key-block tickets (a CTA's key block = the order in which it STARTED)
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct __align__(8) BwdBarriers {
  uint64_t kv_full;
  uint64_t q_full[2], q_empty[2], stat_full[2];
  uint64_t do_full, do_empty;
  uint64_t s_full, p_ready, dp_full, ds_ready, dq_full, dq_free, dkv_full;
  uint32_t tmem_base;
  int key_block;  // deterministic mode: this CTA's ticket
};

// smem carve-up (bytes from the 1 KiB-aligned base)
// head dim kD (64 or 128): an operand tile (K, V, Q, dO) is [128 rows][kD] = kD/64 boxes
template <int kD>
struct BwdLayout {
  static_assert(kD == 64 || kD == 128, "head dim 64 or 128");
  static constexpr int kTileB = kTile * kD * 2;
  static constexpr int kBoxes = kD / 64;
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kOffK + kTileB;
  static constexpr int kOffQ = kOffV + kTileB;        // 2 stages
  static constexpr int kOffDO = kOffQ + 2 * kTileB;   // 1 stage
  static constexpr int kOffDS = kOffDO + kTileB;
  static constexpr int kOffDQ = kOffDS + kDsTileB;    // 2 staging boxes
  static constexpr int kOffStat = kOffDQ + 2 * kDqStageB;  // fold: [128 q rows][16 B] row-stat operand tile; else [2][lse2|delta] fp32
  static constexpr int kOffFoldA = kOffStat + 2 * 2 * kTile * 4;  // fold: 4 constant one-patterns + 1 zero core matrix, 128 B each
  static constexpr int kOffBar = kOffFoldA + 640;
  static constexpr int kSmemBytes = kOffBar + 256;  // no align slack: the dynamic smem base is checked to be 1 KiB aligned
  static_assert(kSmemBytes <= 232448, "backward kernel exceeds 227 KiB of shared memory");
};

template <bool kBF16, bool kFold, int kD>
__global__ void __launch_bounds__(kBwdThreads, 1)
bwd_chunk_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                 const __grid_constant__ CUtensorMap tmDQ, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B atoms need a 1 KiB-aligned base
  using L = BwdLayout<kD>;
  constexpr int kTileB = L::kTileB, kBoxes = L::kBoxes, kKSteps = kD / 16, kChunks = kD / 32;
  constexpr int kOffK = L::kOffK, kOffV = L::kOffV, kOffQ = L::kOffQ, kOffDO = L::kOffDO, kOffDS = L::kOffDS,
                kOffDQ = L::kOffDQ, kOffStat = L::kOffStat, kOffFoldA = L::kOffFoldA, kOffBar = L::kOffBar;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sDO = smem + kOffDO;
  uint8_t* sDS = smem + kOffDS;
  uint8_t* sDQ = smem + kOffDQ;
  float* sStat = reinterpret_cast<float*>(smem + kOffStat);
  BwdBarriers* bars = reinterpret_cast<BwdBarriers*>(smem + kOffBar);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  // Key block of this CTA.  Deterministic mode orders the dQ reductions by key block and makes a CTA wait for
  // all lower key blocks; to make that wait deadlock-free without assuming anything about the order in which
  // the hardware dispatches blockIdx.x, the key block is a ticket drawn when the CTA starts: every lower
  // ticket then belongs to a CTA that is already resident.
  int kb = blockIdx.x;
  if (p.sem) {
    if (threadIdx.x == 0) bars->key_block = atomicAdd(p.ticket + b * p.H + h, 1);
    __syncthreads();
    kb = bars->key_block;
  }
  const int k0 = kb * kTile;
  const int nQ = (p.Sq + kTile - 1) / kTile;
  // first Q block that can see any key of this block: q >= k0 - off
  const int i_begin = p.causal ? max(0, k0 - p.causal_off) / kTile : 0;
  const int n_it = max(0, nQ - i_begin);
  if (n_it == 0) return;  // nothing visible: dK/dV contributions are zero (uniform exit, no barriers yet)

  if (warp == 13 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
  }
  if (warp == 12) {
    if (lane == 0) {
      mbar_init(&bars->kv_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&bars->q_full[s], 1);
        mbar_init(&bars->q_empty[s], 1);
        mbar_init(&bars->stat_full[s], 1);
      }
      mbar_init(&bars->do_full, 1);
      mbar_init(&bars->do_empty, 1);
      mbar_init(&bars->s_full, 1);
      mbar_init(&bars->p_ready, 8);   // one elected arrive per compute warp
      mbar_init(&bars->dp_full, 1);
      mbar_init(&bars->ds_ready, 8);
      mbar_init(&bars->dq_full, 1);
      mbar_init(&bars->dq_free, 4);   // one elected arrive per reduce warp
      mbar_init(&bars->dkv_full, 1);
      fence_mbar_init();
    }
    if constexpr (kFold) {  // constant A operands of the row-stat K steps: 5 core matrices of 8 rows x 16 B
      constexpr uint32_t one2 = kBF16 ? 0x3F803F80u : 0x3C003C00u;
      for (int i = lane; i < 40; i += 32) {
        const int m = i >> 3;  // 0: lse/even (slots 0,1)  1: delta/even (2,3)  2: lse/odd (4,5)  3: delta/odd (6,7)  4: zeros
        *reinterpret_cast<uint4*>(smem + kOffFoldA + i * 16) =
            make_uint4(m == 0 ? one2 : 0u, m == 1 ? one2 : 0u, m == 2 ? one2 : 0u, m == 3 ? one2 : 0u);
      }
      fence_proxy_async_smem();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDK = tmem_base + 256, tDV = tmem_base + 384;

  // register re-distribution (512 threads x 128 at launch): the MMA/load warpgroup donates to the
  // dQ-reduce warpgroup, which needs a whole 128-column TMEM row in registers to free TMEM early

  if (warp == 13) {
    // ============================================================ loader
    reg_dec<64>();
    if constexpr (kFold) {  // (see the fold comment: unwritten slots must read as 0)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(smem + kOffStat + (lane + 32 * j) * 16) = make_uint4(0u, 0u, 0u, 0u);
      fence_proxy_async_smem();
      __syncwarp();
    }
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars->kv_full, 2 * kTileB);
      for (int half = 0; half < kBoxes; ++half) {
        tma_load_4d(sK + half * kBoxB, &tmK, &bars->kv_full, half * 64, h, k0, b);
        tma_load_4d(sV + half * kBoxB, &tmV, &bars->kv_full, half * 64, h, k0, b);
      }
    }
    for (int it = 0; it < n_it; ++it) {
      const int q0 = (i_begin + it) * kTile;
      const int st = it & 1;
      // row statistics of this Q block (lane handles rows lane + 32 j): lse, delta
      float sl[4], sd[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = q0 + lane + 32 * j;
        float l = INFINITY, dl = 0.f;  // +inf: padding row, or a row that saw no key at all -> P = 0
        if (row < p.Sq) {
          l = __ldg(p.lse + (int64_t)b * p.lse_sb + (int64_t)h * p.lse_sh + row);
          dl = __ldg(p.delta + (int64_t)b * p.dl_sb + (int64_t)h * p.dl_sh + row);
          if (l == -INFINITY) l = INFINITY;
        }
        sl[j] = l, sd[j] = dl;
      }
      mbar_wait(&bars->q_empty[st], ((it >> 1) & 1) ^ 1);  // also: the compute warps are done with stat stage st
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars->q_full[st], kTileB);
        for (int half = 0; half < kBoxes; ++half)
          tma_load_4d(sQ + st * kTileB + half * kBoxB, &tmQ, &bars->q_full[st], half * 64, h, q0, b);
      }
      if constexpr (kFold) {
        // this block's half of the operand rows (slots 4 (it & 1) ..): last read by the MMAs of block it - 2, which
        // completed before q_empty[st] above was committed
        constexpr float kBig = kBF16 ? 1e30f : 60000.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = lane + 32 * j;
          const float xl = fminf(fmaxf(-sl[j] * p.inv_scale, -kBig), kBig);  // S^T + xl = (S^T c - lse2) / c
          const float xd = fminf(fmaxf(-sd[j], -kBig), kBig);
          uint32_t l0, l1, d0, d1;
          split2<kBF16>(xl, l0, l1);
          split2<kBF16>(xd, d0, d1);
          *reinterpret_cast<uint2*>(smem + kOffStat + (r >> 3) * 128 + (r & 7) * 16 + st * 8) =
              make_uint2(l0 | (l1 << 16), d0 | (d1 << 16));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->stat_full[st]);
      } else {
        // lse in log2 units, delta: read by the compute warps (broadcast LDS)
        float* stat = sStat + st * 2 * kTile;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          stat[lane + 32 * j] = sl[j] * kBwdLog2e;
          stat[kTile + lane + 32 * j] = sd[j];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->stat_full[st]);
      }
      if (lane == 0) {
        mbar_wait(&bars->do_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->do_full, kTileB);
        for (int half = 0; half < kBoxes; ++half)
          tma_load_4d(sDO + half * kBoxB, &tmDO, &bars->do_full, half * 64, h, q0, b);
      }
      __syncwarp();
    }
  } else if (warp == 12) {
    // ============================================================ MMA issuer
    // (A/B on one box, round 1: a fully unrolled constant-descriptor issue path like the forward's made this
    //  kernel 7 % SLOWER -- 953 vs 1026 TFLOP/s -- so the compact rolled form below is kept.  What bounds the
    //  kernel is the dependency chain between this warp and its consumers)
    reg_dec<64>();
    {
      constexpr uint32_t id_kk = make_idesc(kBF16, 128, 128, false, false);  // A K-major, B K-major
      constexpr uint32_t id_kn = make_idesc(kBF16, 128, kD, false, true);    // A K-major/TMEM, B MN-major (N = d)
      constexpr uint32_t id_nn = make_idesc(kBF16, 128, kD, true, true);     // A MN-major, B MN-major (N = d)
      const uint64_t dK_k = make_smem_desc(smem_u32(sK), 16, 1024);          // K tile as K-major A
      const uint64_t dV_k = make_smem_desc(smem_u32(sV), 16, 1024);
      const uint64_t dK_n = make_smem_desc(smem_u32(sK), kBoxB, 1024);       // K tile as MN-major B
      const uint64_t dDO_k = make_smem_desc(smem_u32(sDO), 16, 1024);
      const uint64_t dDO_n = make_smem_desc(smem_u32(sDO), kBoxB, 1024);
      const uint64_t dDS_k = make_smem_desc(smem_u32(sDS), 16, 1024);        // dS^T [key][q] as K-major A
      const uint64_t dDS_n = make_smem_desc(smem_u32(sDS), kBoxB, 1024);     // ... as MN-major A (M = q)

      // row-stat K steps (kFold): no-swizzle K-major tiles, see the comment at make_smem_desc_noswz
      const uint32_t sFA = smem_u32(smem + kOffFoldA);
      // pattern i at sFA + 128 i (lse/even, delta/even, lse/odd, delta/odd), second K chunk = the zero matrix at + 512
      auto dA_pat = [&](int i) -> uint64_t { return make_smem_desc_noswz(sFA + 128 * i, (4 - i) * 128, 0); };
      const uint64_t dB_stat = make_smem_desc_noswz(smem_u32(smem + kOffStat), 0, 128);

      auto kstep_k = [](int kk) -> uint32_t { return (kk >> 2) * kBoxB + (kk & 3) * 32; };  // K-major k-step
      auto kstep_n = [](int kk) -> uint32_t { return kk * 16 * 128; };                      // MN-major k-step

      auto issue_S = [&](int st, int it_of_block) {  // S^T = K Q^T  (kFold: ... - lse/scale)
        const uint64_t dQ_k = make_smem_desc(smem_u32(sQ + st * kTileB), 16, 1024);
#pragma unroll
        for (int kk = 0; kk < kKSteps; ++kk)
          umma_ss(tS, desc_advance(dK_k, kstep_k(kk)), desc_advance(dQ_k, kstep_k(kk)), id_kk, kk > 0);
        if constexpr (kFold) {
          mbar_wait(&bars->stat_full[it_of_block & 1], (it_of_block >> 1) & 1);
          tc_fence_after();
          umma_ss(tS, dA_pat(2 * (it_of_block & 1)), dB_stat, id_kk, 1);
        }
      };
      auto issue_dP = [&](int it_of_block) {  // dP^T = V dO^T  (kFold: ... - delta; S^T of the block was issued before)
#pragma unroll
        for (int kk = 0; kk < kKSteps; ++kk)
          umma_ss(tDP, desc_advance(dV_k, kstep_k(kk)), desc_advance(dDO_k, kstep_k(kk)), id_kk, kk > 0);
        if constexpr (kFold) umma_ss(tDP, dA_pat(2 * (it_of_block & 1) + 1), dB_stat, id_kk, 1);
      };
      auto issue_dV = [&](bool acc) {  // dV += P^T dO ; P^T cols: q 0..63 at [0,32), q 64..127 at [64,96)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ts(tDV, tS + (kk >> 2) * 64 + (kk & 3) * 8, desc_advance(dDO_n, kstep_n(kk)), id_kn, acc || kk > 0);
      };
      // dK += dS^T Q with dS^T (16-bit) read from TMEM, where the compute warps left it over the
      // dP^T columns they had just consumed (q 0..63 at [128,160), q 64..127 at [192,224)).  TS form: a K = 16
      // step with A in TMEM costs N/2 = 64 clk, with A in shared memory max(74, N/2) (tools/ubench.py).  The dQ MMA
      // issued right behind overwrites these columns; the tensor pipe executes in order, so that is safe.
      auto issue_dK = [&](int st, bool acc) {
        const uint64_t dQ_n = make_smem_desc(smem_u32(sQ + st * kTileB), kBoxB, 1024);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ts(tDK, tDP + (kk >> 2) * 64 + (kk & 3) * 8, desc_advance(dQ_n, kstep_n(kk)), id_kn, acc || kk > 0);
      };
      auto issue_dQ = [&]() {  // dQ = dS K
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ss(tDP, desc_advance(dDS_n, kstep_n(kk)), desc_advance(dK_n, kstep_n(kk)), id_nn, kk > 0);
      };

      mbar_wait(&bars->kv_full, 0);
      mbar_wait(&bars->q_full[0], 0);
      tc_fence_after();
      issue_S(0, 0);
      umma_commit(&bars->s_full);
      // Issue order per Q block: dV, S^T(next), dK, dQ, dP^T(next).  dP^T(next) reuses the TMEM columns of dQ and
      // waits for the drain warps.  (Measured alternative, same box: dV, dP^T, S^T(next), dK, dQ -- which runs dV under
      // that wait but delivers dP^T later to the compute warps -- is 5 % slower: 1027 vs 1082 TFLOP/s.)
      mbar_wait(&bars->do_full, 0);
      tc_fence_after();
      issue_dP(0);
      umma_commit(&bars->dp_full);
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        const bool have_next = it + 1 < n_it;
        mbar_wait(&bars->p_ready, it & 1);
        tc_fence_after();
        issue_dV(it > 0);
        umma_commit(&bars->do_empty);
        if (have_next) {
          mbar_wait(&bars->q_full[st ^ 1], ((it + 1) >> 1) & 1);
          tc_fence_after();
          issue_S(st ^ 1, it + 1);
          umma_commit(&bars->s_full);
        }
        mbar_wait(&bars->ds_ready, it & 1);
        tc_fence_after();
        issue_dK(st, it > 0);
        umma_commit(&bars->q_empty[st]);
        issue_dQ();
        umma_commit(&bars->dq_full);
        if (have_next) {
          mbar_wait(&bars->do_full, (it + 1) & 1);
          mbar_wait(&bars->dq_free, it & 1);
          tc_fence_after();
          issue_dP(it + 1);
          umma_commit(&bars->dp_full);
        }
      }
      umma_commit(&bars->dkv_full);
    }
  } else if (warp >= 12) {
    reg_dec<64>();  // idle register donors (warps 14, 15)
  } else if (warp >= 8) {
    // ============================================================ dQ reduce warps (thread = q row)
    reg_inc<160>();
    const int t = threadIdx.x - 256;  // 0..127 == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const bool issuer = (t == 0);
    for (int it = 0; it < n_it; ++it) {
      const int q0 = (i_begin + it) * kTile;
      mbar_wait(&bars->dq_full, it & 1);
      tc_fence_after();
      // whole dQ row -> registers, then hand the TMEM region straight back to the MMA warp
      // (the next dP^T is waiting for it); staging to smem / TMA happens from registers
      uint32_t v[kD];
#pragma unroll
      for (int c = 0; c < kChunks; ++c) tmem_ld_x32(tDP + lane_base + c * 32, v + c * 32);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->dq_free);
      // deterministic mode: the fp32 adds into dq_acc[q block] happen in key-block order.  Key blocks that
      // see a given Q block are 0..x_max and lower key blocks are tickets of CTAs that started earlier
      // (see the top of the kernel), so waiting for our turn cannot deadlock.
      int* turn = p.sem ? p.sem + ((int64_t)b * p.H + h) * nQ + (i_begin + it) : nullptr;
      if (turn && issuer) {
        while (ld_acquire_gpu(turn) != kb) __nanosleep(64);
      }
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        uint8_t* stage = sDQ + (c & 1) * kDqStageB;
        if (issuer) tma_store_wait_read<1>();  // the reduce that last read this staging box has finished
        named_bar_sync(1, 128);
        // row t: 8 x 16-byte chunks, chunk j stored at position j ^ (t % 8)  (SWIZZLE_128B)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o;
          o.x = __uint_as_float(v[c * 32 + j * 4 + 0]) * p.scale;
          o.y = __uint_as_float(v[c * 32 + j * 4 + 1]) * p.scale;
          o.z = __uint_as_float(v[c * 32 + j * 4 + 2]) * p.scale;
          o.w = __uint_as_float(v[c * 32 + j * 4 + 3]) * p.scale;
          *reinterpret_cast<float4*>(stage + t * 128 + ((j ^ (t & 7)) << 4)) = o;
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (issuer) {
          tma_reduce_add_4d(&tmDQ, stage, c * 32, h, q0, b);
          tma_store_commit();
        }
      }
      if (turn && issuer) {
        tma_store_wait<0>();  // our four reductions have been performed ...
        __threadfence();
        st_release_gpu(turn, kb + 1);  // ... next key block's turn
      }
    }
    if (issuer) tma_store_wait<0>();
  } else {
    // ============================================================ compute warps (thread = key row, half the q cols)
    reg_inc<144>();
    const int r = (warp & 3) * 32 + lane;  // key row within the block == TMEM lane
    const int hf = warp >> 2;              // which 64 query columns
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int key = k0 + r;
    const bool key_valid = key < p.Sk;
    const float scale_log2 = p.scale_log2;
    // additive bias of this thread's key, in log2 units (scores = q k^T scale + bias[key]; reference lao.py:155-173,
    // "vector" bias): a per-thread scalar in this key-row layout, folded into the exponent's FMA
    const float bias2 = (p.bias && key_valid)
                            ? __ldg(p.bias + (int64_t)b * p.bias_sb + (int64_t)h * p.bias_sh + key) * kBwdLog2e
                            : 0.f;
    uint8_t* ds_row = sDS + hf * kBoxB + r * 128;
    for (int it = 0; it < n_it; ++it) {
      const int q0 = (i_begin + it) * kTile;
      const int st = it & 1;
      const float* stat = sStat + st * 2 * kTile + hf * 64;
      if constexpr (!kFold) mbar_wait(&bars->stat_full[st], (it >> 1) & 1);
      mbar_wait(&bars->s_full, it & 1);
      tc_fence_after();
      float pr[64];
      {
        uint32_t* sr = reinterpret_cast<uint32_t*>(pr);
        tmem_ld_x32(tS + lane_base + hf * 64, sr);
        tmem_ld_x32(tS + lane_base + hf * 64 + 32, sr + 32);
        tmem_wait_ld();
      }
      // visible iff key <= q + off  <=>  q >= key - off ; whole block visible when q0 + off >= k0 + 127
      const bool need_mask = p.causal && (q0 + p.causal_off < k0 + kTile - 1);
      const int qmin = key - p.causal_off - q0 - hf * 64;  // first visible column index (local to this half)
      if constexpr (kFold) {  // the tensor core already subtracted lse/scale per column
#pragma unroll
        for (int c = 0; c < 64; ++c) pr[c] = ex2(fmaf(pr[c], scale_log2, bias2));
      } else {
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          const float4 l2 = *reinterpret_cast<const float4*>(stat + c4 * 4);
          pr[c4 * 4 + 0] = ex2(fmaf(pr[c4 * 4 + 0], scale_log2, bias2 - l2.x));
          pr[c4 * 4 + 1] = ex2(fmaf(pr[c4 * 4 + 1], scale_log2, bias2 - l2.y));
          pr[c4 * 4 + 2] = ex2(fmaf(pr[c4 * 4 + 2], scale_log2, bias2 - l2.z));
          pr[c4 * 4 + 3] = ex2(fmaf(pr[c4 * 4 + 3], scale_log2, bias2 - l2.w));
        }
      }
      if (!key_valid) {
#pragma unroll
        for (int c = 0; c < 64; ++c) pr[c] = 0.f;
      } else if (need_mask) {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c < qmin) pr[c] = 0.f;
      }
      {
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) pk[c] = pack2<kBF16>(pr[2 * c], pr[2 * c + 1]);
        tmem_st_x32(tS + lane_base + hf * 64, pk);  // P^T for this half aliases its own S^T columns
        tmem_wait_st();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_ready);

      // dS^T = P^T o (dP^T - delta)
      mbar_wait(&bars->dp_full, it & 1);
      tc_fence_after();
      if (it > 0) mbar_wait(&bars->dq_full, (it - 1) & 1);  // dK(it-1), dQ(it-1) finished reading sDS
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t dp[32];
        tmem_ld_x32(tDP + lane_base + hf * 64 + half * 32, dp);
        tmem_wait_ld();
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          float4 dl = make_float4(0.f, 0.f, 0.f, 0.f);  // kFold: the tensor core already subtracted delta
          if constexpr (!kFold) dl = *reinterpret_cast<const float4*>(stat + kTile + half * 32 + c4 * 4);
          const int c = half * 32 + c4 * 4;
          const float d0 = pr[c + 0] * (__uint_as_float(dp[c4 * 4 + 0]) - dl.x);
          const float d1 = pr[c + 1] * (__uint_as_float(dp[c4 * 4 + 1]) - dl.y);
          const float d2 = pr[c + 2] * (__uint_as_float(dp[c4 * 4 + 2]) - dl.z);
          const float d3 = pr[c + 3] * (__uint_as_float(dp[c4 * 4 + 3]) - dl.w);
          dp[c4 * 2 + 0] = pack2<kBF16>(d0, d1);  // in-place: slot 2*c4 <= 4*c4 already consumed
          dp[c4 * 2 + 1] = pack2<kBF16>(d2, d3);
        }
        // 32 columns = 64 bytes = 4 x 16-byte chunks (chunk index within the 128-byte row: half*4 + j)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o = make_uint4(dp[j * 4 + 0], dp[j * 4 + 1], dp[j * 4 + 2], dp[j * 4 + 3]);
          *reinterpret_cast<uint4*>(ds_row + (((half * 4 + j) ^ (r & 7)) << 4)) = o;
        }
        // ... and the same 16 packed registers into TMEM (A operand of the dK MMA), over dP^T columns
        // this thread has already read
        tmem_st_x16(tDP + lane_base + hf * 64 + half * 16, dp);
      }
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->ds_ready);
    }

    // ---------------------------------------------------------- epilogue: dk_acc += scale*dK, dv_acc += dV
    mbar_wait(&bars->dkv_full, 0);
    tc_fence_after();
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      float* base = which == 0 ? p.dk_acc : p.dv_acc;
      const int64_t sb = which == 0 ? p.dk_sb : p.dv_sb, ss = which == 0 ? p.dk_ss : p.dv_ss,
                    sh = which == 0 ? p.dk_sh : p.dv_sh;
      const float mul = which == 0 ? p.scale : 1.f;
      float* dst = base + (int64_t)b * sb + (int64_t)key * ss + (int64_t)h * sh + hf * (kD / 2);
#pragma unroll
      for (int c = 0; c < kD / 64; ++c) {
        uint32_t v[32];
        tmem_ld_x32((which == 0 ? tDK : tDV) + lane_base + hf * (kD / 2) + c * 32, v);
        tmem_wait_ld();
        if (key_valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4* ptr = reinterpret_cast<float4*>(dst + c * 32 + j * 4);
            float4 a = *ptr;
            a.x = fmaf(__uint_as_float(v[j * 4 + 0]), mul, a.x);
            a.y = fmaf(__uint_as_float(v[j * 4 + 1]), mul, a.y);
            a.z = fmaf(__uint_as_float(v[j * 4 + 2]), mul, a.z);
            a.w = fmaf(__uint_as_float(v[j * 4 + 3]), mul, a.w);
            *ptr = a;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, 512);
}

template <bool kBF16, bool kFold, int kD>
static int launch_bwd(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                      const CUtensorMap& tmDO, const CUtensorMap& tmDQ, const BwdParams& p, cudaStream_t stream) {
  auto kern = bwd_chunk_kernel<kBF16, kFold, kD>;
  constexpr int smem = BwdLayout<kD>::kSmemBytes;
  BA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((p.Sk + kTile - 1) / kTile, p.H, p.B);
  kern<<<grid, kBwdThreads, smem, stream>>>(tmQ, tmK, tmV, tmDO, tmDQ, p);
  BA_CHECK_CUDA(cudaGetLastError());
  return BA_OK;
}

template <bool kFold, int kD>
static int launch_bwd_dt(int dtype, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                         const CUtensorMap& tmDO, const CUtensorMap& tmDQ, const BwdParams& p, cudaStream_t stream) {
  return dtype == BA_DTYPE_BF16 ? launch_bwd<true, kFold, kD>(tmQ, tmK, tmV, tmDO, tmDQ, p, stream)
                                : launch_bwd<false, kFold, kD>(tmQ, tmK, tmV, tmDO, tmDQ, p, stream);
}

// Deterministic-mode workspace (turn counters + tickets), one per (device, stream), grown on demand, zeroed on
// the launching stream before every launch.  Launches on one stream are ordered, so they can share a
// workspace; different streams / devices / host threads never do (a mutex guards the table).
struct BwdWorkspace {
  int device;
  cudaStream_t stream;
  int* ptr;
  size_t cap;
};
static std::mutex g_ws_mutex;
static std::vector<BwdWorkspace> g_ws;

static int* bwd_sem_workspace(size_t n_ints, cudaStream_t stream) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  BwdWorkspace* w = nullptr;
  for (auto& e : g_ws)
    if (e.device == dev && e.stream == stream) w = &e;
  if (!w) {
    g_ws.push_back(BwdWorkspace{dev, stream, nullptr, 0});
    w = &g_ws.back();
  }
  if (w->cap < n_ints) {
    // the old buffer may still be in use by a launch in flight on this stream: free it in stream order
    if (w->ptr && cudaFreeAsync(w->ptr, stream) != cudaSuccess) return nullptr;
    w->ptr = nullptr, w->cap = 0;
    if (cudaMallocAsync(reinterpret_cast<void**>(&w->ptr), n_ints * sizeof(int), stream) != cudaSuccess) return nullptr;
    w->cap = n_ints;
  }
  if (cudaMemsetAsync(w->ptr, 0, n_ints * sizeof(int), stream) != cudaSuccess) return nullptr;
  return w->ptr;
}

static bool f32_view_ok(const ba_tensor4& t) {
  return t.ptr && (reinterpret_cast<uintptr_t>(t.ptr) & 15) == 0 && t.stride_b % 4 == 0 && t.stride_s % 4 == 0 &&
         t.stride_h % 4 == 0;
}

}  // namespace ba

extern "C" int ba_bwd_chunk(ba_tensor4 d_o, ba_tensor4 q, ba_tensor4 k, ba_tensor4 v, ba_rowstat delta,
                            ba_rowstat lse, ba_tensor4 dq_acc, ba_tensor4 dk_acc, ba_tensor4 dv_acc, int B, int Sq,
                            int Sk, int H, int D, float scale, int mask_mode, int causal_offset, int flags, int dtype,
                            void* stream) {
  ba_rowstat none = {nullptr, 0, 0};
  return ba_bwd_chunk_bias(d_o, q, k, v, delta, lse, none, dq_acc, dk_acc, dv_acc, B, Sq, Sk, H, D, scale, mask_mode,
                           causal_offset, flags, dtype, stream);
}

extern "C" int ba_bwd_chunk_bias(ba_tensor4 d_o, ba_tensor4 q, ba_tensor4 k, ba_tensor4 v, ba_rowstat delta,
                                 ba_rowstat lse, ba_rowstat key_bias, ba_tensor4 dq_acc, ba_tensor4 dk_acc,
                                 ba_tensor4 dv_acc, int B, int Sq, int Sk, int H, int D, float scale, int mask_mode,
                                 int causal_offset, int flags, int dtype, void* stream) {
  using namespace ba;
  BA_REQUIRE(D == 128 || D == 64, "ba_bwd_chunk: head dim %d unsupported (64 or 128)", D);
  BA_REQUIRE(B > 0 && Sq > 0 && Sk > 0 && H > 0, "ba_bwd_chunk: empty problem B=%d Sq=%d Sk=%d H=%d", B, Sq, Sk, H);
  BA_REQUIRE(dtype == BA_DTYPE_FP16 || dtype == BA_DTYPE_BF16, "ba_bwd_chunk: bad dtype %d", dtype);
  BA_REQUIRE(mask_mode == BA_MASK_NONE || mask_mode == BA_MASK_CAUSAL, "ba_bwd_chunk: bad mask mode %d", mask_mode);
  BA_REQUIRE(scale > 0.f && isfinite(scale), "ba_bwd_chunk: softmax scale must be positive and finite");
  BA_REQUIRE(d_o.ptr && q.ptr && k.ptr && v.ptr && delta.ptr && lse.ptr, "ba_bwd_chunk: null input");
  BA_REQUIRE(f32_view_ok(dq_acc) && f32_view_ok(dk_acc) && f32_view_ok(dv_acc),
             "ba_bwd_chunk: fp32 accumulators must be non-null, 16-byte aligned, strides multiple of 4");
  BA_REQUIRE(H <= 65535 && B <= 65535, "ba_bwd_chunk: H and B must be <= 65535");

  CUtensorMap tmQ, tmK, tmV, tmDO, tmDQ;
  const CUtensorMapDataType dt = lowp_dtype(dtype);
  int rc;
  if ((rc = make_tensor_map(&tmQ, q, B, Sq, H, D, dt, 2, 64, kTile, true))) return rc;
  if ((rc = make_tensor_map(&tmDO, d_o, B, Sq, H, D, dt, 2, 64, kTile, true))) return rc;
  if ((rc = make_tensor_map(&tmK, k, B, Sk, H, D, dt, 2, 64, kTile, true))) return rc;
  if ((rc = make_tensor_map(&tmV, v, B, Sk, H, D, dt, 2, 64, kTile, true))) return rc;
  if ((rc = make_tensor_map(&tmDQ, dq_acc, B, Sq, H, D, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 32, kTile, true)))
    return rc;

  BwdParams p;
  p.lse = lse.ptr, p.lse_sb = lse.stride_b, p.lse_sh = lse.stride_h;
  p.delta = delta.ptr, p.dl_sb = delta.stride_b, p.dl_sh = delta.stride_h;
  p.dk_acc = static_cast<float*>(dk_acc.ptr);
  p.dk_sb = dk_acc.stride_b, p.dk_ss = dk_acc.stride_s, p.dk_sh = dk_acc.stride_h;
  p.dv_acc = static_cast<float*>(dv_acc.ptr);
  p.dv_sb = dv_acc.stride_b, p.dv_ss = dv_acc.stride_s, p.dv_sh = dv_acc.stride_h;
  p.bias = key_bias.ptr, p.bias_sb = key_bias.stride_b, p.bias_sh = key_bias.stride_h;
  p.B = B, p.Sq = Sq, p.Sk = Sk, p.H = H;
  p.scale = scale;
  p.scale_log2 = scale * kBwdLog2e;
  p.inv_scale = 1.f / scale;
  p.causal = mask_mode == BA_MASK_CAUSAL;
  p.causal_off = causal_offset;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p.sem = p.ticket = nullptr;
  if (flags & BA_BWD_DETERMINISTIC) {
    const size_t n_turn = (size_t)B * H * ((Sq + kTile - 1) / kTile);
    const size_t n = n_turn + (size_t)B * H;
    p.sem = bwd_sem_workspace(n, st);
    if (p.sem) p.ticket = p.sem + n_turn;
    if (!p.sem) {
      set_error("ba_bwd_chunk: could not allocate the deterministic-mode workspace (%zu ints)", n);
      return BA_ERR_CUDA;
    }
  }
  // BA_BWD_FOLD=0 (A/B knob, read once): row statistics through shared memory instead of the extra K steps
  static const bool fold = [] {
    const char* e = getenv("BA_BWD_FOLD");
    return !e || atoi(e) != 0;
  }();
  if (D == 64)
    return fold ? launch_bwd_dt<true, 64>(dtype, tmQ, tmK, tmV, tmDO, tmDQ, p, st)
                : launch_bwd_dt<false, 64>(dtype, tmQ, tmK, tmV, tmDO, tmDQ, p, st);
  return fold ? launch_bwd_dt<true, 128>(dtype, tmQ, tmK, tmV, tmDO, tmDQ, p, st)
              : launch_bwd_dt<false, 128>(dtype, tmQ, tmK, tmV, tmDO, tmDQ, p, st);
}
