// Self tests of the sm_100a building blocks used by the attention kernels
// (called from tests/ only, through ba_selftest): they isolate the TMA box /
// swizzle layout, the K-major and MN-major shared-memory descriptors, the
// instruction descriptor, the TMEM load/store lane mapping and the TS-form MMA,
// so a failing attention parity test can be traced to one assumption.
#include "burst_attn_b200_selftest.h"
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace ba {

constexpr int kStTile = 128 * 128 * 2;  // 32 KiB
constexpr int kStBox = kStTile / 2;
constexpr int kStSmem = 2 * kStTile + 1024 + 64;

// mode 0: out = A * B^T (SS, both K-major)      mode 1: out = A(as P in TMEM) * B (TS, B MN-major)
// mode 3: out = A^T * B (SS, A MN-major [k][m], B MN-major [k][n])   -- used by the backward
// mode 2: raw dump of the first TMA box of A
template <bool kBF16>
__global__ void __launch_bounds__(128, 1)
selftest_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const uint16_t* __restrict__ a_raw, void* __restrict__ out, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStTile;
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sB + kStTile);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar_load + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bar_load, 1);
      mbar_init(bar_mma, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_holder, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;

  if (t == 0) {
    mbar_arrive_expect_tx(bar_load, 2 * kStTile);
    for (int half = 0; half < 2; ++half) {
      tma_load_4d(sA + half * kStBox, &tmA, bar_load, half * 64, 0, 0, 0);
      tma_load_4d(sB + half * kStBox, &tmB, bar_load, half * 64, 0, 0, 0);
    }
  }
  mbar_wait(bar_load, 0);

  if (mode == 2) {
    const uint16_t* s16 = reinterpret_cast<const uint16_t*>(sA);
    uint16_t* o16 = static_cast<uint16_t*>(out);
    for (int i = t; i < kStBox / 2; i += 128) o16[i] = s16[i];
  } else {
    if (mode == 1) {
      // thread t stages row t of P into TMEM cols [128, 192) as packed 16-bit pairs
      uint32_t v[64];
      const uint32_t* src = reinterpret_cast<const uint32_t*>(a_raw + t * 128);
#pragma unroll
      for (int c = 0; c < 64; ++c) v[c] = src[c];
      tmem_st_x32(tmem_base + lane_base + 128, v);
      tmem_st_x32(tmem_base + lane_base + 128 + 32, v + 32);
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      if (mode == 0) {
        constexpr uint32_t idesc = make_idesc(kBF16, 128, 128, false, false);
        const uint64_t a0 = make_smem_desc(smem_u32(sA), 16, 1024);
        const uint64_t b0 = make_smem_desc(smem_u32(sB), 16, 1024);
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * kStBox + (kk & 3) * 32;
          umma_ss(tmem_base, desc_advance(a0, off), desc_advance(b0, off), idesc, kk > 0);
        }
      } else if (mode == 1) {
        constexpr uint32_t idesc = make_idesc(kBF16, 128, 128, false, true);
        const uint64_t b0 = make_smem_desc(smem_u32(sB), kStBox, 1024);
        for (int kk = 0; kk < 8; ++kk)
          umma_ts(tmem_base, tmem_base + 128 + kk * 8, desc_advance(b0, kk * 2048), idesc, kk > 0);
      } else {  // mode 3: both operands MN-major: A stored [k][m], B stored [k][n]
        constexpr uint32_t idesc = make_idesc(kBF16, 128, 128, true, true);
        const uint64_t a0 = make_smem_desc(smem_u32(sA), kStBox, 1024);
        const uint64_t b0 = make_smem_desc(smem_u32(sB), kStBox, 1024);
        for (int kk = 0; kk < 8; ++kk)
          umma_ss(tmem_base, desc_advance(a0, kk * 2048), desc_advance(b0, kk * 2048), idesc, kk > 0);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, 0);
    tc_fence_after();
    float* o = static_cast<float*>(out) + t * 128;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld_x32(tmem_base + lane_base + c * 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) o[c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}


// ---- CTA-pair self tests (cluster of 2, 128 threads per CTA) ---------------------------------------
// mode 4: out[256,128] = A[256,128] * B[128,128]^T   (SS, cta_group::2, M = 256; CTA r holds A rows
//         [128r,128r+128) and B rows [64r,64r+64), i.e. half of N)
// mode 5: out[256,128] = A[256,128] * B[128,128]     (TS, A staged in each CTA's TMEM; B = [k][n] read
//         MN-major, CTA r holds n columns [64r, 64r+64))
// mode 6: probe of the M = 128 cta_group::2 accumulator layout (needed by the pair backward's dQ): a is
//         [128,128]; CTA r holds A rows [64r,64r+64) and B rows [64r,64r+64); every TMEM cell of columns
//         [0,128) is first set to the sentinel 12345.0, then out[r][lane][col] = raw dump of CTA r's TMEM.
// mode 7: mode 4 (M = 256 SS) with the B halves delivered by the PEER's threads through DSMEM stores
//         (st.shared::cluster + fence.proxy.async + cluster barrier) instead of TMA: the hand-off the pair
//         backward uses for dS.
constexpr int kSt2Smem = kStTile + kStBox + 1024 + 64;

template <bool kBF16>
__global__ void __launch_bounds__(128, 1)
selftest2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                 const uint16_t* __restrict__ a_raw, const uint16_t* __restrict__ b_raw, float* __restrict__ out,
                 int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw2[];
  uint8_t* smem = smem_raw2;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sA = smem;             // 32 KiB: this CTA's 128 rows of A (mode 4)
  uint8_t* sB = smem + kStTile;   // 16 KiB: this CTA's half of B
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sB + kStBox);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar_load + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bar_load, 1);
      mbar_init(bar_mma, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2cta(tmem_holder, 256);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote complete_tx / arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;

  if (t == 0) {
    const uint32_t bytes_per_cta = mode == 4 ? kStTile + kStBox : (mode == 7 ? kStTile : (mode == 6 ? 2 * kStBox : kStBox));
    if (rank == 0) mbar_arrive_expect_tx(bar_load, 2 * bytes_per_cta);
    if (mode == 4 || mode == 7) {
      for (int half = 0; half < 2; ++half)
        tma_load_4d_2cta(sA + half * kStBox, &tmA, bar_load, half * 64, 0, (int)rank * 128, 0);
      if (mode == 4)
        for (int half = 0; half < 2; ++half)  // B rows [64r, 64r+64): two 64x64 boxes of 8 KiB
          tma_load_4d_2cta(sB + half * (kStBox / 2), &tmBh, bar_load, half * 64, 0, (int)rank * 64, 0);
    } else if (mode == 6) {
      for (int half = 0; half < 2; ++half) {  // A rows and B rows [64r, 64r+64): 64x64 boxes
        tma_load_4d_2cta(sA + half * (kStBox / 2), &tmA, bar_load, half * 64, 0, (int)rank * 64, 0);
        tma_load_4d_2cta(sB + half * (kStBox / 2), &tmBh, bar_load, half * 64, 0, (int)rank * 64, 0);
      }
    } else {
      // B[k][n]: this CTA's n columns [64r, 64r+64), all 128 k rows: one 128x64 box of 16 KiB
      tma_load_4d_2cta(sB, &tmBh, bar_load, (int)rank * 64, 0, 0, 0);
    }
  }
  if (mode == 5) {
    // stage this CTA's 128 rows of A into its TMEM cols [128,192) (row t <-> lane t)
    uint32_t v[64];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a_raw + ((int)rank * 128 + t) * 128);
#pragma unroll
    for (int c = 0; c < 64; ++c) v[c] = src[c];
    tmem_st_x32(tmem_base + lane_base + 128, v);
    tmem_st_x32(tmem_base + lane_base + 128 + 32, v + 32);
    tmem_wait_st();
  }
  if (mode == 6) {  // sentinel in every cell the MMA could own
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(12345.0f);
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_st_x32(tmem_base + lane_base + c * 32, v);
    tmem_wait_st();
  }
  if (mode == 7) {
    // thread t writes row (t % 64), 64-column box (t / 64) of the PEER's B half into the peer's sB, in the
    // SWIZZLE_128B K-major layout a TMA box load would have produced (16-byte chunk j of row r at j ^ (r % 8))
    const uint32_t peer = rank ^ 1u;
    const int row = t & 63, box = t >> 6;
    const uint4* src = reinterpret_cast<const uint4*>(b_raw + ((int)peer * 64 + row) * 128 + box * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      st_shared_remote_v4(sB + box * (kStBox / 2) + row * 128 + ((j ^ (row & 7)) << 4), peer, src[j]);
    fence_proxy_async_all();
  }
  tc_fence_before();
  cluster_sync_all();  // A rows of BOTH CTAs are in TMEM (mode 5) / the DSMEM writes have landed (mode 7)
  tc_fence_after();

  if (rank == 0 && warp == 0) {
    mbar_wait(bar_load, 0);
    tc_fence_after();
    if (mode == 4 || mode == 7) {
      constexpr uint32_t idesc = make_idesc(kBF16, 256, 128, false, false);
      const uint64_t a0 = make_smem_desc(smem_u32(sA), 16, 1024);
      const uint64_t b0 = make_smem_desc(smem_u32(sB), 16, 1024);
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t offA = (kk >> 2) * kStBox + (kk & 3) * 32;
        const uint32_t offB = (kk >> 2) * (kStBox / 2) + (kk & 3) * 32;
        umma_ss_2cta(tmem_base, desc_advance(a0, offA), desc_advance(b0, offB), idesc, kk > 0);
      }
    } else if (mode == 6) {
      constexpr uint32_t idesc = make_idesc(kBF16, 128, 128, false, false);  // M = 128 over the pair: 64 rows per CTA
      const uint64_t a0 = make_smem_desc(smem_u32(sA), 16, 1024);
      const uint64_t b0 = make_smem_desc(smem_u32(sB), 16, 1024);
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = (kk >> 2) * (kStBox / 2) + (kk & 3) * 32;
        umma_ss_2cta(tmem_base, desc_advance(a0, off), desc_advance(b0, off), idesc, kk > 0);
      }
    } else {
      constexpr uint32_t idesc = make_idesc(kBF16, 256, 128, false, true);
      const uint64_t b0 = make_smem_desc(smem_u32(sB), kStBox, 1024);
      for (int kk = 0; kk < 8; ++kk)
        umma_ts_2cta(tmem_base, tmem_base + 128 + kk * 8, desc_advance(b0, kk * 2048), idesc, kk > 0);
    }
    umma_commit_2cta(bar_mma, 0x3);
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  float* o = out + ((int)rank * 128 + t) * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld_x32(tmem_base + lane_base + c * 32, v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) o[c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_2cta(tmem_base, 256);
}

}  // namespace ba

extern "C" int ba_selftest(int mode, const void* a, const void* b, void* out, int dtype, void* stream) {
  using namespace ba;
  BA_REQUIRE(mode >= 0 && mode <= 7, "ba_selftest: bad mode %d", mode);
  BA_REQUIRE(a && b && out, "ba_selftest: null pointer");
  if (mode >= 4) {  // CTA-pair tests: a is [256,128] ([128,128] in mode 6), b [128,128], out fp32 [256,128]
    const int a_rows = mode == 6 ? 128 : 256;
    ba_tensor4 ta{const_cast<void*>(a), (int64_t)a_rows * 128, 128, 128};
    ba_tensor4 tb{const_cast<void*>(b), 128 * 128, 128, 128};
    CUtensorMap tmA, tmBh;
    int rc;
    if ((rc = make_tensor_map(&tmA, ta, 1, a_rows, 1, 128, lowp_dtype(dtype), 2, 64, mode == 6 ? 64 : 128, true)))
      return rc;
    // modes 4, 6, 7: boxes of 64 rows x 64 cols (half of N); mode 5: 128 rows (k) x 64 cols (half of n)
    if ((rc = make_tensor_map(&tmBh, tb, 1, 128, 1, 128, lowp_dtype(dtype), 2, 64, mode == 5 ? 128 : 64, true)))
      return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = kSt2Smem;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const uint16_t* a16 = static_cast<const uint16_t*>(a);
    const uint16_t* b16 = static_cast<const uint16_t*>(b);
    float* o32 = static_cast<float*>(out);
    if (dtype == BA_DTYPE_BF16) {
      BA_CHECK_CUDA(cudaFuncSetAttribute(selftest2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSt2Smem));
      BA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, selftest2_kernel<true>, tmA, tmBh, a16, b16, o32, mode));
    } else {
      BA_CHECK_CUDA(cudaFuncSetAttribute(selftest2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSt2Smem));
      BA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, selftest2_kernel<false>, tmA, tmBh, a16, b16, o32, mode));
    }
    return BA_OK;
  }
  ba_tensor4 ta{const_cast<void*>(a), 128 * 128, 128, 128};
  ba_tensor4 tb{const_cast<void*>(b), 128 * 128, 128, 128};
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tensor_map(&tmA, ta, 1, 128, 1, 128, lowp_dtype(dtype), 2, 64, 128, true))) return rc;
  if ((rc = make_tensor_map(&tmB, tb, 1, 128, 1, 128, lowp_dtype(dtype), 2, 64, 128, true))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == BA_DTYPE_BF16) {
    BA_CHECK_CUDA(cudaFuncSetAttribute(selftest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStSmem));
    selftest_kernel<true><<<1, 128, kStSmem, st>>>(tmA, tmB, static_cast<const uint16_t*>(a), out, mode);
  } else {
    BA_CHECK_CUDA(cudaFuncSetAttribute(selftest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStSmem));
    selftest_kernel<false><<<1, 128, kStSmem, st>>>(tmA, tmB, static_cast<const uint16_t*>(a), out, mode);
  }
  BA_CHECK_CUDA(cudaGetLastError());
  return BA_OK;
}
