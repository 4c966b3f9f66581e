// Micro-benchmarks of the sm_100a building blocks (tools/ubench.py only, through ba_ubench): the numbers
// the kernel designs in DESIGN.md hinge on -- tcgen05.mma dispatch rates by operand source and shape
// (SS / TS, N = 64 / 128 / 256, cta_group 1 / 2), tcgen05.ld / st throughput and whether TMEM reads slow a
// concurrent MMA chain, MUFU ex2 throughput, commit -> mbarrier and cluster-remote arrive latencies.
// Operand CONTENTS are whatever shared memory / TMEM hold (timing only); every address is in bounds.
#include "burst_attn_b200_selftest.h"
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace ba {

constexpr int kUbThreads = 288;                 // warps 0-7 workers, warp 8 MMA issuer
constexpr uint32_t kUbOffA = 0;                 // 32 KiB: A tile [128 x 128] 16-bit, two SW128 boxes
constexpr uint32_t kUbOffB = 32768;             // 64 KiB: B tile up to [256 x 128]
constexpr uint32_t kUbOffBar = 98304;
constexpr int kUbSmem = kUbOffBar + 256;

struct UbBars {
  uint64_t mma_done, go, back;
  uint32_t tmem_base;
};

BA_DEVICE long long clk() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}

// ---- single-CTA modes
// 0 SS N=128 | 1 TS N=128 | 2 SS N=256 | 3 TS N=256 | 4 SS N=64 | 8 LDTM 4 warps | 9 LDTM 8 warps | 16 LDTM 1 warp
// 10 STTM 4 warps | 11 MUFU ex2, 8 warps | 12 one MMA + commit + wait (latency) | 14 TS N=128 chain with 4 warps
// of LDTM running beside it (out[0] = MMA cycles, out[1] = LDTM cycles for the same number of groups)
__global__ void __launch_bounds__(kUbThreads, 1) ubench_kernel(int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t ub_smem[];
  uint8_t* smem = ub_smem;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  UbBars* bars = reinterpret_cast<UbBars*>(smem + kUbOffBar);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (warp == 8) {
    if (lane == 0) {
      mbar_init(&bars->mma_done, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = bars->tmem_base;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  long long cyc = 0;

  const bool mma_mode = mode <= 4 || mode == 12 || mode == 14;
  if (warp == 8 && mma_mode) {
    const bool ts = mode == 1 || mode == 3 || mode == 14;
    const int N = (mode == 2 || mode == 3) ? 256 : (mode == 4 ? 64 : 128);
    const uint32_t idesc = N == 256 ? make_idesc(true, 128, 256, false, false)
                                    : (N == 64 ? make_idesc(true, 128, 64, false, false)
                                               : make_idesc(true, 128, 128, false, false));
    const uint64_t a0 = make_smem_desc(smem_u32(smem + kUbOffA), 16, 1024);
    const uint64_t b0 = make_smem_desc(smem_u32(smem + kUbOffB), 16, 1024);
    const uint32_t box_b = N * 128;  // bytes of one 64-column SW128 box of B (N rows x 128 B)
    const int n = mode == 12 ? 1 : iters;
    // lean issue path (descriptor halves as constants, 8 k-steps unrolled) so that the chain measures the tensor
    // pipe and its operand fetch, not this warp's address arithmetic (the first version of this loop was
    // issue-bound at ~109 clk per SS dispatch and ~80 per TS dispatch whatever N was)
    const uint32_t a_lo0 = (smem_u32(smem + kUbOffA) >> 4) + desc_lo_lbo(16);
    const uint32_t b_lo0 = (smem_u32(smem + kUbOffB) >> 4) + desc_lo_lbo(16);
    constexpr uint32_t hi = desc_hi(1024);
    const long long t0 = clk();
    if (n == 1) {
      umma_ss_lh(tb, a_lo0, hi, b_lo0, hi, idesc, 1);
    } else {
      for (int i = 0; i < n; i += 8) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t offa = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4, offb = ((kk >> 2) * box_b + (kk & 3) * 32) >> 4;
          if (ts)
            umma_ts_lh(tb, tb + 256 + kk * 8, b_lo0 + offb, hi, idesc, 1);
          else
            umma_ss_lh(tb, a_lo0 + offa, hi, b_lo0 + offb, hi, idesc, 1);
        }
      }
    }
    umma_commit(&bars->mma_done);
    mbar_wait(&bars->mma_done, 0);
    cyc = clk() - t0;
    if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)cyc);
  }
  const int ld_warps = (mode == 9) ? 8 : (mode == 16 ? 1 : 4);
  if (warp < ld_warps && (mode == 8 || mode == 9 || mode == 14 || mode == 16)) {
    uint32_t v[32], x = 0;
    const uint32_t col0 = tb + lane_base + 384;  // columns no MMA of this benchmark touches
    const long long t0 = clk();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_x32(col0 + c * 32, v);
        x ^= v[c];
      }
      tmem_wait_ld();
    }
    cyc = clk() - t0;
    if (x == 0x12345678u) out[3] = 1;  // keep the loads observable
    if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(out + (mode == 14 ? 1 : 0)), (unsigned long long)cyc);
  }
  if (warp < 4 && mode == 10) {
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = lane + j;
    const uint32_t col0 = tb + lane_base + 384;
    const long long t0 = clk();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st_x32(col0 + c * 32, v);
      tmem_wait_st();
    }
    cyc = clk() - t0;
    if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)cyc);
  }
  if (warp < 8 && mode == 11) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = -0.001f * (lane + j + 1);
    const long long t0 = clk();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = ex2(a[j]) - 1.5f;  // 8 independent MUFU chains per thread
    }
    cyc = clk() - t0;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 123.f) out[3] = 2;
    if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)cyc);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tb, 512);
}

// ---- CTA-pair modes (cluster of 2): 5 SS M=256 N=128 | 6 TS M=256 N=128 | 7 SS M=256 N=256 (each CTA holds
// half of B's rows) | 13 remote mbarrier arrive round trip (CTA1 -> CTA0 -> CTA1) | 15 CTA1 remote arrive ->
// leader issues one pair MMA -> multicast commit seen by CTA1
__global__ void __launch_bounds__(kUbThreads, 1) ubench2_kernel(int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t ub_smem2[];
  uint8_t* smem = ub_smem2;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  UbBars* bars = reinterpret_cast<UbBars*>(smem + kUbOffBar);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (warp == 8) {
    if (lane == 0) {
      mbar_init(&bars->mma_done, 1);
      mbar_init(&bars->go, 1);
      mbar_init(&bars->back, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2cta(&bars->tmem_base, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = bars->tmem_base;
  if (mode >= 5 && mode <= 7) {
    if (warp == 8 && rank == 0) {
      const bool ts = mode == 6;
      const int N = mode == 7 ? 256 : 128;
      const uint32_t idesc = N == 256 ? make_idesc(true, 256, 256, false, false) : make_idesc(true, 256, 128, false, false);
      const uint64_t a0 = make_smem_desc(smem_u32(smem + kUbOffA), 16, 1024);
      const uint64_t b0 = make_smem_desc(smem_u32(smem + kUbOffB), 16, 1024);
      const uint32_t box_b = (N / 2) * 128;  // this CTA's half of B: N/2 rows x 128 B per box
      const long long t0 = clk();
      for (int i = 0; i < iters; ++i) {
        const int kk = i & 7;
        const uint32_t offa = (kk >> 2) * 16384 + (kk & 3) * 32, offb = (kk >> 2) * box_b + (kk & 3) * 32;
        if (ts)
          umma_ts_2cta(tb, tb + 256 + kk * 8, desc_advance(b0, offb), idesc, 1);
        else
          umma_ss_2cta(tb, desc_advance(a0, offa), desc_advance(b0, offb), idesc, 1);
      }
      umma_commit_2cta(&bars->mma_done, 0x3);
      mbar_wait(&bars->mma_done, 0);
      const long long cyc = clk() - t0;
      if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)cyc);
    } else if (warp == 8) {
      mbar_wait(&bars->mma_done, 0);  // multicast commit: CTA 1 must not leave before the MMAs retire
    }
  } else if (mode == 13 || mode == 15) {
    if (warp == 8) {
      long long total = 0;
      for (int i = 0; i < iters; ++i) {
        const uint32_t ph = i & 1;
        if (rank == 1) {
          const long long t0 = clk();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&bars->go, 0);
          mbar_wait(mode == 13 ? &bars->back : &bars->mma_done, ph);
          total += clk() - t0;
        } else {
          mbar_wait(&bars->go, ph);
          if (mode == 13) {
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&bars->back, 1);
          } else {
            tc_fence_after();
            const uint64_t a0 = make_smem_desc(smem_u32(smem + kUbOffA), 16, 1024);
            const uint64_t b0 = make_smem_desc(smem_u32(smem + kUbOffB), 16, 1024);
            umma_ss_2cta(tb, a0, b0, make_idesc(true, 256, 128, false, false), 1);
            umma_commit_2cta(&bars->mma_done, 0x3);
            mbar_wait(&bars->mma_done, ph);
          }
        }
      }
      if (rank == 1 && lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)total);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 8) tmem_dealloc_2cta(tb, 512);
}

}  // namespace ba

extern "C" int ba_ubench(int mode, int iters, int grid, int64_t* out4_host, void* stream) {
  using namespace ba;
  BA_REQUIRE(mode >= 0 && mode <= 16 && iters > 0 && iters <= (1 << 20) && grid >= 1 && out4_host,
             "ba_ubench: bad arguments (mode %d, iters %d, grid %d)", mode, iters, grid);
  const bool pair = (mode >= 5 && mode <= 7) || mode == 13 || mode == 15;
  BA_REQUIRE(pair || mode <= 4 || (mode >= 8 && mode <= 12) || mode == 14 || mode == 16, "ba_ubench: unknown mode %d",
             mode);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long* d = nullptr;
  BA_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&d), 4 * sizeof(long long)));
  BA_CHECK_CUDA(cudaMemsetAsync(d, 0, 4 * sizeof(long long), st));
  if (pair) {
    BA_CHECK_CUDA(cudaFuncSetAttribute(ubench2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kUbSmem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * ((grid + 1) / 2));
    cfg.blockDim = dim3(kUbThreads);
    cfg.dynamicSmemBytes = kUbSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    BA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ubench2_kernel, mode, iters, d));
  } else {
    BA_CHECK_CUDA(cudaFuncSetAttribute(ubench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kUbSmem));
    ubench_kernel<<<grid, kUbThreads, kUbSmem, st>>>(mode, iters, d);
    BA_CHECK_CUDA(cudaGetLastError());
  }
  cudaError_t e = cudaStreamSynchronize(st);
  long long h[4] = {0, 0, 0, 0};
  if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  BA_CHECK_CUDA(e);
  for (int i = 0; i < 4; ++i) out4_host[i] = h[i];
  return BA_OK;
}
