// Shared between the forward kernel variants (fwd_sm100.cu, fwd_pair_sm100.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ba {

constexpr int kBlockM = 128;
constexpr int kBlockN = 128;
constexpr int kHeadDim = 128;
constexpr int kKStages = 2;
constexpr int kVStages = 2;
constexpr int kTileBytes = kBlockN * kHeadDim * 2;  // 32 KiB: one 128x128 16-bit tile
constexpr int kBoxBytes = kTileBytes / 2;           // 16 KiB: one 128 x 64 SW128 TMA box
constexpr int kFwdThreads = 320;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units: P stays <= 2^8

struct FwdParams {
  const uint16_t* q;  // raw 16-bit Q view (v3: rows are staged by the softmax threads into TMEM)
  int64_t q_sb, q_ss, q_sh;
  float* o_acc;
  int64_t oacc_sb, oacc_ss, oacc_sh;
  float* lse;
  int64_t lse_sb, lse_sh;
  void* o_out;
  int64_t oout_sb, oout_ss, oout_sh;
  int B, Sq, Sk, H;
  float scale_log2;
  int causal;
  int causal_off;
  int load_state;
  int store_lowp;
};


// CTA-pair forward (fwd_pair_sm100.cu): opt-in, BA_FWD_IMPL=5
int launch_fwd_pair(int dtype, const CUtensorMap& tmK64, const CUtensorMap& tmV, const FwdParams& p,
                    cudaStream_t stream);

// second CTA-pair variant (independent even/odd key streams per warpgroup): opt-in, BA_FWD_IMPL=6
int launch_fwd_pair6(int dtype, const CUtensorMap& tmQ, const CUtensorMap& tmK64, const CUtensorMap& tmV,
                     const FwdParams& p, cudaStream_t stream);

}  // namespace ba
