// State shared by the two ring transports (ring_nccl.cu, ring_ce.cu).
#pragma once
#include "host_common.h"

namespace ba {
typedef struct ncclComm* ncclComm_t;

// Copy-engine transport: ring-owned receive arena (symmetric across ranks), mapped into both
// neighbours with CUDA IPC; flags live in the arena's first page.
struct CeState {
  uint8_t* base = nullptr;      // local arena (cudaMalloc): [0, kCeHeader) flags, then data
  uint8_t* retired = nullptr;   // previous arena, kept until no neighbour can still have it mapped
  int64_t bytes = 0;            // data bytes (without the header)
  uint8_t* next_map = nullptr;  // next rank's arena in this process' address space
  uint8_t* prev_map = nullptr;  // previous rank's arena (== next_map when world == 2)
  uint32_t hop = 0;             // hops posted so far (flag values are hop numbers, monotonic)
  uint32_t* host_vals = nullptr;  // pinned: host_vals[k % kCeVals] = k, source of the 4-byte flag copies
  bool connected = false;
  bool write_value = false;     // BA_CE_FLAG=wv: cuStreamWriteValue32 instead of a 4-byte peer copy
};
constexpr int64_t kCeHeader = 4096;
constexpr int64_t kCeOffReady = 0;      // written by the NEXT rank: "ready to receive hop k"
constexpr int64_t kCeOffArrived = 128;  // written by the PREVIOUS rank: "hop k landed in your arena"
constexpr uint32_t kCeVals = 1u << 16;

int ce_post(struct ::ba_ring* ring, const void* const* src, void* const* dst, const int64_t* nbytes, int n);
void ce_destroy(struct ::ba_ring* ring);
}  // namespace ba

struct ba_ring {
  ba::ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  ba::CeState ce;
};
